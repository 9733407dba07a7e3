#!/usr/bin/env python3
"""Generate tests/golden/sycl2020_golden.json from the UNMODIFIED reference reduction trio.

Runs oracle/_ref/{atomic_reduction,reduce_over_group,kernel_reduction} — the translation units
/root/reference/sycl2020/{atomics,group_algorithms,kernel_reduction}/*.cpp compiled as they lie against
the host-only SYCL shim (oracle/Makefile) — and records the value each benchmark leaves in out[0]
(the shim dumps the buffer verify() reads) together with upstream's own verdict.  Run here, in the
container that mounts /root/reference:

    make -C oracle ref && python tests/golden/make_golden_sycl2020.py
"""
import json
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

FMT = {"i32": "<i", "i64": "<q", "f32": "<f", "f64": "<d"}
NAME = {"i32": "int32", "i64": "int64", "f32": "fp32", "f64": "fp64"}

# (size, local); 2^25 pins the fp32 saturation of ReductionAtomic at 2^24
ND_CASES = [(3072, 256), (65536, 256), (1 << 20, 256), (1 << 20, 1024), (1000, 100), (4096, 64), (1 << 25, 256)]
# kernel_reduction ignores --local; sizes that are not multiples of 4 / 8 drop a tail
KR_CASES = [3072, 65536, 1 << 20, 1 << 24, 1000003, 4099, 7]


def value(t, raw):
    s = struct.calcsize(FMT[t])
    return raw[:s].hex(), struct.unpack(FMT[t], raw[:s])[0]


def main():
    out = {"generator": "tests/golden/make_golden_sycl2020.py",
           "reference": "unisa-hpc/sycl-bench 2.0.4 sycl2020/{atomics/atomic_reduction,group_algorithms/reduce_over_group,"
                        "kernel_reduction/kernel_reduction}.cpp on oracle/shim",
           "atomic_reduction": [], "reduce_over_group": [], "kernel_reduction": []}
    for which, prefix in (("atomic_reduction", "ReductionAtomic_"), ("reduce_over_group", "ReduceGroupAlgorithm_")):
        for size, local in ND_CASES:
            rows, dumps = O.run_reference(which, size, local, num_runs=1, ndrange=True, dump_max_bytes=8, timeout=1800)
            assert len(dumps) == 4, (which, size, len(dumps))
            for t, raw in zip(O.REF_REDUCTION_TYPES, dumps):
                hx, val = value(t, raw)
                out[which].append({"size": size, "local": local, "dtype": t, "result_hex": hx, "result": val,
                                   "reference_verify": rows[prefix + NAME[t]]["Verification"]})
    for size in KR_CASES:
        rows, dumps = O.run_reference("kernel_reduction", size, 256, num_runs=1, ndrange=True, dump_max_bytes=8)
        assert len(dumps) == 12, (size, len(dumps))
        k = 0
        for t in O.REF_REDUCTION_TYPES:
            for cf in (1, 4, 8):
                hx, val = value(t, dumps[k])
                k += 1
                out["kernel_reduction"].append({"size": size, "cf": cf, "dtype": t, "result_hex": hx, "result": val,
                                                "reference_verify": rows[f"KernelReduction_{NAME[t]}_plus_cf{cf}"]["Verification"]})
    path = os.path.join(ROOT, "tests", "golden", "sycl2020_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(f"wrote {path}: " + ", ".join(f"{len(out[k])} {k}" for k in ("atomic_reduction", "reduce_over_group", "kernel_reduction")))


if __name__ == "__main__":
    main()
