#!/usr/bin/env python3
"""Generate tests/golden/reference_golden.json from the UNMODIFIED reference.

Runs oracle/_ref/{reduction,segmentedreduction} — /root/reference/pattern/*.cpp compiled as they
lie against the host-only SYCL shim (oracle/Makefile) — and records the RESULT VALUES the reference
computes (it never prints them; the shim dumps the buffer that verify() reads).  Run here, in the
container that mounts /root/reference:

    make -C oracle ref && python tests/golden/make_golden.py

The JSON travels to the GPU box; tests/test_oracle.py pins oracle/sb_oracle.c to it and
tests/test_parity_gpu.py pins libsb_cuda to it.
"""
import hashlib
import json
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

FMT = {"i16": "<h", "i32": "<i", "i64": "<q", "f32": "<f", "f64": "<d"}

REDUCTION_CASES = [  # (size, local, also run the NDRange variant through fibers)
    (3072, 256, True), (65536, 256, True), (4096, 64, True), (1 << 20, 256, False), (1 << 20, 1024, False),
    (1 << 24, 256, False), (1 << 22, 128, False),
    (1000, 100, True),      # local not a power of two: the halving tree skips elements
    (3000, 256, True),      # size % local != 0: the tail is dropped (assert compiled out)
    (1 << 16, 96, False),
]
SEGRED_CASES = [(1024, 256, True), (4096, 64, True), (1 << 16, 1024, False), (1 << 20, 256, False), (1 << 20, 64, False),
                (1 << 18, 512, False), (1200, 100, True), (1 << 22, 128, False), (1 << 24, 256, False)]


def hexbits(t, raw):
    return raw[: struct.calcsize(FMT[t])].hex()


def main():
    out = {"generator": "tests/golden/make_golden.py", "reference": "unisa-hpc/sycl-bench 2.0.4 (pattern/reduction.cpp, pattern/segmentedreduction.cpp) on oracle/shim",
           "reduction": [], "segmentedreduction": []}
    for size, local, nd in REDUCTION_CASES:
        rows, dumps = O.run_reference("reduction", size, local, num_runs=1, ndrange=nd, dump_max_bytes=8)
        names = list(rows.keys())
        types = O.REF_REDUCTION_TYPES * (2 if nd else 1)
        assert len(dumps) == len(types), (len(dumps), names)
        variants = (["NDRange"] * 4 + ["Hierarchical"] * 4) if nd else ["Hierarchical"] * 4
        for t, v, raw in zip(types, variants, dumps):
            name = f"Pattern_Reduction_{v}_" + {"i32": "int32", "i64": "int64", "f32": "fp32", "f64": "fp64"}[t]
            out["reduction"].append({"size": size, "local": local, "variant": v, "dtype": t,
                                     "result_hex": hexbits(t, raw), "result": struct.unpack(FMT[t], raw[: struct.calcsize(FMT[t])])[0],
                                     "reference_verify": rows[name]["Verification"]})
    for size, local, nd in SEGRED_CASES:
        rows, dumps = O.run_reference("segmentedreduction", size, local, num_runs=1, ndrange=nd)
        types = O.REF_SEGREDUCTION_TYPES * (2 if nd else 1)
        assert len(dumps) == len(types), (len(dumps), list(rows))
        variants = (["NDRange"] * 5 + ["Hierarchical"] * 5) if nd else ["Hierarchical"] * 5
        for t, v, raw in zip(types, variants, dumps):
            s = struct.calcsize(FMT[t])
            assert len(raw) == size * s
            heads = [raw[g * local * s: g * local * s + s].hex() for g in range(min(4, size // local))]
            name = f"Pattern_SegmentedReduction_{v}_" + {"i16": "int16", "i32": "int32", "i64": "int64", "f32": "fp32", "f64": "fp64"}[t]
            out["segmentedreduction"].append({"size": size, "local": local, "variant": v, "dtype": t,
                                              "sha256": hashlib.sha256(raw).hexdigest(), "first_heads_hex": heads,
                                              "last_head_hex": raw[(size - local) * s: (size - local) * s + s].hex(),
                                              "reference_verify": rows[name]["Verification"]})
    path = os.path.join(ROOT, "tests", "golden", "reference_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(f"wrote {path}: {len(out['reduction'])} reduction + {len(out['segmentedreduction'])} segmentedreduction vectors")


if __name__ == "__main__":
    main()
