#!/usr/bin/env python3
"""Generate tests/golden/single_kernel_golden.json from the UNMODIFIED reference.

Runs oracle/_ref/{scalar_prod,lin_reg_coeff} — /root/reference/single-kernel/{scalar_prod,lin_reg_coeff}.cpp
compiled as they lie against the host-only SYCL shim (oracle/Makefile) — on ONE OpenMP thread (both
benchmarks overlap reads and writes between work-groups of a launch; one thread = ascending group order,
see oracle/sb_oracle_single_kernel.c) and records the values they compute:
  scalar_prod    output[0] after run() (the buffer verify() reads, scalar_prod.cpp:186)
  lin_reg_coeff  the four values reduce() reads back (lin_reg_coeff.cpp:119-120): sum xy, sum xx, sum x, sum y
Run here, in the container that mounts /root/reference:

    make -C oracle ref && python tests/golden/make_golden_single_kernel.py
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

# (size, local, also run the NDRange variant through fibers)
SCALAR_PROD_CASES = [(3072, 256, True), (65536, 256, True), (4096, 64, True), (1000, 100, True), (1 << 20, 256, False),
                     (1 << 20, 1024, False), (1 << 22, 128, False), (1 << 24, 256, False), (1 << 25, 256, False)]
LINREG_CASES = [(1024, 256), (4096, 64), (65536, 256), (1 << 18, 128), (1 << 20, 256), (1 << 20, 1024)]


def value(t, raw):
    s = struct.calcsize(FMT[t])
    return raw[:s].hex(), struct.unpack(FMT[t], raw[:s])[0]


def main():
    out = {"generator": "tests/golden/make_golden_single_kernel.py",
           "reference": "unisa-hpc/sycl-bench 2.0.4 single-kernel/{scalar_prod,lin_reg_coeff}.cpp on oracle/shim, OMP_NUM_THREADS=1",
           "scalar_prod": [], "lin_reg_coeff": []}
    for size, local, nd in SCALAR_PROD_CASES:
        rows, dumps = O.run_reference("scalar_prod", size, local, num_runs=1, ndrange=nd, dump_max_bytes=8, threads=1, timeout=1800)
        variants = (["NDRange"] * 4 + ["Hierarchical"] * 4) if nd else ["Hierarchical"] * 4
        types = O.REF_REDUCTION_TYPES * (2 if nd else 1)
        assert len(dumps) == len(types), (size, local, len(dumps))
        for t, v, raw in zip(types, variants, dumps):
            hx, val = value(t, raw)
            out["scalar_prod"].append({"size": size, "local": local, "variant": v, "dtype": t, "result_hex": hx, "result": val,
                                       "reference_verify": rows[f"ScalarProduct_{v}_{NAME[t]}"]["Verification"]})
    for size, local in LINREG_CASES:
        rows, dumps = O.run_reference("lin_reg_coeff", size, local, num_runs=1, ndrange=True, dump_max_bytes=8, threads=1, timeout=1800)
        assert len(dumps) == 8, (size, local, len(dumps))
        for k, t in enumerate(("f32", "f64")):
            sums = [value(t, raw) for raw in dumps[4 * k: 4 * k + 4]]
            out["lin_reg_coeff"].append({"size": size, "local": local, "dtype": t,
                                         "sums_hex": [h for h, _ in sums], "sums": [v for _, v in sums],
                                         "reference_verify": rows[f"LinearRegressionCoeff_{NAME[t]}"]["Verification"]})
    path = os.path.join(ROOT, "tests", "golden", "single_kernel_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(f"wrote {path}: {len(out['scalar_prod'])} scalar_prod + {len(out['lin_reg_coeff'])} lin_reg_coeff vectors")


if __name__ == "__main__":
    main()
