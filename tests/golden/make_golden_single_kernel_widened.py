#!/usr/bin/env python3
"""Generate tests/golden/single_kernel_widened_golden.json from the UNMODIFIED reference.

Runs oracle/_ref/{lin_reg_error,kmeans} — /root/reference/single-kernel/{lin_reg_error,kmeans}.cpp compiled as they lie
against the host-only SYCL shim (oracle/Makefile) — and records what their verify() reads back:
  lin_reg_error  the output buffer (lin_reg_error.cpp:84): every element up to 1024, else 256 evenly spaced ones
  kmeans         the membership buffer (kmeans.cpp:85): sha256 + first / last elements
Run here, in the container that mounts /root/reference:

    make -C oracle ref && python tests/golden/make_golden_single_kernel_widened.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

NAME = {"f32": "fp32", "f64": "fp64"}
LINREG_ERROR_SIZES = [256, 1000, 1024, 3072, 4096, 16384, 65536]  # 65536 = upstream's suite size (bin/run-suite:70-72)
KMEANS_SIZES = [1024, 3072, 65536, 1 << 20]                        # 2^20 = upstream's suite size (bin/run-suite:64-66)


def main():
    out = {"generator": "tests/golden/make_golden_single_kernel_widened.py",
           "reference": "unisa-hpc/sycl-bench 2.0.4 single-kernel/{lin_reg_error,kmeans}.cpp on oracle/shim",
           "lin_reg_error": [], "kmeans": []}
    for size in LINREG_ERROR_SIZES:
        rows, dumps = O.run_reference("lin_reg_error", size, 256, num_runs=1, ndrange=True, timeout=3600)
        assert len(dumps) == 2, (size, len(dumps))
        for k, t in enumerate(("f32", "f64")):
            res = np.frombuffer(dumps[k], dtype=O.NP[t])
            assert res.size == size
            idx = np.arange(size) if size <= 1024 else np.arange(0, size, size // 256)
            out["lin_reg_error"].append({"size": size, "dtype": t, "rand_skip": 0 if t == "f32" else 4 * size,
                                         "indices": "all" if size <= 1024 else [int(i) for i in idx],
                                         "values_hex": res[idx].tobytes().hex(),
                                         "reference_verify": rows[f"LinearRegression_{NAME[t]}"]["Verification"]})
    for size in KMEANS_SIZES:
        rows, dumps = O.run_reference("kmeans", size, 256, num_runs=1, ndrange=True, timeout=3600)
        assert len(dumps) == 2, (size, len(dumps))
        for k, t in enumerate(("f32", "f64")):
            m = np.frombuffer(dumps[k], dtype=np.int32)
            assert m.size == size
            out["kmeans"].append({"size": size, "dtype": t, "sha256": hashlib.sha256(m.tobytes()).hexdigest(),
                                  "first": [int(v) for v in m[:4]], "last": [int(v) for v in m[-4:]],
                                  "reference_verify": rows[f"Kmeans_{NAME[t]}"]["Verification"]})
    path = os.path.join(ROOT, "tests", "golden", "single_kernel_widened_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(f"wrote {path}: {len(out['lin_reg_error'])} lin_reg_error + {len(out['kmeans'])} kmeans vectors")


if __name__ == "__main__":
    main()
