#!/usr/bin/env python3
"""Pretty-print the JSON lines of profiles/tune_scan.py."""
import json
import sys

for path in sys.argv[1:]:
    for l in open(path):
        try:
            r = json.loads(l)
        except Exception:
            print(l.rstrip())
            continue
        env = " ".join(f"{k}={v}" for k, v in sorted(r.get("env", {}).items()) if k != "SB_CUDA_LIB")
        print(f"{r['lib']:30s} {env:22s} {r['op']:18s} {r['dtype']} {r['ms_min']:8.4f} ms {r['GBps_min']:8.1f} GB/s (med {r['GBps_med']:.1f})")
