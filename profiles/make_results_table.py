#!/usr/bin/env python3
"""Turn a run-suite CSV (reference schema) into the BASELINE.md results table."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.oracle import parse_csv  # CSV parsing only


def main():
    path = sys.argv[1]
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6453.7
    rows = []
    header, cur = None, []
    # the CSV repeats its header per benchmark and benchmark names repeat across invocations: parse block-wise
    import csv
    for line in open(path):
        line = line.rstrip("\n")
        if line.startswith("# Benchmark name"):
            header = line.split(",")[1:]
            continue
        cells = next(csv.reader([line]))
        rows.append((cells[0], dict(zip(header, cells[1:]))))
    print("| benchmark row | N | L | kernel-time-min | elements/s | GB/s (algorithmic) | %% of 8000 GB/s | %% of measured %.1f GB/s | Verification |" % peak)
    print("|---|---|---|---|---|---|---|---|---|")
    for name, r in rows:
        n, L = int(r["problem-size"]), int(r["local-size"])
        t = int(r["kernel-time-min-ns"]) * 1e-9
        gib = float(r["throughput-metric"])
        gbps = gib * 1.073741824 / t
        print(f"| {name} | 2^{n.bit_length()-1} | {L} | {t*1e6:.1f} us | {n/t:.3e} | {gbps:.0f} | {gbps/80:.1f} | {gbps/peak*100:.1f} | {r['Verification']} |")


if __name__ == "__main__":
    main()
