#!/usr/bin/env python3
"""Top stalled SASS instructions of an `ncu --page source --csv` export (the per-instruction view behind the stall
breakdown quoted in DESIGN.md).  usage: ncu_source_top.py file.source.csv [N]"""
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        try:
            samples = int(r[idx["# Samples"]] or 0)
        except ValueError:
            continue
        data.append((samples, r))
    total = sum(s for s, _ in data) or 1
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    print(f"{path}: {len(data)} SASS instructions, {total} warp samples")
    agg = {h: sum(int(r[idx[h]] or 0) for _, r in data) for h in stall_cols}
    print("  stall samples: " + ", ".join(f"{h[6:]} {100 * v / total:.1f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 200 > total))
    data.sort(key=lambda t: -t[0])
    for s, r in data[:top]:
        reasons = sorted(((int(r[idx[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
        print(f"  {100 * s / total:5.1f}%  {r[idx['Address']][-5:]}  {r[idx['Source']][:90]:90s} {reasons[0][1]}:{reasons[0][0]} {reasons[1][1]}:{reasons[1][0]}")


if __name__ == "__main__":
    main()
