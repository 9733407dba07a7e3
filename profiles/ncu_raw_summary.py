#!/usr/bin/env python3
"""Summarise `ncu --page raw --csv` exports (the .ncu-rep files are too large to bring back from the GPU box: they are
exported there, see profiles/r02_call1.sh) with the key list of profiles/ncu_summary.py.
usage: ncu_raw_summary.py name.raw.csv [name.source.csv] ..."""
import csv
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_summary import KEYS  # noqa: E402

EXTRA = [("lts__t_sector_hit_rate.pct", "L2 hit rate"), ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps/cycle"),
         ("sm__inst_executed.avg.per_cycle_active", "IPC active")]


def main():
    for path in sys.argv[1:]:
        if path.endswith(".source.csv"):
            out = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_source_top.py"), path, "12"], capture_output=True, text=True).stdout
            print(out)
            continue
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        print(f"== {os.path.basename(path)}")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            print(f"-- {d['Kernel Name'][:150]}")
            for k, label in KEYS + EXTRA:
                if k in d:
                    print(f"   {label:24s} {d[k]:>18s} {units[hdr.index(k)]}")


if __name__ == "__main__":
    main()
