#!/usr/bin/env python3
"""Per-kernel SASS evidence of libsb_cuda.so: counts of the mnemonics that prove the Blackwell data path
(B200_PROFILING.md: TMA = UBLKCP / UTMALDG / UTMASTG, mbarrier = SYNCS.*, redux.sync = REDUX, 128-bit strong stores of the
tile descriptors, system-scope mailbox accesses, L2 cache-policy operands).
usage: python profiles/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sycl-bench_b200", "lib", "libsb_cuda.so")
PATTERNS = [("UBLKCP", r"\bUBLKCP"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("SYNCS.*", r"\bSYNCS\."), ("REDUX", r"\bREDUX"),
            ("SHFL", r"\bSHFL\."), ("LDG.E.128", r"\bLDG\.E\.[A-Z.]*128"), ("STG.E.128", r"\bSTG\.E\.[A-Z.]*128"),
            ("STG.E.128.STRONG.GPU", r"\bSTG\.E\.128\.STRONG\.GPU"), ("LDG/STG .STRONG.SYS", r"\b(LDG|STG)\.E\.[A-Z0-9.]*STRONG\.SYS"),
            ("LDS.128", r"\bLDS\.128"), ("ATOMG/RED", r"\b(ATOMG|RED)\."), ("HMMA/UTC*MMA", r"\b(HMMA|UTC[A-Z]*MMA)")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for name, pat in PATTERNS:
            if re.search(pat, line):
                counts[cur][name] += 1
    names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print(f"# cuobjdump -sass sycl-bench_b200/lib/libsb_cuda.so: {len(counts)} kernels, cubin architectures {arch}")
    total = collections.Counter()
    for c in counts.values():
        total.update(c)
    print("# whole library: " + ", ".join(f"{k} {total[k]}" for k, _ in PATTERNS))
    print("# (no tensor-core instruction is expected: none of these paths is a contraction)\n")
    agg = collections.OrderedDict()
    for mangled, dem in zip(counts, names):
        short = re.sub(r"\(anonymous namespace\)::", "", dem)
        short = re.sub(r"^void ", "", short).split("(")[0]
        fam = short.split("<")[0]
        agg.setdefault(fam, []).append((short, counts[mangled]))
    for fam, items in agg.items():
        tot = collections.Counter()
        for _, c in items:
            tot.update(c)
        print(f"{fam}  ({len(items)} instantiations)")
        print("    " + ", ".join(f"{k} {tot[k]}" for k, _ in PATTERNS if tot[k]))
        for short, c in items[:3]:
            print(f"      e.g. {short[:90]}: " + ", ".join(f"{k} {c[k]}" for k, _ in PATTERNS if c[k]))


if __name__ == "__main__":
    main()
