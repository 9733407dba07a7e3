#!/usr/bin/env python3
"""Kernel-only timing of the libsb_cuda entry points (CUDA events, inputs larger than L2).

    python profiles/tune_scan.py                 # the in-tree lib, all ops
    python profiles/tune_scan.py --variants      # every lib/tune/libsb_cuda_*.so built by `make tune`

Prints one line per (lib, op, dtype): best / median ms and algorithmic GB/s.
"""
import argparse
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(args):
    import torch
    import __graft_entry__ as ge

    pkg = ge.load_package()
    torch.cuda.set_device(0)
    ctx = pkg.Context(0, torch.cuda.current_stream().cuda_stream)
    tdt = {"i16": torch.int16, "i32": torch.int32, "i64": torch.int64, "f32": torch.float32, "f64": torch.float64}
    size = {"i16": 2, "i32": 4, "i64": 8, "f32": 4, "f64": 8}
    n = 1 << args.log2n
    res = []
    for t in args.dtypes.split(","):
        x = torch.empty(n, dtype=tdt[t], device="cuda")
        y = torch.empty_like(x)
        out = torch.zeros(1, dtype=tdt[t], device="cuda")
        ctx.iota(x)
        if t == "i16":  # the C ABI has int16 for the segmented reduction only
            cases = {}
            for L in (64, 128, 256, 1024):
                cases[f"segreduce{L}"] = ((lambda L=L: ctx.segreduce(y, L)), n * 2 + n // L * 2)
        else:
            cases = {
            "reduce": (lambda: ctx.reduce(x, out), n * size[t]),
            "scan": (lambda: ctx.scan(x, y), 2 * n * size[t]),
            "scan_cyclic": (lambda: ctx.scan_cyclic(x, y), 2 * n * size[t]),
            "scan_group256": (lambda: ctx.scan(x, y, exclusive=True, group_len=256), 2 * n * size[t]),
            "segscan_L65536": (lambda: ctx.segscan(x, y, seg_len=65536), 2 * n * size[t]),
            "segscan_L16384": (lambda: ctx.segscan(x, y, seg_len=16384), 2 * n * size[t]),
            "segscan_L1M": (lambda: ctx.segscan(x, y, seg_len=1 << 20), 2 * n * size[t]),
            "segscan_L50000": (lambda: ctx.segscan(x, y, seg_len=50000), 2 * n * size[t]),
            "segreduce256": (lambda: ctx.segreduce(y, 256), n * size[t] + n // 256 * size[t]),
            "segreduce64": (lambda: ctx.segreduce(y, 64), n * size[t] + n // 64 * size[t]),
            "segreduce128": (lambda: ctx.segreduce(y, 128), n * size[t] + n // 128 * size[t]),
            "segreduce1024": (lambda: ctx.segreduce(y, 1024), n * size[t] + n // 1024 * size[t]),
            "carry_add": (lambda: ctx.carry_add(y, out), 2 * n * size[t]),
            "iota": (lambda: ctx.iota(y), n * size[t]),
        }
        if args.flags and t != "i16":
            fl = torch.zeros(n, dtype=torch.uint8, device="cuda")
            fl[::256] = 1
            cases["segscan_flags256"] = (lambda: ctx.segscan(x, y, head_flags=fl), 2 * n * size[t] + n)
            # heads at random positions, mean segment length 256: no tile starts on a head
            fr = (torch.rand(n, device="cuda") < 1.0 / 256).to(torch.uint8)
            cases["segscan_flagsrand"] = (lambda: ctx.segscan(x, y, head_flags=fr), 2 * n * size[t] + n)
            for dist in (32, 1024, 4096, 65536):
                if f"segscan_flagsrand{dist}" in args.ops.split(","):
                    frd = (torch.rand(n, device="cuda") < 1.0 / dist).to(torch.uint8)
                    cases[f"segscan_flagsrand{dist}"] = ((lambda frd=frd: ctx.segscan(x, y, head_flags=frd)), 2 * n * size[t] + n)
        for name in args.ops.split(","):
            if name not in cases:
                continue
            fn, nbytes = cases[name]
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ms = []
            for _ in range(args.iters):
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                e1.synchronize()
                ms.append(e0.elapsed_time(e1))
            ms.sort()
            res.append({"lib": os.path.basename(pkg.lib_path()), "env": {k: v for k, v in os.environ.items() if k.startswith("SB_")}, "op": name, "dtype": t, "log2n": args.log2n, "ms_min": round(ms[0], 4),
                        "ms_med": round(ms[len(ms) // 2], 4), "GBps_min": round(nbytes / ms[0] / 1e6, 1), "GBps_med": round(nbytes / ms[len(ms) // 2] / 1e6, 1)})
        del x, y
    for r in res:
        print(json.dumps(r), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", action="store_true")
    ap.add_argument("--worker", action="store_true")
    ap.add_argument("--log2n", type=int, default=28)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--dtypes", default="i32,i64,f32,f64")
    ap.add_argument("--ops", default="reduce,scan,scan_group256,segscan_L65536,segscan_flags256,segreduce64,segreduce256,segreduce1024,carry_add,iota")
    ap.add_argument("--flags", action="store_true", default=True)
    args = ap.parse_args()
    if args.worker:
        return worker(args)
    libs = [None]
    if args.variants:
        libs = sorted(glob.glob(os.path.join(ROOT, "sycl-bench_b200", "lib", "tune", "libsb_cuda_*.so")))
    for lib in libs:
        env = dict(os.environ)
        if lib:
            env["SB_CUDA_LIB"] = lib
        cmd = [sys.executable, os.path.abspath(__file__), "--worker", "--log2n", str(args.log2n), "--iters", str(args.iters), "--dtypes", args.dtypes, "--ops", args.ops]
        subprocess.run(cmd, env=env, check=False)


if __name__ == "__main__":
    main()
