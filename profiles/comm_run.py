"""Times sb_comm_reduce / sb_comm_scan (single process, N GPUs, include/sb_cuda.h) with the rank-0 timer.

    python profiles/comm_run.py --gpus 2 [--log2n 28] [--reps 10] > gpurun_out/comm.jsonl

One JSON line per (op, dtype, strategy): best and median ms over --reps, aggregate algorithmic GB/s
(reduce: N*s, scan: 2*N*s over ALL shards).  n = 2^log2n elements PER GPU (weak scaling, like bench.py).
"""
import argparse
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

TORCH = {"i32": torch.int32, "i64": torch.int64, "f32": torch.float32, "f64": torch.float64}
SIZE = {"i32": 4, "i64": 8, "f32": 4, "f64": 8}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
    ap.add_argument("--log2n", type=int, default=28)
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    pkg = load_package()
    ctx = pkg.Context(0, pkg.Context.OWNED_STREAM)
    comm = pkg.Comm(ctx, args.gpus)
    per = 1 << args.log2n
    for dtype in ("i32", "i64", "f32", "f64"):
        shards, outs, totals = [], [], []
        for r in range(args.gpus):
            x = torch.empty(per, dtype=TORCH[dtype], device=f"cuda:{r}")
            y = torch.empty_like(x)
            t = torch.zeros(1, dtype=TORCH[dtype], device=f"cuda:{r}")
            torch.cuda.synchronize(r)
            comm.rank_ctx(r).iota(x, offset=r * per)
            comm.rank_ctx(r).sync()
            shards.append(x), outs.append(y), totals.append(t)
        cases = [("reduce", None), ("scan", 0), ("scan", 1)]
        for op, strategy in cases:
            def step():
                if op == "reduce":
                    comm.reduce(shards, totals, dtype=dtype)
                else:
                    comm.scan(shards, outs, strategy=strategy, dtype=dtype)
            for _ in range(3):
                step()
            ctx.sync()
            ms = []
            for _ in range(args.reps):
                ctx.timer_begin()
                step()
                ms.append(ctx.timer_end())
            n_total = per * args.gpus
            alg = n_total * SIZE[dtype] * (1 if op == "reduce" else 2)
            rec = {
                "op": op, "dtype": dtype, "gpus": args.gpus, "n_per_gpu": per,
                "strategy": None if strategy is None else ("reduce_then_scan", "scan_then_carry")[strategy],
                "best_ms": round(min(ms), 4), "median_ms": round(statistics.median(ms), 4),
                "GBps_best": round(alg / min(ms) / 1e6, 1), "GBps_median": round(alg / statistics.median(ms) / 1e6, 1),
            }
            print(json.dumps(rec), flush=True)
        del shards, outs, totals
        torch.cuda.empty_cache()
    comm.close()
    ctx.close()


if __name__ == "__main__":
    main()
