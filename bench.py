#!/usr/bin/env python3
"""bench.py — headline benchmark of the pattern/ reduce + prefix-sum hot path on B200.

Workload (BASELINE.json configs[1]): pattern/reduction + pattern/prefixsum, fp32 and int64,
--size=2^28 per GPU, iota input (pattern/reduction.cpp:31-34 of the reference).  One *step* is one
pass of the hot path over that batch: 4 libsb_cuda launches

    sb_reduce_f32   (1 GiB algorithmic)      sb_scan_f32 inclusive (2 GiB)
    sb_reduce_i64   (2 GiB)                  sb_scan_i64 inclusive (4 GiB)

Metric: effective HBM GB/s = algorithmic bytes (reduce N*s + s, scan 2*N*s) / time.
`value` is timed with inputs resident in HBM; `e2e` goes through the host-buffer C-ABI calls
(sb_host_reduce / sb_host_scan) from pinned host memory, H2D + D2H inside the timed region.

N > 1 (torchrun, one rank per GPU, weak scaling): every rank owns a 2^28-element shard of a global
N*2^28 problem.  reduction = local sb_reduce + NCCL all-reduce of one element; prefix sum = local
sb_reduce -> NCCL all-gather of the shard totals -> sb_exclusive_carry -> seeded sb_scan (the shard is
read twice and written once: 3 instead of the 4 passes of scan + carry-add).

`--impl reference` times the reference's own CPU implementation on the host cores instead:
oracle/_ref/reduction (the unmodified pattern/reduction.cpp on the host SYCL shim) for the two
reduction legs, and the oracle's OpenMP port for the two prefix-sum legs (upstream prefixsum.cpp
is a TODO stub), on a bounded 2^26-element sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG2_N = 28
GIB = float(1 << 30)
LEGS = (("reduce", "f32"), ("reduce", "i64"), ("scan", "f32"), ("scan", "i64"))
SIZE = {"f32": 4, "i64": 8, "i32": 4, "f64": 8}
METRIC = "pattern reduce/scan effective HBM GB/s"


def algorithmic_bytes(op: str, t: str, n: int) -> int:
    s = SIZE[t]
    return n * s + s if op == "reduce" else 2 * n * s


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def traffic_from_profiles(kernel_key: str):
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kernel_key)
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, f"/tmp/sb_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for line in open(self.path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge

    pkg = ge.load_package()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libsb_cuda has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = pkg.Context(local_rank, torch.cuda.current_stream().cuda_stream)

    n = 1 << args.log2n
    tdt = {"f32": torch.float32, "i64": torch.int64}
    x = {t: torch.empty(n, dtype=tdt[t], device=dev) for t in ("f32", "i64")}
    y = {t: torch.empty(n, dtype=tdt[t], device=dev) for t in ("f32", "i64")}
    red = {t: torch.zeros(1, dtype=tdt[t], device=dev) for t in ("f32", "i64")}
    totals = {t: torch.zeros(world, dtype=tdt[t], device=dev) for t in ("f32", "i64")}
    carry = {t: torch.zeros(1, dtype=tdt[t], device=dev) for t in ("f32", "i64")}
    for t in x:
        ctx.iota(x[t], offset=rank * n)  # global iota: shard r starts at r * n
    torch.cuda.synchronize()

    from sycl_bench_b200 import sharded

    ops = sharded.for_context(ctx, dist, rank, world)

    def step(events=None):
        """one pass of the hot path; `events` collects (leg, start, end) CUDA events"""
        for op, t in LEGS:
            if events is not None:
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
            if op == "reduce":
                ops.reduce(x[t], red[t])  # N>1: + ncclSum all-reduce of one element per rank
            else:
                ops.prefixsum(x[t], y[t], totals[t], carry[t], strategy=args.scan_strategy)
            if events is not None:
                e1.record()
                events.append(((op, t), e0, e1))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launches
    events = []
    t_begin = torch.cuda.Event(enable_timing=True); t_end = torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin.record()
    for _ in range(args.steps):
        step(events)
    t_end.record()
    barrier()
    ms_total = t_begin.elapsed_time(t_end)
    launches = ctx.launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tt = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt.item())
    ms_per_step = ms_total / args.steps

    # correctness of what was timed (integers exact via closed forms; global problem of world * n)
    mism = torch.zeros(1, dtype=torch.int64, device=dev)
    ctx.check_iota(y["i64"], mism, offset=rank * n, kind=0)
    ntot = world * n
    want_sum = (ntot * (ntot - 1) // 2) & ((1 << 64) - 1)
    got_sum = int(red["i64"].item()) & ((1 << 64) - 1)
    ok = int(mism.item()) == 0 and got_sum == want_sum

    per_leg = {}
    for leg, e0, e1 in events:
        per_leg.setdefault(leg, []).append(e0.elapsed_time(e1))
    leg_ms = {f"{op}_{t}": sum(v) / len(v) for (op, t), v in per_leg.items()}

    step_bytes = sum(algorithmic_bytes(op, t, n) for op, t in LEGS) * world
    value = step_bytes / (ms_per_step * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    dom = "scan_i64"
    dom_bytes = algorithmic_bytes("scan", "i64", n)
    achieved = dom_bytes / (leg_ms[dom] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "scan_kernel<u64, plain> (sb_scan_i64, 2^28 inclusive)", "achieved": round(achieved, 1),
                "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic_from_profiles(dom), "algorithmic_bytes_per_launch": dom_bytes,
                "launch_ms": round(leg_ms[dom], 4),
                "per_leg": {k: {"ms": round(v, 4), "GBps": round(algorithmic_bytes(k.split('_')[0], k.split('_')[1], n) / (v * 1e-3) / 1e9, 1),
                                "frac": round(algorithmic_bytes(k.split('_')[0], k.split('_')[1], n) / (v * 1e-3) / 1e9 / peak, 4)} for k, v in leg_ms.items()}}

    # ---- end to end through the host-buffer C ABI (pinned host memory, H2D + D2H timed) --------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, pkg, ctx, torch, dist, world, rank, dev, n)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference(args, sample_log2=args.cpu_log2n, steps=1, warmup=0)["cpu_baseline"]

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32+i64", "data": "synthetic (reference iota: in[i]=(T)i, generated on device)",
            "config": {"workload": f"pattern/reduction + pattern/prefixsum, fp32 and int64, --size=2^{args.log2n} per GPU (BASELINE configs[1])",
                       "elements_per_gpu": n, "legs": [f"{op}_{t}" for op, t in LEGS],
                       "l2": "inputs 1-4 GiB per leg, larger than the 126 MB L2; no flush needed",
                       "parallelism": "1 GPU" if world == 1 else f"{world} contiguous shards; reduce: NCCL all-reduce(1 elt); prefixsum: {args.scan_strategy} with NCCL all-gather of shard totals"},
            "elements_per_s": round(4 * n * world / (ms_per_step * 1e-3), 1),
            "verified": bool(ok),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0 if ok else 1


def run_e2e(args, pkg, ctx, torch, dist, world, rank, dev, n):
    """Same step through sb_host_reduce / sb_host_scan with pinned host buffers."""
    tdt = {"f32": torch.float32, "i64": torch.int64}
    h_in = {t: torch.empty(n, dtype=tdt[t]).pin_memory() for t in ("f32", "i64")}
    h_out = {t: torch.empty(n, dtype=tdt[t]).pin_memory() for t in ("f32", "i64")}
    h_red = {t: torch.zeros(1, dtype=tdt[t]).pin_memory() for t in ("f32", "i64")}
    for t in h_in:  # reference input, built on the host like generate_input()
        h_in[t].copy_(torch.arange(rank * n, rank * n + n, dtype=torch.int64).to(tdt[t]))

    def estep():
        for op, t in LEGS:
            if op == "reduce":
                ctx.host_reduce(h_in[t], h_red[t])
            else:
                ctx.host_scan(h_in[t], h_out[t])

    steps = max(1, min(args.steps, args.e2e_steps))
    estep()  # warm-up (grows the staging buffers)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        estep()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
    step_bytes = sum(algorithmic_bytes(op, t, n) for op, t in LEGS) * world
    h2d = sum(n * SIZE[t] for _, t in LEGS)
    d2h = sum(SIZE[t] if op == "reduce" else n * SIZE[t] for op, t in LEGS)
    lo, hi = rank * n, rank * n + n  # shard-local results (host calls are per shard)
    want = (hi * (hi - 1) // 2 - lo * (lo - 1) // 2) & ((1 << 64) - 1)
    ok = (int(h_out["i64"][-1].item()) & ((1 << 64) - 1)) == want and (int(h_red["i64"].item()) & ((1 << 64) - 1)) == want
    return {"value": round(step_bytes / (dt / steps) / 1e9, 2), "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "ms_per_step": round(dt / steps * 1e3, 2), "steps": steps, "verified": bool(ok),
            "api": "sb_host_reduce / sb_host_scan (C ABI, pinned host buffers; per-rank shards are independent at N>1)"}


# ------------------------------------------------------------------------------------------------------
def cpu_reference(args, sample_log2: int, steps: int, warmup: int):
    """The reference's CPU path on the host cores, bounded sample of the same workload."""
    import numpy as np
    from oracle import oracle as O

    n = 1 << sample_log2
    threads = O.num_threads()
    O.lib()
    have_ref = O.ref_available("reduction")
    x = {"f32": O.iota("f32", n), "i64": O.iota("i64", n)}
    out = {t: np.empty_like(v) for t, v in x.items()}
    name = {"f32": "Pattern_Reduction_Hierarchical_fp32", "i64": "Pattern_Reduction_Hierarchical_int64"}

    def one_step():
        t_leg = {}
        if have_ref:
            # unmodified pattern/reduction.cpp on the host shim; its own run-time (include/common.h:69-81)
            rows, _ = O.run_reference("reduction", n, 256, num_runs=1, ndrange=False, verification=False, dump=False)
            for t in ("f32", "i64"):
                t_leg[f"reduce_{t}"] = float(rows[name[t]]["run-time-min"])
        else:
            for t in ("f32", "i64"):
                t0 = time.perf_counter(); O.reduce(t, x[t], 256); t_leg[f"reduce_{t}"] = time.perf_counter() - t0
        for t in ("f32", "i64"):
            t0 = time.perf_counter(); O.cpu_scan(t, x[t], out[t]); t_leg[f"scan_{t}"] = time.perf_counter() - t0
        return t_leg

    for _ in range(warmup):
        one_step()
    acc = []
    for _ in range(steps):
        acc.append(one_step())
    leg = {k: sum(a[k] for a in acc) / len(acc) for k in acc[0]}
    sec = sum(leg.values())
    step_bytes = sum(algorithmic_bytes(op, t, n) for op, t in LEGS)
    value = step_bytes / sec / 1e9
    cpu = {"value": round(value, 3), "unit": "GB/s", "cores": threads, "kind": "reference" if have_ref else "port",
           "sample": f"same 4 legs at 2^{sample_log2} elements (bounded sample of the 2^{LOG2_N} workload; GB/s is size-independent); "
                     + ("reduction legs: unmodified reference pattern/reduction.cpp on the host SYCL shim (oracle/_ref, OpenMP), its own run-time; "
                        if have_ref else "reduction legs: oracle restatement of the reference tree (OpenMP); ")
                     + "prefixsum legs: oracle OpenMP two-pass scan (upstream prefixsum.cpp is a TODO stub)",
           "leg_seconds": {k: round(v, 4) for k, v in leg.items()}}
    return {"cpu_baseline": cpu, "seconds_per_step": sec, "value": value}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    r = cpu_reference(args, sample_log2=args.cpu_log2n, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(r["value"], 3), "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(r["seconds_per_step"] * 1e3, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32+i64", "data": "synthetic (reference iota)",
        "config": {"workload": f"pattern/reduction + pattern/prefixsum, fp32 and int64, CPU sample 2^{args.cpu_log2n} of --size=2^{LOG2_N} (BASELINE configs[1])",
                   "legs": [f"{op}_{t}" for op, t in LEGS]},
        "cpu_baseline": r["cpu_baseline"],
        "e2e": {"value": round(r["value"], 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2n", type=int, default=LOG2_N, help="elements per GPU = 2^log2n (default: BASELINE configs[1])")
    ap.add_argument("--cpu-log2n", type=int, default=26, help="CPU-baseline sample size")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--scan-strategy", default="reduce_then_scan", choices=["reduce_then_scan", "scan_then_carry"],
                    help="N>1 prefix sum: 3-pass reduce -> exchange -> seeded scan (default) or the 4-pass scan -> exchange -> carry-add")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            # convenience: re-exec under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                   "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
            return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
