#!/usr/bin/env python3
"""bench.py — headline benchmark of the pattern/ reduce + prefix-sum hot path on B200.

Default workload = BASELINE.json configs[1]: pattern/reduction + pattern/prefixsum, fp32 and int64,
--size=2^28 per GPU, iota input (pattern/reduction.cpp:31-34 of the reference).  One *step* is one
pass of the hot path over that batch: 4 libsb_cuda launches

    sb_reduce_f32   (1 GiB algorithmic)      sb_scan_f32 inclusive (2 GiB)
    sb_reduce_i64   (2 GiB)                  sb_scan_i64 inclusive (4 GiB)

Metric: effective HBM GB/s = algorithmic bytes (reduce N*s + s, scan 2*N*s) / time.
`value` is timed with inputs resident in HBM; `e2e` goes through the host-buffer C-ABI calls
(sb_host_reduce / sb_host_scan) from pinned host memory, H2D + D2H inside the timed region.

N > 1 (torchrun, one rank per GPU): every rank owns a contiguous shard.  reduction = local sb_reduce
+ NCCL all-reduce of one element; prefix sum = local sb_reduce -> NCCL all-gather of the shard totals
-> sb_exclusive_carry -> seeded sb_scan (the shard is read twice and written once: 3 instead of the 4
passes of scan + carry-add; --scan-strategy scan_then_carry runs the 4-pass recipe).
`--workload config2` (default): weak scaling, 2^28 elements per GPU.
`--workload config5`: BASELINE configs[4], reduction + prefixsum int32, 2^33 elements in total
(strong scaling: 2^33 / N per GPU).

`--impl reference` times the reference's own CPU implementation on the host cores instead:
oracle/_ref/reduction (the unmodified pattern/reduction.cpp on the host SYCL shim) for the
reduction legs, and the oracle's OpenMP port for the prefix-sum legs (upstream prefixsum.cpp is a
TODO stub), on a bounded 2^26-element sample.
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

SIZE = {"f32": 4, "i64": 8, "i32": 4, "f64": 8}
METRIC = "pattern reduce/scan effective HBM GB/s"
WORKLOADS = {
    "config2": {"legs": (("reduce", "f32"), ("reduce", "i64"), ("scan", "f32"), ("scan", "i64")), "log2_per_gpu": 28, "scaling": "weak",
                "name": "pattern/reduction + pattern/prefixsum, fp32 and int64, --size=2^28 per GPU (BASELINE configs[1])"},
    "config5": {"legs": (("reduce", "i32"), ("scan", "i32")), "log2_total": 33, "scaling": "strong",
                "name": "pattern/reduction + pattern/prefixsum, int32, --size=2^33 in total, sharded (BASELINE configs[4])"},
}
REF_ROW = {"f32": "Pattern_Reduction_Hierarchical_fp32", "i64": "Pattern_Reduction_Hierarchical_int64",
           "i32": "Pattern_Reduction_Hierarchical_int32", "f64": "Pattern_Reduction_Hierarchical_fp64"}


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
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
            time.sleep(0.3)  # let the sampler come up before the timed region starts
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
    from sycl_bench_b200 import sharded

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

    wl = WORKLOADS[args.workload]
    legs = wl["legs"]
    if args.log2n:
        n = 1 << args.log2n
    elif "log2_per_gpu" in wl:
        n = 1 << wl["log2_per_gpu"]
    else:
        n = (1 << wl["log2_total"]) // world
    first = rank * n  # global index of this shard's first element
    types = sorted({t for _, t in legs})
    tdt = {"f32": torch.float32, "i64": torch.int64, "i32": torch.int32, "f64": torch.float64}
    x = {t: torch.empty(n, dtype=tdt[t], device=dev) for t in types}
    y = {t: torch.empty(n, dtype=tdt[t], device=dev) for t in types}
    red = {t: torch.zeros(1, dtype=tdt[t], device=dev) for t in types}
    totals = {t: torch.zeros(world, dtype=tdt[t], device=dev) for t in types}
    carry = {t: torch.zeros(1, dtype=tdt[t], device=dev) for t in types}
    for t in types:
        ctx.iota(x[t], offset=first)
    torch.cuda.synchronize()
    ops = sharded.for_context(ctx, dist, rank, world)

    def step(events=None):
        """one pass of the hot path; `events` collects (leg, start, end) CUDA events"""
        for op, t in legs:
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

    # correctness of what was timed: integer legs exactly, through closed forms of the GLOBAL problem
    ntot = world * n
    ok = True
    for op, t in legs:
        if t[0] != "i":
            continue
        bits = 8 * SIZE[t]
        if op == "reduce":
            want = (ntot * (ntot - 1) // 2) & ((1 << bits) - 1)
            ok = ok and (int(red[t].item()) & ((1 << bits) - 1)) == want
        else:
            mism = torch.zeros(1, dtype=torch.int64, device=dev)
            ctx.check_iota(y[t], mism, offset=first, kind=0)
            ok = ok and int(mism.item()) == 0
    if world > 1:
        okt = torch.tensor([1 if ok else 0], dtype=torch.int64, device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        ok = bool(okt.item())

    per_leg = {}
    for leg, e0, e1 in events:
        per_leg.setdefault(leg, []).append(e0.elapsed_time(e1))
    leg_ms = {f"{op}_{t}": sum(v) / len(v) for (op, t), v in per_leg.items()}

    step_bytes = sum(algorithmic_bytes(op, t, n) for op, t in legs) * world
    value = step_bytes / (ms_per_step * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    dom = max((f"{op}_{t}" for op, t in legs), key=lambda k: leg_ms[k])
    dom_op, dom_t = dom.split("_")
    dom_bytes = algorithmic_bytes(dom_op, dom_t, n)
    achieved = dom_bytes / (leg_ms[dom] * 1e-3) / 1e9

    def leg_info(k, v):
        o, t = k.split("_")
        g = algorithmic_bytes(o, t, n) / (v * 1e-3) / 1e9
        return {"ms": round(v, 4), "GBps": round(g, 1), "frac": round(g / peak, 4)}

    kern = {"scan": "scan_ws_kernel<%s, plain> (sb_scan_%s, persistent warp-specialised TMA ring, round look-back)",
            "reduce": "reduce_kernel<%s, 512, 4> (sb_reduce_%s)"}[dom_op] % (dom_t, dom_t)
    roofline = {"bound": "hbm", "kernel": kern + ("" if world == 1 else f" + the N>1 exchange ({args.scan_strategy})"),
                "achieved": round(achieved, 1), "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "frac_of_nominal_8000": round(achieved / 8000.0, 4),
                "traffic": traffic_from_profiles(dom) if (world == 1 and n == 1 << 28) else None,
                "algorithmic_bytes_per_launch": dom_bytes, "launch_ms": round(leg_ms[dom], 4),
                "per_leg": {k: leg_info(k, v) for k, v in leg_ms.items()}}

    # ---- end to end through the host-buffer C ABI (pinned host memory, H2D + D2H timed) --------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, legs, ctx, torch, dist, world, rank, dev, min(n, 1 << 28), first)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference(legs, sample_log2=args.cpu_log2n, steps=1, warmup=1)["cpu_baseline"]

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
            "dtype": "+".join(types), "data": "synthetic (reference iota: in[i]=(T)i, generated on device)",
            "config": {"workload": wl["name"], "elements_per_gpu": n, "elements_total": ntot, "legs": [f"{op}_{t}" for op, t in legs],
                       "l2": "every leg streams >= 1 GiB, larger than the 126 MB L2; no flush needed",
                       "parallelism": "1 GPU" if world == 1 else f"{world} contiguous shards; reduce: NCCL all-reduce(1 elt); prefixsum: {args.scan_strategy} with NCCL all-gather of shard totals"},
            "elements_per_s": round(len(legs) * ntot / (ms_per_step * 1e-3), 1),
            "frac_of_aggregate_measured_peak": round(value / (peak * world), 4),
            "verified": bool(ok),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0 if ok else 1


def run_e2e(args, legs, ctx, torch, dist, world, rank, dev, n, first):
    """Same step through sb_host_reduce / sb_host_scan with pinned host buffers (per rank: its shard)."""
    tdt = {"f32": torch.float32, "i64": torch.int64, "i32": torch.int32, "f64": torch.float64}
    types = sorted({t for _, t in legs})
    h_in = {t: torch.empty(n, dtype=tdt[t]).pin_memory() for t in types}
    h_out = {t: torch.empty(n, dtype=tdt[t]).pin_memory() for t in types}
    h_red = {t: torch.zeros(1, dtype=tdt[t]).pin_memory() for t in types}
    for t in types:  # reference input, built on the host like generate_input()
        h_in[t].copy_(torch.arange(first, first + n, dtype=torch.int64).to(tdt[t]))

    def estep():
        for op, t in legs:
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
    step_bytes = sum(algorithmic_bytes(op, t, n) for op, t in legs) * world
    h2d = sum(n * SIZE[t] for _, t in legs)
    d2h = sum(SIZE[t] if op == "reduce" else n * SIZE[t] for op, t in legs)
    ok = True
    lo, hi = first, first + n  # shard-local results (host calls are per shard)
    for op, t in legs:
        if t[0] != "i":
            continue
        mask = (1 << (8 * SIZE[t])) - 1
        want = (hi * (hi - 1) // 2 - lo * (lo - 1) // 2) & mask
        got = int(h_red[t].item()) if op == "reduce" else int(h_out[t][-1].item())
        ok = ok and (got & mask) == want
    return {"value": round(step_bytes / (dt / steps) / 1e9, 2), "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "ms_per_step": round(dt / steps * 1e3, 2), "steps": steps, "elements_per_gpu": n, "verified": bool(ok),
            "api": "sb_host_reduce / sb_host_scan (C ABI, pinned host buffers, 64 MiB chunks, D2H of chunk k-1 overlaps H2D of chunk k; per-rank shards are independent at N>1)"}


# ------------------------------------------------------------------------------------------------------
def cpu_reference(legs, sample_log2: int, steps: int, warmup: int):
    """The reference's CPU path on the host cores, bounded sample of the same workload.
    Reduction legs: ONE invocation of oracle/_ref/reduction with --num-runs = warmup + steps; a leg's
    time is the mean of its `steps` fastest run-time samples (the CSV lists them sorted).
    Prefix-sum legs: in-process OpenMP scan, `warmup` untimed + `steps` timed calls."""
    import numpy as np
    from oracle import oracle as O

    n = 1 << sample_log2
    threads = O.num_threads()
    O.lib()
    have_ref = O.ref_available("reduction")
    types = sorted({t for _, t in legs})
    x = {t: O.iota(t, n) for t in types}
    out = {t: np.empty_like(v) for t, v in x.items()}
    leg = {}
    red_types = [t for op, t in legs if op == "reduce"]
    if red_types:
        if have_ref:
            rows, _ = O.run_reference("reduction", n, 256, num_runs=warmup + steps, ndrange=False, verification=False, dump=False, timeout=1800)
            for t in red_types:
                samples = sorted(float(v) for v in rows[REF_ROW[t]]["run-time-samples"].strip('"').split())
                leg[f"reduce_{t}"] = sum(samples[:steps]) / steps
        else:
            for t in red_types:
                for _ in range(warmup):
                    O.reduce(t, x[t], 256)
                t0 = time.perf_counter()
                for _ in range(steps):
                    O.reduce(t, x[t], 256)
                leg[f"reduce_{t}"] = (time.perf_counter() - t0) / steps
    for op, t in legs:
        if op != "scan":
            continue
        for _ in range(warmup):
            O.cpu_scan(t, x[t], out[t])
        t0 = time.perf_counter()
        for _ in range(steps):
            O.cpu_scan(t, x[t], out[t])
        leg[f"scan_{t}"] = (time.perf_counter() - t0) / steps
    sec = sum(leg.values())
    step_bytes = sum(algorithmic_bytes(op, t, n) for op, t in legs)
    value = step_bytes / sec / 1e9
    scan_legs = [f"{op}_{t}" for op, t in legs if op == "scan"]
    legs_kind = {f"{op}_{t}": ("reference" if (op == "reduce" and have_ref) else "port") for op, t in legs}
    cpu = {"value": round(value, 3), "unit": "GB/s", "cores": threads,
           # "reference" only if EVERY leg ran the reference's own code; the prefix-sum legs never can (upstream has none)
           "kind": "reference" if all(k == "reference" for k in legs_kind.values()) else "port", "legs_kind": legs_kind,
           "sample": f"the same {len(legs)} legs at 2^{sample_log2} elements (bounded sample; GB/s is size-independent); "
                     + ("reduction legs: unmodified reference pattern/reduction.cpp on the host SYCL shim (oracle/_ref, OpenMP over work-groups), its own run-time; "
                        if have_ref else "reduction legs: oracle restatement of the reference tree (OpenMP); ")
                     + "prefixsum legs: oracle OpenMP two-pass scan (upstream prefixsum.cpp is a TODO stub)",
           "leg_seconds": {k: round(v, 4) for k, v in leg.items()}}
    return {"cpu_baseline": cpu, "seconds_per_step": sec, "value": value}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = WORKLOADS[args.workload]
    r = cpu_reference(wl["legs"], sample_log2=args.cpu_log2n, steps=max(1, args.steps), warmup=max(0, args.warmup))
    types = sorted({t for _, t in wl["legs"]})
    line = {
        "impl": "reference", "metric": METRIC, "value": round(r["value"], 3), "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(r["seconds_per_step"] * 1e3, 2), "higher_is_better": True, "scaling": wl["scaling"],
        "vs_baseline": None, "dtype": "+".join(types), "data": "synthetic (reference iota)",
        "config": {"workload": wl["name"] + f" — CPU sample of 2^{args.cpu_log2n} elements", "legs": [f"{op}_{t}" for op, t in wl["legs"]]},
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
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--log2n", type=int, default=0, help="override: elements per GPU = 2^log2n")
    ap.add_argument("--cpu-log2n", type=int, default=26, help="CPU-baseline sample size")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scan-strategy", default="reduce_then_scan", choices=["reduce_then_scan", "scan_then_carry"],
                    help="N>1 prefix sum: 3-pass reduce -> exchange -> seeded scan (default) or the 4-pass scan -> exchange -> carry-add")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world and world == 1 and args.gpus > 1:
        # convenience: re-exec under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
