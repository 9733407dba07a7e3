#!/usr/bin/env python3
"""bench.py — headline benchmark of the pattern/ reduce + prefix-sum hot path on B200.

Default workload = BASELINE.json configs[1]: pattern/reduction + pattern/prefixsum, fp32 and int64,
--size=2^28 per GPU, iota input (pattern/reduction.cpp:31-34 of the reference).  One *step* is one
pass of the hot path over that batch: 4 libsb_cuda launches

    sb_reduce_f32   (1 GiB algorithmic)      prefix sum f32, inclusive (2 GiB)
    sb_reduce_i64   (2 GiB)                  prefix sum i64, inclusive (4 GiB)

Metric: effective HBM GB/s = algorithmic bytes (reduce N*s + s, scan 2*N*s) / time.
`value` is timed with inputs resident in HBM; `e2e` is the same step from pinned HOST buffers with the
H2D and D2H copies inside the timed region (N=1: sb_host_reduce / sb_host_scan of the C ABI; N>1: host
shard -> device -> the SHARDED global operation -> host).

N > 1 (torchrun, one rank per GPU).  reduction = sb_reduce_all: the all-reduce of the per-GPU sums is the tail of the
reduce kernel itself (peer-memory mailboxes over NVLink); --reduce-exchange nccl = sb_reduce + NCCL all-reduce of one element.
Prefix sum, --scan-strategy:
    cyclic (default)   block-cyclic shards, sb_scan_cyclic: the round totals cross NVLink inside the kernel
                       (peer-memory mailboxes), every element is read from HBM once (2*N*s/P per GPU)
    reduce_then_scan   contiguous shards: sb_reduce -> NCCL all-gather of the totals -> seeded sb_scan (3*N*s/P)
    scan_then_carry    contiguous shards: sb_scan -> all-gather -> sb_carry_add, the literal north-star recipe (4*N*s/P)
Every leg of every run is verified on the device: integers bit-exact against the closed forms of the GLOBAL
problem, floating point within BASELINE.json's tolerance (1e-5 log2 N fp32, 1e-12 log2 N fp64).

Sub-records on the same JSON line (each with its own steps, per-leg ms and `verified`):
    config5   BASELINE configs[4]: reduction + prefixsum int32, 2^33 elements in total, strong scaling; at N>1 rank 0
              also times the same problem on ONE GPU, so `speedup_vs_1gpu` is measured inside the run
    replicas  BASELINE configs[2] / [3] legs that shard with no exchange (group-scope scan, segmented scan, segmented
              reduction): every rank works on its own shard cut on segment boundaries (weak scaling)

`--impl reference` times the reference's own CPU implementation on the host cores instead: oracle/_ref/reduction
(the unmodified pattern/reduction.cpp on the host SYCL shim) for the reduction legs, and the oracle's OpenMP port
for the prefix-sum legs (upstream prefixsum.cpp is a TODO stub), on the same 2^28-element configuration.
"""
from __future__ import annotations

import argparse
import json
import math
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
    "config2": {"legs": (("reduce", "f32", 0), ("reduce", "i64", 0), ("scan", "f32", 0), ("scan", "i64", 0)), "log2_per_gpu": 28, "scaling": "weak",
                "name": "pattern/reduction + pattern/prefixsum, fp32 and int64, --size=2^28 per GPU (BASELINE configs[1])"},
    "config5": {"legs": (("reduce", "i32", 0), ("scan", "i32", 0)), "log2_total": 33, "scaling": "strong",
                "name": "pattern/reduction + pattern/prefixsum, int32, --size=2^33 in total, sharded (BASELINE configs[4])"},
    "config3": {"legs": (("gscan", "i32", 256), ("segscan", "i32", 256), ("segscan_flags", "i32", 256)), "log2_per_gpu": 30, "scaling": "weak",
                "name": "pattern/scan --scan-scope=group + pattern/segmentedscan --no-flags + pattern/segmentedscan (explicit head flags), int32, "
                        "--size=2^30 per GPU, --local=256 (BASELINE configs[2], replicas)"},
    "config4": {"legs": tuple(("segreduce", "f64", L) for L in (64, 128, 256, 512, 1024)), "log2_per_gpu": 29, "scaling": "weak",
                "name": "pattern/segmentedreduction fp64, --size=2^29 per GPU, --local 64..1024 (BASELINE configs[3], replicas)"},
}
REF_ROW = {"f32": "Pattern_Reduction_Hierarchical_fp32", "i64": "Pattern_Reduction_Hierarchical_int64",
           "i32": "Pattern_Reduction_Hierarchical_int32", "f64": "Pattern_Reduction_Hierarchical_fp64"}


def leg_name(leg) -> str:
    op, t, L = leg
    return f"{op}_{t}" + (f"_L{L}" if L else "")


def algorithmic_bytes(op: str, t: str, n: int, L: int = 0) -> int:
    s = SIZE[t]
    if op == "reduce":
        return n * s + s
    if op == "segreduce":
        return n * s + (n // L) * s
    if op == "segscan_flags":
        return 2 * n * s + n  # one flag byte per element
    return 2 * n * s  # scan, gscan, segscan


def rel_tol(t: str, n: int) -> float:
    """BASELINE.json north_star: fp32 <= 1e-5 * log2 N, fp64 <= 1e-12 * log2 N (relative)."""
    return (1e-5 if t == "f32" else 1e-12) * max(math.log2(max(n, 2)), 1.0)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def traffic_from_profiles(kernel_key: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu capture of the same kernel
    and size (profiles/traffic.json names the capture).  DRAM counters cannot be read outside a profiler, so this is
    the recorded number, not one measured in this run; the run's own evidence is the per-leg rate."""
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


def bind_to_gpu_numa_node(bus_id: str):
    """Pin this process (and the pinned host memory it allocates afterwards) to the NUMA node the GPU hangs off.
    Returns a short description for the JSON line; silently does nothing where sysfs does not say."""
    try:
        node = int(open(f"/sys/bus/pci/devices/{bus_id.lower()}/numa_node").read().strip())
        if node < 0:
            return {"node": None, "note": "sysfs reports no NUMA affinity for the GPU"}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"node": node, "note": "no allowed CPU on that node"}
        os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": len(cpus)}
    except Exception as e:  # noqa: BLE001
        return {"node": None, "note": f"not bound ({type(e).__name__})"}


# ------------------------------------------------------------------------------------------------------
class Bench:
    """Per-process state of the GPU arm: context, sharded operations, layout of the prefix-sum shards."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import __graft_entry__ as ge

        self.torch, self.dist, self.args = torch, dist, args
        self.pkg = ge.load_package()
        from sycl_bench_b200 import sharded

        self.sharded = sharded
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; libsb_cuda has no CPU fallback (use --impl reference for the CPU arm)")
        self.affinity0 = os.sched_getaffinity(0)
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.ctx = self.pkg.Context(self.local_rank, torch.cuda.current_stream().cuda_stream)
        self.numa = bind_to_gpu_numa_node(self._bus_id()) if not args.no_numa_bind else {"node": None, "note": "--no-numa-bind"}
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.strategy = args.scan_strategy if self.world > 1 else "single"
        self.strategy_note = None
        self.peer = None
        if self.strategy == "cyclic" or (self.world > 1 and args.reduce_exchange == "fused"):
            try:
                self.peer = sharded.make_peer_group(self.ctx, dist, self.rank, self.world)
                ok = 1
            except self.pkg.SbError as e:
                ok, self.strategy_note = 0, f"peer memory unavailable on this box ({e}); fell back to reduce_then_scan / NCCL"
            okt = torch.tensor([ok], dtype=torch.int64, device=self.dev)
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
            if int(okt.item()) == 0:
                self.peer = None
                if self.strategy == "cyclic":
                    self.strategy = "reduce_then_scan"
                self.strategy_note = self.strategy_note or "a peer rank could not map the mailboxes; fell back to reduce_then_scan / NCCL"
        self.ops = sharded.for_context(self.ctx, dist, self.rank, self.world, peer=self.peer)
        self.tdt = {"f32": torch.float32, "i64": torch.int64, "i32": torch.int32, "f64": torch.float64}

    def _bus_id(self) -> str:
        import ctypes as C

        buf = C.create_string_buffer(64)
        L = self.pkg.lib()
        L.sb_ctx_pci_bus_id.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        if L.sb_ctx_pci_bus_id(self.ctx._h, buf, 64) != 0:
            return ""
        return buf.value.decode()

    # -- helpers
    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def all_ok(self, ok: bool) -> bool:
        if self.world > 1:
            t = self.torch.tensor([1 if ok else 0], dtype=self.torch.int64, device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
            ok = bool(t.item())
        return ok

    def max_over_ranks(self, v: float) -> float:
        if self.world > 1:
            t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            v = float(t.item())
        return v

    def empty(self, t: str, n: int):
        return self.torch.empty(n, dtype=self.tdt[t], device=self.dev)

    def cyclic(self) -> bool:
        return self.strategy == "cyclic"

    def fill_input(self, x, t: str, n: int, ops=None):
        """The reference's iota, in[i] = (T)i over the GLOBAL index space, in the layout the scan strategy uses."""
        ops = ops or self
        if ops.cyclic():
            self.ctx.iota_cyclic(x, self.ctx.cyclic_block_elems(t), self.world, self.rank)
        else:
            self.ctx.iota(x, offset=self.rank * n)

    # -- one leg = (run, verify)
    def make_leg(self, leg, n: int, x, y, solo: bool = False):
        """`solo`: the leg runs on this GPU alone (world size 1 semantics) even inside a multi-rank job."""
        op, t, L = leg
        torch, ctx = self.torch, self.ctx
        world = 1 if solo else self.world
        rank = 0 if solo else self.rank
        ntot = n * world
        first = rank * n
        mism = torch.zeros(1, dtype=torch.int64, device=self.dev)
        if op == "reduce":
            red = torch.zeros(1, dtype=self.tdt[t], device=self.dev)

            def run():
                if solo:
                    ctx.reduce(x, red)
                else:
                    self.ops.reduce(x, red, exchange=self.args.reduce_exchange)  # N>1: all-reduce inside the kernel (peer mailboxes) or NCCL

            def verify():
                exact = ntot * (ntot - 1) // 2
                if t[0] == "i":
                    mask = (1 << (8 * SIZE[t])) - 1
                    return (int(red.item()) & mask) == (exact & mask)
                return abs(float(red.item()) - float(exact)) <= rel_tol(t, ntot) * float(exact)
            return run, verify
        if op == "scan":
            totals = torch.zeros(max(world, 1), dtype=self.tdt[t], device=self.dev)
            carry = torch.zeros(1, dtype=self.tdt[t], device=self.dev)
            cyc = self.cyclic() and not solo

            def run():
                if solo:
                    ctx.scan(x, y)
                else:
                    self.ops.prefixsum(x, y, totals, carry, strategy=self.strategy if self.world > 1 else "reduce_then_scan")

            def verify():
                if cyc:
                    ctx.check_iota_cyclic(y, mism, ctx.cyclic_block_elems(t), world, rank, kind=0, rel_tol=rel_tol(t, ntot))
                else:
                    ctx.check_iota(y, mism, offset=first, kind=0, rel_tol=rel_tol(t, ntot))
                return int(mism.item()) == 0
            return run, verify
        # replicas: every rank works on its own shard, cut on segment boundaries; nothing is exchanged
        if op == "gscan":
            def run():
                self.ops.foreach_group_scan(x, y, L, exclusive=True)

            def verify():
                ctx.check_iota(y, mism, offset=first, kind=1, seg_len=L, rel_tol=rel_tol(t, ntot))
                return int(mism.item()) == 0
            return run, verify
        if op == "segscan":
            def run():
                self.ops.foreach_segscan(x, y, L, exclusive=False)

            def verify():
                ctx.check_iota(y, mism, offset=first, kind=0, seg_len=L, rel_tol=rel_tol(t, ntot))
                return int(mism.item()) == 0
            return run, verify
        if op == "segscan_flags":
            # explicit head flags, one byte per element, a head every L (what pattern/segmentedscan generates by default)
            fl = self.torch.zeros(n, dtype=self.torch.uint8, device=self.dev)
            fl[::L] = 1

            def run():
                ctx.segscan(x, y, head_flags=fl)

            def verify():
                ctx.check_iota(y, mism, offset=first, kind=0, seg_len=L, rel_tol=rel_tol(t, ntot))
                return int(mism.item()) == 0
            return run, verify
        if op == "segreduce":
            def run():  # in place: timed on a buffer whose heads already hold sums (same traffic)
                self.ops.foreach_segreduce(y, L)

            def verify():  # what the reference does per run: fresh input (setup()), one reduction, check
                ctx.iota(y, offset=first)
                self.ops.foreach_segreduce(y, L)
                ctx.check_iota(y, mism, offset=first, kind=2, seg_len=L, rel_tol=rel_tol(t, ntot))
                return int(mism.item()) == 0
            return run, verify
        raise ValueError(op)

    def time_legs(self, legs, n: int, steps: int, warmup: int, solo: bool = False, sample_clocks: bool = False):
        """Allocate + fill the inputs of `legs`, run warm-up and `steps` timed steps, verify every leg.
        Returns dict(ms_per_step, leg_ms, verified, launches, clocks)."""
        torch = self.torch
        types = sorted({t for _, t, _ in legs})
        x = {t: self.empty(t, n) for t in types}
        y = {t: self.empty(t, n) for t in types}
        # the block-cyclic layout belongs to the global reduce / prefix-sum legs; replicas own contiguous shards
        replicas = any(op in ("gscan", "segscan", "segscan_flags", "segreduce") for op, _, _ in legs)
        for t in types:
            if solo:
                self.ctx.iota(x[t], offset=0)
            elif replicas:
                self.ctx.iota(x[t], offset=self.rank * n)
            else:
                self.fill_input(x[t], t, n)
            if any(op == "segreduce" for op, tt, _ in legs if tt == t):
                self.ctx.iota(y[t], offset=(0 if solo else self.rank * n))
            # scratch of the scans sized outside the timed region (sb_ctx_reserve: setup(), not run())
            self.ctx.reserve(1, t, n)
        runs = [(leg,) + self.make_leg(leg, n, x[leg[1]], y[leg[1]], solo=solo) for leg in legs]
        sync = torch.cuda.synchronize if solo else self.barrier

        def step(events=None):
            for leg, run, _ in runs:
                if events is not None:
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    e0.record()
                run()
                if events is not None:
                    e1.record()
                    events.append((leg, e0, e1))

        for _ in range(warmup):
            step()
        sync()
        sampler = ClockSampler(self.local_rank) if (sample_clocks and self.rank == 0) else None
        if sampler:
            sampler.start()
        launches0 = self.ctx.launches
        events = []
        t_begin = torch.cuda.Event(enable_timing=True); t_end = torch.cuda.Event(enable_timing=True)
        sync()
        t_begin.record()
        for _ in range(steps):
            step(events)
        t_end.record()
        sync()
        ms_total = t_begin.elapsed_time(t_end)
        launches = self.ctx.launches - launches0
        clocks = sampler.stop() if sampler else None
        if not solo:
            ms_total = self.max_over_ranks(ms_total)
            if self.world > 1:
                lt = torch.tensor([launches], dtype=torch.int64, device=self.dev)
                self.dist.all_reduce(lt)
                launches = int(lt.item())
        verified = {}
        for leg, run, verify in runs:
            run()  # legs of one dtype share the output buffer: check what THIS leg writes
            ok = bool(verify())
            verified[leg_name(leg)] = ok if solo else self.all_ok(ok)
        per_leg = {}
        for leg, e0, e1 in events:
            per_leg.setdefault(leg_name(leg), []).append(e0.elapsed_time(e1))
        leg_ms = {k: sum(v) / len(v) for k, v in per_leg.items()}
        if not solo and self.world > 1:  # a leg takes as long as its slowest rank
            lm = torch.tensor([leg_ms[leg_name(l)] for l in legs], dtype=torch.float64, device=self.dev)
            self.dist.all_reduce(lm, op=self.dist.ReduceOp.MAX)
            leg_ms = {leg_name(l): float(v) for l, v in zip(legs, lm.tolist())}
        del x, y, runs
        torch.cuda.empty_cache()
        return {"ms_per_step": ms_total / steps, "leg_ms": leg_ms, "verified": verified, "launches": int(launches), "clocks": clocks}


def leg_table(legs, n, leg_ms, peak, world=1):
    out = {}
    for leg in legs:
        k = leg_name(leg)
        g = algorithmic_bytes(leg[0], leg[1], n, leg[2]) * world / (leg_ms[k] * 1e-3) / 1e9
        out[k] = {"ms": round(leg_ms[k], 4), "GBps": round(g, 1), "frac": round(g / (peak * world), 4)}
    return out


def run_ours(args):
    B = Bench(args)
    torch, world, rank = B.torch, B.world, B.rank
    wl = WORKLOADS[args.workload]
    legs = wl["legs"]
    if args.log2n:
        n = 1 << args.log2n
    elif "log2_per_gpu" in wl:
        n = 1 << wl["log2_per_gpu"]
    else:
        n = (1 << wl["log2_total"]) // world
    ntot = n * world
    peak, peak_src = measured_peak()

    main = B.time_legs(legs, n, args.steps, args.warmup, sample_clocks=True)
    ms_per_step, leg_ms = main["ms_per_step"], main["leg_ms"]
    ok = all(main["verified"].values())
    step_bytes = sum(algorithmic_bytes(op, t, n, L) for op, t, L in legs) * world
    value = step_bytes / (ms_per_step * 1e-3) / 1e9
    dom = max(legs, key=lambda l: leg_ms[leg_name(l)])
    dom_key = leg_name(dom)
    dom_bytes = algorithmic_bytes(dom[0], dom[1], n, dom[2])
    achieved = dom_bytes / (leg_ms[dom_key] * 1e-3) / 1e9
    kernels = {
        "single": "scan_ws_kernel<%s, plain> (sb_scan_%s: persistent warp-specialised TMA ring, round look-back over independent round totals)",
        "cyclic": "scan_cyclic_kernel<%s> (sb_scan_cyclic_%s: two-phase persistent TMA-ring kernel, tile re-read from L2, round totals exchanged over NVLink peer memory)",
        "reduce_then_scan": "reduce_kernel<%s> + NCCL all-gather + seeded scan_ws_kernel (sb_reduce_%s -> sb_scan)",
        "scan_then_carry": "scan_ws_kernel<%s> + NCCL all-gather + carry_add_kernel (sb_scan_%s -> sb_carry_add)",
    }
    kern = {"scan": kernels[B.strategy], "reduce": "reduce_kernel<%s, 512, 4> (sb_reduce_%s)" + ("" if world == 1 else (" with the all-reduce fused into its last block (peer mailboxes over NVLink)" if (B.peer is not None and args.reduce_exchange == "fused") else " + NCCL all-reduce of one element")),
            "gscan": "scan_independent_kernel<%s> (sb_scan_%s, group scope)", "segscan": "scan_independent_kernel<%s> (sb_segscan_%s)",
            "segscan_flags": "flags_tile_heads_kernel + scan_independent_kernel<%s, FLAGS> (sb_segscan_%s, explicit head flags)",
            "segreduce": "segreduce_rows_kernel<%s> (sb_segreduce_%s)"}[dom[0]] % (dom[1], dom[1])
    roofline = {"bound": "hbm", "kernel": kern, "achieved": round(achieved, 1), "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "frac_of_nominal_8000": round(achieved / 8000.0, 4),
                "traffic": traffic_from_profiles(dom_key) if (world == 1 and n == 1 << 28) else None,
                "traffic_source": "profiles/traffic.json (ncu capture of this kernel at this size; DRAM counters are not readable outside a profiler)",
                "algorithmic_bytes_per_launch": dom_bytes, "launch_ms": round(leg_ms[dom_key], 4),
                "per_leg": leg_table(legs, n, leg_ms, peak)}

    # ---- sub-records --------------------------------------------------------------------------------
    sub_steps = max(3, min(args.steps, args.sub_steps))
    config5 = None
    if not args.no_config5 and args.workload == "config2":
        config5 = run_config5(B, sub_steps, peak)
        ok = ok and config5["verified"]
    replicas = None
    if not args.no_replicas and args.workload == "config2":
        replicas = run_replicas(B, sub_steps, peak)
        ok = ok and replicas["verified"]

    # ---- end to end from pinned host memory (H2D + D2H timed) ------------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(B, legs, min(n, 1 << 28))
        ok = ok and e2e["verified"]

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, B.affinity0)  # the CPU leg uses every core the job may use, not one NUMA node
        cpu = cpu_reference(legs, sample_log2=args.cpu_log2n, steps=1, warmup=1, budget_s=40.0)["cpu_baseline"]

    if rank == 0:
        red = "sb_reduce_all (all-reduce inside the kernel, peer mailboxes)" if (B.peer is not None and args.reduce_exchange == "fused") else "sb_reduce + NCCL all-reduce(1 elt)"
        par = {"single": "1 GPU",
               "cyclic": f"{world} GPUs, block-cyclic shards (block = {B.ctx.sm_count} tiles of 70 KB); reduce: {red}; prefixsum: sb_scan_cyclic, round totals over NVLink peer memory inside the kernel, 2 passes",
               "reduce_then_scan": f"{world} contiguous shards; reduce: {red}; prefixsum: sb_reduce -> NCCL all-gather -> seeded sb_scan, 3 passes",
               "scan_then_carry": f"{world} contiguous shards; reduce: {red}; prefixsum: sb_scan -> NCCL all-gather -> sb_carry_add, 4 passes"}[B.strategy]
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
            "dtype": "+".join(sorted({t for _, t, _ in legs})), "data": "synthetic (reference iota: in[i]=(T)i over the global index space, generated on device)",
            "config": {"workload": wl["name"], "elements_per_gpu": n, "elements_total": ntot, "legs": [leg_name(l) for l in legs],
                       "l2": "every leg streams >= 1 GiB, larger than the 126 MB L2; no flush needed",
                       "parallelism": par, "scan_strategy": B.strategy, "reduce_exchange": ("fused" if (B.peer is not None and args.reduce_exchange == "fused" and world > 1) else ("nccl" if world > 1 else None)), "strategy_note": B.strategy_note, "numa": B.numa},
            "elements_per_s": round(len(legs) * ntot / (ms_per_step * 1e-3), 1),
            "frac_of_aggregate_measured_peak": round(value / (peak * world), 4),
            "verified": bool(ok), "verified_legs": main["verified"],
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "config5": config5, "replicas": replicas,
            "gpu_launches": main["launches"], "clocks": main["clocks"],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        B.dist.barrier()
        if B.peer is not None:
            B.peer.close()
        B.dist.destroy_process_group()
    return 0 if ok else 1


def run_config5(B, steps, peak):
    """BASELINE configs[4]: reduction + prefixsum int32 over 2^33 elements in total (strong scaling)."""
    wl = WORKLOADS["config5"]
    total = 1 << (B.args.config5_log2_total or wl["log2_total"])
    n = total // B.world
    r = B.time_legs(wl["legs"], n, steps, 3)
    rec = {"workload": wl["name"], "elements_total": total, "elements_per_gpu": n, "steps": steps, "scan_strategy": B.strategy,
           "ms_per_step": round(r["ms_per_step"], 4), "per_leg": leg_table(wl["legs"], n, r["leg_ms"], peak, B.world),
           "value": round(sum(algorithmic_bytes(op, t, total, L) for op, t, L in wl["legs"]) / (r["ms_per_step"] * 1e-3) / 1e9, 1), "unit": "GB/s",
           "verified": all(r["verified"].values()), "verified_legs": r["verified"]}
    if B.world > 1:
        # the same 2^33 elements on ONE GPU (rank 0; 32 GiB in + 32 GiB out), so that the speed-up is measured in this run
        solo = None
        if B.rank == 0:
            solo = B.time_legs(wl["legs"], total, steps, 3, solo=True)
        B.barrier()
        if B.rank == 0:
            rec["one_gpu_ms_per_step"] = round(solo["ms_per_step"], 4)
            rec["one_gpu_per_leg_ms"] = {k: round(v, 4) for k, v in solo["leg_ms"].items()}
            rec["one_gpu_verified"] = all(solo["verified"].values())
            rec["speedup_vs_1gpu"] = round(solo["ms_per_step"] / r["ms_per_step"], 3)
            rec["speedup_per_leg"] = {k: round(solo["leg_ms"][k] / r["leg_ms"][k], 3) for k in r["leg_ms"]}
            rec["verified"] = rec["verified"] and rec["one_gpu_verified"]
    return rec


def run_replicas(B, steps, peak):
    """Legs that shard with no exchange at all: every rank runs the 1-GPU kernel on its own shard."""
    out = {"note": "independent per-rank shards cut on segment boundaries; no collective in the data path", "steps": steps, "verified": True}
    for key in ("config3", "config4"):
        wl = WORKLOADS[key]
        n = 1 << (B.args.replicas_log2n or wl["log2_per_gpu"])
        r = B.time_legs(wl["legs"], n, steps, 3)
        byts = sum(algorithmic_bytes(op, t, n, L) for op, t, L in wl["legs"]) * B.world
        out[key] = {"workload": wl["name"], "elements_per_gpu": n, "ms_per_step": round(r["ms_per_step"], 4),
                    "value": round(byts / (r["ms_per_step"] * 1e-3) / 1e9, 1), "unit": "GB/s",
                    "frac_of_aggregate_measured_peak": round(byts / (r["ms_per_step"] * 1e-3) / 1e9 / (peak * B.world), 4),
                    "per_leg": leg_table(wl["legs"], n, r["leg_ms"], peak, B.world), "verified_legs": r["verified"]}
        out["verified"] = out["verified"] and all(r["verified"].values())
    return out


def pcie_probe(B, nbytes=1 << 28, reps=3):
    """Concurrent H2D + D2H cudaMemcpyAsync on every rank at once: the ceiling the end-to-end path runs under."""
    torch = B.torch
    h_a = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_b = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(nbytes, dtype=torch.uint8, device=B.dev)
    d_b = torch.empty(nbytes, dtype=torch.uint8, device=B.dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for mode in ("h2d", "d2h", "both"):
        B.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_a.copy_(h_a, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_b.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize()
        dt = B.max_over_ranks(time.perf_counter() - t0)
        res[mode] = reps * nbytes / dt / 1e9  # per GPU and direction
    return {"h2d_GBps_per_gpu": round(res["h2d"], 1), "d2h_GBps_per_gpu": round(res["d2h"], 1), "bidir_GBps_per_gpu_and_direction": round(res["both"], 1),
            "how": f"{reps} x {nbytes >> 20} MiB cudaMemcpyAsync per direction from/to pinned host memory, all {B.world} ranks at once, wall clock, max over ranks"}


def run_e2e(B, legs, n):
    """The same step from pinned HOST buffers.  N=1: sb_host_reduce / sb_host_scan (C ABI: chunked, D2H of chunk k-1
    overlaps H2D of chunk k).  N>1: host shard -> device -> the SHARDED global operation (ShardedOps) -> host."""
    torch, ctx, world, rank = B.torch, B.ctx, B.world, B.rank
    types = sorted({t for _, t, _ in legs})
    ntot = n * world
    h_in = {t: torch.empty(n, dtype=B.tdt[t]).pin_memory() for t in types}
    h_out = {t: torch.empty(n, dtype=B.tdt[t]).pin_memory() for t in types}
    h_red = {t: torch.zeros(1, dtype=B.tdt[t]).pin_memory() for t in types}
    d_in = {t: B.empty(t, n) for t in types}
    d_out = {t: B.empty(t, n) for t in types}
    d_red = {t: torch.zeros(1, dtype=B.tdt[t], device=B.dev) for t in types}
    totals = {t: torch.zeros(world, dtype=B.tdt[t], device=B.dev) for t in types}
    carry = {t: torch.zeros(1, dtype=B.tdt[t], device=B.dev) for t in types}
    for t in types:  # the reference's generate_input(), in the layout of the scan strategy; staged through the device generator
        B.fill_input(d_in[t], t, n)
        h_in[t].copy_(d_in[t])
        d_in[t].zero_()
    torch.cuda.synchronize()

    def estep():
        for op, t, _ in legs:
            if world == 1:
                if op == "reduce":
                    ctx.host_reduce(h_in[t], h_red[t])
                else:
                    ctx.host_scan(h_in[t], h_out[t])
            else:
                d_in[t].copy_(h_in[t], non_blocking=True)  # H2D on the ctx stream (torch's current stream)
                if op == "reduce":
                    B.ops.reduce(d_in[t], d_red[t], exchange=B.args.reduce_exchange)
                    h_red[t].copy_(d_red[t], non_blocking=True)
                else:
                    B.ops.prefixsum(d_in[t], d_out[t], totals[t], carry[t], strategy=B.strategy)
                    h_out[t].copy_(d_out[t], non_blocking=True)
        torch.cuda.synchronize()

    steps = max(1, min(B.args.steps, B.args.e2e_steps))
    estep()  # warm-up (grows the staging buffers)
    B.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        estep()
    torch.cuda.synchronize()
    dt = B.max_over_ranks(time.perf_counter() - t0)
    # verification of what came back to the HOST: re-upload the result buffers and run the device checkers
    ok = True
    mism = torch.zeros(1, dtype=torch.int64, device=B.dev)
    for op, t, _ in legs:
        if op == "reduce":
            exact = ntot * (ntot - 1) // 2
            if t[0] == "i":
                mask = (1 << (8 * SIZE[t])) - 1
                ok = ok and (int(h_red[t].item()) & mask) == (exact & mask)
            else:
                ok = ok and abs(float(h_red[t].item()) - float(exact)) <= rel_tol(t, ntot) * float(exact)
        else:
            d_out[t].copy_(h_out[t])
            if B.cyclic():
                ctx.check_iota_cyclic(d_out[t], mism, ctx.cyclic_block_elems(t), world, rank, kind=0, rel_tol=rel_tol(t, ntot))
            else:
                ctx.check_iota(d_out[t], mism, offset=rank * n, kind=0, rel_tol=rel_tol(t, ntot))
            ok = ok and int(mism.item()) == 0
    ok = B.all_ok(ok)
    step_bytes = sum(algorithmic_bytes(op, t, n, L) for op, t, L in legs) * world
    h2d = sum(n * SIZE[t] for _, t, _ in legs)
    d2h = sum(SIZE[t] if op == "reduce" else n * SIZE[t] for op, t, _ in legs)
    ms = dt / steps * 1e3
    probe = pcie_probe(B)
    if world == 1:  # both directions overlap inside sb_host_scan
        floor_ms = max(h2d / probe["bidir_GBps_per_gpu_and_direction"], d2h / probe["bidir_GBps_per_gpu_and_direction"]) / 1e6
        api = "sb_host_reduce / sb_host_scan (C ABI, pinned host buffers, 64 MiB chunks, D2H of chunk k-1 overlaps H2D of chunk k)"
    else:  # the carry of a rank needs every rank's upload: H2D and D2H of a leg cannot overlap
        floor_ms = (h2d / probe["h2d_GBps_per_gpu"] + d2h / probe["d2h_GBps_per_gpu"]) / 1e6
        api = f"pinned host shard -> cudaMemcpyAsync -> ShardedOps.reduce / ShardedOps.prefixsum(strategy={B.strategy}) over all {world} GPUs -> cudaMemcpyAsync -> host"
    return {"value": round(step_bytes / (dt / steps) / 1e9, 2), "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "ms_per_step": round(ms, 2), "steps": steps, "elements_per_gpu": n, "verified": bool(ok), "api": api,
            "roofline": {"bound": "pcie", "probe": probe, "floor_ms": round(floor_ms, 2), "frac": round(floor_ms / ms, 3),
                         "note": "floor = this step's H2D and D2H bytes at the probe's rates (overlapped at N=1, serial per leg at N>1); the kernels are <2 % of the step"}}


# ------------------------------------------------------------------------------------------------------
def host_threads() -> int:
    return max(1, len(os.sched_getaffinity(0)))


def cpu_reference(legs, sample_log2: int, steps: int, warmup: int, budget_s: float = 150.0):
    """The reference's CPU path on the host cores, on the same legs.
    Reduction legs: ONE invocation of oracle/_ref/reduction (the unmodified pattern/reduction.cpp) with --num-runs R;
    a leg's time is the mean of its min(steps, R) fastest run-time samples.  R = warmup + steps unless a probe run says
    that would exceed `budget_s`.  Prefix-sum legs: in-process OpenMP scan, `warmup` untimed + `steps` timed calls.
    The thread count is set explicitly: torchrun exports OMP_NUM_THREADS=1, which would run this arm on one core."""
    threads = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(threads)
    import numpy as np
    from oracle import oracle as O

    n = 1 << sample_log2
    O.lib()
    threads = O.num_threads()
    have_ref = O.ref_available("reduction")
    types = sorted({t for _, t, _ in legs})
    x = {t: O.iota(t, n) for t in types}
    out = {t: np.empty_like(v) for t, v in x.items()}
    leg = {}
    note = ""
    red_types = [t for op, t, _ in legs if op == "reduce"]
    if red_types:
        if have_ref:
            t0 = time.perf_counter()
            O.run_reference("reduction", n, 256, num_runs=1, ndrange=False, verification=False, dump=False, threads=threads, timeout=1800)
            t1 = time.perf_counter() - t0
            runs = max(2, min(warmup + steps, int(budget_s / max(t1, 1e-3))))
            rows, _ = O.run_reference("reduction", n, 256, num_runs=runs, ndrange=False, verification=False, dump=False, threads=threads, timeout=3600)
            for t in red_types:
                samples = sorted(float(v) for v in rows[REF_ROW[t]]["run-time-samples"].strip('"').split())
                k = max(1, min(steps, len(samples)))
                leg[f"reduce_{t}"] = sum(samples[:k]) / k
            note = f"{runs} runs of the reference binary (one probe run took {t1:.1f} s; all four dtypes run each time), mean of the {min(steps, runs)} fastest run-time samples; "
        else:
            for t in red_types:
                for _ in range(warmup):
                    O.reduce(t, x[t], 256)
                t0 = time.perf_counter()
                for _ in range(steps):
                    O.reduce(t, x[t], 256)
                leg[f"reduce_{t}"] = (time.perf_counter() - t0) / steps
    for op, t, _ in legs:
        if op != "scan":
            continue
        for _ in range(warmup):
            O.cpu_scan(t, x[t], out[t])
        t0 = time.perf_counter()
        for _ in range(steps):
            O.cpu_scan(t, x[t], out[t])
        leg[f"scan_{t}"] = (time.perf_counter() - t0) / steps
    sec = sum(leg.values())
    step_bytes = sum(algorithmic_bytes(op, t, n, L) for op, t, L in legs)
    value = step_bytes / sec / 1e9
    legs_kind = {f"{op}_{t}": ("reference" if (op == "reduce" and have_ref) else "port") for op, t, _ in legs}
    cpu = {"value": round(value, 3), "unit": "GB/s", "cores": threads,
           # "reference" only if EVERY leg ran the reference's own code; the prefix-sum legs never can (upstream has none)
           "kind": "reference" if all(k == "reference" for k in legs_kind.values()) else "port", "legs_kind": legs_kind,
           "sample": f"the same {len(legs)} legs at 2^{sample_log2} elements; "
                     + ("reduction legs: unmodified reference pattern/reduction.cpp on the host SYCL shim (oracle/_ref, OpenMP over work-groups), its own run-time, " + note
                        if have_ref else "reduction legs: oracle restatement of the reference tree (OpenMP); ")
                     + "prefixsum legs: oracle OpenMP two-pass scan (upstream prefixsum.cpp is a TODO stub)",
           "leg_seconds": {k: round(v, 4) for k, v in leg.items()}}
    return {"cpu_baseline": cpu, "seconds_per_step": sec, "value": value}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = WORKLOADS[args.workload]
    legs = [l for l in wl["legs"] if l[0] in ("reduce", "scan")]
    log2n = args.cpu_log2n or wl.get("log2_per_gpu", 28)
    log2n = min(log2n, 28)  # host RAM: 2^28 int64 in + out + the reference's own copies = 10 GiB
    r = cpu_reference(legs, sample_log2=log2n, steps=max(1, args.steps), warmup=max(0, args.warmup))
    n = 1 << log2n
    same = "log2_per_gpu" in wl and log2n == wl["log2_per_gpu"]
    line = {
        "impl": "reference", "metric": METRIC, "value": round(r["value"], 3), "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(r["seconds_per_step"] * 1e3, 2), "higher_is_better": True, "scaling": wl["scaling"],
        "vs_baseline": None, "dtype": "+".join(sorted({t for _, t, _ in legs})), "data": "synthetic (reference iota)",
        "config": {"workload": wl["name"] if same else wl["name"] + f" — CPU sample of 2^{log2n} elements", "elements_per_gpu": n, "elements_total": n,
                   "legs": [leg_name(l) for l in legs],
                   "note": "one host runs the 2^28-element step of the GPU arm's per-GPU configuration; the CPU arm does not scale with --gpus"},
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
    ap.add_argument("--cpu-log2n", type=int, default=0, help="CPU sample size (default: 26 for the cpu_baseline leg, the workload's own 2^28 for --impl reference)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--sub-steps", type=int, default=5, help="timed steps of the config5 / replicas sub-records")
    ap.add_argument("--config5-log2-total", type=int, default=0, help="override the 2^33 total of the config5 sub-record (tests)")
    ap.add_argument("--replicas-log2n", type=int, default=0, help="override the per-GPU size of the replicas sub-record (tests)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config5", action="store_true")
    ap.add_argument("--no-replicas", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true")
    ap.add_argument("--reduce-exchange", default="fused", choices=["fused", "nccl"],
                    help="N>1 reduction: all-reduce fused into the reduce kernel over NVLink peer memory (default; needs the peer group of "
                         "--scan-strategy cyclic, else NCCL is used) or sb_reduce + NCCL all-reduce of one element")
    ap.add_argument("--scan-strategy", default="cyclic", choices=["cyclic", "reduce_then_scan", "scan_then_carry"],
                    help="N>1 prefix sum: block-cyclic 2-pass kernel with in-kernel NVLink exchange (default), 3-pass reduce -> all-gather -> seeded scan, "
                         "or the 4-pass scan -> all-gather -> carry-add")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if not args.cpu_log2n:
        args.cpu_log2n = 26
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
