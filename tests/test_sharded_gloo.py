"""N > 1 host logic on CPU: world_size 2 and 3 over gloo.  The local operations are the oracle's
(there is no GPU here); what is under test is the sharding, the exchange and the carry arithmetic of
sycl-bench_b200/sharded.py, which bench.py runs over NCCL with libsb_cuda as the local backend."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, dtype_name, strategy, q):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    from oracle import oracle as O

    pkg = ge.load_package()
    from sycl_bench_b200 import sharded

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t = {"i32": torch.int32, "i64": torch.int64, "f64": torch.float64}[dtype_name]

    def o_reduce(x, out):
        out[0] = int(O.cpu_sum(dtype_name, x.numpy())) if dtype_name[0] == "i" else float(O.cpu_sum(dtype_name, x.numpy()))

    def o_scan(x, y, exclusive, seed):
        r = O.scan(dtype_name, x.numpy(), exclusive)
        y.copy_(torch.from_numpy(r))
        if seed is not None:
            y.add_(seed)

    def o_carry_add(y, c):
        y.add_(c)

    def o_excl(totals, r, c):
        c[0] = totals[:r].sum() if r else 0

    ops = sharded.ShardedOps(dist, rank, world, o_reduce, o_scan, o_carry_add, o_excl)
    b, e = sharded.shard_bounds(n, world, rank)
    x = torch.from_numpy(O.iota(dtype_name, e - b, b))
    out = torch.zeros(1, dtype=t)
    ops.reduce(x, out)
    y = torch.empty_like(x)
    totals, carry = torch.zeros(world, dtype=t), torch.zeros(1, dtype=t)
    ops.prefixsum(x, y, totals, carry, strategy=strategy)
    q.put((rank, b, e, out.item(), y.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,dtype_name,strategy", [
    (2, "i32", "reduce_then_scan"), (2, "i32", "scan_then_carry"), (2, "i64", "reduce_then_scan"),
    (2, "f64", "scan_then_carry"), (3, "i64", "scan_then_carry"), (3, "i32", "reduce_then_scan"),
])
def test_sharded_reduce_and_prefixsum(world, dtype_name, strategy):
    from oracle import oracle as O

    n = 100003 if dtype_name != "i32" else (1 << 17) + 5  # int32 wraps: the carry arithmetic must wrap too
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, dtype_name, strategy, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = O.iota(dtype_name, n)
    want_scan = O.scan(dtype_name, full)
    want_sum = O.wrap(O.kat_reduce(n), dtype_name) if dtype_name[0] == "i" else float(n) * (n - 1) / 2
    got = np.concatenate([r[4] for r in results])
    assert [r[1] for r in results][0] == 0 and results[-1][2] == n
    for r in results:
        assert r[3] == want_sum
    assert np.array_equal(got, want_scan)


def _worker_cyclic_and_replicas(rank, world, port, n, block, dtype_name, q):
    """Block-cyclic prefix sum and the exchange-free replicas through ShardedOps, local ops = the oracle.
    The oracle stands in for sb_scan_cyclic by gathering the shards over gloo and scanning the re-assembled
    global array: what is under test is the layout arithmetic and the plumbing of sharded.py."""
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    from oracle import oracle as O

    ge.load_package()
    from sycl_bench_b200 import sharded

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t = {"i32": torch.int32, "i64": torch.int64, "f64": torch.float64}[dtype_name]
    gidx = [sharded.cyclic_global_indices(n, block, world, r) for r in range(world)]

    def o_scan_cyclic(x, y, exclusive):
        parts = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(parts, x)
        full = np.empty(n * world, dtype=x.numpy().dtype)
        for r in range(world):
            full[gidx[r]] = parts[r].numpy()
        y.copy_(torch.from_numpy(O.scan(dtype_name, full, exclusive)[gidx[rank]]))

    def o_scan(x, y, exclusive, seed, group_len=0):
        y.copy_(torch.from_numpy(O.scan(dtype_name, x.numpy(), exclusive, group_len)))

    ops = sharded.ShardedOps(dist, rank, world, None, o_scan, None, None, scan_cyclic=o_scan_cyclic,
                             segreduce=lambda x, L: x.copy_(torch.from_numpy(O.segreduce(dtype_name, x.numpy(), L))),
                             segscan=lambda x, y, L, ex: y.copy_(torch.from_numpy(O.scan(dtype_name, x.numpy(), ex, L))))
    # input: the reference's iota over the GLOBAL index space, in the block-cyclic layout
    x = torch.from_numpy(gidx[rank].astype(x_np := {"i32": np.int32, "i64": np.int64, "f64": np.float64}[dtype_name]))
    y = torch.empty_like(x)
    ops.prefixsum(x, y, None, None, strategy="cyclic")
    cyc = y.numpy().copy()
    # replicas: contiguous shards cut on segment boundaries
    L = 64
    b, e = sharded.shard_bounds(n * world // L * L, world, rank, align=L)
    xs = torch.from_numpy(O.iota(dtype_name, e - b, b))
    ys = torch.empty_like(xs)
    ops.foreach_segscan(xs, ys, L)
    segscan = ys.numpy().copy()
    ops.foreach_group_scan(xs, ys, L, exclusive=True)
    gscan = ys.numpy().copy()
    ops.foreach_segreduce(xs, L)
    try:
        ops.foreach_segscan(xs[:-1], ys[:-1], L)
        raised = False
    except ValueError:
        raised = True
    q.put((rank, cyc, (b, e), segscan, gscan, xs.numpy().copy(), raised))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,dtype_name,n,block", [(2, "i32", 5000, 1000), (2, "i64", 2500, 1000), (3, "f64", 3000, 1000), (3, "i32", 999, 1000)])
def test_cyclic_prefixsum_and_replicas(world, dtype_name, n, block):
    from oracle import oracle as O
    import __graft_entry__ as ge

    ge.load_package()
    from sycl_bench_b200.sharded import cyclic_global_index, cyclic_global_indices

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_cyclic_and_replicas, args=(r, world, port, n, block, dtype_name, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=120) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = n * world
    # the layout is a bijection onto [0, total), rank-major inside a round, and the scalar form agrees with the vector form
    allidx = np.concatenate([cyclic_global_indices(n, block, world, r) for r in range(world)])
    assert np.array_equal(np.sort(allidx), np.arange(total))
    for r in range(world):
        for i in (0, n // 2, n - 1):
            assert cyclic_global_index(i, n, block, world, r) == cyclic_global_indices(n, block, world, r)[i]
    want = O.scan(dtype_name, O.iota(dtype_name, total))
    L = 64
    seg_total = total // L * L
    full = O.iota(dtype_name, seg_total)
    for rank, cyc, (b, e), segscan, gscan, segred, raised in results:
        assert np.array_equal(cyc, want[cyclic_global_indices(n, block, world, rank)])
        assert np.array_equal(segscan, O.scan(dtype_name, full, False, L)[b:e])
        assert np.array_equal(gscan, O.scan(dtype_name, full, True, L)[b:e])
        assert np.array_equal(segred, O.segreduce(dtype_name, full, L)[b:e])
        assert raised  # a shard that is not cut on a segment boundary is refused
    assert results[0][2][0] == 0 and results[-1][2][1] == seg_total


def test_shard_bounds():
    import __graft_entry__ as ge

    ge.load_package()
    from sycl_bench_b200.sharded import shard_bounds

    for n, world, align in [(1 << 33, 8, 1), (1000, 3, 1), (1 << 29, 8, 1024), (10, 4, 4), (5, 8, 1)]:
        spans = [shard_bounds(n, world, r, align) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for (b0, e0), (b1, e1) in zip(spans, spans[1:]):
            assert e0 == b1 and b0 <= e0
        for (b, e), (b1, e1) in zip(spans, spans[1:]):
            if e1 > b1:  # a shard followed by a non-empty shard ends on a segment boundary
                assert (e - b) % align == 0
    assert shard_bounds(1 << 33, 8, 3) == (3 << 30, 4 << 30)
