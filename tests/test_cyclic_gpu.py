"""GPU parity tests of the block-cyclic two-phase prefix sum (sb_scan_cyclic_<T>, sb_peer_*, SB_SCAN_CYCLIC).

Upstream has no scan code (pattern/prefixsum.cpp:1 is a TODO stub): the oracle is the sequential scan of
oracle/sb_oracle.c (parity unpinned) and the closed forms of the iota input evaluated on the device.
Integers bit-exact, fp32 within 1e-5*log2 N, fp64 within 1e-12*log2 N.  World size 1 runs on one GPU; the
multi-rank cases need >= 2 GPUs (driver-side proof of those: bench.py under torchrun verifies every leg).
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TORCH = {"i32": torch.int32, "i64": torch.int64, "f32": torch.float32, "f64": torch.float64}
ALL4 = ("i32", "i64", "f32", "f64")


def tol(t, n):
    return (1e-5 if t == "f32" else 1e-12) * max(np.log2(max(n, 2)), 1.0)


def assert_close(t, got, want, n, what=""):
    if t[0] == "i":
        bad = np.flatnonzero(got != want)
        assert bad.size == 0, f"{what} {t}: {bad.size} integer mismatches, first at {bad[:5]}: got {got[bad[:5]]} want {want[bad[:5]]}"
    else:
        g, w = got.astype(np.float64), want.astype(np.float64)
        err = np.abs(g - w) / np.maximum(np.abs(w), 1.0)
        assert err.max(initial=0.0) <= tol(t, n), f"{what} {t}: rel err {err.max()} > {tol(t, n)} at {err.argmax()}"


def sizes(ctx, t):
    block = ctx.cyclic_block_elems(t)
    tile = block // ctx.sm_count
    return block, tile, [1, 1000, tile, tile + 1, 5 * tile + 17, block - 3, block, block + 5, 2 * block + tile // 2 + 3, 3 * block]


@pytest.mark.parametrize("t", ALL4)
def test_block_is_one_round_of_72kb_tiles(ctx, t):
    block = ctx.cyclic_block_elems(t)
    assert block == ctx.sm_count * (70 * 1024 // TORCH[t].itemsize)  # one tile (14 workers x 10 rows x 512 B) per SM and round


@pytest.mark.parametrize("t", ALL4)
@pytest.mark.parametrize("exclusive", [False, True])
def test_world1_matches_oracle(ctx, O, t, exclusive):
    """One GPU: the block-cyclic layout is the contiguous one; ragged tiles, partial rounds, several rounds."""
    _, _, ns = sizes(ctx, t)
    for n in ns:
        hx = O.iota(t, n, 3)
        x = torch.from_numpy(hx).cuda()
        y = torch.empty_like(x)
        ctx.scan_cyclic(x, y, exclusive=exclusive)
        assert_close(t, y.cpu().numpy(), O.scan(t, hx, exclusive), n, f"scan_cyclic n={n} excl={exclusive}")


@pytest.mark.parametrize("t", ALL4)
def test_world1_random_input_seed_and_in_place(ctx, O, t):
    rng = np.random.default_rng(7)
    block, tile, _ = sizes(ctx, t)
    n = 2 * block + 3 * tile + 11
    npdt = {"i32": np.int32, "i64": np.int64, "f32": np.float32, "f64": np.float64}[t]
    hx = rng.integers(-(1 << 15), 1 << 15, size=n).astype(npdt) if t[0] == "i" else (rng.random(n) * 2 - 1).astype(npdt)
    x = torch.from_numpy(hx).cuda()
    seed = torch.tensor([5], dtype=TORCH[t], device="cuda")
    ctx.scan_cyclic(x, x, seed=seed)  # in place
    want = O.scan(t, hx, False)
    want = want + want.dtype.type(5)
    got = x.cpu().numpy()
    if t[0] == "i":
        assert np.array_equal(got, want)
    else:  # random signs: compare against the magnitude of the partial sums of |x|
        scale = np.maximum(np.cumsum(np.abs(hx.astype(np.float64))), 1.0)
        assert np.max(np.abs(got.astype(np.float64) - want.astype(np.float64)) / scale) <= tol(t, n)


@pytest.mark.parametrize("t", ALL4)
def test_world1_equals_lookback_kernel_bits(ctx, t):
    """Integers: bit-identical to sb_scan_<T>; back-to-back calls exercise both mailbox planes."""
    n = (1 << 24) + 77
    x = torch.empty(n, dtype=TORCH[t], device="cuda")
    ctx.iota(x, 11)
    y0, y1 = torch.empty_like(x), torch.empty_like(x)
    ctx.scan(x, y0)
    for _ in range(3):
        y1.zero_()
        ctx.scan_cyclic(x, y1)
        if t[0] == "i":
            assert torch.equal(y0, y1)
        else:
            a, b = y0.double(), y1.double()
            assert float(((a - b).abs() / a.abs().clamp(min=1.0)).max()) <= tol(t, n)


@pytest.mark.parametrize("t", ALL4)
def test_world1_deterministic(ctx, t):
    """The summation order is fixed: floating-point results are bit-reproducible run to run."""
    n = (1 << 23) + 5
    x = torch.empty(n, dtype=TORCH[t], device="cuda")
    ctx.iota(x)
    a, b = torch.empty_like(x), torch.empty_like(x)
    ctx.scan_cyclic(x, a)
    ctx.scan_cyclic(x, b)
    assert torch.equal(a, b)


@pytest.mark.parametrize("t,log2n", [("f32", 28), ("i64", 28), ("i32", 30), ("f64", 28)])
def test_world1_baseline_sizes_closed_form(ctx, t, log2n):
    """BASELINE configs[1] / [2] sizes, EVERY element against the closed form on the device."""
    n = 1 << log2n
    x = torch.empty(n, dtype=TORCH[t], device="cuda")
    ctx.iota(x)
    y = torch.empty_like(x)
    mism = torch.zeros(1, dtype=torch.int64, device="cuda")
    for exclusive in (False, True):
        ctx.scan_cyclic(x, y, exclusive=exclusive)
        ctx.check_iota(y, mism, kind=int(exclusive), rel_tol=tol(t, n))
        ctx.sync()
        assert mism.item() == 0, (t, exclusive)


def test_cyclic_layout_helpers(ctx, pkg):
    """sb_iota_cyclic / sb_check_iota_cyclic agree with the layout function of sharded.py."""
    from sycl_bench_b200.sharded import cyclic_global_indices

    block, world = 1000, 3
    for n in (1, 999, 1000, 2500, 3000):
        for rank in range(world):
            x = torch.empty(n, dtype=torch.int64, device="cuda")
            ctx.iota_cyclic(x, block, world, rank)
            assert np.array_equal(x.cpu().numpy(), cyclic_global_indices(n, block, world, rank))
            # a buffer holding the inclusive scan of iota at those global indices passes the check, one wrong element fails
            g = cyclic_global_indices(n, block, world, rank)
            y = torch.from_numpy(g * (g + 1) // 2).cuda()
            mism = torch.zeros(1, dtype=torch.int64, device="cuda")
            ctx.check_iota_cyclic(y, mism, block, world, rank, kind=0)
            assert mism.item() == 0
            y[n // 2] += 1
            ctx.check_iota_cyclic(y, mism, block, world, rank, kind=0)
            assert mism.item() == 1


@pytest.mark.parametrize("t", ["f32", "f64"])
def test_fp_closed_form_checker(ctx, O, t):
    """sb_check_iota_f32/f64: accepts the oracle's scan, rejects a perturbed element and NaN."""
    n = 1 << 20
    hx = O.iota(t, n)
    mism = torch.zeros(1, dtype=torch.int64, device="cuda")
    for kind, seg in ((0, 0), (1, 0), (0, 256), (1, 256)):
        y = torch.from_numpy(O.scan(t, hx, bool(kind), seg)).cuda()
        ctx.check_iota(y, mism, kind=kind, seg_len=seg, rel_tol=tol(t, n))
        assert mism.item() == 0, (kind, seg)
        y[12345] *= 1.01
        y[777] = float("nan")
        ctx.check_iota(y, mism, kind=kind, seg_len=seg, rel_tol=tol(t, n))
        assert mism.item() == 2, (kind, seg)
    z = torch.from_numpy(O.segreduce(t, hx, 64)).cuda()
    ctx.check_iota(z, mism, kind=2, seg_len=64, rel_tol=tol(t, n))
    assert mism.item() == 0
    z[65] += 1  # an element the segmented reduction must not touch
    ctx.check_iota(z, mism, kind=2, seg_len=64, rel_tol=tol(t, n))
    assert mism.item() == 1


def test_missing_peer_times_out_instead_of_hanging(pkg):
    """A rank whose peer never calls gives up after SB_CYCLIC_TIMEOUT_MS and reports SB_ERR_TIMEOUT."""
    code = r"""
import sys
sys.path.insert(0, %r)
import torch
import __graft_entry__ as ge
pkg = ge.load_package()
torch.cuda.set_device(0)
ctx = pkg.Context(0, torch.cuda.current_stream().cuda_stream)
p0, p1 = pkg.Peer(ctx, 0, 2), pkg.Peer(ctx, 1, 2)
p0.connect_local(p1); p1.connect_local(p0)
n = 3 * ctx.cyclic_block_elems("i32")
x = torch.ones(n, dtype=torch.int32, device="cuda"); y = torch.empty_like(x)
ctx.scan_cyclic(x, y, peer=p1)          # rank 1 needs rank 0's round totals; rank 0 never runs
try:
    ctx.sync()
    print("NO-ERROR")
except pkg.SbError as e:
    print("STATUS", e.status)
ctx.scan(x, y); ctx.sync()               # the ctx is usable afterwards
print("AFTER", int(y[-1].item()) == n)
""" % ROOT
    env = dict(os.environ, SB_CYCLIC_TIMEOUT_MS="300")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env)
    assert "STATUS -9" in r.stdout, r.stdout + r.stderr
    assert "AFTER True" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("t", ALL4)
def test_reduce_all_two_ranks_on_one_gpu(pkg, O, t):
    """sb_reduce_all: the all-reduce is the tail of the reduce kernel (peer mailboxes).  Two ranks as two contexts /
    streams on ONE device: the last block of one kernel waits for the other kernel's sum.  Integers exact; floating
    point identical on both ranks (the sums are added in rank order) and within tolerance; several calls in a row use
    both mailbox planes."""
    torch.cuda.set_device(0)
    c0, c1 = pkg.Context(0, pkg.Context.OWNED_STREAM), pkg.Context(0, pkg.Context.OWNED_STREAM)
    p0, p1 = pkg.Peer(c0, 0, 2), pkg.Peer(c1, 1, 2)
    p0.connect_local(p1)
    p1.connect_local(p0)
    try:
        for n in (1000, (1 << 22) + 77):
            h0, h1 = O.iota(t, n, 5), O.iota(t, n, 5 + n)
            x0, x1 = torch.from_numpy(h0).cuda(), torch.from_numpy(h1).cuda()
            o0, o1 = torch.zeros(1, dtype=TORCH[t], device="cuda"), torch.zeros(1, dtype=TORCH[t], device="cuda")
            torch.cuda.synchronize()
            for _ in range(3):
                c0.reduce_all(x0, o0, peer=p0)
                c1.reduce_all(x1, o1, peer=p1)
                c0.sync()
                c1.sync()
                g0, g1 = o0.cpu().numpy()[0], o1.cpu().numpy()[0]
                assert g0 == g1
                if t[0] == "i":
                    lo, hi = 5, 5 + 2 * n
                    assert int(g0) == O.wrap(hi * (hi - 1) // 2 - lo * (lo - 1) // 2, t)
                else:
                    want = float(np.sum(h0.astype(np.float64)) + np.sum(h1.astype(np.float64)))
                    assert abs(float(g0) - want) <= tol(t, 2 * n) * abs(want)
    finally:
        p0.close()
        p1.close()
        c0.close()
        c1.close()


# ---------------------------------------------------------------------------------------- >= 2 GPUs
def _ngpus():
    return torch.cuda.device_count()


def _cyclic_shards(pkg, comm, t, n, world):
    xs, ys = [], []
    for r in range(world):
        c = comm.rank_ctx(r)
        block = c.cyclic_block_elems(t)
        with torch.cuda.device(r):
            x = torch.empty(n, dtype=TORCH[t], device=f"cuda:{r}")
            c.iota_cyclic(x, block, world, r)
            xs.append(x)
            ys.append(torch.empty_like(x))
    return xs, ys, block


@pytest.mark.parametrize("t", ALL4)
@pytest.mark.parametrize("world", [2, 4, 8])
def test_comm_scan_cyclic(pkg, O, t, world):
    """Single-process group (SB_SCAN_CYCLIC): several rounds + a short last round, against the oracle's scan of the
    global array, re-assembled with the layout function; repeated to use both mailbox planes."""
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    from sycl_bench_b200.sharded import cyclic_global_indices

    torch.cuda.set_device(0)
    ctx0 = pkg.Context(0, pkg.Context.OWNED_STREAM)
    comm = pkg.Comm(ctx0, world)
    try:
        block = ctx0.cyclic_block_elems(t)
        for n in (block // 3, 2 * block + 12345):
            xs, ys, _ = _cyclic_shards(pkg, comm, t, n, world)
            want = O.scan(t, O.iota(t, n * world), False)
            for exclusive in (False, True, False):
                comm.scan(xs, ys, exclusive=exclusive, strategy=pkg.Comm.CYCLIC)
                ctx0.sync()
                w = O.scan(t, O.iota(t, n * world), True) if exclusive else want
                for r in range(world):
                    g = cyclic_global_indices(n, block, world, r)
                    assert_close(t, ys[r].cpu().numpy(), w[g], n * world, f"rank {r} n={n} excl={exclusive}")
    finally:
        comm.close()
        ctx0.close()


@pytest.mark.parametrize("op", ["segreduce", "segscan", "group_scan"])
def test_comm_foreach_replicas(pkg, O, op):
    """sb_comm_foreach: shards cut on segment boundaries, independent per rank (no exchange)."""
    world = 2
    if _ngpus() < world:
        pytest.skip("needs 2 GPUs")
    torch.cuda.set_device(0)
    ctx0 = pkg.Context(0, pkg.Context.OWNED_STREAM)
    comm = pkg.Comm(ctx0, world)
    try:
        L, per = 256, 256 * 5000
        for t in ("i32", "f64"):
            full = O.iota(t, per * world)
            xs, ys = [], []
            for r in range(world):
                with torch.cuda.device(r):
                    xs.append(torch.from_numpy(full[r * per:(r + 1) * per].copy()).to(f"cuda:{r}"))
                    ys.append(torch.empty_like(xs[-1]))
            if op == "segreduce":
                comm.foreach(pkg.Comm.OP_SEGREDUCE, None, xs, L)
                want, got = O.segreduce(t, full, L), xs
            elif op == "segscan":
                comm.foreach(pkg.Comm.OP_SEGSCAN, xs, ys, L, exclusive=False)
                want, got = O.scan(t, full, False, L), ys
            else:
                comm.foreach(pkg.Comm.OP_SCAN, xs, ys, L, exclusive=True)
                want, got = O.scan(t, full, True, L), ys
            ctx0.sync()
            for r in range(world):
                assert_close(t, got[r].cpu().numpy(), want[r * per:(r + 1) * per], per * world, f"{op} rank {r}")
    finally:
        comm.close()
        ctx0.close()
