"""sb_comm_* (single-process multi-GPU, include/sb_cuda.h) on a real multi-GPU box.

With one visible GPU only the size-1 group is exercised (no NCCL involved); the 2-GPU cases are skipped.
Results are checked against the closed forms of the iota input on the device (sb_check_iota) and against
the CPU oracle at small sizes.
"""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

TORCH = {"i32": torch.int32, "i64": torch.int64, "f32": torch.float32, "f64": torch.float64}


def _bounds(n, world, rank):
    per = -(-n // world)
    per = -(-per // 16) * 16
    b = min(n, rank * per)
    return b, min(n, b + per)


def _run(pkg, world, dtype, n, strategy, exclusive=False):
    ctx = pkg.Context(0)
    comm = pkg.Comm(ctx, world)
    shards, outs, totals = [], [], []
    for r in range(world):
        b, e = _bounds(n, world, r)
        with torch.cuda.device(r):
            x = torch.empty(max(e - b, 1), dtype=TORCH[dtype], device=f"cuda:{r}")[: e - b]
            y = torch.empty_like(x)
            t = torch.zeros(1, dtype=TORCH[dtype], device=f"cuda:{r}")
            torch.cuda.synchronize(r)
        comm.rank_ctx(r).iota(x, offset=b)
        comm.rank_ctx(r).sync()
        shards.append(x)
        outs.append(y)
        totals.append(t)
    comm.reduce(shards, totals, dtype=dtype)
    comm.scan(shards, outs, exclusive=exclusive, strategy=strategy, dtype=dtype)
    ctx.sync()
    for r in range(world):
        comm.rank_ctx(r).sync()
    got_total = [t.cpu().numpy()[0] for t in totals]
    got_scan = np.concatenate([y.cpu().numpy() for y in outs]) if n else np.zeros(0, TORCH_NP[dtype])
    comm.close()
    ctx.close()
    return got_total, got_scan


TORCH_NP = {"i32": np.int32, "i64": np.int64, "f32": np.float32, "f64": np.float64}


def _expected(dtype, n, exclusive):
    i = np.arange(n, dtype=np.uint64)
    if dtype in ("i32", "i64"):
        inc = np.cumsum(i, dtype=np.uint64)
        exc = inc - i
        want = (exc if exclusive else inc).astype(np.uint32 if dtype == "i32" else np.uint64)
        return want.view(np.int32 if dtype == "i32" else np.int64), int(i.sum(dtype=np.uint64))
    inc = np.cumsum(i.astype(np.float64))
    want = inc - i if exclusive else inc
    return want, float(i.astype(np.float64).sum())


def _assert_matches(dtype, n, exclusive, totals, scan):
    want, total = _expected(dtype, n, exclusive)
    if dtype in ("i32", "i64"):
        assert np.array_equal(scan, want)
        mask = 0xFFFFFFFF if dtype == "i32" else 0xFFFFFFFFFFFFFFFF
        for t in totals:
            assert int(t) & mask == total & mask
    else:
        tol = (1e-5 if dtype == "f32" else 1e-12) * max(1.0, np.log2(max(n, 2)))
        np.testing.assert_allclose(scan.astype(np.float64), want, rtol=tol, atol=tol)
        for t in totals:
            assert abs(float(t) - total) <= tol * max(total, 1.0)


@pytest.mark.parametrize("dtype", ["i32", "i64", "f32", "f64"])
@pytest.mark.parametrize("n", [0, 1, 1000, 1 << 20])
def test_comm_of_one_matches_closed_form(pkg, dtype, n):
    if dtype == "f32" and n > 4096:
        n = 4096  # fp32 represents i exactly, and the sums stay < 2^24
    totals, scan = _run(pkg, 1, dtype, n, 0)
    _assert_matches(dtype, n, False, totals, scan)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("strategy", [0, 1])
@pytest.mark.parametrize("dtype", ["i32", "i64", "f32", "f64"])
@pytest.mark.parametrize("n", [0, 5, 1000, (1 << 22) + 3])
def test_comm_of_two_matches_closed_form(pkg, dtype, n, strategy):
    if dtype == "f32" and n > 4096:
        n = 4096
    totals, scan = _run(pkg, 2, dtype, n, strategy)
    _assert_matches(dtype, n, False, totals, scan)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_comm_exclusive_scan_two_gpus(pkg):
    totals, scan = _run(pkg, 2, "i64", (1 << 20) + 17, 0, exclusive=True)
    _assert_matches("i64", (1 << 20) + 17, True, totals, scan)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("binary,extra,rows", [
    ("reduction", [], 4),
    ("prefixsum", [], 4),  # default strategy: block-cyclic shards, exchange inside the kernel
    ("prefixsum", ["--scan-strategy=reduce_then_scan"], 4),
    ("prefixsum", ["--scan-strategy=scan_then_carry"], 4),
    ("scan", [], 8),  # global scope through sb_comm_scan (exclusive), group scope as replicas
    ("segmentedscan", ["--no-flags"], 4),
    ("segmentedreduction", [], 5),  # the int16 row stays on one GPU
])
def test_pattern_binaries_num_gpus(binary, extra, rows):
    exe = os.path.join(ROOT, "sycl-bench_b200", "bin", binary)
    if not os.path.exists(exe):
        pytest.skip("pattern binaries not built")
    out = subprocess.run([exe, "--size=16777216", "--local=256", "--num-gpus=2", "--num-runs=2", "--strict"] + extra, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("Verification: PASS") == rows and "FAIL" not in out.stdout, out.stdout
