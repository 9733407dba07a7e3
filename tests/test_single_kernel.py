"""single-kernel/scalar_prod and single-kernel/lin_reg_coeff: sb_dot_<T>, sb_linreg_<T>.

CPU: oracle/sb_oracle_single_kernel.c is pinned bit for bit to the golden vectors of the UNMODIFIED
reference (tests/golden/single_kernel_golden.json, made by tests/golden/make_golden_single_kernel.py) and
to the live oracle/_ref binaries.  GPU: the fused kernels through the C ABI against the golden vectors, the
oracle on random inputs, and the drop-in binaries.
"""
import json
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "sycl-bench_b200", "bin")
FMT = {"i32": "<i", "i64": "<q", "f32": "<f", "f64": "<d"}
NAME = {"i32": "int32", "i64": "int64", "f32": "fp32", "f64": "fp64"}
NP = {"i32": np.int32, "i64": np.int64, "f32": np.float32, "f64": np.float64}


@pytest.fixture(scope="module")
def golden_sk():
    with open(os.path.join(ROOT, "tests", "golden", "single_kernel_golden.json")) as f:
        return json.load(f)


def hexof(t, value):
    return struct.pack(FMT[t], value).hex()


def tol(t, n):
    return (1e-5 if t == "f32" else 1e-12) * max(np.log2(max(n, 2)), 1.0)


def is_pow2(x):
    return x > 0 and x & (x - 1) == 0


# ------------------------------------------------------------------------------------------- CPU: oracle
def test_oracle_scalar_prod_reproduces_golden(O, golden_sk):
    assert len(golden_sk["scalar_prod"]) >= 40
    for g in golden_sk["scalar_prod"]:
        t, n = g["dtype"], g["size"]
        got = O.scalar_prod(t, np.ones(n, NP[t]), np.full(n, 2, NP[t]), g["local"])
        assert hexof(t, got) == g["result_hex"], g


def test_scalar_prod_halving_tree_quirk_is_in_the_golden_vectors(golden_sk):
    """--local=100: the stride loop 50, 25, 12, 6, 3, 1 (scalar_prod.cpp:120-127) skips elements, the
    reference computes 1024 instead of 2000 and fails its own verify; every power-of-two --local gives 2N."""
    for g in golden_sk["scalar_prod"]:
        if is_pow2(g["local"]):
            assert g["result"] == 2 * g["size"] and g["reference_verify"] == "PASS", g
        else:
            assert g["result"] == 1024 and g["reference_verify"] == "FAIL", g


def test_oracle_linreg_reproduces_golden(O, golden_sk):
    assert len(golden_sk["lin_reg_coeff"]) >= 12
    for g in golden_sk["lin_reg_coeff"]:
        t, n = g["dtype"], g["size"]
        out = O.linreg(t, np.ones(n, NP[t]), np.full(n, 2, NP[t]), g["local"])
        assert [hexof(t, v) for v in out[:4]] == g["sums_hex"], g
        assert out[4] == 2 and out[5] == 0 and g["reference_verify"] == "PASS"


def test_oracle_on_random_inputs_matches_numpy(O):
    rng = np.random.default_rng(7)
    n = 1 << 14
    a = rng.integers(-1000, 1000, n).astype(np.int32)
    b = rng.integers(-1000, 1000, n).astype(np.int32)
    exact = int(np.dot(a.astype(np.int64), b.astype(np.int64)))
    assert int(O.scalar_prod("i32", a, b, 128)) == (exact + 2**31) % 2**32 - 2**31  # wraps like the device
    assert int(O.scalar_prod("i64", a.astype(np.int64), b.astype(np.int64), 128)) == exact
    x = rng.random(n)
    y = 3.0 * x + 0.5 + 0.01 * rng.standard_normal(n)
    out = O.linreg("f64", x, y, 64)
    np.testing.assert_allclose(out[:4], [np.dot(x, y), np.dot(x, x), x.sum(), y.sum()], rtol=1e-12)


@pytest.mark.parametrize("which", ["scalar_prod", "lin_reg_coeff"])
def test_live_reference_matches_oracle(O, which):
    if not O.ref_available(which):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    size, local = 8192, 128
    rows, dumps = O.run_reference(which, size, local, num_runs=1, ndrange=True, dump_max_bytes=8, threads=1)
    if which == "scalar_prod":
        assert len(dumps) == 8
        for t, raw in zip(O.REF_REDUCTION_TYPES * 2, dumps):
            want = O.scalar_prod(t, np.ones(size, NP[t]), np.full(size, 2, NP[t]), local)
            assert raw[: struct.calcsize(FMT[t])].hex() == hexof(t, want)
    else:
        assert len(dumps) == 8
        for k, t in enumerate(("f32", "f64")):
            want = O.linreg(t, np.ones(size, NP[t]), np.full(size, 2, NP[t]), local)
            got = [raw[: struct.calcsize(FMT[t])].hex() for raw in dumps[4 * k: 4 * k + 4]]
            assert got == [hexof(t, v) for v in want[:4]]
    assert all(r["Verification"] == "PASS" for r in rows.values())


def test_binaries_built():
    for b in ("scalar_prod", "lin_reg_coeff"):
        assert os.access(os.path.join(BIN, b), os.X_OK), f"{b} not built: make -C sycl-bench_b200 pattern"


# ------------------------------------------------------------------------------------------- GPU: C ABI
def _dev(x):
    import torch

    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.mark.gpu
def test_gpu_dot_matches_golden(ctx, golden_sk):
    import torch

    for g in golden_sk["scalar_prod"]:
        if not is_pow2(g["local"]) or g["variant"] != "Hierarchical":
            continue  # non-power-of-two --local: the reference drops elements (documented, not emulated)
        t, n = g["dtype"], g["size"]
        a = _dev(np.ones(n, NP[t]))
        b = _dev(np.full(n, 2, NP[t]))
        out = torch.zeros(1, dtype=a.dtype, device="cuda")
        ctx.dot(a, b, out)
        ctx.sync()
        assert hexof(t, out.cpu().numpy()[0].item()) == g["result_hex"], g


@pytest.mark.gpu
@pytest.mark.parametrize("t", ["i32", "i64", "f32", "f64"])
@pytest.mark.parametrize("n,offset", [(1, 0), (7, 0), (1000, 1), (4099, 3), (1 << 16, 0), ((1 << 20) + 5, 0), ((1 << 20) + 5, 1)])
def test_gpu_dot_matches_oracle_on_random_inputs(ctx, O, t, n, offset):
    import torch

    rng = np.random.default_rng(n + offset)
    if t[0] == "i":
        a = rng.integers(-(1 << 15), 1 << 15, n + offset).astype(NP[t])
        b = rng.integers(-(1 << 15), 1 << 15, n + offset).astype(NP[t])
    else:
        a = (rng.random(n + offset) * 2 - 1).astype(NP[t])
        b = (rng.random(n + offset) * 2 - 1).astype(NP[t])
    da, db = _dev(a), _dev(b)
    out = torch.zeros(1, dtype=da.dtype, device="cuda")
    ctx.dot(da[offset:], db[offset:], out, n=n)  # offset != 0: element-aligned only -> the scalar-load path
    ctx.sync()
    got = out.cpu().numpy()[0]
    want = O.scalar_prod(t, np.ascontiguousarray(a[offset:]), np.ascontiguousarray(b[offset:]), 256)
    if t[0] == "i":
        assert got == want
    else:
        scale = max(float(np.abs(a[offset:].astype(np.float64) * b[offset:].astype(np.float64)).sum()), 1.0)
        assert abs(float(got) - float(want)) <= tol(t, n) * scale


@pytest.mark.gpu
@pytest.mark.parametrize("t", ["f32", "f64"])
@pytest.mark.parametrize("n,offset", [(2, 0), (1024, 0), (4099, 1), (1 << 18, 0), ((1 << 20) + 3, 2)])
def test_gpu_linreg_matches_oracle_on_random_inputs(ctx, O, t, n, offset):
    import torch

    rng = np.random.default_rng(n)
    x = rng.random(n + offset).astype(NP[t])
    y = (3.0 * x + 0.5 + 0.01 * rng.standard_normal(n + offset)).astype(NP[t])
    dx, dy = _dev(x), _dev(y)
    out = torch.zeros(6, dtype=dx.dtype, device="cuda")
    ctx.linreg(dx[offset:], dy[offset:], out, n=n)
    ctx.sync()
    got = out.cpu().numpy()
    xs, ys = np.ascontiguousarray(x[offset:]), np.ascontiguousarray(y[offset:])
    # sums: against the exact (float64 / long double) values within the north-star tolerance
    exact = [np.dot(xs.astype(np.longdouble), ys.astype(np.longdouble)), np.dot(xs.astype(np.longdouble), xs.astype(np.longdouble)),
             xs.astype(np.longdouble).sum(), ys.astype(np.longdouble).sum()]
    for g, w in zip(got[:4], exact):
        assert abs(float(g) - float(w)) <= tol(t, n) * max(abs(float(w)), 1.0)
    # coefficients: lin_reg_coeff.cpp:134-142 applied to the device sums, operation for operation in T
    T = NP[t]
    sxy, sxx, sx, sy = (T(v) for v in got[:4])
    mean_x, mean_y = T(sx / T(n)), T(sy / T(n))
    ss_xy, ss_xx = T(sxy - T(mean_x * mean_y)), T(sxx - T(mean_x * mean_x))
    b1 = T(ss_xy / ss_xx)
    b0 = T(mean_y - T(b1 * mean_x))
    assert hexof(t, got[4].item()) == hexof(t, b1.item()) and hexof(t, got[5].item()) == hexof(t, b0.item())
    # and the oracle (reference tree order) agrees on the sums within the same tolerance.  Only for power-of-two
    # sizes: the reference kernel reads g[2*gid + 1] without a bound check (lin_reg_coeff.cpp:95-97), so an odd
    # intermediate size folds a stale element of the previous pass into the sum — faithfully restated by the
    # oracle, not a property of the sums themselves.
    want = O.linreg(t, xs, ys, 256) if is_pow2(n) and n >= 2 else None
    if want is not None:
        for g, w in zip(got[:4], want[:4]):
            assert abs(float(g) - float(w)) <= 2 * tol(t, n) * max(abs(float(w)), 1.0)


@pytest.mark.gpu
def test_gpu_fused_does_not_modify_inputs_and_is_deterministic_on_request(ctx):
    import torch

    n = (1 << 22) + 11
    x = torch.rand(n, device="cuda", dtype=torch.float32)
    y = torch.rand(n, device="cuda", dtype=torch.float32)
    x0, y0 = x.clone(), y.clone()
    ctx.set_deterministic(True)
    try:
        outs = []
        for _ in range(3):
            out = torch.zeros(6, dtype=torch.float32, device="cuda")
            ctx.linreg(x, y, out)
            ctx.sync()
            outs.append(out.cpu().numpy().tobytes())
        assert outs[0] == outs[1] == outs[2]
    finally:
        ctx.set_deterministic(False)
    assert torch.equal(x, x0) and torch.equal(y, y0)


@pytest.mark.gpu
def test_gpu_baseline_size(ctx):
    """2^28 elements, x = 1, y = 2 (the reference's inputs): dot = 2^29 exactly in every type; b1 = 2, b0 = 0."""
    import torch

    n = 1 << 28
    for t, dt in (("i32", torch.int32), ("i64", torch.int64), ("f32", torch.float32), ("f64", torch.float64)):
        a = torch.empty(n, dtype=dt, device="cuda")
        b = torch.empty(n, dtype=dt, device="cuda")
        ctx.fill(a, 1)
        ctx.fill(b, 2)
        out = torch.zeros(6, dtype=dt, device="cuda")
        ctx.dot(a, b, out)
        ctx.sync()
        assert float(out.cpu()[0]) == float(2 * n)
        if t[0] == "f":
            ctx.linreg(a, b, out)
            ctx.sync()
            got = out.cpu().numpy()
            assert list(got[:4]) == [2.0 * n, 1.0 * n, 1.0 * n, 2.0 * n] and got[4] == 2.0 and got[5] == 0.0
        del a, b


# ------------------------------------------------------------------------------------------- GPU: binaries
def _run(bench, tmp_path, *flags):
    from oracle import oracle as O

    out = tmp_path / f"{bench}.csv"
    r = subprocess.run([os.path.join(BIN, bench), f"--output={out}", *flags], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    return (O.parse_csv(str(out)) if out.exists() else {}), r


@pytest.mark.gpu
def test_binaries_rows_and_verdicts(tmp_path):
    rows, _ = _run("scalar_prod", tmp_path, "--size=1048576", "--local=256", "--num-runs=3")
    assert list(rows) == [f"ScalarProduct_NDRange_{NAME[t]}" for t in FMT]
    assert all(r["Verification"] == "PASS" and r["kernel-launches-per-run"] == "1" for r in rows.values())
    rows, _ = _run("lin_reg_coeff", tmp_path, "--size=1048576", "--local=256", "--num-runs=3")
    assert list(rows) == ["LinearRegressionCoeff_fp32", "LinearRegressionCoeff_fp64"]
    assert all(r["Verification"] == "PASS" and r["kernel-launches-per-run"] == "1" for r in rows.values())
