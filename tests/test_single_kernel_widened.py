"""single-kernel/lin_reg_error and single-kernel/kmeans (the rows the path widens into next): sb_linreg_error_<T>, sb_kmeans_<T>.

CPU: oracle/sb_oracle_single_kernel.c against the golden vectors of the UNMODIFIED reference
(tests/golden/single_kernel_widened_golden.json, made by tests/golden/make_golden_single_kernel_widened.py) — kmeans bit for
bit, lin_reg_error within tolerance (the reference is built with -ffast-math: the order of its inner sum is the compiler's) —
and live against oracle/_ref where the binaries exist.  GPU: the kernels through the C ABI against the golden vectors and
against the oracle on random inputs, and the drop-in binaries.

Tolerance (north star): |got - want| <= (1e-5 fp32 | 1e-12 fp64) * log2(N) * max(|want|, 1).
"""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "sycl-bench_b200", "bin")
NP = {"f32": np.float32, "f64": np.float64}
NAME = {"f32": "fp32", "f64": "fp64"}


@pytest.fixture(scope="module")
def golden_w():
    with open(os.path.join(ROOT, "tests", "golden", "single_kernel_widened_golden.json")) as f:
        return json.load(f)


def tol(t, n):
    return (1e-5 if t == "f32" else 1e-12) * max(np.log2(max(n, 2)), 1.0)


def assert_within(t, n, got, want, what, abs_floor=0.0):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    bad = np.abs(got - want) > tol(t, n) * np.maximum(np.abs(want), 1.0) + abs_floor
    assert not bad.any(), f"{what}: {int(bad.sum())} of {bad.size} outside tolerance, first at {int(np.argmax(bad))}: {got[np.argmax(bad)]} vs {want[np.argmax(bad)]}"


def golden_values(g):
    t, n = g["dtype"], g["size"]
    idx = np.arange(n) if g["indices"] == "all" else np.asarray(g["indices"])
    return idx, np.frombuffer(bytes.fromhex(g["values_hex"]), dtype=NP[t])


def exact_errors(in1, in2, alpha, beta):
    """The N x M squared errors in float64, from the sums of the quadratic form (long double sums)."""
    x, y = in1.astype(np.longdouble), in2.astype(np.longdouble)
    a, b = alpha.astype(np.longdouble), beta.astype(np.longdouble)
    n = np.longdouble(x.size)
    sx, sy, sxx, sxy, syy = x.sum(), y.sum(), (x * x).sum(), (x * y).sum(), (y * y).sum()
    return (a * a * sxx + 2 * a * b * sx - 2 * a * sxy + n * b * b - 2 * b * sy + syy).astype(np.float64)


# ------------------------------------------------------------------------------------------- CPU: oracle
def test_oracle_lin_reg_error_reproduces_golden(O, golden_w):
    assert len(golden_w["lin_reg_error"]) >= 12
    for g in golden_w["lin_reg_error"]:
        t, n = g["dtype"], g["size"]
        assert g["reference_verify"] == "PASS"
        idx, want = golden_values(g)
        ins = O.lin_reg_error_inputs(t, n, g["rand_skip"])
        if n <= 16384:  # the literal N^2 loop (65536: 4e9 iterations per type — left to the closed form below)
            assert_within(t, n, O.lin_reg_error(t, *ins)[idx], want, f"lin_reg_error loop {t} n={n}")
        # The correctly rounded values (quadratic form, long double sums) pin the regenerated inputs and meet the reference's
        # OWN criterion (lin_reg_error.cpp:80-103: relative L2 error below 1e-6) at every size.  The literal loop in fp32
        # does not from N = 16384 on (1.9e-6): the reference's vectorised sum is more accurate than a sequential one.
        exact = exact_errors(*ins)[idx]
        assert_within(t, n, exact, want, f"lin_reg_error closed form {t} n={n}")
        assert np.linalg.norm(exact - want) / np.linalg.norm(want.astype(np.float64)) < 1e-6


def test_oracle_lin_reg_error_loop_equals_quadratic_form(O):
    for t in ("f32", "f64"):
        n = 3000
        ins = O.lin_reg_error_inputs(t, n, 17)
        assert_within(t, n, O.lin_reg_error(t, *ins), exact_errors(*ins), f"loop vs closed form {t}")


def test_oracle_kmeans_reproduces_golden(O, golden_w):
    assert len(golden_w["kmeans"]) >= 8
    for g in golden_w["kmeans"]:
        t, n = g["dtype"], g["size"]
        m = O.kmeans(t, np.full(2 * n, 2, NP[t]), np.ones(3 * n, NP[t]), n)
        assert hashlib.sha256(m.tobytes()).hexdigest() == g["sha256"], g
        assert list(m[:4]) == g["first"] and list(m[-4:]) == g["last"] and g["reference_verify"] == "PASS"


def test_oracle_kmeans_index_quirk(O):
    """`index = gid` (kmeans.cpp:74): the point's own index when a distance is below FLT_MAX, 0 when none is (overflow, NaN)."""
    n = 8
    f = np.array([0, 1, 2, 3e30, np.nan, 5, 6, 7] + [0] * 8, np.float32)  # feature 0 | feature 1
    c = np.array([1, 1, 2, 2, -3e30, 0], np.float32)
    # point 3: distances to clusters 0 / 1 overflow to inf, cluster 2: (3e30 + 3e30)^2 = inf as well -> 0; point 4: NaN -> 0
    assert list(O.kmeans("f32", f, c, n)) == [0, 1, 2, 0, 0, 5, 6, 7]
    assert list(O.kmeans("f64", f.astype(np.float64), c.astype(np.float64), n)) == [0, 1, 2, 0, 0, 5, 6, 7]  # 3.6e61 > FLT_MAX too


@pytest.mark.parametrize("which", ["lin_reg_error", "kmeans"])
def test_live_reference_matches_oracle(O, which):
    if not O.ref_available(which):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    size = 2048
    rows, dumps = O.run_reference(which, size, 256, num_runs=1, ndrange=True)
    assert len(dumps) == 2 and all(r["Verification"] == "PASS" for r in rows.values())
    for k, t in enumerate(("f32", "f64")):
        if which == "lin_reg_error":
            ins = O.lin_reg_error_inputs(t, size, 0 if k == 0 else 4 * size)
            assert_within(t, size, O.lin_reg_error(t, *ins), np.frombuffer(dumps[k], NP[t]), f"live {which} {t}")
        else:
            want = O.kmeans(t, np.full(2 * size, 2, NP[t]), np.ones(3 * size, NP[t]), size)
            assert dumps[k] == want.tobytes()


def test_binaries_built():
    for b in ("lin_reg_error", "kmeans"):
        assert os.access(os.path.join(BIN, b), os.X_OK), f"{b} not built: make -C sycl-bench_b200 pattern"


# ------------------------------------------------------------------------------------------- GPU: C ABI
def _dev(x):
    import torch

    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.mark.gpu
def test_gpu_lin_reg_error_matches_golden(ctx, O, golden_w):
    import torch

    for g in golden_w["lin_reg_error"]:
        t, n = g["dtype"], g["size"]
        idx, want = golden_values(g)
        d = [_dev(a) for a in O.lin_reg_error_inputs(t, n, g["rand_skip"])]
        out = torch.zeros(n, dtype=d[0].dtype, device="cuda")
        ctx.linreg_error(*d, out)
        ctx.sync()
        got = out.cpu().numpy()[idx]
        assert_within(t, n, got, want, f"gpu lin_reg_error {t} n={n}")
        assert np.linalg.norm(got.astype(np.float64) - want) / np.linalg.norm(want.astype(np.float64)) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("t", ["f32", "f64"])
@pytest.mark.parametrize("n,offset", [(1, 0), (2, 0), (7, 1), (1000, 0), (4099, 3), (8192, 0), (12289, 1)])
def test_gpu_lin_reg_error_matches_oracle_on_random_inputs(ctx, O, t, n, offset):
    import torch

    rng = np.random.default_rng(n * 7 + offset)
    ins = [(rng.random(n + offset) * 4 - 2).astype(NP[t]) for _ in range(4)]
    d = [_dev(a) for a in ins]
    out = torch.full((n + offset,), -1, dtype=d[0].dtype, device="cuda")
    ctx.linreg_error(*[a[offset:] for a in d], out[offset:], n=n)  # offset != 0: element-aligned only -> scalar paths
    ctx.sync()
    got = out.cpu().numpy()
    assert (got[:offset] == -1).all()
    want = O.lin_reg_error(t, *[np.ascontiguousarray(a[offset:]) for a in ins])
    assert_within(t, n, got[offset:], want, f"gpu lin_reg_error random {t} n={n}")


@pytest.mark.gpu
@pytest.mark.parametrize("t", ["f32", "f64"])
def test_gpu_lin_reg_error_large_and_exact_fit(ctx, t):
    """2^22 points against the quadratic form in long double (the literal loop would be 1.8e13 iterations), on the worst
    input for that form: ALL points lie exactly on y = 0.5 x + 0.25, so candidates near (0.5, 0.25) have an error that is
    tiny against the O(N) terms it is the difference of.  The five sums carry double rounding errors, so such candidates
    are resolved to an ABSOLUTE accuracy of a few N * 2^-52 (1e-10 here, measured) rather than to the relative tolerance —
    the conditioning of the closed form, stated in include/sb_cuda.h; the reference's own criterion (relative L2 error over
    the whole output below 1e-6) is met with ten orders of magnitude to spare.  The exact candidate itself gives ~0."""
    import torch

    n = 1 << 22
    rng = np.random.default_rng(3)
    x = rng.random(n).astype(NP[t])
    y = (x * NP[t](0.5) + NP[t](0.25)).astype(NP[t])
    a, b = rng.random(n).astype(NP[t]), rng.random(n).astype(NP[t])
    a[5], b[5] = 0.5, 0.25
    d = [_dev(v) for v in (x, y, a, b)]
    out = torch.zeros(n, dtype=d[0].dtype, device="cuda")
    ctx.linreg_error(*d, out)
    ctx.sync()
    got = out.cpu().numpy()
    want = exact_errors(x, y, a, b)
    assert_within(t, n, got, want, f"gpu lin_reg_error 2^22 {t}", abs_floor=n * 2.0 ** -48)
    assert np.linalg.norm(got.astype(np.float64) - want) / np.linalg.norm(want) < (1e-14 if t == "f64" else 1e-6)
    assert abs(float(got[5])) <= n * 2.0 ** -48


@pytest.mark.gpu
def test_gpu_kmeans_matches_golden(ctx, golden_w):
    import torch

    for g in golden_w["kmeans"]:
        t, n = g["dtype"], g["size"]
        f, c = _dev(np.full(2 * n, 2, NP[t])), _dev(np.ones(3 * n, NP[t]))
        m = torch.full((n,), -7, dtype=torch.int32, device="cuda")
        ctx.kmeans(f, c, m, n)
        ctx.sync()
        assert hashlib.sha256(m.cpu().numpy().tobytes()).hexdigest() == g["sha256"], g


@pytest.mark.gpu
@pytest.mark.parametrize("t", ["f32", "f64"])
@pytest.mark.parametrize("n,nf,nc,offset", [(1, 2, 3, 0), (5, 2, 3, 0), (1000, 2, 3, 0), (1003, 2, 3, 0), (4096, 2, 3, 1), (65536, 2, 3, 0),
                                            (70001, 1, 1, 0), (4096, 4, 7, 0), (4100, 6, 5, 0), (2048, 9, 300, 0), (512, 3, 0, 0), (512, 0, 2, 0)])
def test_gpu_kmeans_matches_oracle_on_random_inputs(ctx, O, t, n, nf, nc, offset):
    """Random features with a sprinkling of huge values (distance overflows to inf: membership 0) and NaNs; vector path,
    scalar path (row pitch not a multiple of 16 bytes / unaligned base), features in registers (<= 4) and re-read (> 4),
    cluster table in shared memory and (300 x 9 > 2048 elements) in global memory, degenerate shapes."""
    import torch

    rng = np.random.default_rng(n + nf * 31 + nc)
    f = (rng.standard_normal(max(nf * n, 1) + offset) * 3).astype(NP[t])
    big = NP[t](3e30) if t == "f32" else NP[t](1e200)
    f[rng.random(f.size) < 0.02] = big
    f[rng.random(f.size) < 0.01] = np.nan
    c = rng.standard_normal(max(nc * nf, 1)).astype(NP[t])
    df, dc = _dev(f), _dev(c)
    m = torch.full((n + 1,), -7, dtype=torch.int32, device="cuda")
    ctx.kmeans(df[offset:], dc, m, n, nf, nc)
    ctx.sync()
    got = m.cpu().numpy()
    assert got[n] == -7
    want = O.kmeans(t, np.ascontiguousarray(f[offset:]), c, n, nf, nc)
    np.testing.assert_array_equal(got[:n], want)
    if nf and nc:
        assert (want == 0).sum() > (1 if n > 100 else 0) or n < 100  # the overflow / NaN path is exercised


# ------------------------------------------------------------------------------------------- GPU: binaries
def _run(bench, tmp_path, *flags):
    from oracle import oracle as O

    out = tmp_path / f"{bench}.csv"
    r = subprocess.run([os.path.join(BIN, bench), f"--output={out}", *flags], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    return (O.parse_csv(str(out)) if out.exists() else {}), r


@pytest.mark.gpu
def test_binaries_rows_and_verdicts(tmp_path):
    rows, _ = _run("lin_reg_error", tmp_path, "--size=65536", "--local=256", "--num-runs=3")  # upstream's suite size
    assert list(rows) == ["LinearRegression_fp32", "LinearRegression_fp64"]
    assert all(r["Verification"] == "PASS" and r["kernel-launches-per-run"] == "2" for r in rows.values())
    rows, _ = _run("kmeans", tmp_path, "--size=1048576", "--local=256", "--num-runs=3")
    assert list(rows) == ["Kmeans_fp32", "Kmeans_fp64"]
    assert all(r["Verification"] == "PASS" and r["kernel-launches-per-run"] == "1" for r in rows.values())
    rows, _ = _run("kmeans", tmp_path, "--size=1000003", "--local=1", "--num-runs=1")  # odd size: the scalar path
    assert all(r["Verification"] == "PASS" for r in rows.values())
