"""Property tests of the oracle (test infrastructure) on random integer inputs — CPU only.

Integers wrap mod 2^w, so every restated reference algorithm with a power-of-two --local must agree with a plain
modular sum whatever the order: these properties pin the restatements independently of the golden vectors."""
import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

NP = {"i16": np.int16, "i32": np.int32, "i64": np.int64}
BITS = {"i16": 16, "i32": 32, "i64": 64}


def wrap(v, t):
    b = BITS[t]
    return (int(v) + (1 << (b - 1))) % (1 << b) - (1 << (b - 1))


def rand(rng, t, n):
    b = BITS[t]
    return rng.integers(-(1 << (b - 2)), 1 << (b - 2), size=n).astype(NP[t])


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 2**31), groups=st.integers(1, 40), logl=st.integers(1, 8), t=st.sampled_from(["i32", "i64"]))  # --local=1 never terminates upstream (G stays n, reduction.cpp:83-92)
def test_reduce_is_the_modular_sum(O, seed, groups, logl, t):
    L = 1 << logl
    x = rand(np.random.default_rng(seed), t, groups * L)
    assert int(O.reduce(t, x, L)) == wrap(sum(int(v) for v in x), t)


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 2**31), groups=st.integers(1, 40), logl=st.integers(0, 8), t=st.sampled_from(["i16", "i32", "i64"]))
def test_segreduce_heads_are_segment_sums_and_the_rest_is_untouched(O, seed, groups, logl, t):
    L = 1 << logl
    x = rand(np.random.default_rng(seed), t, groups * L)
    y = O.segreduce(t, x, L)
    for g in range(groups):
        assert int(y[g * L]) == wrap(sum(int(v) for v in x[g * L:(g + 1) * L]), t)
    mask = np.ones(x.size, bool)
    mask[::L] = False
    assert np.array_equal(y[mask], x[mask])


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 2**31), n=st.integers(1, 3000), group=st.integers(0, 300), exclusive=st.booleans(), t=st.sampled_from(["i32", "i64"]))
def test_scan_matches_a_python_loop(O, seed, n, group, exclusive, t):
    x = rand(np.random.default_rng(seed), t, n)
    y = O.scan(t, x, exclusive=exclusive, group_len=group)
    run = 0
    for i in range(n):
        if group and i % group == 0:
            run = 0
        want = run if exclusive else run + int(x[i])
        assert int(y[i]) == wrap(want, t), i
        run += int(x[i])


@settings(max_examples=30, deadline=None)
@given(seed=st.integers(0, 2**31), n=st.integers(1, 3000), density=st.floats(0.0, 0.3))
def test_flagged_scan_restarts_exactly_at_heads(O, seed, n, density):
    rng = np.random.default_rng(seed)
    x = rand(rng, "i64", n)
    flags = (rng.random(n) < density).astype(np.uint8)
    y = O.scan("i64", x, flags=flags)
    run = 0
    for i in range(n):
        if flags[i] or i == 0:
            run = 0
        run += int(x[i])
        assert int(y[i]) == wrap(run, "i64"), i


@settings(max_examples=30, deadline=None)
@given(seed=st.integers(0, 2**31), groups=st.integers(1, 30), logl=st.integers(1, 7), t=st.sampled_from(["i32", "i64"]))
def test_scalar_prod_is_the_modular_dot_product(O, seed, groups, logl, t):
    L = 1 << logl
    rng = np.random.default_rng(seed)
    a, b = rand(rng, t, groups * L) >> 8, rand(rng, t, groups * L) >> 8
    want = wrap(sum(int(u) * int(v) for u, v in zip(a, b)), t)
    assert int(O.scalar_prod(t, a, b, L)) == want


@pytest.mark.parametrize("t", ["f32", "f64"])
def test_linreg_recovers_an_exact_line(O, t):
    """y = 3x + 0.5 on a power-of-two grid with x in multiples of 1/8: every sum is exact in both types."""
    n = 1 << 12
    x = (np.arange(n) % 64 / 8.0).astype(np.float32 if t == "f32" else np.float64)
    y = (3.0 * x + 0.5).astype(x.dtype)
    out = O.linreg(t, x, y, 64)
    sx, sy, sxy, sxx = x.astype(np.float64).sum(), y.astype(np.float64).sum(), (x.astype(np.float64) * y).sum(), (x.astype(np.float64) ** 2).sum()
    assert [float(v) for v in out[:4]] == [sxy, sxx, sx, sy]
    # the reference's formula is not the textbook one (it subtracts mean_x*mean_y, not n*mean_x*mean_y): check it as written
    mx, my = sx / n, sy / n
    b1 = (sxy - mx * my) / (sxx - mx * mx)
    assert abs(float(out[4]) - b1) <= 1e-5 * abs(b1) and abs(float(out[5]) - (my - b1 * mx)) <= 1e-4


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 2**31), n=st.integers(1, 300), t=st.sampled_from(["f32", "f64"]))
def test_lin_reg_error_on_exactly_representable_points(O, seed, n, t):
    """Points and candidates on a coarse binary grid: every product and partial sum of the reference's loop
    (lin_reg_error.cpp:64-77) is exact in both types, so the restatement must equal the integer arithmetic bit for bit."""
    rng = np.random.default_rng(seed)
    dt = np.float32 if t == "f32" else np.float64
    x, y, a, b = [(rng.integers(-8, 9, n) / 4.0).astype(dt) for _ in range(4)]
    got = O.lin_reg_error(t, x, y, a, b)
    xi, yi, ai, bi = [np.round(v.astype(np.float64) * 4).astype(np.int64) for v in (x, y, a, b)]
    for g in range(n):
        e = ai[g] * xi + 4 * bi[g] - 4 * yi  # 16 * ((a x + b) - y)
        assert float(got[g]) == float((e * e).sum()) / 256.0


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 2**31), n=st.integers(1, 200), nf=st.integers(0, 5), nc=st.integers(0, 4), t=st.sampled_from(["f32", "f64"]))
def test_kmeans_membership_is_the_point_index_or_zero(O, seed, n, nf, nc, t):
    """kmeans.cpp:61-79 stores `gid` when a distance is below FLT_MAX and 0 otherwise — never a cluster index."""
    rng = np.random.default_rng(seed)
    dt = np.float32 if t == "f32" else np.float64
    f = rng.standard_normal(max(nf * n, 1)).astype(dt)
    f[rng.random(f.size) < 0.1] = np.nan
    c = rng.standard_normal(max(nf * nc, 1)).astype(dt)
    m = O.kmeans(t, f, c, n, nf, nc)
    for g in range(n):
        dists = [sum((float(f[l * n + g]) - float(c[i * nf + l])) ** 2 for l in range(nf)) for i in range(nc)]
        finite = any(d < 3.402823466e38 for d in dists)  # NaN compares false
        assert m[g] == (g if finite else 0)
