"""The sycl2020/ reduction trio: ReductionAtomic, ReduceGroupAlgorithm, KernelReduction.

CPU: oracle/sb_oracle_sycl2020.c is pinned to the golden vectors the UNMODIFIED reference produced
(tests/golden/sycl2020_golden.json, made by tests/golden/make_golden_sycl2020.py) and to the live
oracle/_ref binaries.  GPU: sb_fill + sb_reduce through the C ABI, and the three drop-in binaries, against
the same vectors.
"""
import json
import os
import struct
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "sycl-bench_b200", "bin")
FMT = {"i32": "<i", "i64": "<q", "f32": "<f", "f64": "<d"}
NAME = {"i32": "int32", "i64": "int64", "f32": "fp32", "f64": "fp64"}
TRIO = ["atomic_reduction", "reduce_over_group", "kernel_reduction"]


@pytest.fixture(scope="module")
def golden2020():
    with open(os.path.join(ROOT, "tests", "golden", "sycl2020_golden.json")) as f:
        return json.load(f)


def hexof(t, value):
    return struct.pack(FMT[t], value).hex()


def oracle_value(O, which, g):
    L = O.lib()
    if which == "atomic_reduction":
        return getattr(L, f"ref2020_atomic_{g['dtype']}")(g["size"])
    if which == "reduce_over_group":
        return getattr(L, f"ref2020_group_{g['dtype']}")(g["size"], g["local"])
    return getattr(L, f"ref2020_kernel_{g['dtype']}")(g["size"], g["cf"])


# ------------------------------------------------------------------------------------------- CPU: oracle
@pytest.mark.parametrize("which", TRIO)
def test_oracle_reproduces_golden_bit_exact(O, golden2020, which):
    assert len(golden2020[which]) >= 28
    for g in golden2020[which]:
        assert hexof(g["dtype"], oracle_value(O, which, g)) == g["result_hex"], (which, g)


def test_golden_holds_the_fp32_saturation(golden2020):
    """atomic_reduction.cpp:33-40 with N = 2^25 ones: fp32 stops at 2^24, every other type counts to N."""
    rows = {g["dtype"]: g for g in golden2020["atomic_reduction"] if g["size"] == 1 << 25}
    assert rows["f32"]["result"] == float(1 << 24)
    assert rows["i32"]["result"] == rows["i64"]["result"] == 1 << 25 and rows["f64"]["result"] == float(1 << 25)
    group = {g["dtype"]: g for g in golden2020["reduce_over_group"] if g["size"] == 1 << 25}
    assert group["f32"]["result"] == float(1 << 25)  # 2^17 additions of 256.0f stay exact


def test_oracle_verify_matches_reference_verdict(O, golden2020):
    """`T verified_results = problem_size` + percentDiff <= 0.05 (kernel_reduction.cpp:43-53): a dropped
    coarsening tail fails only when it is a visible fraction of N (size 7: FAIL, size 1000003: PASS)."""
    L = O.lib()
    for which in ("reduce_over_group", "kernel_reduction"):
        for g in golden2020[which]:
            ok = getattr(L, f"ref2020_verify_size_{g['dtype']}")(g["result"], g["size"])
            assert ("PASS" if ok else "FAIL") == g["reference_verify"], (which, g)
    for g in golden2020["atomic_reduction"]:
        if g["dtype"] == "f32" and g["size"] > 1 << 24:
            # upstream's expected value is accumulated on the host in T (atomic_reduction.cpp:46-47): built
            # with its Release flags (-ffast-math) the loop is vectorised and reaches N, so the saturated
            # device result FAILS; a scalar build saturates too and PASSES.  Flag dependent: not asserted.
            continue
        ok = getattr(L, f"ref2020_verify_atomic_{g['dtype']}")(g["result"], g["size"])
        assert ("PASS" if ok else "FAIL") == g["reference_verify"], g


def test_percent_diff_restatement(O):
    L = O.lib()
    assert L.ref2020_percent_diff(0.001, 0.002) == 0.0
    assert abs(L.ref2020_percent_diff(100.0, 99.0) - 1.0) < 1e-5
    assert abs(L.ref2020_percent_diff(4.0, 7.0) - 75.0) < 1e-4


@pytest.mark.parametrize("which", TRIO)
def test_live_reference_matches_oracle(O, which):
    if not O.ref_available(which):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    size, local = 8192, 128
    rows, dumps = O.run_reference(which, size, local, num_runs=1, ndrange=True, dump_max_bytes=8)
    types = [(t, cf) for t in O.REF_REDUCTION_TYPES for cf in ((1, 4, 8) if which == "kernel_reduction" else (None,))]
    assert len(dumps) == len(types)
    for (t, cf), raw in zip(types, dumps):
        g = {"dtype": t, "size": size, "local": local, "cf": cf}
        assert raw[: struct.calcsize(FMT[t])].hex() == hexof(t, oracle_value(O, which, g))
    assert all(r["Verification"] == "PASS" for r in rows.values())


def test_trio_binaries_built():
    for b in TRIO:
        assert os.access(os.path.join(BIN, b), os.X_OK), f"{b} not built: make -C sycl-bench_b200 pattern"


# ------------------------------------------------------------------------------------------- GPU: C ABI
@pytest.mark.gpu
@pytest.mark.parametrize("which", TRIO)
def test_gpu_matches_golden(ctx, golden2020, which):
    import torch

    TORCH = {"i32": torch.int32, "i64": torch.int64, "f32": torch.float32, "f64": torch.float64}
    for g in golden2020[which]:
        t, n = g["dtype"], g["size"]
        effective = (n // g["cf"]) * g["cf"] if which == "kernel_reduction" else n
        x = torch.empty(max(n, 1), dtype=TORCH[t], device="cuda")
        out = torch.full((1,), -7, dtype=TORCH[t], device="cuda")
        ctx.fill(x, 1, n=n)
        ctx.reduce(x, out, n=effective)
        ctx.sync()
        got = out.cpu().numpy()[0]
        if which == "atomic_reduction" and t == "f32" and n > 1 << 24:
            # the reference saturates at 2^24 (see test_golden_holds_the_fp32_saturation); sb_reduce counts
            assert g["result"] == float(1 << 24) and float(got) == float(n)
            continue
        assert hexof(t, got.item()) == g["result_hex"], (which, g, got)


@pytest.mark.gpu
@pytest.mark.parametrize("t", ["i32", "i64", "f32", "f64"])
@pytest.mark.parametrize("n,value", [(0, 1), (1, 3), (1000, -2), (4099, 5), ((1 << 20) + 1, 1)])
def test_gpu_fill(ctx, t, n, value):
    import torch

    TORCH = {"i32": torch.int32, "i64": torch.int64, "f32": torch.float32, "f64": torch.float64}
    buf = torch.zeros(n + 9, dtype=TORCH[t], device="cuda")
    ctx.fill(buf[4:], value, n=n)  # element-aligned but not 16-byte aligned for the 4-byte types
    ctx.sync()
    host = buf.cpu()
    assert (host[4: 4 + n] == value).all() and (host[:4] == 0).all() and (host[4 + n:] == 0).all()


@pytest.mark.gpu
def test_gpu_ones_at_baseline_size(ctx):
    """2^28 ones per type: integers and fp64 count exactly, fp32 within the north-star tolerance
    (the reference's ReductionAtomic_fp32 would stop at 2^24 here)."""
    import torch

    n = 1 << 28
    for t, dt in (("i32", torch.int32), ("i64", torch.int64), ("f32", torch.float32), ("f64", torch.float64)):
        x = torch.empty(n, dtype=dt, device="cuda")
        out = torch.zeros(1, dtype=dt, device="cuda")
        ctx.fill(x, 1)
        ctx.reduce(x, out)
        ctx.sync()
        got = float(out.cpu()[0])
        assert got == float(n) if t != "f32" else abs(got - n) <= 1e-5 * 28 * n
        del x


# ------------------------------------------------------------------------------------------- GPU: binaries
def _run(bench, tmp_path, *flags):
    from oracle import oracle as O

    out = tmp_path / f"{bench}.csv"
    r = subprocess.run([os.path.join(BIN, bench), f"--output={out}", *flags], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    return (O.parse_csv(str(out)) if out.exists() else {}), r


@pytest.mark.gpu
def test_binaries_rows_names_and_verdicts(tmp_path):
    rows, _ = _run("atomic_reduction", tmp_path, "--size=1048576", "--local=256", "--num-runs=3")
    assert list(rows) == [f"ReductionAtomic_{NAME[t]}" for t in FMT] and all(r["Verification"] == "PASS" for r in rows.values())
    rows, _ = _run("reduce_over_group", tmp_path, "--size=1048576", "--local=1024", "--num-runs=3")
    assert list(rows) == [f"ReduceGroupAlgorithm_{NAME[t]}" for t in FMT] and all(r["Verification"] == "PASS" for r in rows.values())
    rows, _ = _run("kernel_reduction", tmp_path, "--size=1000003", "--num-runs=3")
    assert list(rows) == [f"KernelReduction_{NAME[t]}_plus_cf{cf}" for t in FMT for cf in (1, 4, 8)]
    assert all(r["Verification"] == "PASS" for r in rows.values())
    assert all(r["kernel-launches-per-run"] == "1" for r in rows.values())


@pytest.mark.gpu
def test_binaries_follow_upstream_on_edge_cases(tmp_path, golden2020):
    """size 7: the cf4 / cf8 rows FAIL upstream's 0.05 % criterion (golden), and so do they here;
    an nd_range with size % local != 0 is an error upstream (the row is discarded), and here."""
    rows, _ = _run("kernel_reduction", tmp_path, "--size=7", "--num-runs=1")
    want = {f"KernelReduction_{NAME[g['dtype']]}_plus_cf{g['cf']}": g["reference_verify"] for g in golden2020["kernel_reduction"] if g["size"] == 7}
    assert {k: v["Verification"] for k, v in rows.items()} == want
    rows, r = _run("atomic_reduction", tmp_path, "--size=1000", "--local=300", "--num-runs=1")
    assert rows == {} and "nd_range" in (r.stderr + r.stdout)
