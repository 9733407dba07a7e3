"""The C++ plugin surface: binaries exist and fail fast without a GPU (CPU); on the B200 every
pattern/ binary runs through BenchmarkApp/BenchmarkManager, verifies, and writes the reference's
CSV schema (include/time_metrics.h:104-144, include/result_consumer.h:67-92 upstream)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "sycl-bench_b200", "bin")
PATTERNS = ["reduction", "prefixsum", "scan", "segmentedreduction", "segmentedscan"]

REFERENCE_COLUMNS = ["Verification", "device-name", "local-size", "problem-size", "sycl-implementation", "throughput-metric"] + [
    f"{t}-{s}" for t in ("kernel-time", "run-time", "submit-time", "system-time") for s in ("mean", "median", "min", "samples", "stddev", "throughput")
]


def test_binaries_built():
    for p in PATTERNS:
        assert os.access(os.path.join(BIN, p), os.X_OK), f"{p} not built: python -c 'import __graft_entry__ as g; g.build()'"


def test_fail_fast_on_bad_flags():
    import torch

    r = subprocess.run([os.path.join(BIN, "reduction"), "--device=cpu"], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU path" in r.stderr
    r = subprocess.run([os.path.join(BIN, "reduction"), "--size=8", "--size=9"], capture_output=True, text=True)
    assert r.returncode == 2 and "several times" in r.stderr
    if not torch.cuda.is_available():
        r = subprocess.run([os.path.join(BIN, "scan")], capture_output=True, text=True)
        assert r.returncode == 2 and "no CPU fallback" in r.stderr


def _run(bench, tmp_path, *flags):
    from oracle import oracle as O

    out = tmp_path / f"{bench}.csv"
    r = subprocess.run([os.path.join(BIN, bench), f"--output={out}", *flags], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    return O.parse_csv(str(out)), r


@pytest.mark.gpu
@pytest.mark.parametrize("bench,rows", [("reduction", 4), ("prefixsum", 4), ("scan", 8), ("segmentedreduction", 5), ("segmentedscan", 8)])
def test_benchmark_rows_and_schema(tmp_path, bench, rows):
    table, _ = _run(bench, tmp_path, "--size=1048576", "--local=256", "--num-runs=3", "--warmup-run", "--strict")
    assert len(table) == rows
    for name, row in table.items():
        assert name.startswith("Pattern_")
        for c in REFERENCE_COLUMNS:
            assert c in row, (name, c)
        assert row["Verification"] == "PASS", name
        assert row["problem-size"] == "1048576" and row["local-size"] == "256"
        assert len(row["kernel-time-samples"].strip('"').split()) == 3  # the warm-up run is dropped
        assert float(row["kernel-time-min"]) > 0 and float(row["run-time-min"]) >= float(row["kernel-time-min"]) * 0.5
        # one launch per run — except explicit head flags from 64 tiles on: tile-head check + the gated kernel pair
        assert row["kernel-launches-per-run"] == ("3" if "_Flags_" in name else "1"), name
        assert "NVIDIA" in row["device-name"]


@pytest.mark.gpu
def test_reference_row_names(tmp_path):
    table, _ = _run("reduction", tmp_path, "--size=65536", "--num-runs=1")
    assert sorted(table) == [f"Pattern_Reduction_NDRange_{t}" for t in ("fp32", "fp64", "int32", "int64")]
    table, _ = _run("segmentedreduction", tmp_path, "--size=65536", "--num-runs=1")
    assert sorted(table) == [f"Pattern_SegmentedReduction_NDRange_{t}" for t in ("fp32", "fp64", "int16", "int32", "int64")]


@pytest.mark.gpu
def test_reference_verify_mode_reproduces_F3(tmp_path):
    """--reference-verify applies upstream's int-accumulator check: PASS at 65 536, FAIL above."""
    table, _ = _run("reduction", tmp_path, "--size=65536", "--num-runs=1", "--reference-verify")
    assert all(r["Verification"] == "PASS" for r in table.values())
    table, _ = _run("reduction", tmp_path, "--size=1048576", "--num-runs=1", "--reference-verify")
    assert all(r["Verification"] == "FAIL" for r in table.values())
    table, _ = _run("reduction", tmp_path, "--size=1048576", "--num-runs=1")
    assert all(r["Verification"] == "PASS" for r in table.values())


@pytest.mark.gpu
def test_dropped_tail_and_odd_local(tmp_path):
    # N % L != 0: the reference drops the tail; non-power-of-two --local for segmentedreduction follows the reference tree
    table, _ = _run("reduction", tmp_path, "--size=3000", "--local=256", "--num-runs=1")
    assert all(r["Verification"] == "PASS" for r in table.values())
    table, _ = _run("segmentedreduction", tmp_path, "--size=1200", "--local=100", "--num-runs=1")
    assert all(r["Verification"] == "PASS" for r in table.values())
    table, _ = _run("scan", tmp_path, "--size=100000", "--local=96", "--num-runs=1", "--scan-scope=group")
    assert len(table) == 4 and all(r["Verification"] == "PASS" for r in table.values())


@pytest.mark.gpu
def test_no_ndrange_and_no_verification(tmp_path):
    out = tmp_path / "x.csv"
    r = subprocess.run([os.path.join(BIN, "reduction"), f"--output={out}", "--no-ndrange-kernels"], capture_output=True, text=True)
    assert r.returncode == 0 and "nothing to run" in r.stderr
    table, _ = _run("prefixsum", tmp_path, "--size=4096", "--num-runs=1", "--no-verification")
    assert all(r["Verification"] == "N/A" for r in table.values())


@pytest.mark.gpu
def test_stdio_consumer():
    r = subprocess.run([os.path.join(BIN, "segmentedscan"), "--size=8192", "--local=64", "--num-runs=1", "--no-flags"], capture_output=True, text=True)
    assert r.returncode == 0
    assert "********** Results for Pattern_SegmentedScan_NDRange_int32**********" in r.stdout
    assert "kernel-time-min: " in r.stdout and "[s]" in r.stdout and "Verification: PASS" in r.stdout


@pytest.mark.gpu
def test_nvml_energy_hook_columns(tmp_path):
    """harness/nv_energy_meas.h (the reference's empty include/nv_energy_meas.h): clock / power / throttle columns
    from NVML, matched to the CUDA device by PCI bus id; --no-nvml removes them."""
    table, _ = _run("reduction", tmp_path, "--size=268435456", "--num-runs=20")
    for row in table.values():
        assert float(row["sm-clock-mhz"]) > 0 and float(row["sm-clock-max-mhz"]) >= float(row["sm-clock-mhz"])
        assert float(row["power-w"]) > 0
        assert row["throttle-reasons"] == "none" or all(p in ("gpu_idle", "applications_clocks_setting", "sw_power_cap", "hw_slowdown", "sync_boost",
                                                              "sw_thermal_slowdown", "hw_thermal_slowdown", "hw_power_brake_slowdown",
                                                              "display_clock_setting") for p in row["throttle-reasons"].split("|"))
    os.remove(tmp_path / "reduction.csv")
    table, _ = _run("reduction", tmp_path, "--size=65536", "--num-runs=1", "--no-nvml")
    assert all("sm-clock-mhz" not in row for row in table.values())


@pytest.mark.gpu
def test_run_suite_quicktest(tmp_path):
    out = tmp_path / "suite.csv"
    r = subprocess.run([os.path.join(ROOT, "sycl-bench_b200", "run-suite"), "quicktest", f"--output={out}"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    from oracle import oracle as O

    table = O.parse_csv(str(out))
    assert len(table) == 4 + 4 + 8 + 5 + 8 + 4 + 4 + 12 + 4 + 2 + 2 + 2 + 6  # pattern/ + sycl2020/ trio + single-kernel/ + micro/DRAM rows
    assert all(row["Verification"] == "PASS" for row in table.values())
    r = subprocess.run([os.path.join(ROOT, "sycl-bench_b200", "run-suite"), "quicktest", f"--output={out}"], capture_output=True, text=True)
    assert r.returncode == 1 and "already exists" in r.stdout
