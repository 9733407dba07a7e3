"""bench.py prints ONE JSON line with the contract's keys (small sizes; the driver runs the real one)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
        "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"]


def test_reference_arm_runs_on_cpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-log2n", "20"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["gpu_launches"] == 0
    assert line["value"] > 0 and line["unit"] == "GB/s" and line["metric"].startswith("pattern reduce/scan")
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    cb = line["cpu_baseline"]
    assert cb["cores"] >= 1 and cb["kind"] in ("reference", "port") and "sample" in cb
    assert set(cb["legs_kind"]) == {"reduce_f32", "reduce_i64", "scan_f32", "scan_i64"}
    assert all(cb["legs_kind"][k] == "port" for k in ("scan_f32", "scan_i64"))


@pytest.mark.gpu
def test_bench_line_contract():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--log2n", "22", "--cpu-log2n", "20",
                        "--config5-log2-total", "24", "--replicas-log2n", "22"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in KEYS:
        assert k in line, k
    assert line["n_gpus"] == 1 and line["steps"] == 3 and line["verified"] is True
    assert line["gpu_launches"] == 3 * 4  # four libsb_cuda launches per step
    rf = line["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-3
    e = line["e2e"]
    assert e["verified"] and e["h2d_bytes_per_step"] == (4 + 8) * 2 * (1 << 22) and e["d2h_bytes_per_step"] == (4 + 8) * (1 << 22) + 12
    assert e["value"] < line["value"]  # PCIe-bound end to end vs HBM-resident
    assert "workload" in line["config"] and "model" not in line["config"]
    # every leg of the timed step is verified, floating point included
    assert line["verified_legs"] == {"reduce_f32": True, "reduce_i64": True, "scan_f32": True, "scan_i64": True}
    # sub-records: BASELINE configs[4] (2^33 sharded; scaled down here) and the exchange-free legs of configs[2] / [3]
    c5 = line["config5"]
    assert c5["verified"] and set(c5["per_leg"]) == {"reduce_i32", "scan_i32"} and c5["elements_total"] == 1 << 24
    rp = line["replicas"]
    assert rp["verified"] and set(rp["config3"]["per_leg"]) == {"gscan_i32_L256", "segscan_i32_L256", "segscan_flags_i32_L256"}
    assert set(rp["config4"]["per_leg"]) == {f"segreduce_f64_L{L}" for L in (64, 128, 256, 512, 1024)}
    assert e["roofline"]["bound"] == "pcie" and 0 < e["roofline"]["frac"] <= 1.5
