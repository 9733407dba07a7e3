"""micro/DRAM: sb_copy through the C ABI and the MicroBench_DRAM rows.

The reference kernel is `out[gid] = in[gid]` over a buffer of --size^3 bytes filled with 33
(/root/reference/micro/DRAM.cpp:11-22,42,54-62); parity is byte equality of output and input.
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "sycl-bench_b200", "bin")


def test_binary_built():
    assert os.access(os.path.join(BIN, "DRAM"), os.X_OK), "DRAM not built: make -C sycl-bench_b200 pattern"


@pytest.mark.gpu
@pytest.mark.parametrize("nbytes,src_off,dst_off", [(0, 0, 0), (1, 0, 0), (15, 3, 3), (4096, 0, 0), (100003, 5, 5), (100003, 1, 2),
                                                    ((1 << 24) + 7, 0, 0), ((1 << 24) + 7, 8, 8), (65536, 0, 4)])
def test_gpu_copy_is_byte_exact(ctx, nbytes, src_off, dst_off):
    import torch

    g = torch.Generator(device="cuda").manual_seed(nbytes + src_off)
    src = torch.randint(0, 256, (nbytes + 64,), dtype=torch.uint8, device="cuda", generator=g)
    dst = torch.full((nbytes + 64,), 0xA5, dtype=torch.uint8, device="cuda")
    ctx.copy(dst[dst_off:], src[src_off:], nbytes)
    ctx.sync()
    assert torch.equal(dst[dst_off: dst_off + nbytes], src[src_off: src_off + nbytes])
    assert (dst[:dst_off] == 0xA5).all() and (dst[dst_off + nbytes:] == 0xA5).all()  # nothing outside the range is touched


@pytest.mark.gpu
def test_gpu_copy_large(ctx):
    import torch

    n = 1 << 28
    src = torch.empty(n, dtype=torch.float32, device="cuda")
    ctx.fill(src, 33.0)
    dst = torch.zeros_like(src)
    ctx.copy(dst, src)
    ctx.sync()
    assert torch.equal(dst, src) and float(dst[-1]) == 33.0


@pytest.mark.gpu
def test_dram_binary_rows(tmp_path):
    from oracle import oracle as O

    out = tmp_path / "dram.csv"
    r = subprocess.run([os.path.join(BIN, "DRAM"), "--size=256", "--num-runs=3", f"--output={out}"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    rows = O.parse_csv(str(out))
    assert list(rows) == [f"MicroBench_DRAM_{t}_{d}" for t in ("fp32", "fp64") for d in (1, 2, 3)]
    for row in rows.values():
        assert row["Verification"] == "PASS" and row["kernel-launches-per-run"] == "1"
        assert abs(float(row["throughput-metric"]) - 2 * 256**3 / 2**30) < 1e-9  # 2 x buffer GiB, DRAM.cpp:45-50


@pytest.mark.gpu
def test_gpu_copy_rejects_partial_overlap(ctx, pkg):
    """The copy kernel reads through the read-only path: overlapping ranges are refused (dst == src is a no-op)."""
    import torch

    buf = torch.zeros(4096, dtype=torch.uint8, device="cuda")
    with pytest.raises(pkg.SbError) as e:
        ctx.copy(buf[16:], buf, 1024)
    assert e.value.status == -1
    with pytest.raises(pkg.SbError):
        ctx.copy(buf, buf[1000:], 1024)
    ctx.copy(buf, buf, 1024)
    ctx.copy(buf[2048:], buf, 1024)  # disjoint halves are fine


@pytest.mark.gpu
def test_dram_element_count_follows_upstream_per_dims(tmp_path):
    """--size=1001: upstream's getBufferSize divides by sizeof(T) at a different place per dimensionality
    (/root/reference/micro/DRAM.cpp:11-22), so the three rows of a type copy different element counts.  The CSV prints
    the throughput metric (2 x buffer GiB) with six decimals, enough to tell the counts apart at this size."""
    from oracle import oracle as O

    out = tmp_path / "dram1001.csv"
    r = subprocess.run([os.path.join(BIN, "DRAM"), "--size=1001", "--num-runs=1", f"--output={out}"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    rows = O.parse_csv(str(out))
    p = 1001
    for t, size in (("fp32", 4), ("fp64", 8)):
        want = (p * p * p // size, (p * p // size) * p, (p // size) * p * p)
        assert len(set(want)) == 3
        for d in (1, 2, 3):
            row = rows[f"MicroBench_DRAM_{t}_{d}"]
            assert row["Verification"] == "PASS"
            assert abs(float(row["throughput-metric"]) - 2 * want[d - 1] * size / 2**30) < 6e-7, (t, d, row["throughput-metric"])
