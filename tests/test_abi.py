"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol
include/sb_cuda.h declares, and fails loudly (no CPU fallback) without a device."""
import ctypes as C
import os

import pytest


def test_library_is_in_tree(pkg):
    assert os.path.exists(pkg.lib_path()), "build with `python -c 'import __graft_entry__ as g; g.build()'`"
    assert pkg.lib_path().startswith(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def test_exports_every_declared_symbol(pkg):
    L = pkg.lib()
    names = pkg.declared_symbols()
    assert len(names) >= 55
    for family in ("sb_reduce_", "sb_scan_", "sb_segreduce_", "sb_segscan_"):
        for t in ("i32", "i64", "f32", "f64"):
            assert family + t in names
    assert "sb_segreduce_i16" in names  # the reference runs short too (pattern/segmentedreduction.cpp:169)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_version_and_strerror(pkg):
    L = pkg.lib()
    assert L.sb_version() == 100
    assert L.sb_strerror(0) == b"ok"
    assert b"no CPU fallback" in L.sb_strerror(-5)
    assert L.sb_strerror(-4).startswith(b"sb: invalid segment length")


def test_no_cpu_fallback(pkg):
    """Without a CUDA device the product path must fail loudly, never compute on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    with pytest.raises(pkg.SbError) as e:
        pkg.Context(0)
    assert e.value.status == -5
    n = C.c_int(-1)
    assert pkg.lib().sb_device_count(C.byref(n)) == -5 and n.value == 0


def test_workspace_bytes(pkg):
    L = pkg.lib()
    assert L.sb_workspace_bytes(0, 1, 1 << 28) > 0
    # scan: an upper bound — the smallest tile any kernel uses (16 KB), one descriptor per tile + one per round
    # (8 B for 4-byte, 16 B for 8-byte types)
    assert L.sb_workspace_bytes(1, 1, 1 << 30) == (2 * ((1 << 30) // 4096) + 64) * 8 + 256
    assert L.sb_workspace_bytes(1, 2, 1 << 28) == (2 * ((1 << 28) // 2048) + 64) * 16 + 256
    assert L.sb_workspace_bytes(2, 4, 1 << 29) == 0


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under sycl-bench_b200/ or include/ may mention it."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bad = []
    for base in ("sycl-bench_b200", "include"):
        for dp, dn, fn in os.walk(os.path.join(root, base)):
            if any(part in dp for part in ("build", "lib", "bin", "__pycache__")):
                continue
            for f in fn:
                if f.endswith((".cu", ".cuh", ".h", ".cpp", ".py")):
                    text = open(os.path.join(dp, f), errors="ignore").read()
                    for line in text.splitlines():
                        s = line.strip()
                        if ("liboracle" in s or "import oracle" in s or "from oracle" in s or "#include \"sb_oracle" in s):
                            bad.append((f, s))
    assert not bad, bad
