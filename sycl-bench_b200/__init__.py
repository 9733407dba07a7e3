"""sycl-bench_b200 — B200-native pattern/ reduce/scan benchmarks behind sycl-bench's plugin surface.

The product is native: ``lib/libsb_cuda.so`` (C ABI, ``include/sb_cuda.h``; hand-written CUDA for
sm_100a) plus the C++ harness and the five ``pattern/`` binaries.  This Python package is only the
ctypes view of that C ABI used by ``tests/``, ``bench.py`` and ``__graft_entry__.py``; it contains
no compute and no fallback: if the shared library is missing, importing :mod:`sbcuda` symbols
raises.

The directory name contains a hyphen (it is the reference's name), so it is imported through
``__graft_entry__.load_package()`` under the module name ``sycl_bench_b200``.
"""
from .sbcuda import (  # noqa: F401
    DTYPES,
    Comm,
    Context,
    Peer,
    SbError,
    lib,
    lib_path,
    declared_symbols,
)

__all__ = ["DTYPES", "Comm", "Context", "Peer", "SbError", "lib", "lib_path", "declared_symbols"]
