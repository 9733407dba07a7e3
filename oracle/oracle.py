"""ctypes + numpy view of oracle/_build/liboracle.so (oracle/sb_oracle.c) and of the reference
binaries under oracle/_ref.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py — never by the product path.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
REF_DIR = os.path.join(_HERE, "_ref")

NP = {"i16": np.int16, "i32": np.int32, "i64": np.int64, "f32": np.float32, "f64": np.float64}
CT = {"i16": C.c_int16, "i32": C.c_int32, "i64": C.c_int64, "f32": C.c_float, "f64": C.c_double}
# order in which the reference binaries run their types (pattern/reduction.cpp:210-224,
# pattern/segmentedreduction.cpp:168-184)
REF_REDUCTION_TYPES = ("i32", "i64", "f32", "f64")
REF_SEGREDUCTION_TYPES = ("i16", "i32", "i64", "f32", "f64")

_lib = None


def build() -> None:
    subprocess.run(["make", "-s", "-C", _HERE, "oracle"], check=True)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        vp, sz, u64, i = C.c_void_p, C.c_size_t, C.c_uint64, C.c_int
        L.ref_tree_member.argtypes = [sz, sz]
        L.oracle_num_threads.restype = i
        for t in NP:
            getattr(L, f"ref_iota_{t}").argtypes = [vp, sz, u64]
            getattr(L, f"ref_iota_{t}").restype = None
            getattr(L, f"ref_reduce_{t}").argtypes = [vp, sz, sz, vp]
            getattr(L, f"ref_segreduce_{t}").argtypes = [vp, sz, sz]
            getattr(L, f"ref_segment_sequential_sum_{t}").argtypes = [vp, sz]
            getattr(L, f"ref_segment_sequential_sum_{t}").restype = CT[t]
            getattr(L, f"ref_scan_{t}").argtypes = [vp, vp, sz, i, sz, vp]
            getattr(L, f"cpu_sum_{t}").argtypes = [vp, sz]
            getattr(L, f"cpu_sum_{t}").restype = CT[t]
            getattr(L, f"cpu_scan_{t}").argtypes = [vp, vp, sz, i]
        # sycl2020 reduction trio (oracle/sb_oracle_sycl2020.c)
        L.ref2020_percent_diff.argtypes = [C.c_double, C.c_double]
        L.ref2020_percent_diff.restype = C.c_float
        for t in REF_REDUCTION_TYPES:
            getattr(L, f"ref2020_atomic_{t}").argtypes = [u64]
            getattr(L, f"ref2020_atomic_{t}").restype = CT[t]
            getattr(L, f"ref2020_group_{t}").argtypes = [u64, u64]
            getattr(L, f"ref2020_group_{t}").restype = CT[t]
            getattr(L, f"ref2020_kernel_{t}").argtypes = [u64, u64]
            getattr(L, f"ref2020_kernel_{t}").restype = CT[t]
            getattr(L, f"ref2020_verify_atomic_{t}").argtypes = [CT[t], u64]
            getattr(L, f"ref2020_verify_size_{t}").argtypes = [CT[t], u64]
        # single-kernel/scalar_prod, lin_reg_coeff (oracle/sb_oracle_single_kernel.c)
        for t in REF_REDUCTION_TYPES:
            getattr(L, f"ref_scalar_prod_{t}").argtypes = [vp, vp, sz, sz, vp]
        for t in ("f32", "f64"):
            getattr(L, f"ref_linreg_{t}").argtypes = [vp, vp, sz, sz, vp]
            # single-kernel/lin_reg_error, kmeans
            getattr(L, f"ref_lin_reg_error_inputs_{t}").argtypes = [sz, sz, vp, vp, vp, vp]
            getattr(L, f"ref_lin_reg_error_inputs_{t}").restype = None
            getattr(L, f"ref_lin_reg_error_{t}").argtypes = [vp, vp, vp, vp, sz, vp]
            getattr(L, f"ref_lin_reg_error_{t}").restype = None
            getattr(L, f"ref_kmeans_{t}").argtypes = [vp, vp, sz, i, i, vp]
            getattr(L, f"ref_kmeans_{t}").restype = None
        L.kat_reduce.argtypes = [u64]
        L.kat_reduce.restype = u64
        L.kat_scan.argtypes = [u64, i, u64]
        L.kat_scan.restype = u64
        L.kat_segreduce_head.argtypes = [u64, u64]
        L.kat_segreduce_head.restype = u64
        _lib = L
    return _lib


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def iota(t: str, n: int, offset: int = 0) -> np.ndarray:
    out = np.empty(n, dtype=NP[t])
    getattr(lib(), f"ref_iota_{t}")(out.ctypes.data, n, offset)
    return out


def reduce(t: str, x: np.ndarray, L: int):
    """Reference multi-pass tree reduction (pattern/reduction.cpp:71-126)."""
    x = np.ascontiguousarray(x, dtype=NP[t])
    out = np.zeros(1, dtype=NP[t])
    rc = getattr(lib(), f"ref_reduce_{t}")(x.ctypes.data, x.size, L, out.ctypes.data)
    if rc < 0:
        raise RuntimeError(f"ref_reduce failed: {rc}")
    return out[0]


def segreduce(t: str, x: np.ndarray, L: int) -> np.ndarray:
    """Reference in-place segmented tree reduction (pattern/segmentedreduction.cpp:35-61); returns a copy."""
    buf = np.array(x, dtype=NP[t], copy=True)
    rc = getattr(lib(), f"ref_segreduce_{t}")(buf.ctypes.data, buf.size, L)
    if rc != 0:
        raise ValueError("ref_segreduce: bad segment length")
    return buf


def scan(t: str, x: np.ndarray, exclusive: bool = False, group_len: int = 0, flags: np.ndarray | None = None) -> np.ndarray:
    """Sequential (segmented) sum scan in a wide accumulator — DESIGN.md §Scan semantics."""
    x = np.ascontiguousarray(x, dtype=NP[t])
    out = np.empty_like(x)
    fp = None
    if flags is not None:
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        fp = flags.ctypes.data
    getattr(lib(), f"ref_scan_{t}")(x.ctypes.data, out.ctypes.data, x.size, int(exclusive), group_len, fp)
    return out


def scalar_prod(t: str, a: np.ndarray, b: np.ndarray, L: int):
    """single-kernel/scalar_prod.cpp:54-180 restated (ref_scalar_prod_<t>)."""
    out = np.zeros(1, dtype=NP[t])
    rc = getattr(lib(), f"ref_scalar_prod_{t}")(a.ctypes.data, b.ctypes.data, a.size, L, out.ctypes.data)
    if rc != 0:
        raise ValueError("ref_scalar_prod: n and L must be positive")
    return out[0]


def lin_reg_error_inputs(t: str, n: int, skip: int = 0):
    """single-kernel/lin_reg_error.cpp:30-43 restated: (in1, in2, alpha, beta) from the C rand() stream, default seed,
    after discarding `skip` draws (the fp64 benchmark of one process starts 4 * size draws in)."""
    arrs = [np.empty(n, dtype=NP[t]) for _ in range(4)]
    getattr(lib(), f"ref_lin_reg_error_inputs_{t}")(n, skip, *[a.ctypes.data for a in arrs])
    return arrs


def lin_reg_error(t: str, in1: np.ndarray, in2: np.ndarray, alpha: np.ndarray, beta: np.ndarray) -> np.ndarray:
    """single-kernel/lin_reg_error.cpp:64-77 restated: out[g] = sum_i ((alpha[g] * in1[i] + beta[g]) - in2[i])^2 in T."""
    n = in1.size
    assert in2.size == n and alpha.size == n and beta.size == n
    out = np.empty(n, dtype=NP[t])
    getattr(lib(), f"ref_lin_reg_error_{t}")(in1.ctypes.data, in2.ctypes.data, alpha.ctypes.data, beta.ctypes.data, n, out.ctypes.data)
    return out


def kmeans(t: str, features: np.ndarray, clusters: np.ndarray, n: int, nfeatures: int = 2, nclusters: int = 3) -> np.ndarray:
    """single-kernel/kmeans.cpp:61-79 restated: membership[g] = g if any cluster distance is below FLT_MAX, else 0."""
    assert features.size >= nfeatures * n and clusters.size >= nclusters * nfeatures
    out = np.empty(n, dtype=np.int32)
    getattr(lib(), f"ref_kmeans_{t}")(features.ctypes.data, clusters.ctypes.data, n, nfeatures, nclusters, out.ctypes.data)
    return out


def linreg(t: str, x: np.ndarray, y: np.ndarray, L: int) -> np.ndarray:
    """single-kernel/lin_reg_coeff.cpp:126-146 restated: [sum xy, sum xx, sum x, sum y, b1, b0]."""
    out = np.zeros(6, dtype=NP[t])
    rc = getattr(lib(), f"ref_linreg_{t}")(x.ctypes.data, y.ctypes.data, x.size, L, out.ctypes.data)
    if rc != 0:
        raise ValueError("ref_linreg: n >= 1 and L >= 2")
    return out


def cpu_sum(t: str, x: np.ndarray):
    return getattr(lib(), f"cpu_sum_{t}")(x.ctypes.data, x.size)


def cpu_scan(t: str, x: np.ndarray, out: np.ndarray, exclusive: bool = False) -> np.ndarray:
    getattr(lib(), f"cpu_scan_{t}")(x.ctypes.data, out.ctypes.data, x.size, int(exclusive))
    return out


def tree_member(j: int, L: int) -> bool:
    return bool(lib().ref_tree_member(j, L))


def kat_reduce(n: int) -> int:
    return int(lib().kat_reduce(n))


def kat_scan(g: int, exclusive: bool, L: int = 0) -> int:
    return int(lib().kat_scan(g, int(exclusive), L))


def kat_segreduce_head(q: int, L: int) -> int:
    return int(lib().kat_segreduce_head(q, L))


def wrap(value: int, t: str):
    """value mod 2^w as the signed integer type t."""
    bits = {"i16": 16, "i32": 32, "i64": 64}[t]
    value &= (1 << bits) - 1
    if value >= 1 << (bits - 1):
        value -= 1 << bits
    return value


# ---- the unmodified reference, compiled against the host SYCL shim (oracle/_ref) -------------------

def ref_available(which: str = "reduction") -> bool:
    return os.access(os.path.join(REF_DIR, which), os.X_OK)


def run_reference(which: str, size: int, local: int, num_runs: int = 1, ndrange: bool = False,
                  verification: bool = True, dump: bool = True, dump_max_bytes: int | None = None,
                  threads: int | None = None, timeout: float = 600.0):
    """Run oracle/_ref/<which> like bin/run-suite does (bin/run-suite:269-279) and return
    (rows, dumps): rows = {benchmark name: {column: value}} parsed from its CSV, dumps = list of raw
    byte strings, one per host access (verify() call), in execution order."""
    exe = os.path.join(REF_DIR, which)
    with tempfile.TemporaryDirectory() as tmp:
        csv = os.path.join(tmp, "out.csv")
        env = dict(os.environ)
        if dump:
            env["SYCL_SHIM_DUMP"] = tmp
            if dump_max_bytes is not None:
                env["SYCL_SHIM_DUMP_MAX_BYTES"] = str(dump_max_bytes)
        if threads is not None:
            env["OMP_NUM_THREADS"] = str(threads)
        cmd = [exe, f"--size={size}", f"--local={local}", "--device=cpu", f"--num-runs={num_runs}", f"--output={csv}"]
        if not ndrange:
            cmd.append("--no-ndrange-kernels")
        if not verification:
            cmd.append("--no-verification")
        subprocess.run(cmd, check=True, env=env, timeout=timeout, stdout=subprocess.DEVNULL)
        rows = parse_csv(csv)
        dumps = [open(p, "rb").read() for p in sorted(glob.glob(os.path.join(tmp, "*.bin")))]
    return rows, dumps


def parse_csv(path: str) -> dict:
    """Parse the AppendingCsvResultConsumer format (include/result_consumer.h:67-92): a
    '# Benchmark name,<cols>' header before every row; the samples column is quoted."""
    rows, header = {}, None
    for line in open(path):
        line = line.rstrip("\n")
        if not line:
            continue
        cells, cur, quoted = [], "", False
        for ch in line:
            if ch == '"':
                quoted = not quoted
                cur += ch
            elif ch == "," and not quoted:
                cells.append(cur)
                cur = ""
            else:
                cur += ch
        cells.append(cur)
        if line.startswith("# Benchmark name"):
            header = cells[1:]
        elif header is not None:
            rows[cells[0]] = dict(zip(header, cells[1:]))
    return rows
