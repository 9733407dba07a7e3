"""ctypes view of libsb_cuda (include/sb_cuda.h).  Plumbing only — no compute, no fallback."""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)

DTYPES = ("i16", "i32", "i64", "f32", "f64")
_DTYPE_ENUM = {"i16": 0, "i32": 1, "i64": 2, "f32": 3, "f64": 4}
_DTYPE_SIZE = {"i16": 2, "i32": 4, "i64": 8, "f32": 4, "f64": 8}


class SbError(RuntimeError):
    def __init__(self, status: int, what: str):
        super().__init__(f"libsb_cuda error {status}: {what}")
        self.status = status


def lib_path() -> str:
    # SB_CUDA_LIB selects a tuning variant built by `make tune` (profiles/tune_scan.py); it must still be
    # an in-tree libsb_cuda build — there is no other backend.
    override = os.environ.get("SB_CUDA_LIB")
    if override:
        return os.path.abspath(override)
    return os.path.join(_HERE, "lib", "libsb_cuda.so")


def header_path() -> str:
    return os.path.join(_ROOT, "include", "sb_cuda.h")


def declared_symbols() -> list[str]:
    """Every function include/sb_cuda.h declares."""
    text = open(header_path()).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", text)))


_lib = None


def lib() -> C.CDLL:
    """Load libsb_cuda.so from the tree.  Raises loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `make -C sycl-bench_b200` or __graft_entry__.build(). "
            "There is no CPU or PyTorch fallback."
        )
    L = C.CDLL(path)
    vp, sz, i, u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint64
    L.sb_version.restype = i
    L.sb_strerror.restype = C.c_char_p
    L.sb_strerror.argtypes = [i]
    L.sb_device_count.argtypes = [C.POINTER(i)]
    L.sb_ctx_create.argtypes = [i, vp, C.POINTER(vp)]
    L.sb_ctx_destroy.argtypes = [vp]
    L.sb_ctx_set_stream.argtypes = [vp, vp]
    L.sb_ctx_stream.argtypes = [vp]
    L.sb_ctx_stream.restype = vp
    L.sb_ctx_sync.argtypes = [vp]
    L.sb_ctx_device_name.argtypes = [vp, C.c_char_p, sz]
    L.sb_ctx_sm_count.argtypes = [vp]
    L.sb_ctx_set_deterministic.argtypes = [vp, i]
    L.sb_ctx_launch_count.argtypes = [vp]
    L.sb_ctx_launch_count.restype = u64
    L.sb_workspace_bytes.argtypes = [i, i, sz]
    L.sb_workspace_bytes.restype = sz
    L.sb_timer_begin.argtypes = [vp]
    L.sb_timer_end.argtypes = [vp, C.POINTER(C.c_float)]
    L.sb_malloc.argtypes = [vp, C.POINTER(vp), sz]
    L.sb_free.argtypes = [vp, vp]
    L.sb_host_alloc.argtypes = [vp, C.POINTER(vp), sz]
    L.sb_host_free.argtypes = [vp, vp]
    L.sb_memcpy_h2d.argtypes = [vp, vp, vp, sz]
    L.sb_memcpy_d2h.argtypes = [vp, vp, vp, sz]
    L.sb_memset.argtypes = [vp, vp, i, sz]
    for t in DTYPES:
        getattr(L, f"sb_iota_{t}").argtypes = [vp, vp, sz, u64]
        getattr(L, f"sb_segreduce_{t}").argtypes = [vp, vp, sz, sz]
    for t in DTYPES[1:]:
        getattr(L, f"sb_reduce_{t}").argtypes = [vp, vp, sz, vp]
        getattr(L, f"sb_scan_{t}").argtypes = [vp, vp, vp, sz, i, sz, vp]
        getattr(L, f"sb_segscan_{t}").argtypes = [vp, vp, vp, sz, sz, vp, i]
        getattr(L, f"sb_carry_add_{t}").argtypes = [vp, vp, sz, vp]
        getattr(L, f"sb_exclusive_carry_{t}").argtypes = [vp, vp, i, vp]
    for t in ("i32", "i64"):
        getattr(L, f"sb_check_iota_{t}").argtypes = [vp, vp, sz, u64, i, sz, vp]
    for t in ("f32", "f64"):
        getattr(L, f"sb_check_iota_{t}").argtypes = [vp, vp, sz, u64, i, sz, C.c_double, vp]
    for t in DTYPES[1:]:
        getattr(L, f"sb_iota_cyclic_{t}").argtypes = [vp, vp, sz, sz, i, i]
        getattr(L, f"sb_check_iota_cyclic_{t}").argtypes = [vp, vp, sz, sz, i, i, i, C.c_double, vp]
        getattr(L, f"sb_scan_cyclic_{t}").argtypes = [vp, vp, vp, vp, sz, i, vp]
    L.sb_scan_cyclic.argtypes = [vp, vp, i, vp, vp, sz, i]
    for t in DTYPES[1:]:
        getattr(L, f"sb_reduce_all_{t}").argtypes = [vp, vp, vp, sz, vp]
    L.sb_reduce_all.argtypes = [vp, vp, i, vp, sz, vp]
    L.sb_cyclic_block_elems.argtypes = [vp, i]
    L.sb_cyclic_block_elems.restype = sz
    L.sb_peer_create.argtypes = [vp, i, i, C.POINTER(vp)]
    L.sb_peer_destroy.argtypes = [vp]
    L.sb_peer_handle.argtypes = [vp, vp, sz]
    L.sb_peer_connect.argtypes = [vp, i, vp, sz]
    L.sb_peer_connect_local.argtypes = [vp, vp]
    L.sb_peer_rank.argtypes = [vp]
    L.sb_peer_world.argtypes = [vp]
    L.sb_ctx_reserve.argtypes = [vp, i, i, sz]
    L.sb_comm_foreach.argtypes = [vp, i, i, C.POINTER(vp), C.POINTER(vp), C.POINTER(sz), sz, i]
    for t, ct in (("i32", C.c_int32), ("i64", C.c_int64), ("f32", C.c_float), ("f64", C.c_double)):
        getattr(L, f"sb_fill_{t}").argtypes = [vp, vp, sz, ct]
    L.sb_copy.argtypes = [vp, vp, vp, sz]
    for t in ("i32", "i64", "f32", "f64"):
        getattr(L, f"sb_dot_{t}").argtypes = [vp, vp, vp, sz, vp]
    for t in ("f32", "f64"):
        getattr(L, f"sb_linreg_{t}").argtypes = [vp, vp, vp, sz, vp]
        getattr(L, f"sb_linreg_error_{t}").argtypes = [vp, vp, vp, vp, vp, sz, vp]
        getattr(L, f"sb_kmeans_{t}").argtypes = [vp, vp, vp, sz, i, i, vp]
    L.sb_host_reduce.argtypes = [vp, i, vp, sz, vp]
    L.sb_host_scan.argtypes = [vp, i, vp, vp, sz, i, sz]
    L.sb_host_segreduce.argtypes = [vp, i, vp, sz, sz]
    L.sb_host_segscan.argtypes = [vp, i, vp, vp, sz, sz, vp, i]
    L.sb_comm_create.argtypes = [vp, i, C.POINTER(vp)]
    L.sb_comm_destroy.argtypes = [vp]
    L.sb_comm_size.argtypes = [vp]
    L.sb_comm_ctx.argtypes = [vp, i]
    L.sb_comm_ctx.restype = vp
    L.sb_comm_join.argtypes = [vp]
    L.sb_comm_reduce.argtypes = [vp, i, C.POINTER(vp), C.POINTER(sz), C.POINTER(vp)]
    L.sb_comm_scan.argtypes = [vp, i, C.POINTER(vp), C.POINTER(vp), C.POINTER(sz), i, i]
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc != 0:
        raise SbError(rc, lib().sb_strerror(rc).decode())


def _ptr(x) -> Optional[int]:
    if x is None:
        return None
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):  # numpy
        return x.ctypes.data
    return int(x)


def dtype_of(x) -> str:
    """Map a torch / numpy dtype to the C-ABI suffix."""
    name = str(getattr(x, "dtype", x)).replace("torch.", "")
    table = {"int16": "i16", "int32": "i32", "int64": "i64", "float32": "f32", "float64": "f64"}
    if name not in table:
        raise TypeError(f"unsupported dtype {name}")
    return table[name]


class Context:
    """One (device, stream) pair — the counterpart of the reference's BenchmarkArgs::device_queue
    (/root/reference/include/command_line.h:129-140).  Methods take torch tensors (or raw device
    addresses) and forward to the C ABI; they are asynchronous on the ctx stream."""

    OWNED_STREAM = C.c_void_p(-1).value  # SB_STREAM_OWNED

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        """stream: a cudaStream_t handle as int (0 / None = the CUDA default stream, which is what
        torch.cuda.current_stream().cuda_stream is by default); Context.OWNED_STREAM = new stream."""
        self._h = C.c_void_p()
        _check(lib().sb_ctx_create(device, stream, C.byref(self._h)))
        self.device = device

    def close(self) -> None:
        if self._h:
            lib().sb_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing
    def sync(self) -> None:
        _check(lib().sb_ctx_sync(self._h))

    def set_deterministic(self, on: bool) -> None:
        """Bit-reproducible floating-point reductions (fixed summation order, ~3 % slower)."""
        _check(lib().sb_ctx_set_deterministic(self._h, int(on)))

    def set_stream(self, stream: Optional[int]) -> None:
        _check(lib().sb_ctx_set_stream(self._h, stream))

    @property
    def stream(self) -> int:
        return lib().sb_ctx_stream(self._h) or 0

    @property
    def launches(self) -> int:
        return int(lib().sb_ctx_launch_count(self._h))

    @property
    def sm_count(self) -> int:
        return int(lib().sb_ctx_sm_count(self._h))

    def device_name(self) -> str:
        buf = C.create_string_buffer(256)
        _check(lib().sb_ctx_device_name(self._h, buf, 256))
        return buf.value.decode()

    def timer_begin(self) -> None:
        _check(lib().sb_timer_begin(self._h))

    def timer_end(self) -> float:
        ms = C.c_float()
        _check(lib().sb_timer_end(self._h, C.byref(ms)))
        return float(ms.value)

    # -- ops on device buffers
    def iota(self, out, offset: int = 0, n: Optional[int] = None, dtype: Optional[str] = None):
        t = dtype or dtype_of(out)
        n = out.numel() if n is None else n
        _check(getattr(lib(), f"sb_iota_{t}")(self._h, _ptr(out), n, offset))
        return out

    def fill(self, out, value, n: Optional[int] = None, dtype: Optional[str] = None):
        t = dtype or dtype_of(out)
        n = out.numel() if n is None else n
        _check(getattr(lib(), f"sb_fill_{t}")(self._h, _ptr(out), n, value))
        return out

    def reduce(self, x, out, n: Optional[int] = None, dtype: Optional[str] = None):
        t = dtype or dtype_of(x)
        n = x.numel() if n is None else n
        _check(getattr(lib(), f"sb_reduce_{t}")(self._h, _ptr(x), n, _ptr(out)))
        return out

    def reduce_all(self, x, out, peer: "Optional[Peer]" = None, n: Optional[int] = None, dtype: Optional[str] = None):
        """out[0] = sum over the inputs of ALL ranks of the peer group (all-reduce fused into the reduce kernel)."""
        t = dtype or dtype_of(x)
        n = x.numel() if n is None else n
        _check(getattr(lib(), f"sb_reduce_all_{t}")(self._h, peer._h if peer is not None else None, _ptr(x), n, _ptr(out)))
        return out

    def copy(self, dst, src, nbytes: Optional[int] = None):
        """dst = src, device to device (the micro/DRAM copy kernel)."""
        if nbytes is None:
            nbytes = src.numel() * src.element_size()
        _check(lib().sb_copy(self._h, _ptr(dst), _ptr(src), nbytes))
        return dst

    def dot(self, a, b, out, n: Optional[int] = None, dtype: Optional[str] = None):
        """out[0] = sum a[i] * b[i] (one launch, both inputs read once)."""
        t = dtype or dtype_of(a)
        n = a.numel() if n is None else n
        _check(getattr(lib(), f"sb_dot_{t}")(self._h, _ptr(a), _ptr(b), n, _ptr(out)))
        return out

    def linreg(self, x, y, out6, n: Optional[int] = None, dtype: Optional[str] = None):
        """out6 = [sum xy, sum xx, sum x, sum y, b1, b0] (one launch)."""
        t = dtype or dtype_of(x)
        n = x.numel() if n is None else n
        _check(getattr(lib(), f"sb_linreg_{t}")(self._h, _ptr(x), _ptr(y), n, _ptr(out6)))
        return out6

    def linreg_error(self, x, y, alpha, beta, out, n: Optional[int] = None, dtype: Optional[str] = None):
        """out[g] = sum_i ((alpha[g] * x[i] + beta[g]) - y[i])^2 (single-kernel/lin_reg_error.cpp:64-77)."""
        t = dtype or dtype_of(x)
        n = x.numel() if n is None else n
        _check(getattr(lib(), f"sb_linreg_error_{t}")(self._h, _ptr(x), _ptr(y), _ptr(alpha), _ptr(beta), n, _ptr(out)))
        return out

    def kmeans(self, features, clusters, membership, n: int, nfeatures: int = 2, nclusters: int = 3, dtype: Optional[str] = None):
        """membership[g] = g if any cluster distance < FLT_MAX else 0 (single-kernel/kmeans.cpp:61-79)."""
        t = dtype or dtype_of(features)
        _check(getattr(lib(), f"sb_kmeans_{t}")(self._h, _ptr(features), _ptr(clusters), n, nfeatures, nclusters, _ptr(membership)))
        return membership

    def scan(self, x, out, exclusive: bool = False, group_len: int = 0, seed=None, n: Optional[int] = None, dtype: Optional[str] = None):
        t = dtype or dtype_of(x)
        n = x.numel() if n is None else n
        _check(getattr(lib(), f"sb_scan_{t}")(self._h, _ptr(x), _ptr(out), n, int(exclusive), group_len, _ptr(seed)))
        return out

    def segreduce(self, x, seg_len: int, n: Optional[int] = None, dtype: Optional[str] = None):
        t = dtype or dtype_of(x)
        n = x.numel() if n is None else n
        _check(getattr(lib(), f"sb_segreduce_{t}")(self._h, _ptr(x), n, seg_len))
        return x

    def segscan(self, x, out, seg_len: int = 0, head_flags=None, exclusive: bool = False, n: Optional[int] = None, dtype: Optional[str] = None):
        t = dtype or dtype_of(x)
        n = x.numel() if n is None else n
        _check(getattr(lib(), f"sb_segscan_{t}")(self._h, _ptr(x), _ptr(out), n, seg_len, _ptr(head_flags), int(exclusive)))
        return out

    def carry_add(self, x, carry, n: Optional[int] = None, dtype: Optional[str] = None):
        t = dtype or dtype_of(x)
        n = x.numel() if n is None else n
        _check(getattr(lib(), f"sb_carry_add_{t}")(self._h, _ptr(x), n, _ptr(carry)))
        return x

    def exclusive_carry(self, totals, rank: int, out, dtype: Optional[str] = None):
        t = dtype or dtype_of(totals)
        _check(getattr(lib(), f"sb_exclusive_carry_{t}")(self._h, _ptr(totals), rank, _ptr(out)))
        return out

    def check_iota(self, data, mismatches, offset: int = 0, kind: int = 0, seg_len: int = 0, n: Optional[int] = None, dtype: Optional[str] = None,
                   rel_tol: float = 0.0):
        """Count elements that differ from the closed form of the iota input (integers: exactly; floating
        point: by more than rel_tol * max(|closed form|, 1))."""
        t = dtype or dtype_of(data)
        n = data.numel() if n is None else n
        if t[0] == "f":
            _check(getattr(lib(), f"sb_check_iota_{t}")(self._h, _ptr(data), n, offset, kind, seg_len, rel_tol, _ptr(mismatches)))
        else:
            _check(getattr(lib(), f"sb_check_iota_{t}")(self._h, _ptr(data), n, offset, kind, seg_len, _ptr(mismatches)))
        return mismatches

    def reserve(self, op: int, dtype: str, n: int) -> None:
        """Pre-size the ctx scratch for `op` (0 reduce, 1 scan, 2 segreduce, 3 segscan) on n elements."""
        _check(lib().sb_ctx_reserve(self._h, op, _DTYPE_ENUM[dtype], n))

    # -- block-cyclic prefix sum (sb_scan_cyclic.cu)
    def cyclic_block_elems(self, dtype: str) -> int:
        return int(lib().sb_cyclic_block_elems(self._h, _DTYPE_ENUM[dtype]))

    def iota_cyclic(self, out, block: int, world: int, rank: int, n: Optional[int] = None, dtype: Optional[str] = None):
        t = dtype or dtype_of(out)
        n = out.numel() if n is None else n
        _check(getattr(lib(), f"sb_iota_cyclic_{t}")(self._h, _ptr(out), n, block, world, rank))
        return out

    def check_iota_cyclic(self, data, mismatches, block: int, world: int, rank: int, kind: int = 0, rel_tol: float = 0.0,
                          n: Optional[int] = None, dtype: Optional[str] = None):
        t = dtype or dtype_of(data)
        n = data.numel() if n is None else n
        _check(getattr(lib(), f"sb_check_iota_cyclic_{t}")(self._h, _ptr(data), n, block, world, rank, kind, rel_tol, _ptr(mismatches)))
        return mismatches

    def scan_cyclic(self, x, out, peer: "Optional[Peer]" = None, exclusive: bool = False, seed=None, n: Optional[int] = None, dtype: Optional[str] = None):
        """This rank's part of the global prefix sum over block-cyclic shards (peer=None: one GPU)."""
        t = dtype or dtype_of(x)
        n = x.numel() if n is None else n
        _check(getattr(lib(), f"sb_scan_cyclic_{t}")(self._h, peer._h if peer is not None else None, _ptr(x), _ptr(out), n, int(exclusive), _ptr(seed)))
        return out

    # -- host-buffer (end-to-end) calls; synchronous
    def host_reduce(self, h_in, h_out, n: Optional[int] = None, dtype: Optional[str] = None):
        t = dtype or dtype_of(h_in)
        n = _numel(h_in) if n is None else n
        _check(lib().sb_host_reduce(self._h, _DTYPE_ENUM[t], _ptr(h_in), n, _ptr(h_out)))
        return h_out

    def host_scan(self, h_in, h_out, exclusive: bool = False, group_len: int = 0, n: Optional[int] = None, dtype: Optional[str] = None):
        t = dtype or dtype_of(h_in)
        n = _numel(h_in) if n is None else n
        _check(lib().sb_host_scan(self._h, _DTYPE_ENUM[t], _ptr(h_in), _ptr(h_out), n, int(exclusive), group_len))
        return h_out

    def host_segreduce(self, h_inout, seg_len: int, n: Optional[int] = None, dtype: Optional[str] = None):
        t = dtype or dtype_of(h_inout)
        n = _numel(h_inout) if n is None else n
        _check(lib().sb_host_segreduce(self._h, _DTYPE_ENUM[t], _ptr(h_inout), n, seg_len))
        return h_inout

    def host_segscan(self, h_in, h_out, seg_len: int = 0, head_flags=None, exclusive: bool = False, n: Optional[int] = None, dtype: Optional[str] = None):
        t = dtype or dtype_of(h_in)
        n = _numel(h_in) if n is None else n
        _check(lib().sb_host_segscan(self._h, _DTYPE_ENUM[t], _ptr(h_in), _ptr(h_out), n, seg_len, _ptr(head_flags), int(exclusive)))
        return h_out


class Peer:
    """Mailbox of one rank of a block-cyclic scan group (sb_peer_* in include/sb_cuda.h).  One process per
    GPU: create, exchange `handle()` bytes with every other rank (any transport), `connect` them."""

    HANDLE_BYTES = 64

    def __init__(self, ctx: Context, rank: int, world: int):
        self._h = C.c_void_p()
        self._ctx = ctx
        _check(lib().sb_peer_create(ctx._h, rank, world, C.byref(self._h)))
        self.rank, self.world = rank, world

    def handle(self) -> bytes:
        buf = C.create_string_buffer(self.HANDLE_BYTES)
        _check(lib().sb_peer_handle(self._h, buf, self.HANDLE_BYTES))
        return buf.raw

    def connect(self, peer_rank: int, handle: bytes) -> None:
        buf = C.create_string_buffer(bytes(handle), self.HANDLE_BYTES)
        _check(lib().sb_peer_connect(self._h, peer_rank, buf, self.HANDLE_BYTES))

    def connect_local(self, other: "Peer") -> None:
        _check(lib().sb_peer_connect_local(self._h, other._h))

    def close(self) -> None:
        if self._h:
            lib().sb_peer_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _BorrowedContext(Context):
    """A Context whose handle belongs to an sb_comm."""

    def close(self) -> None:
        self._h = C.c_void_p()


class Comm:
    """Single-process multi-GPU group (sb_comm_* in include/sb_cuda.h): rank r works on device
    ctx.device + r with its own stream; one NCCL collective of one element per rank joins the shards.
    bench.py uses the one-process-per-GPU form of the same recipe (sharded.ShardedOps) instead."""

    REDUCE_THEN_SCAN = 0
    SCAN_THEN_CARRY = 1
    CYCLIC = 2
    OP_SCAN, OP_SEGREDUCE, OP_SEGSCAN = 1, 2, 3

    def __init__(self, ctx: Context, ngpus: int):
        self._h = C.c_void_p()
        self._root = ctx  # keeps rank 0 alive
        _check(lib().sb_comm_create(ctx._h, ngpus, C.byref(self._h)))
        self.size = ngpus

    def close(self) -> None:
        if self._h:
            lib().sb_comm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def rank_ctx(self, rank: int) -> Context:
        """A borrowed Context view of rank `rank` (do not close it)."""
        h = lib().sb_comm_ctx(self._h, rank)
        if not h:
            raise IndexError(rank)
        view = _BorrowedContext.__new__(_BorrowedContext)
        view._h = C.c_void_p(h)
        view.device = self._root.device + rank
        return view

    def join(self) -> None:
        _check(lib().sb_comm_join(self._h))

    @staticmethod
    def _arr(ctype, values):
        return (ctype * len(values))(*values)

    def reduce(self, shards, outs, dtype: Optional[str] = None):
        """shards[r] / outs[r] live on device r; every outs[r][0] receives the global sum."""
        t = dtype or dtype_of(shards[0])
        ins = self._arr(C.c_void_p, [_ptr(x) for x in shards])
        ns = self._arr(C.c_size_t, [_numel(x) for x in shards])
        os_ = self._arr(C.c_void_p, [_ptr(x) for x in outs])
        _check(lib().sb_comm_reduce(self._h, _DTYPE_ENUM[t], ins, ns, os_))
        return outs

    def scan(self, shards, outs, exclusive: bool = False, strategy: int = 0, dtype: Optional[str] = None):
        """Global prefix sum over the concatenation of the shards (rank order)."""
        t = dtype or dtype_of(shards[0])
        ins = self._arr(C.c_void_p, [_ptr(x) for x in shards])
        ns = self._arr(C.c_size_t, [_numel(x) for x in shards])
        os_ = self._arr(C.c_void_p, [_ptr(x) for x in outs])
        _check(lib().sb_comm_scan(self._h, _DTYPE_ENUM[t], ins, os_, ns, int(exclusive), strategy))
        return outs


def _comm_foreach(self, op: int, shards, outs, seg_len: int, exclusive: bool = False, dtype: Optional[str] = None):
    """Independent per-shard segmented reduction (outs in place, shards ignored) / segmented scan / group-scope
    scan on every rank; shards are cut on segment boundaries, nothing is exchanged."""
    t = dtype or dtype_of(outs[0])
    ins = Comm._arr(C.c_void_p, [_ptr(x) for x in (shards if shards is not None else outs)])
    ns = Comm._arr(C.c_size_t, [_numel(x) for x in outs])
    os_ = Comm._arr(C.c_void_p, [_ptr(x) for x in outs])
    _check(lib().sb_comm_foreach(self._h, op, _DTYPE_ENUM[t], ins, os_, ns, seg_len, int(exclusive)))
    return outs


Comm.foreach = _comm_foreach


def _numel(x) -> int:
    if hasattr(x, "numel"):
        return int(x.numel())
    return int(x.size)
