// sb_comm_*: the sharded (multi-GPU) reduction and prefix sum of BASELINE.json north_star, driven from
// one process: per-rank kernels of libsb_cuda + ONE tiny NCCL collective per operation.
//
//   reduction : sb_reduce per shard -> ncclAllReduce(sum) of one element per rank
//   scan      : SB_REDUCE_THEN_SCAN  sb_reduce -> ncclAllGather of the shard totals -> sb_exclusive_carry
//                                    -> sb_scan seeded with the carry                  (3 passes / shard)
//               SB_SCAN_THEN_CARRY   sb_scan -> ncclAllGather -> sb_exclusive_carry -> sb_carry_add (4 passes)
//               SB_SCAN_CYCLIC       block-cyclic shards, sb_scan_cyclic on every rank: 2 passes, the round totals cross
//                                    NVLink inside the kernel (peer-memory mailboxes, sb_scan_cyclic.cu), no NCCL call
//   foreach   : segmented reduction / segmented scan / group-scope scan on shards cut at segment boundaries:
//               independent replicas, nothing is exchanged
//
// The reference has no multi-device code; nothing here replaces a reference function.  NCCL is loaded
// with dlopen("libnccl.so.2") on first use so that libsb_cuda.so itself does not depend on it (inside a
// PyTorch process this resolves to the NCCL build torch already loaded).  bench.py exercises the same
// recipe with one process per GPU through torch.distributed (sycl-bench_b200/sharded.py).
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "sb_common.cuh"

namespace {

struct NcclApi {
  void* handle = nullptr;
  decltype(&ncclCommInitAll) CommInitAll = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api = [] {
    NcclApi a;
    a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if(!a.handle) a.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if(!a.handle) return a;
    a.CommInitAll = reinterpret_cast<decltype(a.CommInitAll)>(dlsym(a.handle, "ncclCommInitAll"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(a.handle, "ncclCommDestroy"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(a.handle, "ncclAllReduce"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(a.handle, "ncclAllGather"));
    a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(dlsym(a.handle, "ncclGroupStart"));
    a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(dlsym(a.handle, "ncclGroupEnd"));
    a.ok = a.CommInitAll && a.CommDestroy && a.AllReduce && a.AllGather && a.GroupStart && a.GroupEnd;
    return a;
  }();
  return api;
}

size_t dtype_size(int dtype) {
  switch(dtype) {
  case SB_I32: case SB_F32: return 4;
  case SB_I64: case SB_F64: return 8;
  default: return 0;
  }
}

ncclDataType_t nccl_type(int dtype) {
  switch(dtype) {
  case SB_I32: return ncclInt32;
  case SB_I64: return ncclInt64;
  case SB_F32: return ncclFloat32;
  default: return ncclFloat64;
  }
}

int reduce_call(sb_ctx* c, int dtype, const void* in, size_t n, void* out) {
  switch(dtype) {
  case SB_I32: return sb_reduce_i32(c, (const int32_t*)in, n, (int32_t*)out);
  case SB_I64: return sb_reduce_i64(c, (const int64_t*)in, n, (int64_t*)out);
  case SB_F32: return sb_reduce_f32(c, (const float*)in, n, (float*)out);
  default: return sb_reduce_f64(c, (const double*)in, n, (double*)out);
  }
}
int scan_call(sb_ctx* c, int dtype, const void* in, void* out, size_t n, int exclusive, const void* seed) {
  switch(dtype) {
  case SB_I32: return sb_scan_i32(c, (const int32_t*)in, (int32_t*)out, n, exclusive, 0, (const int32_t*)seed);
  case SB_I64: return sb_scan_i64(c, (const int64_t*)in, (int64_t*)out, n, exclusive, 0, (const int64_t*)seed);
  case SB_F32: return sb_scan_f32(c, (const float*)in, (float*)out, n, exclusive, 0, (const float*)seed);
  default: return sb_scan_f64(c, (const double*)in, (double*)out, n, exclusive, 0, (const double*)seed);
  }
}
int excl_carry_call(sb_ctx* c, int dtype, const void* totals, int rank, void* carry) {
  switch(dtype) {
  case SB_I32: return sb_exclusive_carry_i32(c, (const int32_t*)totals, rank, (int32_t*)carry);
  case SB_I64: return sb_exclusive_carry_i64(c, (const int64_t*)totals, rank, (int64_t*)carry);
  case SB_F32: return sb_exclusive_carry_f32(c, (const float*)totals, rank, (float*)carry);
  default: return sb_exclusive_carry_f64(c, (const double*)totals, rank, (double*)carry);
  }
}
int carry_add_call(sb_ctx* c, int dtype, void* buf, size_t n, const void* carry) {
  switch(dtype) {
  case SB_I32: return sb_carry_add_i32(c, (int32_t*)buf, n, (const int32_t*)carry);
  case SB_I64: return sb_carry_add_i64(c, (int64_t*)buf, n, (const int64_t*)carry);
  case SB_F32: return sb_carry_add_f32(c, (float*)buf, n, (const float*)carry);
  default: return sb_carry_add_f64(c, (double*)buf, n, (const double*)carry);
  }
}

}  // namespace

struct sb_comm {
  int n = 0;
  std::vector<sb_ctx*> ctx;       // ctx[0] is borrowed, the others are owned
  std::vector<ncclComm_t> comms;
  std::vector<void*> scratch;     // per rank: [0, 8) carry, [64, 64 + 8 n) gathered totals
  std::vector<cudaEvent_t> done;  // per rank: end of its part of an operation
  cudaEvent_t start = nullptr;    // on rank 0: everything enqueued before the operation
  std::vector<sb_peer*> peers;    // per rank: mailbox of the block-cyclic scan (created on first use)
};

namespace {
// Ranks 1.. start after whatever rank 0's stream already holds, so that events recorded on rank 0's stream
// (sb_timer_begin / sb_timer_end) bracket the whole sharded operation.
int comm_fork(sb_comm* c) {
  if(c->n == 1) return SB_OK;
  {
    sb_device_guard guard(c->ctx[0]->device);
    SB_CUDA_TRY(cudaEventRecord(c->start, c->ctx[0]->stream));
  }
  for(int r = 1; r < c->n; ++r) {
    sb_device_guard guard(c->ctx[r]->device);
    SB_CUDA_TRY(cudaStreamWaitEvent(c->ctx[r]->stream, c->start, 0));
  }
  return SB_OK;
}
// error exit after comm_fork: ranks 1.. may already run on buffers of the caller; join them into rank 0's stream
// first, so that the caller's sb_ctx_sync(rank 0) covers everything that was enqueued
int comm_fail(sb_comm* c, int rc) {
  (void)sb_comm_join(c);
  return rc;
}

int ensure_peers(sb_comm* c) {
  if(!c->peers.empty()) return SB_OK;
  std::vector<sb_peer*> peers(c->n, nullptr);
  int rc = SB_OK;
  for(int r = 0; r < c->n && rc == SB_OK; ++r) rc = sb_peer_create(c->ctx[r], r, c->n, &peers[r]);
  for(int r = 0; r < c->n && rc == SB_OK; ++r)
    for(int o = 0; o < c->n && rc == SB_OK; ++o)
      if(o != r) rc = sb_peer_connect_local(peers[r], peers[o]);
  if(rc != SB_OK) {
    for(sb_peer* p : peers) sb_peer_destroy(p);
    return rc;
  }
  c->peers = peers;
  return SB_OK;
}
}  // namespace

extern "C" {

int sb_comm_create(sb_ctx* rank0, int ngpus, sb_comm** out) {
  if(!rank0 || !out || ngpus < 1) return SB_ERR_BAD_ARG;
  *out = nullptr;
  int ndev = 0;
  if(cudaGetDeviceCount(&ndev) != cudaSuccess || rank0->device + ngpus > ndev) return SB_ERR_NO_DEVICE;
  if(ngpus > 1 && !nccl().ok) return SB_ERR_NCCL;
  sb_comm* c = new(std::nothrow) sb_comm();
  if(!c) return SB_ERR_WORKSPACE;
  c->n = ngpus;
  c->ctx.assign(ngpus, nullptr);
  c->comms.assign(ngpus, nullptr);
  c->scratch.assign(ngpus, nullptr);
  c->done.assign(ngpus, nullptr);
  c->ctx[0] = rank0;
  int rc = SB_OK;
  for(int r = 1; r < ngpus && rc == SB_OK; ++r) rc = sb_ctx_create(rank0->device + r, SB_STREAM_OWNED, &c->ctx[r]);
  for(int r = 0; r < ngpus && rc == SB_OK; ++r) {
    sb_device_guard guard(c->ctx[r]->device);
    if(cudaMalloc(&c->scratch[r], 64 + 8 * (size_t)ngpus) != cudaSuccess) rc = SB_ERR_WORKSPACE;
    else if(cudaEventCreateWithFlags(&c->done[r], cudaEventDisableTiming) != cudaSuccess) rc = SB_ERR_WORKSPACE;
  }
  if(rc == SB_OK) {
    sb_device_guard guard(rank0->device);
    if(cudaEventCreateWithFlags(&c->start, cudaEventDisableTiming) != cudaSuccess) rc = SB_ERR_WORKSPACE;
  }
  if(rc == SB_OK && ngpus > 1) {
    std::vector<int> devs(ngpus);
    for(int r = 0; r < ngpus; ++r) devs[r] = c->ctx[r]->device;
    if(nccl().CommInitAll(c->comms.data(), ngpus, devs.data()) != ncclSuccess) rc = SB_ERR_NCCL;
  }
  if(rc != SB_OK) {
    sb_comm_destroy(c);
    return rc;
  }
  *out = c;
  return SB_OK;
}

int sb_comm_destroy(sb_comm* c) {
  if(!c) return SB_OK;
  for(int r = 0; r < c->n; ++r) {  // every rank idle before any rank's mailbox goes away
    if(!c->ctx[r]) continue;
    sb_device_guard guard(c->ctx[r]->device);
    cudaStreamSynchronize(c->ctx[r]->stream);
  }
  for(int r = 0; r < c->n; ++r) {
    if(!c->ctx[r]) continue;
    sb_device_guard guard(c->ctx[r]->device);
    if(r < (int)c->peers.size() && c->peers[r]) sb_peer_destroy(c->peers[r]);
    if(c->comms[r]) nccl().CommDestroy(c->comms[r]);
    if(c->scratch[r]) cudaFree(c->scratch[r]);
    if(c->done[r]) cudaEventDestroy(c->done[r]);
    if(r > 0) sb_ctx_destroy(c->ctx[r]);
  }
  if(c->start) cudaEventDestroy(c->start);
  (void)cudaGetLastError();
  delete c;
  return SB_OK;
}

int sb_comm_size(sb_comm* c) { return c ? c->n : 0; }
sb_ctx* sb_comm_ctx(sb_comm* c, int rank) { return (c && rank >= 0 && rank < c->n) ? c->ctx[rank] : nullptr; }

int sb_comm_join(sb_comm* c) {
  if(!c) return SB_ERR_BAD_ARG;
  for(int r = 1; r < c->n; ++r) {
    {
      sb_device_guard guard(c->ctx[r]->device);
      SB_CUDA_TRY(cudaEventRecord(c->done[r], c->ctx[r]->stream));
    }
    sb_device_guard guard(c->ctx[0]->device);
    SB_CUDA_TRY(cudaStreamWaitEvent(c->ctx[0]->stream, c->done[r], 0));
  }
  return SB_OK;
}

int sb_comm_reduce(sb_comm* c, int dtype, const void* const* in, const size_t* n, void* const* out) {
  if(!c || !in || !n || !out || dtype_size(dtype) == 0) return SB_ERR_BAD_ARG;
  // default: the all-reduce is the tail of the reduce kernel itself (sb_reduce_all: peer-memory mailboxes over NVLink);
  // SB_COMM_REDUCE=nccl, or a box without peer access, takes the literal recipe: sb_reduce + ncclAllReduce of one element
  static const bool force_nccl = getenv("SB_COMM_REDUCE") && !strcmp(getenv("SB_COMM_REDUCE"), "nccl");
  const bool fused = c->n > 1 && !force_nccl && ensure_peers(c) == SB_OK;
  if(int rc = comm_fork(c); rc != SB_OK) return rc;
  for(int r = 0; r < c->n; ++r) {
    int rc = fused ? sb_reduce_all(c->ctx[r], c->peers[r], dtype, in[r], n[r], out[r]) : reduce_call(c->ctx[r], dtype, in[r], n[r], out[r]);
    if(rc != SB_OK) return comm_fail(c, rc);
  }
  if(c->n > 1 && !fused) {
    if(nccl().GroupStart() != ncclSuccess) return comm_fail(c, SB_ERR_NCCL);
    bool ok = true;
    for(int r = 0; r < c->n && ok; ++r)
      ok = nccl().AllReduce(out[r], out[r], 1, nccl_type(dtype), ncclSum, c->comms[r], c->ctx[r]->stream) == ncclSuccess;
    if(nccl().GroupEnd() != ncclSuccess || !ok) return comm_fail(c, SB_ERR_NCCL);  // the group is always closed
  }
  return sb_comm_join(c);
}

int sb_comm_scan(sb_comm* c, int dtype, const void* const* in, void* const* out, const size_t* n, int exclusive, int strategy) {
  const size_t s = dtype_size(dtype);
  if(!c || !in || !n || !out || s == 0) return SB_ERR_BAD_ARG;
  if(strategy == SB_SCAN_THEN_CARRY && exclusive) return SB_ERR_BAD_ARG;
  if(strategy == SB_SCAN_CYCLIC) {
    for(int r = 1; r < c->n; ++r)
      if(n[r] != n[0]) return SB_ERR_BAD_ARG;  // the block-cyclic layout has equal shards
    if(int rc = ensure_peers(c); rc != SB_OK) return rc;
    if(int rc = comm_fork(c); rc != SB_OK) return rc;
    for(int r = 0; r < c->n; ++r) {
      int rc = sb_scan_cyclic(c->ctx[r], c->peers[r], dtype, in[r], out[r], n[r], exclusive);
      if(rc != SB_OK) return comm_fail(c, rc);  // (ranks already launched give up after the kernel's time-out)
    }
    return sb_comm_join(c);
  }
  if(strategy != SB_REDUCE_THEN_SCAN && strategy != SB_SCAN_THEN_CARRY) return SB_ERR_BAD_ARG;
  if(c->n == 1) {
    int rc = scan_call(c->ctx[0], dtype, in[0], out[0], n[0], exclusive, nullptr);
    return rc;
  }
  if(int rc = comm_fork(c); rc != SB_OK) return rc;
  auto carry = [&](int r) { return c->scratch[r]; };
  auto totals = [&](int r) { return (void*)((char*)c->scratch[r] + 64); };
  // 1. shard totals
  for(int r = 0; r < c->n; ++r) {
    int rc;
    if(strategy == SB_REDUCE_THEN_SCAN) {
      rc = reduce_call(c->ctx[r], dtype, in[r], n[r], carry(r));
    } else {
      rc = scan_call(c->ctx[r], dtype, in[r], out[r], n[r], 0, nullptr);
      if(rc == SB_OK) {
        sb_device_guard guard(c->ctx[r]->device);
        cudaError_t e = n[r] ? cudaMemcpyAsync(carry(r), (const char*)out[r] + (n[r] - 1) * s, s, cudaMemcpyDeviceToDevice, c->ctx[r]->stream)
                             : cudaMemsetAsync(carry(r), 0, s, c->ctx[r]->stream);
        if(e != cudaSuccess) rc = (int)e;
      }
    }
    if(rc != SB_OK) return comm_fail(c, rc);
  }
  // 2. one element per rank crosses NVLink
  if(nccl().GroupStart() != ncclSuccess) return comm_fail(c, SB_ERR_NCCL);
  bool ok = true;
  for(int r = 0; r < c->n && ok; ++r)
    ok = nccl().AllGather(carry(r), totals(r), 1, nccl_type(dtype), c->comms[r], c->ctx[r]->stream) == ncclSuccess;
  if(nccl().GroupEnd() != ncclSuccess || !ok) return comm_fail(c, SB_ERR_NCCL);  // the group is always closed
  // 3. exclusive carry on the device, then the seeded scan / the carry-add pass
  for(int r = 0; r < c->n; ++r) {
    int rc = excl_carry_call(c->ctx[r], dtype, totals(r), r, carry(r));
    if(rc != SB_OK) return comm_fail(c, rc);
    if(strategy == SB_REDUCE_THEN_SCAN) rc = scan_call(c->ctx[r], dtype, in[r], out[r], n[r], exclusive, carry(r));
    else rc = carry_add_call(c->ctx[r], dtype, out[r], n[r], carry(r));
    if(rc != SB_OK) return comm_fail(c, rc);
  }
  return sb_comm_join(c);
}

int sb_comm_foreach(sb_comm* c, int op, int dtype, const void* const* in, void* const* out, const size_t* n, size_t seg_len, int exclusive) {
  if(!c || !out || !n || seg_len == 0) return SB_ERR_BAD_ARG;
  if(op != SB_OP_SEGREDUCE && op != SB_OP_SEGSCAN && op != SB_OP_SCAN) return SB_ERR_BAD_ARG;
  if(op != SB_OP_SEGREDUCE && (!in || dtype_size(dtype) == 0)) return SB_ERR_BAD_ARG;
  for(int r = 0; r < c->n; ++r)
    if(n[r] % seg_len != 0) return SB_ERR_SEG_LEN;  // shards are cut on segment boundaries
  if(int rc = comm_fork(c); rc != SB_OK) return rc;
  for(int r = 0; r < c->n; ++r) {
    sb_ctx* x = c->ctx[r];
    int rc;
    if(op == SB_OP_SEGREDUCE) {
      switch(dtype) {
      case SB_I16: rc = sb_segreduce_i16(x, (int16_t*)out[r], n[r], seg_len); break;
      case SB_I32: rc = sb_segreduce_i32(x, (int32_t*)out[r], n[r], seg_len); break;
      case SB_I64: rc = sb_segreduce_i64(x, (int64_t*)out[r], n[r], seg_len); break;
      case SB_F32: rc = sb_segreduce_f32(x, (float*)out[r], n[r], seg_len); break;
      case SB_F64: rc = sb_segreduce_f64(x, (double*)out[r], n[r], seg_len); break;
      default: rc = SB_ERR_BAD_ARG;
      }
    } else if(op == SB_OP_SEGSCAN) {
      switch(dtype) {
      case SB_I32: rc = sb_segscan_i32(x, (const int32_t*)in[r], (int32_t*)out[r], n[r], seg_len, nullptr, exclusive); break;
      case SB_I64: rc = sb_segscan_i64(x, (const int64_t*)in[r], (int64_t*)out[r], n[r], seg_len, nullptr, exclusive); break;
      case SB_F32: rc = sb_segscan_f32(x, (const float*)in[r], (float*)out[r], n[r], seg_len, nullptr, exclusive); break;
      default: rc = sb_segscan_f64(x, (const double*)in[r], (double*)out[r], n[r], seg_len, nullptr, exclusive); break;
      }
    } else {
      switch(dtype) {
      case SB_I32: rc = sb_scan_i32(x, (const int32_t*)in[r], (int32_t*)out[r], n[r], exclusive, seg_len, nullptr); break;
      case SB_I64: rc = sb_scan_i64(x, (const int64_t*)in[r], (int64_t*)out[r], n[r], exclusive, seg_len, nullptr); break;
      case SB_F32: rc = sb_scan_f32(x, (const float*)in[r], (float*)out[r], n[r], exclusive, seg_len, nullptr); break;
      default: rc = sb_scan_f64(x, (const double*)in[r], (double*)out[r], n[r], exclusive, seg_len, nullptr); break;
      }
    }
    if(rc != SB_OK) return comm_fail(c, rc);
  }
  return sb_comm_join(c);
}

}  // extern "C"
