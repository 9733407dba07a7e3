// sb_reduce_<T>: global sum in ONE launch.
//
// Replaces Reduction<T>::submit + local_reduce_{ndrange,hierarchical}
// (/root/reference/pattern/reduction.cpp:71-97, :99-126, :128-167), which launches ceil(log_L N)
// kernels of one element per work-item with a log2(L)-barrier shared-memory tree.
//
// B200 design: HBM-bound read-only stream (algorithmic bytes N*s + s).
//  - persistent grid of sm_count * 4 CTAs of 512 threads; the 32 KB chunks are handed out IN ORDER by
//    an atomic ticket (the next ticket is drawn while the current chunk is in flight).  In-order
//    hand-out keeps the DRAM access front compact: the same copy kernel measures 5.97 TB/s with a
//    static grid-stride assignment and 6.95 TB/s with ticketed chunks (profiles/r01_stream_probe.txt).
//    A flat grid (one chunk per CTA) orders the accesses as well but pays the block epilogue per
//    chunk, which cost fp64 20 % (profiles/r01_ops_v6_flat_grids.jsonl).
//  - all UNROLL 128-bit loads of a thread are issued before the first add
//    (ld.global.nc.L1::no_allocate.v4: streamed once, keeps L1/L2 clean); a resident SM keeps
//    2048 threads * 64 B = 128 KB in flight (B200 needs ~43 KB/SM to cover HBM latency).
//  - warp stage: redux.sync.add for 32-bit integers, xor-shuffle tree for 64-bit / floating point.
//  - block stage through 64 B of shared memory, one partial per CTA to the ctx scratch, and a
//    last-block-done ticket sums the partials in index order.  Chunks are ticketed, so WHICH chunks a
//    CTA sums depends on timing: integer results are exact regardless; floating-point results are
//    within the stated tolerance but not bit-reproducible run to run.  sb_ctx_set_deterministic(ctx, 1)
//    selects the static chunk assignment instead: fixed summation order, bit-reproducible, ~3 % slower
//    (i32 6.76 vs 6.97, f64 7.11 vs 7.25 TB/s).
//  - sb_reduce_all_<T> (several GPUs): the last block does not stop at this GPU's sum.  It stores the sum into the
//    mailbox of every peer (st.relaxed.sys over NVLink, sb_peer.cuh), waits for the other ranks' sums in its own
//    mailbox and adds them in rank order: the all-reduce is the tail of the same kernel, no NCCL call, ~3 us
//    instead of a separate collective launch.
//  - the loads are `volatile` asm so that the batch is issued back to back; with plain asm nvcc sank
//    part of the batch next to its first use for int32 and fp64 (5.85 / 5.6 TB/s).
#include <cstdlib>
#include <cstring>

#include "sb_peer.cuh"

namespace {

constexpr int RED_NT = 512;
constexpr int RED_UNROLL = 4;
constexpr int RED_CTAS_PER_SM = 4;

template <typename A, int NT>
__device__ __forceinline__ A block_sum(A v, A* smem) {
  v = sb::warp_sum<A>(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if(lane == 0) smem[warp] = v;
  __syncthreads();
  A r = A(0);
  if(warp == 0) {
    r = (lane < NT / 32) ? smem[lane] : A(0);
    r = sb::warp_sum<A>(r);
  }
  return r;  // valid in warp 0
}

// tickets[0] = CTAs done, tickets[1] = next chunk
template <typename A, int NT, int UNROLL, bool TICKETED>
__global__ void __launch_bounds__(NT) reduce_kernel(const A* __restrict__ in, size_t n, A* __restrict__ partials,
    unsigned int* __restrict__ tickets, A* __restrict__ out, const __grid_constant__ ReducePeers rp) {
  constexpr int VEC = 16 / sizeof(A);
  __shared__ A smem[NT / 32];
  __shared__ bool is_last;
  __shared__ unsigned int s_chunk[2];

  // peel elements in front of the first 16-byte boundary
  size_t head = ((16 - (reinterpret_cast<uintptr_t>(in) & 15)) & 15) / sizeof(A);
  if(head > n) head = n;
  const A* body = in + head;
  const size_t nvec = (n - head) / VEC;
  const size_t tail_begin = head + nvec * VEC;

  A acc[UNROLL];
#pragma unroll
  for(int u = 0; u < UNROLL; ++u) acc[u] = A(0);

  constexpr size_t CHUNK = (size_t)NT * UNROLL;  // vectors per chunk
  const size_t nchunks = (nvec + CHUNK - 1) / CHUNK;

  auto do_chunk = [&](size_t chunk) {
    const size_t i0 = chunk * CHUNK + threadIdx.x;
    if((chunk + 1) * CHUNK <= nvec) {
      sb::Vec<A> t[UNROLL];
#pragma unroll
      for(int u = 0; u < UNROLL; ++u) t[u] = sb::ldg_nc_ordered<A>(body + (i0 + (size_t)u * NT) * VEC);
#pragma unroll
      for(int u = 0; u < UNROLL; ++u) {
        A s = t[u].v[0];
#pragma unroll
        for(int c = 1; c < VEC; ++c) s += t[u].v[c];
        acc[u] += s;
      }
    } else {
#pragma unroll
      for(int u = 0; u < UNROLL; ++u) {
        const size_t i = i0 + (size_t)u * NT;
        if(i < nvec) {
          sb::Vec<A> t = sb::ldg_nc<A>(body + i * VEC);
          A s = t.v[0];
#pragma unroll
          for(int c = 1; c < VEC; ++c) s += t.v[c];
          acc[u] += s;
        }
      }
    }
  };

  if(TICKETED) {
    // double-buffered ticket: the draw for the next chunk overlaps the loads of the current one
    if(threadIdx.x == 0) s_chunk[0] = atomicAdd(tickets + 1, 1u);
    __syncthreads();
    for(int b = 0;; b ^= 1) {
      const size_t chunk = s_chunk[b];
      if(chunk >= nchunks) break;
      if(threadIdx.x == 0) s_chunk[b ^ 1] = atomicAdd(tickets + 1, 1u);
      do_chunk(chunk);
      __syncthreads();
    }
  } else {
    for(size_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) do_chunk(chunk);
  }
  A v = acc[0];
#pragma unroll
  for(int u = 1; u < UNROLL; ++u) v += acc[u];

  // scalar head/tail (fewer than 2*VEC elements), block 0 only
  if(blockIdx.x == 0) {
    if(threadIdx.x < head) v += in[threadIdx.x];
    const size_t t = tail_begin + threadIdx.x;
    if(t < n) v += in[t];
  }

  A bs = block_sum<A, NT>(v, smem);
  if(threadIdx.x == 0) {
    partials[blockIdx.x] = bs;
    __threadfence();
    const unsigned int done = atomicAdd(tickets, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if(is_last) {
    __threadfence();
    A p = A(0);
    for(unsigned int i = threadIdx.x; i < gridDim.x; i += NT) p += __ldcg(partials + i);
    __syncthreads();  // smem reuse
    A total = block_sum<A, NT>(p, smem);
    if(rp.world > 1 && threadIdx.x < 32) total = mailbox_allreduce<A>(total, rp);  // warp 0 holds the block sum
    if(threadIdx.x == 0) {
      out[0] = total;
      tickets[0] = 0u;  // every CTA has drawn its last ticket and signed off: re-arm for the next launch
      tickets[1] = 0u;
    }
  }
}

template <typename T>
int reduce_impl(sb_ctx* ctx, sb_peer* peer, const T* in, size_t n, T* out) {
  using A = typename sb_arith<T>::type;
  if(!ctx || !out || (n && !in)) return SB_ERR_BAD_ARG;
  if(peer && peer->ctx != ctx) return SB_ERR_BAD_ARG;
  if(!sb::aligned_to(in, sizeof(T)) || !sb::aligned_to(out, sizeof(T))) return SB_ERR_MISALIGNED;
  sb_device_guard guard(ctx->device);
  constexpr int VEC = 16 / sizeof(T);
  const size_t chunk_elems = (size_t)RED_NT * RED_UNROLL * VEC;
  const size_t chunks = (n + chunk_elems - 1) / chunk_elems;
  if(chunks >= 0xfffffff0ull) return SB_ERR_BAD_ARG;  // 32-bit chunk tickets: n below 2^47 bytes
  size_t blocks = chunks;
  const size_t max_blocks = (size_t)ctx->sm_count * RED_CTAS_PER_SM;
  if(blocks > max_blocks) blocks = max_blocks;
  if(blocks < 1) blocks = 1;
  if(blocks > ctx->partials_slots) blocks = ctx->partials_slots;
  static const bool env_deterministic = getenv("SB_REDUCE_DETERMINISTIC") && getenv("SB_REDUCE_DETERMINISTIC")[0] == '1';
  const bool deterministic = env_deterministic || ctx->deterministic;
  const A* ain = reinterpret_cast<const A*>(in);
  A* apart = reinterpret_cast<A*>(ctx->partials);
  A* aout = reinterpret_cast<A*>(out);
  ReducePeers rp;
  memset(&rp, 0, sizeof(rp));
  rp.world = 1;
  if(peer && peer->world > 1) {
    for(int d = 0; d < peer->world; ++d)
      if(!peer->connected[d]) return SB_ERR_PEER;
    // two planes alternate: a rank can enter all-reduce k+2 only after every rank has written k+1, i.e. has finished
    // (and read everything of) all-reduce k, the last user of the plane
    const uint32_t tag = ++peer->red_seq;
    const size_t plane = CY_REDUCE_OFFSET + (size_t)(tag & 1u) * CY_MAX_WORLD * 16;
    for(int d = 0; d < peer->world; ++d)
      rp.slot[d] = reinterpret_cast<unsigned long long*>(static_cast<char*>(peer->peer_base[d]) + plane) + 2 * peer->rank;
    rp.mine = reinterpret_cast<const unsigned long long*>(static_cast<char*>(peer->base) + plane);
    rp.world = peer->world;
    rp.rank = peer->rank;
    rp.tag = tag;
    rp.timeout_ns = sb_peer_timeout_ns();
    rp.err = ctx->err_dev;
  }
  if(deterministic)
    reduce_kernel<A, RED_NT, RED_UNROLL, false><<<(unsigned)blocks, RED_NT, 0, ctx->stream>>>(ain, n, apart, ctx->reduce_ticket, aout, rp);
  else
    reduce_kernel<A, RED_NT, RED_UNROLL, true><<<(unsigned)blocks, RED_NT, 0, ctx->stream>>>(ain, n, apart, ctx->reduce_ticket, aout, rp);
  SB_CHECK_LAUNCH(ctx);
  return SB_OK;
}

}  // namespace

extern "C" {
int sb_reduce_i32(sb_ctx* c, const int32_t* in, size_t n, int32_t* out) { return reduce_impl<int32_t>(c, nullptr, in, n, out); }
int sb_reduce_i64(sb_ctx* c, const int64_t* in, size_t n, int64_t* out) { return reduce_impl<int64_t>(c, nullptr, in, n, out); }
int sb_reduce_f32(sb_ctx* c, const float* in, size_t n, float* out) { return reduce_impl<float>(c, nullptr, in, n, out); }
int sb_reduce_f64(sb_ctx* c, const double* in, size_t n, double* out) { return reduce_impl<double>(c, nullptr, in, n, out); }

int sb_reduce_all_i32(sb_ctx* c, sb_peer* p, const int32_t* in, size_t n, int32_t* out) { return reduce_impl<int32_t>(c, p, in, n, out); }
int sb_reduce_all_i64(sb_ctx* c, sb_peer* p, const int64_t* in, size_t n, int64_t* out) { return reduce_impl<int64_t>(c, p, in, n, out); }
int sb_reduce_all_f32(sb_ctx* c, sb_peer* p, const float* in, size_t n, float* out) { return reduce_impl<float>(c, p, in, n, out); }
int sb_reduce_all_f64(sb_ctx* c, sb_peer* p, const double* in, size_t n, double* out) { return reduce_impl<double>(c, p, in, n, out); }
int sb_reduce_all(sb_ctx* c, sb_peer* p, int dtype, const void* in, size_t n, void* out) {
  switch(dtype) {
  case SB_I32: return sb_reduce_all_i32(c, p, (const int32_t*)in, n, (int32_t*)out);
  case SB_I64: return sb_reduce_all_i64(c, p, (const int64_t*)in, n, (int64_t*)out);
  case SB_F32: return sb_reduce_all_f32(c, p, (const float*)in, n, (float*)out);
  case SB_F64: return sb_reduce_all_f64(c, p, (const double*)in, n, (double*)out);
  default: return SB_ERR_BAD_ARG;
  }
}
}
