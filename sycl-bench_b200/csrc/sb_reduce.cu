// sb_reduce_<T>: global sum in ONE launch.
//
// Replaces Reduction<T>::submit + local_reduce_{ndrange,hierarchical}
// (/root/reference/pattern/reduction.cpp:71-97, :99-126, :128-167), which launches ceil(log_L N)
// kernels of one element per work-item with a log2(L)-barrier shared-memory tree.
//
// B200 design: HBM-bound read-only stream (algorithmic bytes N*s + s).
//  - persistent grid of sm_count * CTAS_PER_SM CTAs, grid-stride over chunks of NT*UNROLL 16-byte
//    vectors; all UNROLL 128-bit loads of a thread are issued before the first add, so a resident
//    SM keeps 2048 threads * 64 B = 128 KB in flight (B200 needs ~43 KB/SM to cover HBM latency).
//  - loads are ld.global.nc.L1::no_allocate.v4 (streamed once, keep L1/L2 clean).
//  - warp stage: redux.sync.add for 32-bit integers, xor-shuffle tree for 64-bit / floating point.
//  - block stage through 64 B of shared memory, one partial per CTA to the ctx scratch, and a
//    last-block-done ticket sums the partials in index order: fixed order => deterministic fp.
#include "sb_common.cuh"

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

template <typename A, int NT, int UNROLL>
__global__ void __launch_bounds__(NT) reduce_kernel(const A* __restrict__ in, size_t n, A* __restrict__ partials,
    unsigned int* __restrict__ ticket, A* __restrict__ out) {
  constexpr int VEC = 16 / sizeof(A);
  __shared__ A smem[NT / 32];
  __shared__ bool is_last;

  // peel elements in front of the first 16-byte boundary
  size_t head = ((16 - (reinterpret_cast<uintptr_t>(in) & 15)) & 15) / sizeof(A);
  if(head > n) head = n;
  const A* body = in + head;
  const size_t nvec = (n - head) / VEC;
  const size_t tail_begin = head + nvec * VEC;

  A acc[UNROLL];
#pragma unroll
  for(int u = 0; u < UNROLL; ++u) acc[u] = A(0);

  constexpr size_t CHUNK = (size_t)NT * UNROLL;
  const size_t stride = (size_t)gridDim.x * CHUNK;
  for(size_t base = (size_t)blockIdx.x * CHUNK; base < nvec; base += stride) {
    const size_t i0 = base + threadIdx.x;
    if(base + CHUNK <= nvec) {
      sb::Vec<A> t[UNROLL];
#pragma unroll
      for(int u = 0; u < UNROLL; ++u) t[u] = sb::ldg_nc<A>(body + (i0 + (size_t)u * NT) * VEC);
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
    const unsigned int done = atomicAdd(ticket, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if(is_last) {
    __threadfence();
    A p = A(0);
    for(unsigned int i = threadIdx.x; i < gridDim.x; i += NT) p += __ldcg(partials + i);
    __syncthreads();  // smem reuse
    A total = block_sum<A, NT>(p, smem);
    if(threadIdx.x == 0) {
      out[0] = total;
      *ticket = 0u;  // ready for the next launch on this stream
    }
  }
}

template <typename T>
int reduce_impl(sb_ctx* ctx, const T* in, size_t n, T* out) {
  using A = typename sb_arith<T>::type;
  if(!ctx || !out || (n && !in)) return SB_ERR_BAD_ARG;
  if(!sb::aligned_to(in, sizeof(T)) || !sb::aligned_to(out, sizeof(T))) return SB_ERR_MISALIGNED;
  sb_device_guard guard(ctx->device);
  constexpr int VEC = 16 / sizeof(T);
  const size_t chunk_elems = (size_t)RED_NT * RED_UNROLL * VEC;
  size_t blocks = (n + chunk_elems - 1) / chunk_elems;
  const size_t max_blocks = (size_t)ctx->sm_count * RED_CTAS_PER_SM;
  if(blocks > max_blocks) blocks = max_blocks;
  if(blocks < 1) blocks = 1;
  if(blocks > ctx->partials_slots) blocks = ctx->partials_slots;
  reduce_kernel<A, RED_NT, RED_UNROLL><<<(unsigned)blocks, RED_NT, 0, ctx->stream>>>(
      reinterpret_cast<const A*>(in), n, reinterpret_cast<A*>(ctx->partials), ctx->reduce_ticket, reinterpret_cast<A*>(out));
  SB_CHECK_LAUNCH(ctx);
  return SB_OK;
}

}  // namespace

extern "C" {
int sb_reduce_i32(sb_ctx* c, const int32_t* in, size_t n, int32_t* out) { return reduce_impl<int32_t>(c, in, n, out); }
int sb_reduce_i64(sb_ctx* c, const int64_t* in, size_t n, int64_t* out) { return reduce_impl<int64_t>(c, in, n, out); }
int sb_reduce_f32(sb_ctx* c, const float* in, size_t n, float* out) { return reduce_impl<float>(c, in, n, out); }
int sb_reduce_f64(sb_ctx* c, const double* in, size_t n, double* out) { return reduce_impl<double>(c, in, n, out); }
}
