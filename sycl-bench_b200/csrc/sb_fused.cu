// sb_dot_<T>, sb_linreg_<T>: two-input streaming reductions in ONE launch.
//
// Replaces, in /root/reference:
//   single-kernel/scalar_prod.cpp:54-175   product kernel (writes N products) + per pass {work-group tree
//                                          kernel, gather kernel} over the product buffer, in place
//   single-kernel/lin_reg_coeff.cpp:52-146 two product kernels + four in-place multi-pass reductions
//                                          (sum xy, sum xx, sum x, sum y), coefficients on the host
// The reference moves 3*N*s for every product and ~N*s(1 + 1/L + ...) for every reduction.  Fused, both
// inputs are read exactly once and nothing but the results is written: algorithmic bytes 2*N*s (+ the
// results).  HBM-bound; same structure as sb_reduce (sb_reduce.cu): persistent grid of sm_count * 4 CTAs,
// 32 KB of loads per chunk (16 KB from each input), chunks handed out in order by a double-buffered atomic
// ticket (static assignment under sb_ctx_set_deterministic), all loads of a chunk issued before the first
// use, one partial per sum and CTA, the last CTA to finish folds the partials in index order.
//
// Arithmetic: products are rounded to T before they are added (__fmul_rn / __dmul_rn, no FMA contraction),
// like the reference, which stores every product to memory first.  Integers wrap.
#include "sb_common.cuh"

namespace {

constexpr int FU_NT = 512;
constexpr int FU_UNROLL = 2;  // 16-byte vectors per input, thread and chunk
constexpr int FU_CTAS_PER_SM = 4;

template <typename A> __device__ __forceinline__ A mul_rn(A a, A b) { return a * b; }
template <> __device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <typename A> __device__ __forceinline__ A sub_rn(A a, A b) { return a - b; }
template <> __device__ __forceinline__ float sub_rn<float>(float a, float b) { return __fsub_rn(a, b); }
template <> __device__ __forceinline__ double sub_rn<double>(double a, double b) { return __dsub_rn(a, b); }

struct OpDot {
  static constexpr int NSUM = 1;
  template <typename A>
  static __device__ __forceinline__ void apply(A (&acc)[NSUM], A x, A y) {
    acc[0] += mul_rn<A>(x, y);
  }
};
// sums in the order lin_reg_coeff.cpp:128-136 produces them: xy, xx, x, y
struct OpLinreg {
  static constexpr int NSUM = 4;
  template <typename A>
  static __device__ __forceinline__ void apply(A (&acc)[NSUM], A x, A y) {
    acc[0] += mul_rn<A>(x, y);
    acc[1] += mul_rn<A>(x, x);
    acc[2] += x;
    acc[3] += y;
  }
};

// VECTOR: both inputs 16-byte aligned, 128-bit loads; otherwise one element per load
template <typename A, typename Op, int NT, int UNROLL, bool TICKETED, bool VECTOR>
__global__ void __launch_bounds__(NT) fused2_kernel(const A* __restrict__ x, const A* __restrict__ y, size_t n, A* __restrict__ partials,
    unsigned int* __restrict__ tickets, A* __restrict__ out) {
  constexpr int VEC = VECTOR ? 16 / sizeof(A) : 1;
  constexpr int NSUM = Op::NSUM;
  __shared__ A smem[NSUM][NT / 32];
  __shared__ bool is_last;
  __shared__ unsigned int s_chunk[2];

  const size_t nvec = n / VEC;
  A acc[NSUM];
#pragma unroll
  for(int k = 0; k < NSUM; ++k) acc[k] = A(0);

  constexpr size_t CHUNK = (size_t)NT * UNROLL;  // vectors per input and chunk
  const size_t nchunks = (nvec + CHUNK - 1) / CHUNK;

  auto do_chunk = [&](size_t chunk) {
    const size_t i0 = chunk * CHUNK + threadIdx.x;
    if constexpr(VECTOR) {
      if((chunk + 1) * CHUNK <= nvec) {
        sb::Vec<A> tx[UNROLL], ty[UNROLL];
#pragma unroll
        for(int u = 0; u < UNROLL; ++u) tx[u] = sb::ldg_nc_ordered<A>(x + (i0 + (size_t)u * NT) * VEC);
#pragma unroll
        for(int u = 0; u < UNROLL; ++u) ty[u] = sb::ldg_nc_ordered<A>(y + (i0 + (size_t)u * NT) * VEC);
#pragma unroll
        for(int u = 0; u < UNROLL; ++u)
#pragma unroll
          for(int c = 0; c < VEC; ++c) Op::apply(acc, tx[u].v[c], ty[u].v[c]);
      } else {
#pragma unroll
        for(int u = 0; u < UNROLL; ++u) {
          const size_t i = i0 + (size_t)u * NT;
          if(i < nvec) {
            const sb::Vec<A> tx = sb::ldg_nc<A>(x + i * VEC), ty = sb::ldg_nc<A>(y + i * VEC);
#pragma unroll
            for(int c = 0; c < VEC; ++c) Op::apply(acc, tx.v[c], ty.v[c]);
          }
        }
      }
    } else {
      A tx[UNROLL], ty[UNROLL];
#pragma unroll
      for(int u = 0; u < UNROLL; ++u) {
        const size_t i = i0 + (size_t)u * NT;
        tx[u] = i < nvec ? x[i] : A(0);
        ty[u] = i < nvec ? y[i] : A(0);
      }
#pragma unroll
      for(int u = 0; u < UNROLL; ++u) Op::apply(acc, tx[u], ty[u]);  // padding adds x = y = 0: no effect on any sum
    }
  };

  if(TICKETED) {
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

  // scalar tail (fewer than VEC elements), block 0 only
  if(blockIdx.x == 0) {
    const size_t t = nvec * VEC + threadIdx.x;
    if(t < n) Op::apply(acc, x[t], y[t]);
  }

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for(int k = 0; k < NSUM; ++k) {
    const A w = sb::warp_sum<A>(acc[k]);
    if(lane == 0) smem[k][warp] = w;
  }
  __syncthreads();
  if(warp == 0) {
#pragma unroll
    for(int k = 0; k < NSUM; ++k) {
      A r = (lane < NT / 32) ? smem[k][lane] : A(0);
      r = sb::warp_sum<A>(r);
      if(lane == 0) partials[(size_t)blockIdx.x * NSUM + k] = r;
    }
  }
  if(threadIdx.x == 0) {
    __threadfence();
    const unsigned int done = atomicAdd(tickets, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if(!is_last) return;

  __threadfence();
  A p[NSUM];
#pragma unroll
  for(int k = 0; k < NSUM; ++k) p[k] = A(0);
  for(unsigned int i = threadIdx.x; i < gridDim.x; i += NT)
#pragma unroll
    for(int k = 0; k < NSUM; ++k) p[k] += __ldcg(partials + (size_t)i * NSUM + k);
  __syncthreads();  // smem reuse
#pragma unroll
  for(int k = 0; k < NSUM; ++k) {
    const A w = sb::warp_sum<A>(p[k]);
    if(lane == 0) smem[k][warp] = w;
  }
  __syncthreads();
  if(warp == 0) {
    A total[NSUM];
#pragma unroll
    for(int k = 0; k < NSUM; ++k) {
      A r = (lane < NT / 32) ? smem[k][lane] : A(0);
      total[k] = sb::warp_sum<A>(r);
    }
    if(lane == 0) {
#pragma unroll
      for(int k = 0; k < NSUM; ++k) out[k] = total[k];
      if constexpr(NSUM == 4) {
        // lin_reg_coeff.cpp:134-146, in T, one rounding per operation
        const A cnt = (A)n;
        const A mean_x = total[2] / cnt, mean_y = total[3] / cnt;
        const A ss_xy = sub_rn<A>(total[0], mul_rn<A>(mean_x, mean_y));
        const A ss_xx = sub_rn<A>(total[1], mul_rn<A>(mean_x, mean_x));
        const A b1 = ss_xy / ss_xx;
        out[4] = b1;
        out[5] = sub_rn<A>(mean_y, mul_rn<A>(b1, mean_x));
      }
      tickets[0] = 0u;
      tickets[1] = 0u;
    }
  }
}

template <typename T, typename Op>
int fused2_impl(sb_ctx* ctx, const T* x, const T* y, size_t n, T* out) {
  using A = typename sb_arith<T>::type;
  if(!ctx || !out || (n && (!x || !y))) return SB_ERR_BAD_ARG;
  if(!sb::aligned_to(x, sizeof(T)) || !sb::aligned_to(y, sizeof(T)) || !sb::aligned_to(out, sizeof(T))) return SB_ERR_MISALIGNED;
  sb_device_guard guard(ctx->device);
  const bool vector = sb::aligned_to(x, 16) && sb::aligned_to(y, 16);
  const size_t vec = vector ? 16 / sizeof(T) : 1;
  const size_t chunk_elems = (size_t)FU_NT * FU_UNROLL * vec;
  const size_t chunks = (n + chunk_elems - 1) / chunk_elems;
  if(chunks >= 0xfffffff0ull) return SB_ERR_BAD_ARG;
  size_t blocks = chunks;
  const size_t max_blocks = (size_t)ctx->sm_count * FU_CTAS_PER_SM;
  if(blocks > max_blocks) blocks = max_blocks;
  if(blocks < 1) blocks = 1;
  if(blocks * Op::NSUM > ctx->partials_slots) blocks = ctx->partials_slots / Op::NSUM;
  const A* ax = reinterpret_cast<const A*>(x);
  const A* ay = reinterpret_cast<const A*>(y);
  A* apart = reinterpret_cast<A*>(ctx->partials);
  A* aout = reinterpret_cast<A*>(out);
  const unsigned g = (unsigned)blocks;
  if(vector) {
    if(ctx->deterministic) fused2_kernel<A, Op, FU_NT, FU_UNROLL, false, true><<<g, FU_NT, 0, ctx->stream>>>(ax, ay, n, apart, ctx->reduce_ticket, aout);
    else fused2_kernel<A, Op, FU_NT, FU_UNROLL, true, true><<<g, FU_NT, 0, ctx->stream>>>(ax, ay, n, apart, ctx->reduce_ticket, aout);
  } else {
    if(ctx->deterministic) fused2_kernel<A, Op, FU_NT, FU_UNROLL, false, false><<<g, FU_NT, 0, ctx->stream>>>(ax, ay, n, apart, ctx->reduce_ticket, aout);
    else fused2_kernel<A, Op, FU_NT, FU_UNROLL, true, false><<<g, FU_NT, 0, ctx->stream>>>(ax, ay, n, apart, ctx->reduce_ticket, aout);
  }
  SB_CHECK_LAUNCH(ctx);
  return SB_OK;
}

}  // namespace

extern "C" {
int sb_dot_i32(sb_ctx* c, const int32_t* a, const int32_t* b, size_t n, int32_t* out) { return fused2_impl<int32_t, OpDot>(c, a, b, n, out); }
int sb_dot_i64(sb_ctx* c, const int64_t* a, const int64_t* b, size_t n, int64_t* out) { return fused2_impl<int64_t, OpDot>(c, a, b, n, out); }
int sb_dot_f32(sb_ctx* c, const float* a, const float* b, size_t n, float* out) { return fused2_impl<float, OpDot>(c, a, b, n, out); }
int sb_dot_f64(sb_ctx* c, const double* a, const double* b, size_t n, double* out) { return fused2_impl<double, OpDot>(c, a, b, n, out); }
int sb_linreg_f32(sb_ctx* c, const float* x, const float* y, size_t n, float* out6) { return fused2_impl<float, OpLinreg>(c, x, y, n, out6); }
int sb_linreg_f64(sb_ctx* c, const double* x, const double* y, size_t n, double* out6) { return fused2_impl<double, OpLinreg>(c, x, y, n, out6); }
}
