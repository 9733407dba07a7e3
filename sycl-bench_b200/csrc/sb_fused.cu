// sb_dot_<T>, sb_linreg_<T>: two-input streaming reductions in ONE launch.
// sb_linreg_error_<T>, sb_kmeans_<T>: the next two single-kernel/ rows (bottom of this file).
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
  template <typename A> using acc = A;
  template <typename A>
  static __device__ __forceinline__ void apply(A (&acc)[NSUM], A x, A y) {
    acc[0] += mul_rn<A>(x, y);
  }
};
// sums in the order lin_reg_coeff.cpp:128-136 produces them: xy, xx, x, y
struct OpLinreg {
  static constexpr int NSUM = 4;
  template <typename A> using acc = A;
  template <typename A>
  static __device__ __forceinline__ void apply(A (&acc)[NSUM], A x, A y) {
    acc[0] += mul_rn<A>(x, y);
    acc[1] += mul_rn<A>(x, x);
    acc[2] += x;
    acc[3] += y;
  }
};
// the five sums lin_reg_error needs (below), accumulated in double for both types: sum x, y, xx, xy, yy
struct OpLinregErr {
  static constexpr int NSUM = 5;
  template <typename A> using acc = double;
  template <typename A>
  static __device__ __forceinline__ void apply(double (&acc)[NSUM], A x, A y) {
    const double dx = (double)x, dy = (double)y;
    acc[0] += dx;
    acc[1] += dy;
    acc[2] += dx * dx;
    acc[3] += dx * dy;
    acc[4] += dy * dy;
  }
};

// VECTOR: both inputs 16-byte aligned, 128-bit loads; otherwise one element per load
template <typename A, typename Op, int NT, int UNROLL, bool TICKETED, bool VECTOR>
__global__ void __launch_bounds__(NT) fused2_kernel(const A* __restrict__ x, const A* __restrict__ y, size_t n,
    typename Op::template acc<A>* __restrict__ partials, unsigned int* __restrict__ tickets, typename Op::template acc<A>* __restrict__ out) {
  using Acc = typename Op::template acc<A>;
  constexpr int VEC = VECTOR ? 16 / sizeof(A) : 1;
  constexpr int NSUM = Op::NSUM;
  __shared__ Acc smem[NSUM][NT / 32];
  __shared__ bool is_last;
  __shared__ unsigned int s_chunk[2];

  const size_t nvec = n / VEC;
  Acc acc[NSUM];
#pragma unroll
  for(int k = 0; k < NSUM; ++k) acc[k] = Acc(0);

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
    const Acc w = sb::warp_sum<Acc>(acc[k]);
    if(lane == 0) smem[k][warp] = w;
  }
  __syncthreads();
  if(warp == 0) {
#pragma unroll
    for(int k = 0; k < NSUM; ++k) {
      Acc r = (lane < NT / 32) ? smem[k][lane] : Acc(0);
      r = sb::warp_sum<Acc>(r);
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
  Acc p[NSUM];
#pragma unroll
  for(int k = 0; k < NSUM; ++k) p[k] = Acc(0);
  for(unsigned int i = threadIdx.x; i < gridDim.x; i += NT)
#pragma unroll
    for(int k = 0; k < NSUM; ++k) p[k] += __ldcg(partials + (size_t)i * NSUM + k);
  __syncthreads();  // smem reuse
#pragma unroll
  for(int k = 0; k < NSUM; ++k) {
    const Acc w = sb::warp_sum<Acc>(p[k]);
    if(lane == 0) smem[k][warp] = w;
  }
  __syncthreads();
  if(warp == 0) {
    Acc total[NSUM];
#pragma unroll
    for(int k = 0; k < NSUM; ++k) {
      Acc r = (lane < NT / 32) ? smem[k][lane] : Acc(0);
      total[k] = sb::warp_sum<Acc>(r);
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
int fused2_impl(sb_ctx* ctx, const T* x, const T* y, size_t n, typename Op::template acc<T>* out) {
  using A = typename sb_arith<T>::type;
  using Acc = typename Op::template acc<A>;
  if(!ctx || !out || (n && (!x || !y))) return SB_ERR_BAD_ARG;
  if(!sb::aligned_to(x, sizeof(T)) || !sb::aligned_to(y, sizeof(T)) || !sb::aligned_to(out, sizeof(Acc))) return SB_ERR_MISALIGNED;
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
  Acc* apart = reinterpret_cast<Acc*>(ctx->partials);
  Acc* aout = reinterpret_cast<Acc*>(out);
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

// ---- sb_linreg_error_<T> ----------------------------------------------------------------------------------------------
// Replaces LinearRegressionBench::run (/root/reference/single-kernel/lin_reg_error.cpp:54-79): one work-item per (alpha, beta)
// pair, each looping over ALL N points — N^2 multiply-adds for 5*N*s bytes (2^16 points: 4.3e9 loop iterations per type).
//     out[g] = sum_i ((a*x_i + b) - y_i)^2 ,  a = alpha[g], b = beta[g]
// is a quadratic form in (a, b) whose coefficients are five sums over the points, so the N^2 loop is one fused streaming
// reduction (fused2_kernel<OpLinregErr>: x and y read once, double accumulators for both types) plus one map over alpha / beta:
// 5*N*s bytes moved, HBM-bound.  The map evaluates the centred form
//     a^2*Cxx - 2a*Cxy + Cyy + N*(a*mx + b - my)^2,   mx = Sx/N, my = Sy/N, Cxx = Sxx - Sx*mx, Cxy = Sxy - Sx*my, Cyy = Syy - Sy*my
// (two non-negative terms: no cancellation between O(N) products) in double and rounds once to T.  The result is the exactly
// rounded value up to a few ulp of double; the reference's own sequential sum in T deviates from it by its rounding error
// (fp32, N = 2^16: ~1e-6 relative), far inside the north-star tolerance (tests/test_single_kernel_widened.py).
template <typename A, bool VECTOR>
__global__ void __launch_bounds__(256) linreg_error_map_kernel(const A* __restrict__ alpha, const A* __restrict__ beta, size_t n,
    const double* __restrict__ sums, A* __restrict__ out) {
  constexpr int VEC = VECTOR ? 16 / sizeof(A) : 1;
  const double cnt = (double)n;
  const double sx = sums[0], sy = sums[1], sxx = sums[2], sxy = sums[3], syy = sums[4];
  const double mx = sx / cnt, my = sy / cnt;
  const double cxx = sxx - sx * mx, cxy = sxy - sx * my, cyy = syy - sy * my;
  auto err = [&](A a_, A b_) -> A {
    const double a = (double)a_, b = (double)b_;
    const double d = a * mx + b - my;
    return (A)(a * a * cxx - 2.0 * a * cxy + cyy + cnt * d * d);
  };
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if constexpr(VECTOR) {
    const size_t nvec = n / VEC;
    if(i < nvec) {
      const sb::Vec<A> va = sb::ldg_stream<A>(alpha + i * VEC), vb = sb::ldg_stream<A>(beta + i * VEC);
      sb::Vec<A> o;
#pragma unroll
      for(int c = 0; c < VEC; ++c) o.v[c] = err(va.v[c], vb.v[c]);
      sb::stg_vec<A>(out + i * VEC, o);
    }
    if(blockIdx.x == 0) {
      const size_t t = nvec * VEC + threadIdx.x;
      if(t < n) out[t] = err(alpha[t], beta[t]);
    }
  } else {
    if(i < n) out[i] = err(alpha[i], beta[i]);
  }
}

template <typename T>
int linreg_error_impl(sb_ctx* ctx, const T* x, const T* y, const T* alpha, const T* beta, size_t n, T* out) {
  if(!ctx || (n && (!x || !y || !alpha || !beta || !out))) return SB_ERR_BAD_ARG;
  if(!sb::aligned_to(alpha, sizeof(T)) || !sb::aligned_to(beta, sizeof(T)) || !sb::aligned_to(out, sizeof(T))) return SB_ERR_MISALIGNED;
  if(n == 0) return SB_OK;
  double* sums = reinterpret_cast<double*>(static_cast<char*>(ctx->stage_small) + 128);  // 5 doubles of the ctx's 256-byte slot
  const int rc = fused2_impl<T, OpLinregErr>(ctx, x, y, n, sums);
  if(rc != SB_OK) return rc;
  sb_device_guard guard(ctx->device);
  const bool vector = sb::aligned_to(alpha, 16) && sb::aligned_to(beta, 16) && sb::aligned_to(out, 16);
  const size_t vec = vector ? 16 / sizeof(T) : 1;
  const size_t threads = vector ? (n / vec > 256 ? n / vec : 256) : n;
  const size_t blocks = (threads + 255) / 256;
  if(blocks > 0x7fffffffull) return SB_ERR_BAD_ARG;
  if(vector) linreg_error_map_kernel<T, true><<<(unsigned)blocks, 256, 0, ctx->stream>>>(alpha, beta, n, sums, out);
  else linreg_error_map_kernel<T, false><<<(unsigned)blocks, 256, 0, ctx->stream>>>(alpha, beta, n, sums, out);
  SB_CHECK_LAUNCH(ctx);
  return SB_OK;
}

// ---- sb_kmeans_<T> ------------------------------------------------------------------------------------------------
// Replaces KmeansBench::run (/root/reference/single-kernel/kmeans.cpp:49-82): one work-item per point, nclusters x nfeatures
// squared distances, features stored feature-major (features[l*N + g], :70).  The reference stores the POINT's index, not
// the cluster's (`index = gid`, :74): membership[g] = g as soon as one distance is below the running minimum, which starts
// at FLT_MAX, and 0 otherwise (all distances >= FLT_MAX or NaN).  Kept literally.  A pure map: nfeatures*N*s bytes in,
// 4*N out, HBM-bound; a thread owns one 16-byte vector of points per feature row (128-bit coalesced loads, 128-bit or
// 2 x 128-bit stores), the cluster table sits in shared memory.  Distances are summed in T in the reference's order
// (d*d + dist contracted to one fma, like the reference's own -ffast-math build; it only matters within an ulp of FLT_MAX).
constexpr int KM_MAX_TABLE = 2048;  // cluster-table elements kept in shared memory

// NFR > 0: the point's features (nfeatures <= NFR) are read once and stay in registers over the cluster loop, U vectors of
// points per thread (all U * nfeatures 128-bit loads of a thread are issued before the first use: with one vector per thread
// the kernel had 32 bytes in flight per thread and ran at 4.0-4.4 TB/s of DRAM traffic).  NFR == 0: any nfeatures, the rows
// are re-read per cluster (they stay in L1 / L2).
template <typename A, bool VECTOR, int NFR, int U>
__global__ void __launch_bounds__(256) kmeans_kernel(const A* __restrict__ features, const A* __restrict__ clusters, size_t n, int nfeatures,
    int nclusters, int* __restrict__ membership) {
  constexpr int VEC = VECTOR ? 16 / sizeof(A) : 1;
  constexpr int NF = NFR > 0 ? NFR : 1;
  __shared__ A table[KM_MAX_TABLE];
  const int tsize = nfeatures * nclusters;
  const bool in_smem = tsize <= KM_MAX_TABLE;
  const A* tab = in_smem ? table : clusters;
  const size_t v0 = (size_t)blockIdx.x * (256 * U) + threadIdx.x;  // the thread's vectors: v0 + u * 256
  A f[U][NF][VEC];
  int cnt[U];
#pragma unroll
  for(int u = 0; u < U; ++u) {
    const size_t g0 = (v0 + (size_t)u * 256) * VEC;
    cnt[u] = g0 >= n ? 0 : (n - g0 < (size_t)VEC) ? (int)(n - g0) : VEC;
  }
  if(NFR > 0) {
    // A block whose 256 * U vectors all lie inside the array (every block but the last) loads without per-thread guards:
    // with the guard the compiler merged the vector and the scalar path through predicated moves behind EVERY load, and
    // the U * nfeatures loads of a thread waited for one another (ncu, fp64: 71 % of the warp samples on those moves,
    // 3.8 TB/s of DRAM traffic; profiles/r02_ncu_kmeans_f64.txt).
    const bool full_block = ((size_t)blockIdx.x + 1) * (256 * U) * VEC <= n;
    if(VECTOR && full_block) {
      sb::Vec<A> t[U][NF];
#pragma unroll
      for(int u = 0; u < U; ++u)
#pragma unroll
        for(int l = 0; l < NF; ++l)
          if(l < nfeatures) t[u][l] = sb::ldg_nc_ordered<A>(features + (size_t)l * n + (v0 + (size_t)u * 256) * VEC);
#pragma unroll
      for(int u = 0; u < U; ++u)
#pragma unroll
        for(int l = 0; l < NF; ++l)
#pragma unroll
          for(int c = 0; c < VEC; ++c) f[u][l][c] = t[u][l].v[c];
    } else {
#pragma unroll
      for(int u = 0; u < U; ++u) {
        const size_t g0 = (v0 + (size_t)u * 256) * VEC;
#pragma unroll
        for(int l = 0; l < NF; ++l) {
          if(l < nfeatures) {
            const A* row = features + (size_t)l * n + g0;
#pragma unroll
            for(int c = 0; c < VEC; ++c) f[u][l][c] = c < cnt[u] ? __ldg(row + c) : A(0);
          }
        }
      }
    }
  }
  // the cluster table is staged AFTER the feature loads have been issued (staging it first put two dependent memory
  // latencies into the life of every block: fp64, 2^28 points 1.72 -> 1.34 ms)
  if(in_smem) {
    for(int k = threadIdx.x; k < tsize; k += blockDim.x) table[k] = clusters[k];
    __syncthreads();
  }
#pragma unroll
  for(int u = 0; u < U; ++u) {
    if(cnt[u] == 0) continue;
    const size_t g0 = (v0 + (size_t)u * 256) * VEC;
    A min_dist[VEC];
    int index[VEC];
    bool set[VEC];
#pragma unroll
    for(int c = 0; c < VEC; ++c) {
      min_dist[c] = (A)3.402823466e+38f;  // FLT_MAX, kmeans.cpp:66
      index[c] = 0;
      set[c] = c >= cnt[u];  // (components past the end of the array count as decided)
    }
    for(int k = 0; k < nclusters; ++k) {
      A dist[VEC];
#pragma unroll
      for(int c = 0; c < VEC; ++c) dist[c] = A(0);
      if(NFR > 0) {
#pragma unroll
        for(int l = 0; l < NF; ++l) {
          if(l < nfeatures) {
            const A cl = tab[k * nfeatures + l];
#pragma unroll
            for(int c = 0; c < VEC; ++c) {
              const A d = sub_rn<A>(f[u][l][c], cl);
              dist[c] = fma(d, d, dist[c]);  // (contracted, as the reference's own -ffast-math build does)
            }
          }
        }
      } else {
        for(int l = 0; l < nfeatures; ++l) {
          const A cl = tab[k * nfeatures + l];
          const A* row = features + (size_t)l * n + g0;
#pragma unroll
          for(int c = 0; c < VEC; ++c) {
            const A d = sub_rn<A>(c < cnt[u] ? __ldg(row + c) : A(0), cl);
            dist[c] += mul_rn<A>(d, d);
          }
        }
      }
#pragma unroll
      for(int c = 0; c < VEC; ++c) {  // kmeans.cpp:72-75
        if(dist[c] < min_dist[c]) {
          min_dist[c] = dist[c];
          index[c] = (int)(g0 + c);  // the reference narrows the size_t global id to int the same way
          set[c] = true;
        }
      }
      bool all_set = true;
#pragma unroll
      for(int c = 0; c < VEC; ++c) all_set = all_set && set[c];
      // Once a distance has been below FLT_MAX the stored value is g whatever the remaining clusters give (`index = gid`
      // again): stop.  For ordinary data that is after the first cluster — a third of the arithmetic at nclusters = 3,
      // which matters for fp64 (21 double operations per point against 20 bytes of traffic).
      if(all_set) break;
    }
    if(VECTOR && cnt[u] == VEC && (reinterpret_cast<uintptr_t>(membership + g0) & (4 * VEC - 1)) == 0) {
      if constexpr(VEC == 4) *reinterpret_cast<int4*>(membership + g0) = make_int4(index[0], index[1], index[2], index[3]);
      else if constexpr(VEC == 2) *reinterpret_cast<int2*>(membership + g0) = make_int2(index[0], index[1]);
      else membership[g0] = index[0];
    } else {
#pragma unroll
      for(int c = 0; c < VEC; ++c)
        if(c < cnt[u]) membership[g0 + c] = index[c];
    }
  }
}

template <typename T, bool VECTOR, int NFR, int U>
void kmeans_launch(sb_ctx* ctx, const T* features, const T* clusters, size_t n, int nfeatures, int nclusters, int32_t* membership) {
  const size_t vec = VECTOR ? 16 / sizeof(T) : 1;
  const size_t per_block = 256 * (size_t)U * vec;
  const size_t blocks = (n + per_block - 1) / per_block;
  kmeans_kernel<T, VECTOR, NFR, U><<<(unsigned)blocks, 256, 0, ctx->stream>>>(features, clusters, n, nfeatures, nclusters, membership);
}

template <typename T>
int kmeans_impl(sb_ctx* ctx, const T* features, const T* clusters, size_t n, int nfeatures, int nclusters, int32_t* membership) {
  if(!ctx || nfeatures < 0 || nclusters < 0 || (n && (!membership || (nfeatures && !features) || (nfeatures && nclusters && !clusters)))) return SB_ERR_BAD_ARG;
  if(!sb::aligned_to(features, sizeof(T)) || !sb::aligned_to(clusters, sizeof(T)) || !sb::aligned_to(membership, 4)) return SB_ERR_MISALIGNED;
  if(n == 0) return SB_OK;
  if(n / 256 > 0x7fffffffull) return SB_ERR_BAD_ARG;
  sb_device_guard guard(ctx->device);
  // 128-bit loads need every feature row 16-byte aligned: base aligned and the row pitch N*s a multiple of 16
  const bool vector = sb::aligned_to(features, 16) && (n * sizeof(T)) % 16 == 0;
  if(vector) {
    if(nfeatures <= 2) kmeans_launch<T, true, 2, 4>(ctx, features, clusters, n, nfeatures, nclusters, membership);
    else if(nfeatures <= 4) kmeans_launch<T, true, 4, 2>(ctx, features, clusters, n, nfeatures, nclusters, membership);
    else kmeans_launch<T, true, 0, 1>(ctx, features, clusters, n, nfeatures, nclusters, membership);
  } else {
    if(nfeatures <= 4) kmeans_launch<T, false, 4, 4>(ctx, features, clusters, n, nfeatures, nclusters, membership);
    else kmeans_launch<T, false, 0, 1>(ctx, features, clusters, n, nfeatures, nclusters, membership);
  }
  SB_CHECK_LAUNCH(ctx);
  return SB_OK;
}

}  // namespace

extern "C" {
int sb_linreg_error_f32(sb_ctx* c, const float* x, const float* y, const float* alpha, const float* beta, size_t n, float* out) { return linreg_error_impl<float>(c, x, y, alpha, beta, n, out); }
int sb_linreg_error_f64(sb_ctx* c, const double* x, const double* y, const double* alpha, const double* beta, size_t n, double* out) { return linreg_error_impl<double>(c, x, y, alpha, beta, n, out); }
int sb_kmeans_f32(sb_ctx* c, const float* f, const float* cl, size_t n, int nf, int nc, int32_t* m) { return kmeans_impl<float>(c, f, cl, n, nf, nc, m); }
int sb_kmeans_f64(sb_ctx* c, const double* f, const double* cl, size_t n, int nf, int nc, int32_t* m) { return kmeans_impl<double>(c, f, cl, n, nf, nc, m); }
int sb_dot_i32(sb_ctx* c, const int32_t* a, const int32_t* b, size_t n, int32_t* out) { return fused2_impl<int32_t, OpDot>(c, a, b, n, out); }
int sb_dot_i64(sb_ctx* c, const int64_t* a, const int64_t* b, size_t n, int64_t* out) { return fused2_impl<int64_t, OpDot>(c, a, b, n, out); }
int sb_dot_f32(sb_ctx* c, const float* a, const float* b, size_t n, float* out) { return fused2_impl<float, OpDot>(c, a, b, n, out); }
int sb_dot_f64(sb_ctx* c, const double* a, const double* b, size_t n, double* out) { return fused2_impl<double, OpDot>(c, a, b, n, out); }
int sb_linreg_f32(sb_ctx* c, const float* x, const float* y, size_t n, float* out6) { return fused2_impl<float, OpLinreg>(c, x, y, n, out6); }
int sb_linreg_f64(sb_ctx* c, const double* x, const double* y, size_t n, double* out6) { return fused2_impl<double, OpLinreg>(c, x, y, n, out6); }
}
