// sb_segreduce_<T>: per-segment sum, in place, result at the segment head.
//
// Replaces SegmentedReduction<T>::submit_ndrange / submit_hierarchical
// (/root/reference/pattern/segmentedreduction.cpp:35-61, :63-96): one work-group of `--local`
// work-items per segment, one element per work-item, log2(L) barrier tree in local memory, then
// `acc[gid] = scratch[0]` for lid == 0.  Layout contract: :112-119 (head = sum, rest untouched).
//
// B200 design: HBM-bound stream, algorithmic bytes N*s + (N/L)*s.
//  - no shared memory and no barriers: a warp streams "super-rows" of U=4 rows x 32 lanes x 16 B
//    (2 KB, four 128-bit loads in flight per lane, 128 KB per resident SM);
//  - a segment is reduced with xor-shuffles over exactly the lanes that hold it
//    (L <= 2 KB / s ... sub-warp groups; larger L ... register accumulation over super-rows then
//    one warp reduce);
//  - the head is rewritten by the lane that owns it (8..2 B scattered stores, one per segment).
//  - grid = sm_count x 8 CTAs of 256 threads, warp-granular grid-stride so that a CTA reads
//    16 KB contiguous per iteration.
// A generic warp-per-segment kernel covers everything else (L not a power of two, L < 16 B,
// unaligned buffers); it sums exactly the subset of a segment the reference's halving tree
// reaches (oracle/sb_oracle.c: ref_tree_member).
#include "sb_common.cuh"

namespace {

constexpr int SR_NT = 256;
constexpr int SR_U = 4;
constexpr int SR_CTAS_PER_SM = 8;

template <typename A> struct acc_of { using type = A; };
template <> struct acc_of<uint16_t> { using type = uint32_t; };

template <typename Acc>
__device__ __forceinline__ Acc group_sum(Acc s, int lanes) {
  for(int d = lanes >> 1; d > 0; d >>= 1) s += __shfl_xor_sync(SB_FULL_MASK, s, d);
  return s;
}

// Four (two) independent warp-wide sums with 6 (5) shuffles instead of 20 (10): every xor step that
// still carries several values halves their number (a lane keeps the values of "its" half and sends
// the others).  On return lane l holds sum number 2*bit4(l) + bit3(l)  (bit4(l) for warp_sum2).
template <typename Acc>
__device__ __forceinline__ Acc warp_sum4(Acc a0, Acc a1, Acc a2, Acc a3, int lane) {
  const bool hi = lane & 16;
  Acc k0 = hi ? a2 : a0, k1 = hi ? a3 : a1;
  k0 += __shfl_xor_sync(SB_FULL_MASK, hi ? a0 : a2, 16);
  k1 += __shfl_xor_sync(SB_FULL_MASK, hi ? a1 : a3, 16);
  const bool hi2 = lane & 8;
  Acc k = hi2 ? k1 : k0;
  k += __shfl_xor_sync(SB_FULL_MASK, hi2 ? k0 : k1, 8);
  k += __shfl_xor_sync(SB_FULL_MASK, k, 4);
  k += __shfl_xor_sync(SB_FULL_MASK, k, 2);
  k += __shfl_xor_sync(SB_FULL_MASK, k, 1);
  return k;
}
template <typename Acc>
__device__ __forceinline__ Acc warp_sum2(Acc a0, Acc a1, int lane) {
  const bool hi = lane & 16;
  Acc k = hi ? a1 : a0;
  k += __shfl_xor_sync(SB_FULL_MASK, hi ? a0 : a1, 16);
  k += __shfl_xor_sync(SB_FULL_MASK, k, 8);
  k += __shfl_xor_sync(SB_FULL_MASK, k, 4);
  k += __shfl_xor_sync(SB_FULL_MASK, k, 2);
  k += __shfl_xor_sync(SB_FULL_MASK, k, 1);
  return k;
}

template <typename A, int NT, int U>
__global__ void __launch_bounds__(NT) segreduce_pow2_kernel(A* buf, size_t n, size_t seg_len) {
  using Acc = typename acc_of<A>::type;
  constexpr int VEC = 16 / sizeof(A);
  constexpr size_t ROW = 32 * VEC;
  constexpr size_t SUPER = U * ROW;
  static_assert(U == 4, "the two-rows-per-segment case is written for U == 4");
  const int lane = threadIdx.x & 31;
  const size_t gw = (size_t)blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
  const size_t nw = (size_t)gridDim.x * (NT / 32);

  if(seg_len <= SUPER) {
    const size_t nsuper = (n + SUPER - 1) / SUPER;
    for(size_t sr = gw; sr < nsuper; sr += nw) {
      const size_t base = sr * SUPER + (size_t)lane * VEC;
      sb::Vec<A> t[U];
      bool valid[U];
#pragma unroll
      for(int u = 0; u < U; ++u) {
        const size_t idx = base + u * ROW;
        valid[u] = idx < n;  // n % VEC == 0, vectors are entirely inside or outside
        if(valid[u]) t[u] = sb::ldg_stream<A>(buf + idx);
      }
      Acc s[U];
#pragma unroll
      for(int u = 0; u < U; ++u) {
        s[u] = Acc(0);
        if(valid[u]) {
#pragma unroll
          for(int c = 0; c < VEC; ++c) s[u] += t[u].v[c];
        }
      }
      const size_t sr_base = sr * SUPER;
      if(seg_len == ROW && !sb::has_redux<Acc>::value) {
        // one row per segment: four warp sums at once (64-bit / floating-point shuffles are the cost here;
        // 32-bit integers use one redux.sync per sum instead)
        const Acc t4 = warp_sum4<Acc>(s[0], s[1], s[2], s[3], lane);
        const int u = lane >> 3;
        if((lane & 7) == 0 && sr_base + u * ROW < n) buf[sr_base + u * ROW] = (A)t4;
      } else if(seg_len <= ROW) {
        const int lps = (int)(seg_len / VEC);  // lanes per segment, power of two in [1, 32]
#pragma unroll
        for(int u = 0; u < U; ++u) {
          s[u] = group_sum<Acc>(s[u], lps);
          if(valid[u] && (lane & (lps - 1)) == 0) buf[base + u * ROW] = (A)s[u];
        }
      } else if(seg_len == 2 * ROW && !sb::has_redux<Acc>::value) {
        const Acc t2 = warp_sum2<Acc>(s[0] + s[1], s[2] + s[3], lane);
        const int u = lane >> 4;
        if((lane & 15) == 0 && sr_base + u * 2 * ROW < n) buf[sr_base + u * 2 * ROW] = (A)t2;
      } else if(seg_len == 2 * ROW) {
        const Acc a = sb::warp_sum<Acc>(s[0] + s[1]), b = sb::warp_sum<Acc>(s[2] + s[3]);
        if(lane == 0) {
          if(valid[0]) buf[base] = (A)a;
          if(valid[2]) buf[base + 2 * ROW] = (A)b;
        }
      } else {  // seg_len == SUPER
        Acc a = sb::warp_sum<Acc>((s[0] + s[1]) + (s[2] + s[3]));
        if(lane == 0 && valid[0]) buf[base] = (A)a;
      }
    }
  } else {
    const size_t nseg = n / seg_len;
    const size_t per = seg_len / SUPER;
    for(size_t seg = gw; seg < nseg; seg += nw) {
      A* p = buf + seg * seg_len + (size_t)lane * VEC;
      Acc a[U];
#pragma unroll
      for(int u = 0; u < U; ++u) a[u] = Acc(0);
      for(size_t k = 0; k < per; ++k) {
        sb::Vec<A> t[U];
#pragma unroll
        for(int u = 0; u < U; ++u) t[u] = sb::ldg_stream<A>(p + k * SUPER + u * ROW);
#pragma unroll
        for(int u = 0; u < U; ++u) {
#pragma unroll
          for(int c = 0; c < VEC; ++c) a[u] += t[u].v[c];
        }
      }
      Acc total = sb::warp_sum<Acc>((a[0] + a[1]) + (a[2] + a[3]));
      if(lane == 0) buf[seg * seg_len] = (A)total;
    }
  }
}

// Which offsets of a segment the reference's tree `for(i = L/2; i > 0; i /= 2) s[l] += s[l+i]`
// (pattern/segmentedreduction.cpp:52-56) actually folds into s[0]: the subset sums of the halving
// sequence.  For a power-of-two L that is every offset.
__device__ __forceinline__ bool ref_tree_member(size_t j, size_t seg_len) {
  for(size_t i = seg_len >> 1; i > 0; i >>= 1)
    if(j >= i) j -= i;
  return j == 0;
}

template <typename A, int NT>
__global__ void __launch_bounds__(NT) segreduce_generic_kernel(A* buf, size_t n, size_t seg_len, int pow2) {
  using Acc = typename acc_of<A>::type;
  const int lane = threadIdx.x & 31;
  const size_t gw = (size_t)blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
  const size_t nw = (size_t)gridDim.x * (NT / 32);
  const size_t nseg = n / seg_len;
  for(size_t seg = gw; seg < nseg; seg += nw) {
    const A* p = buf + seg * seg_len;
    Acc a = Acc(0);
    for(size_t j = lane; j < seg_len; j += 32)
      if(pow2 || ref_tree_member(j, seg_len)) a += p[j];
    a = sb::warp_sum<Acc>(a);
    __syncwarp();
    if(lane == 0) buf[seg * seg_len] = (A)a;
  }
}

template <typename T>
int segreduce_impl(sb_ctx* ctx, T* inout, size_t n, size_t seg_len) {
  using A = typename sb_arith<T>::type;
  if(!ctx || (n && !inout)) return SB_ERR_BAD_ARG;
  if(seg_len == 0 || n % seg_len != 0) return SB_ERR_SEG_LEN;
  if(!sb::aligned_to(inout, sizeof(T))) return SB_ERR_MISALIGNED;
  if(n == 0) return SB_OK;
  sb_device_guard guard(ctx->device);
  constexpr int VEC = 16 / sizeof(T);
  A* buf = reinterpret_cast<A*>(inout);
  const bool pow2 = sb::is_pow2(seg_len);
  const size_t max_blocks = (size_t)ctx->sm_count * SR_CTAS_PER_SM;
  if(pow2 && seg_len >= (size_t)VEC && sb::aligned_to(inout, 16)) {
    const size_t super = (size_t)SR_U * 32 * VEC;
    const size_t tasks = seg_len <= super ? (n + super - 1) / super : n / seg_len;
    size_t blocks = (tasks + SR_NT / 32 - 1) / (SR_NT / 32);
    if(blocks > max_blocks) blocks = max_blocks;
    segreduce_pow2_kernel<A, SR_NT, SR_U><<<(unsigned)blocks, SR_NT, 0, ctx->stream>>>(buf, n, seg_len);
  } else {
    const size_t tasks = n / seg_len;
    size_t blocks = (tasks + SR_NT / 32 - 1) / (SR_NT / 32);
    if(blocks > max_blocks) blocks = max_blocks;
    segreduce_generic_kernel<A, SR_NT><<<(unsigned)blocks, SR_NT, 0, ctx->stream>>>(buf, n, seg_len, pow2 ? 1 : 0);
  }
  SB_CHECK_LAUNCH(ctx);
  return SB_OK;
}

}  // namespace

extern "C" {
int sb_segreduce_i16(sb_ctx* c, int16_t* p, size_t n, size_t L) { return segreduce_impl<int16_t>(c, p, n, L); }
int sb_segreduce_i32(sb_ctx* c, int32_t* p, size_t n, size_t L) { return segreduce_impl<int32_t>(c, p, n, L); }
int sb_segreduce_i64(sb_ctx* c, int64_t* p, size_t n, size_t L) { return segreduce_impl<int64_t>(c, p, n, L); }
int sb_segreduce_f32(sb_ctx* c, float* p, size_t n, size_t L) { return segreduce_impl<float>(c, p, n, L); }
int sb_segreduce_f64(sb_ctx* c, double* p, size_t n, size_t L) { return segreduce_impl<double>(c, p, n, L); }
}
