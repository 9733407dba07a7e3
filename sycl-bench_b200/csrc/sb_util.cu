// Device-side input generation, multi-GPU carry helpers and closed-form verification.
//  - sb_iota_<T>      : generate_input on the device (/root/reference/pattern/reduction.cpp:31-34,
//                       pattern/segmentedreduction.cpp:25-28), needed because 2^33 elements
//                       (32 GiB) cannot be staged through the host.
//  - sb_carry_add_<T> : the "fused carry-add pass" of the multi-GPU prefix sum (BASELINE.json).
//  - sb_check_iota_<T>: verify() against the closed forms of SURVEY.md A.4, on the device.
#include "sb_common.cuh"

namespace {

constexpr int UT_NT = 256;

template <typename A> __device__ __forceinline__ A cast_index(unsigned long long g) { return (A)g; }

// Where local element i of a shard sits in the global array.  block == 0: contiguous shard starting at `offset`.
// block > 0: block-cyclic layout of sb_scan_cyclic (include/sb_cuda.h): blocks of `block` elements, block k on rank
// k % world, a rank stores its blocks back to back; every rank holds the same n elements, so the blocks of the last
// round are all `last_len` = n % block long and follow each other in rank order.
struct Layout {
  unsigned long long offset, block, full, last_len;
  unsigned world, rank;
  __device__ __forceinline__ unsigned long long global(size_t i) const {
    if(block == 0) return offset + i;
    const unsigned long long b = i / block, j = i - b * block;
    if(b < full) return (b * world + rank) * block + j;
    return full * world * block + rank * last_len + j;
  }
};

template <typename A>
__global__ void __launch_bounds__(UT_NT) iota_kernel(A* __restrict__ out, size_t n, Layout lay, int aligned) {
  constexpr int VEC = 16 / sizeof(A);
  const size_t stride = (size_t)gridDim.x * UT_NT * VEC;
  for(size_t i = ((size_t)blockIdx.x * UT_NT + threadIdx.x) * VEC; i < n; i += stride) {
    if(aligned && i + VEC <= n) {
      sb::Vec<A> v;
#pragma unroll
      for(int c = 0; c < VEC; ++c) v.v[c] = cast_index<A>(lay.global(i + c));
      sb::stg_vec<A>(out + i, v);
    } else {
#pragma unroll
      for(int c = 0; c < VEC; ++c)
        if(i + c < n) out[i + c] = cast_index<A>(lay.global(i + c));
    }
  }
}

template <typename A>
__global__ void __launch_bounds__(UT_NT) fill_kernel(A* __restrict__ out, size_t n, A value, int aligned) {
  constexpr int VEC = 16 / sizeof(A);
  const size_t stride = (size_t)gridDim.x * UT_NT * VEC;
  sb::Vec<A> v;
#pragma unroll
  for(int c = 0; c < VEC; ++c) v.v[c] = value;
  for(size_t i = ((size_t)blockIdx.x * UT_NT + threadIdx.x) * VEC; i < n; i += stride) {
    if(aligned && i + VEC <= n) {
      sb::stg_vec<A>(out + i, v);
    } else {
#pragma unroll
      for(int c = 0; c < VEC; ++c)
        if(i + c < n) out[i + c] = value;
    }
  }
}

// sb_copy: flat grid, one 16 KB chunk per CTA (512 threads x 2 x 16 B, both loads issued before the first
// store).  The block scheduler hands the chunks out in address order, which keeps the DRAM access front
// compact: 6.93 TB/s against 5.97 TB/s for a persistent grid-stride loop and 6.58 TB/s for cudaMemcpyAsync
// device-to-device (profiles/r01_stream_probe.txt).
constexpr int CP_NT = 512;
constexpr int CP_U = 2;
__global__ void __launch_bounds__(CP_NT) copy_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t nvec) {
  const size_t b = (size_t)blockIdx.x * CP_NT * CP_U + threadIdx.x;
  uint4 t[CP_U];
#pragma unroll
  for(int u = 0; u < CP_U; ++u)
    if(b + (size_t)u * CP_NT < nvec) t[u] = __ldg(in + b + (size_t)u * CP_NT);
#pragma unroll
  for(int u = 0; u < CP_U; ++u)
    if(b + (size_t)u * CP_NT < nvec) out[b + (size_t)u * CP_NT] = t[u];
}
// bytes that the vector body does not cover (head / tail), or everything when src and dst are not congruent mod 16
__global__ void __launch_bounds__(UT_NT) copy_bytes_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ out, size_t n) {
  for(size_t i = (size_t)blockIdx.x * UT_NT + threadIdx.x; i < n; i += (size_t)gridDim.x * UT_NT) out[i] = in[i];
}

template <typename A>
__global__ void __launch_bounds__(UT_NT) carry_add_kernel(A* __restrict__ buf, size_t n, const A* __restrict__ carry, int aligned) {
  constexpr int VEC = 16 / sizeof(A);
  constexpr int U = 4;
  const A c0 = *carry;
  const size_t nvec = aligned ? n / VEC : 0;
  const size_t stride = (size_t)gridDim.x * UT_NT * U;
  for(size_t b = (size_t)blockIdx.x * UT_NT * U + threadIdx.x; b < nvec; b += stride) {
    sb::Vec<A> t[U];
#pragma unroll
    for(int u = 0; u < U; ++u)
      if(b + (size_t)u * UT_NT < nvec) t[u] = sb::ldg_stream<A>(buf + (b + (size_t)u * UT_NT) * VEC);
#pragma unroll
    for(int u = 0; u < U; ++u)
      if(b + (size_t)u * UT_NT < nvec) {
#pragma unroll
        for(int c = 0; c < VEC; ++c) t[u].v[c] += c0;
        sb::stg_vec<A>(buf + (b + (size_t)u * UT_NT) * VEC, t[u]);
      }
  }
  // scalar remainder
  const size_t rem_begin = nvec * VEC;
  for(size_t i = rem_begin + (size_t)blockIdx.x * UT_NT + threadIdx.x; i < n; i += (size_t)gridDim.x * UT_NT) buf[i] += c0;
}

template <typename A>
__global__ void exclusive_carry_kernel(const A* __restrict__ totals, int rank, A* __restrict__ carry_out) {
  if(threadIdx.x == 0 && blockIdx.x == 0) {
    A c = A(0);
    for(int r = 0; r < rank; ++r) c += totals[r];
    *carry_out = c;
  }
}

// g*(g+1)/2 mod 2^64 without overflowing the intermediate product
__device__ __forceinline__ unsigned long long tri(unsigned long long g) {
  return (g & 1ull) ? g * ((g + 1ull) >> 1) : (g >> 1) * (g + 1ull);
}

// closed forms of SURVEY.md A.4 as exact integers mod 2^64 ...
__device__ __forceinline__ unsigned long long expect_int(unsigned long long g, int kind, unsigned long long seg_len) {
  if(seg_len == 0) return (kind == 0) ? tri(g) : (g == 0 ? 0ull : tri(g - 1ull));
  const unsigned long long q = g / seg_len, j = g - q * seg_len;
  if(kind == 0) return q * seg_len * (j + 1ull) + tri(j);
  if(kind == 1) return q * seg_len * j + (j == 0 ? 0ull : tri(j - 1ull));
  return (j == 0) ? q * seg_len * seg_len + tri(seg_len - 1ull) : g;
}
// ... and in double precision (the values exceed 2^64 for g > 2^32; relative error of the evaluation ~1e-16)
__device__ __forceinline__ double trid(unsigned long long g) { return 0.5 * (double)g * ((double)g + 1.0); }
__device__ __forceinline__ double expect_fp(unsigned long long g, int kind, unsigned long long seg_len, bool& exact) {
  exact = false;
  if(seg_len == 0) return (kind == 0) ? trid(g) : (g == 0 ? 0.0 : trid(g - 1ull));
  const unsigned long long q = g / seg_len, j = g - q * seg_len;
  const double base = (double)q * (double)seg_len;
  if(kind == 0) return base * ((double)j + 1.0) + trid(j);
  if(kind == 1) return base * (double)j + (j == 0 ? 0.0 : trid(j - 1ull));
  if(j == 0) return base * (double)seg_len + trid(seg_len - 1ull);
  exact = true;  // segmented-reduction layout: everything but the head is the untouched input
  return (double)g;
}

template <typename A> struct is_fp { static constexpr bool value = false; };
template <> struct is_fp<float> { static constexpr bool value = true; };
template <> struct is_fp<double> { static constexpr bool value = true; };

// Integer types: bit-exact (mod 2^w).  Floating point: |data - closed form| <= rel_tol * max(|closed form|, 1); the
// untouched elements of the segmented-reduction layout must equal the input (T)g exactly.
template <typename A>
__global__ void __launch_bounds__(UT_NT) check_iota_kernel(const A* __restrict__ data, size_t n, Layout lay, int kind,
    unsigned long long seg_len, double rel_tol, unsigned long long* __restrict__ mismatches) {
  unsigned int bad = 0;
  const size_t stride = (size_t)gridDim.x * UT_NT;
  for(size_t i = (size_t)blockIdx.x * UT_NT + threadIdx.x; i < n; i += stride) {
    const unsigned long long g = lay.global(i);
    if(is_fp<A>::value) {
      bool exact;
      const double want = expect_fp(g, kind, seg_len, exact);
      const double got = (double)data[i];
      if(exact) {
        if(data[i] != cast_index<A>(g)) ++bad;
      } else if(!(fabs(got - want) <= rel_tol * fmax(fabs(want), 1.0))) {  // the negation also catches NaN
        ++bad;
      }
    } else {
      if(data[i] != (A)expect_int(g, kind, seg_len)) ++bad;
    }
  }
  bad = __reduce_add_sync(SB_FULL_MASK, bad);
  if((threadIdx.x & 31) == 0 && bad) atomicAdd(mismatches, (unsigned long long)bad);
}

size_t grid_for(sb_ctx* ctx, size_t work_items, size_t per_block) {
  // flat grids: one chunk per CTA, dispatched in order (an ordered DRAM access front is worth ~15 % on
  // B200 against a persistent grid-stride grid, profiles/r01_stream_probe.txt); the kernels still loop
  // if the cap is ever hit
  (void)ctx;
  size_t blocks = (work_items + per_block - 1) / per_block;
  const size_t cap = size_t(1) << 22;
  if(blocks > cap) blocks = cap;
  return blocks ? blocks : 1;
}

template <typename T>
int iota_impl(sb_ctx* ctx, T* out, size_t n, Layout lay) {
  using A = typename sb_arith<T>::type;
  if(!ctx || (n && !out)) return SB_ERR_BAD_ARG;
  if(!sb::aligned_to(out, sizeof(T))) return SB_ERR_MISALIGNED;
  if(n == 0) return SB_OK;
  sb_device_guard guard(ctx->device);
  constexpr int VEC = 16 / sizeof(T);
  const size_t blocks = grid_for(ctx, n, (size_t)UT_NT * VEC);
  iota_kernel<A><<<(unsigned)blocks, UT_NT, 0, ctx->stream>>>(reinterpret_cast<A*>(out), n, lay, sb::aligned_to(out, 16) ? 1 : 0);
  SB_CHECK_LAUNCH(ctx);
  return SB_OK;
}

template <typename T>
int fill_impl(sb_ctx* ctx, T* out, size_t n, T value) {
  using A = typename sb_arith<T>::type;
  if(!ctx || (n && !out)) return SB_ERR_BAD_ARG;
  if(!sb::aligned_to(out, sizeof(T))) return SB_ERR_MISALIGNED;
  if(n == 0) return SB_OK;
  sb_device_guard guard(ctx->device);
  constexpr int VEC = 16 / sizeof(T);
  const size_t blocks = grid_for(ctx, n, (size_t)UT_NT * VEC);
  fill_kernel<A><<<(unsigned)blocks, UT_NT, 0, ctx->stream>>>(reinterpret_cast<A*>(out), n, (A)value, sb::aligned_to(out, 16) ? 1 : 0);
  SB_CHECK_LAUNCH(ctx);
  return SB_OK;
}

template <typename T>
int carry_add_impl(sb_ctx* ctx, T* inout, size_t n, const T* carry) {
  using A = typename sb_arith<T>::type;
  if(!ctx || !carry || (n && !inout)) return SB_ERR_BAD_ARG;
  if(!sb::aligned_to(inout, sizeof(T)) || !sb::aligned_to(carry, sizeof(T))) return SB_ERR_MISALIGNED;
  if(n == 0) return SB_OK;
  sb_device_guard guard(ctx->device);
  constexpr int VEC = 16 / sizeof(T);
  const size_t blocks = grid_for(ctx, n, (size_t)UT_NT * VEC * 4);
  carry_add_kernel<A><<<(unsigned)blocks, UT_NT, 0, ctx->stream>>>(
      reinterpret_cast<A*>(inout), n, reinterpret_cast<const A*>(carry), sb::aligned_to(inout, 16) ? 1 : 0);
  SB_CHECK_LAUNCH(ctx);
  return SB_OK;
}

template <typename T>
int exclusive_carry_impl(sb_ctx* ctx, const T* totals, int rank, T* carry_out) {
  using A = typename sb_arith<T>::type;
  if(!ctx || !totals || !carry_out || rank < 0) return SB_ERR_BAD_ARG;
  sb_device_guard guard(ctx->device);
  exclusive_carry_kernel<A><<<1, 32, 0, ctx->stream>>>(reinterpret_cast<const A*>(totals), rank, reinterpret_cast<A*>(carry_out));
  SB_CHECK_LAUNCH(ctx);
  return SB_OK;
}

template <typename T>
int check_impl(sb_ctx* ctx, const T* data, size_t n, Layout lay, int kind, size_t seg_len, double rel_tol, uint64_t* mismatches) {
  using A = typename sb_arith<T>::type;
  if(!ctx || !mismatches || (n && !data) || kind < 0 || kind > 2 || !(rel_tol >= 0.0)) return SB_ERR_BAD_ARG;
  if(kind == 2 && (seg_len == 0 || !sb::is_pow2(seg_len))) return SB_ERR_SEG_LEN;
  sb_device_guard guard(ctx->device);
  SB_CUDA_TRY(cudaMemsetAsync(mismatches, 0, sizeof(uint64_t), ctx->stream));
  if(n == 0) return SB_OK;
  const size_t blocks = grid_for(ctx, n, (size_t)UT_NT * 8);
  check_iota_kernel<A><<<(unsigned)blocks, UT_NT, 0, ctx->stream>>>(
      reinterpret_cast<const A*>(data), n, lay, kind, seg_len, rel_tol, reinterpret_cast<unsigned long long*>(mismatches));
  SB_CHECK_LAUNCH(ctx);
  return SB_OK;
}

inline Layout contiguous(uint64_t offset) { return Layout{offset, 0ull, 0ull, 0ull, 1u, 0u}; }
inline bool cyclic_ok(size_t block, int world, int rank) { return block > 0 && world >= 1 && rank >= 0 && rank < world; }
inline Layout cyclic(size_t n, size_t block, int world, int rank) {
  return Layout{0ull, (unsigned long long)block, (unsigned long long)(n / block), (unsigned long long)(n % block), (unsigned)world, (unsigned)rank};
}

}  // namespace

extern "C" {
int sb_iota_i16(sb_ctx* c, int16_t* o, size_t n, uint64_t off) { return iota_impl<int16_t>(c, o, n, contiguous(off)); }
int sb_iota_i32(sb_ctx* c, int32_t* o, size_t n, uint64_t off) { return iota_impl<int32_t>(c, o, n, contiguous(off)); }
int sb_iota_i64(sb_ctx* c, int64_t* o, size_t n, uint64_t off) { return iota_impl<int64_t>(c, o, n, contiguous(off)); }
int sb_iota_f32(sb_ctx* c, float* o, size_t n, uint64_t off) { return iota_impl<float>(c, o, n, contiguous(off)); }
int sb_iota_f64(sb_ctx* c, double* o, size_t n, uint64_t off) { return iota_impl<double>(c, o, n, contiguous(off)); }
#define SB_IOTA_CYCLIC(sfx, T)                                                                              \
  int sb_iota_cyclic_##sfx(sb_ctx* c, T* o, size_t n, size_t block, int world, int rank) {                  \
    if(!cyclic_ok(block, world, rank)) return SB_ERR_BAD_ARG;                                               \
    return iota_impl<T>(c, o, n, cyclic(n, block, world, rank));                                               \
  }
SB_IOTA_CYCLIC(i32, int32_t)
SB_IOTA_CYCLIC(i64, int64_t)
SB_IOTA_CYCLIC(f32, float)
SB_IOTA_CYCLIC(f64, double)
#undef SB_IOTA_CYCLIC
int sb_copy(sb_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if(!ctx || (bytes && (!dst || !src))) return SB_ERR_BAD_ARG;
  if(bytes == 0 || dst == src) return SB_OK;
  sb_device_guard guard(ctx->device);
  const unsigned char* s = static_cast<const unsigned char*>(src);
  unsigned char* d = static_cast<unsigned char*>(dst);
  if(d < s + bytes && s < d + bytes) return SB_ERR_BAD_ARG;  // partially overlapping ranges (the kernel reads through the read-only path)
  const uintptr_t sa = reinterpret_cast<uintptr_t>(s) & 15, da = reinterpret_cast<uintptr_t>(d) & 15;
  if(sa != da) {  // no common 16-byte phase: byte copy
    copy_bytes_kernel<<<(unsigned)grid_for(ctx, bytes, UT_NT), UT_NT, 0, ctx->stream>>>(s, d, bytes);
    SB_CHECK_LAUNCH(ctx);
    return SB_OK;
  }
  size_t head = (16 - sa) & 15;
  if(head > bytes) head = bytes;
  const size_t nvec = (bytes - head) / 16;
  const size_t tail = bytes - head - nvec * 16;
  if(nvec) {
    const size_t chunks = (nvec + (size_t)CP_NT * CP_U - 1) / ((size_t)CP_NT * CP_U);
    if(chunks > 0x7fffffffull) return SB_ERR_BAD_ARG;
    copy_kernel<<<(unsigned)chunks, CP_NT, 0, ctx->stream>>>(reinterpret_cast<const uint4*>(s + head), reinterpret_cast<uint4*>(d + head), nvec);
    SB_CHECK_LAUNCH(ctx);
  }
  if(head) {
    copy_bytes_kernel<<<1, UT_NT, 0, ctx->stream>>>(s, d, head);
    SB_CHECK_LAUNCH(ctx);
  }
  if(tail) {
    copy_bytes_kernel<<<1, UT_NT, 0, ctx->stream>>>(s + head + nvec * 16, d + head + nvec * 16, tail);
    SB_CHECK_LAUNCH(ctx);
  }
  return SB_OK;
}
int sb_fill_i32(sb_ctx* c, int32_t* o, size_t n, int32_t v) { return fill_impl<int32_t>(c, o, n, v); }
int sb_fill_i64(sb_ctx* c, int64_t* o, size_t n, int64_t v) { return fill_impl<int64_t>(c, o, n, v); }
int sb_fill_f32(sb_ctx* c, float* o, size_t n, float v) { return fill_impl<float>(c, o, n, v); }
int sb_fill_f64(sb_ctx* c, double* o, size_t n, double v) { return fill_impl<double>(c, o, n, v); }

int sb_carry_add_i32(sb_ctx* c, int32_t* p, size_t n, const int32_t* k) { return carry_add_impl<int32_t>(c, p, n, k); }
int sb_carry_add_i64(sb_ctx* c, int64_t* p, size_t n, const int64_t* k) { return carry_add_impl<int64_t>(c, p, n, k); }
int sb_carry_add_f32(sb_ctx* c, float* p, size_t n, const float* k) { return carry_add_impl<float>(c, p, n, k); }
int sb_carry_add_f64(sb_ctx* c, double* p, size_t n, const double* k) { return carry_add_impl<double>(c, p, n, k); }

int sb_exclusive_carry_i32(sb_ctx* c, const int32_t* t, int r, int32_t* o) { return exclusive_carry_impl<int32_t>(c, t, r, o); }
int sb_exclusive_carry_i64(sb_ctx* c, const int64_t* t, int r, int64_t* o) { return exclusive_carry_impl<int64_t>(c, t, r, o); }
int sb_exclusive_carry_f32(sb_ctx* c, const float* t, int r, float* o) { return exclusive_carry_impl<float>(c, t, r, o); }
int sb_exclusive_carry_f64(sb_ctx* c, const double* t, int r, double* o) { return exclusive_carry_impl<double>(c, t, r, o); }

int sb_check_iota_i32(sb_ctx* c, const int32_t* d, size_t n, uint64_t off, int kind, size_t L, uint64_t* m) { return check_impl<int32_t>(c, d, n, contiguous(off), kind, L, 0.0, m); }
int sb_check_iota_i64(sb_ctx* c, const int64_t* d, size_t n, uint64_t off, int kind, size_t L, uint64_t* m) { return check_impl<int64_t>(c, d, n, contiguous(off), kind, L, 0.0, m); }
int sb_check_iota_f32(sb_ctx* c, const float* d, size_t n, uint64_t off, int kind, size_t L, double tol, uint64_t* m) { return check_impl<float>(c, d, n, contiguous(off), kind, L, tol, m); }
int sb_check_iota_f64(sb_ctx* c, const double* d, size_t n, uint64_t off, int kind, size_t L, double tol, uint64_t* m) { return check_impl<double>(c, d, n, contiguous(off), kind, L, tol, m); }
#define SB_CHECK_CYCLIC(sfx, T)                                                                                                        \
  int sb_check_iota_cyclic_##sfx(sb_ctx* c, const T* d, size_t n, size_t block, int world, int rank, int kind, double tol, uint64_t* m) { \
    if(!cyclic_ok(block, world, rank) || kind > 1) return SB_ERR_BAD_ARG;                                                              \
    return check_impl<T>(c, d, n, cyclic(n, block, world, rank), kind, 0, tol, m);                                                        \
  }
SB_CHECK_CYCLIC(i32, int32_t)
SB_CHECK_CYCLIC(i64, int64_t)
SB_CHECK_CYCLIC(f32, float)
SB_CHECK_CYCLIC(f64, double)
#undef SB_CHECK_CYCLIC
}
