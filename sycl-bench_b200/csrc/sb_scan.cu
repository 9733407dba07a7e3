// sb_scan_<T> / sb_segscan_<T>: single-pass prefix sums with decoupled look-back.
//
// Upstream pattern/prefixsum.cpp:1, pattern/scan.cpp:1 and pattern/segmentedscan.cpp:1 are
// one-line TODO stubs; the semantics implemented here are the written contract in DESIGN.md
// (§Scan semantics), derived from the two implemented siblings (same iota input,
// pattern/reduction.cpp:31-34; `--local` as the group/segment length,
// pattern/segmentedreduction.cpp:35-61).
//
// B200 design: HBM-bound copy-like stream (algorithmic bytes 2*N*s).
//  - one CTA of 256 threads per 16 KB tile; a warp owns K=4 consecutive rows of 32 x 16 B, all
//    128-bit loads are issued before the first add (64 B in flight per thread, ~80 KB per SM).
//  - tile ids come from an atomic ticket (forward progress does not depend on CTA dispatch order).
//  - tile descriptors {status, value} are written/read with ONE relaxed 64-bit (4-byte T) or
//    128-bit (8-byte T, st/ld .b128) access, so no fence separates flag and value.
//  - status carries a per-launch epoch: descriptors are never reset between launches.
//  - the segmented variants run the same kernel on (value, head-flag) pairs:
//    (v1,f1)(+)(v2,f2) = (f2 ? v2 : v1+v2, f1|f2).  A tile that contains a head publishes its
//    INCLUSIVE state immediately (it does not depend on its predecessors); a tile that starts on
//    a head skips the look-back; if every tile starts on a head (power-of-two segment <= tile)
//    neither tickets nor descriptors are touched.
#include "sb_common.cuh"

namespace {

constexpr int SCAN_NT = 256;
constexpr int SCAN_NW = SCAN_NT / 32;
constexpr int SCAN_K = 4;

enum : uint32_t { ST_AGG = 1u, ST_INCL = 2u };
enum { MODE_PLAIN = 0, MODE_IMPLICIT = 1, MODE_FLAGS = 2 };

template <typename A> struct ScanCfg {
  static constexpr int VEC = 16 / sizeof(A);
  static constexpr int ROW = 32 * VEC;
  static constexpr int WCHUNK = SCAN_K * ROW;
  static constexpr int TILE = SCAN_NW * WCHUNK;
};

// ---- descriptors ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bits_of(uint32_t v) { return v; }
__device__ __forceinline__ uint32_t bits_of(float v) { return __float_as_uint(v); }
__device__ __forceinline__ unsigned long long bits_of(unsigned long long v) { return v; }
__device__ __forceinline__ unsigned long long bits_of(double v) { return (unsigned long long)__double_as_longlong(v); }
template <typename A> __device__ __forceinline__ A from_bits32(uint32_t b);
template <> __device__ __forceinline__ uint32_t from_bits32<uint32_t>(uint32_t b) { return b; }
template <> __device__ __forceinline__ float from_bits32<float>(uint32_t b) { return __uint_as_float(b); }
template <typename A> __device__ __forceinline__ A from_bits64(unsigned long long b);
template <> __device__ __forceinline__ unsigned long long from_bits64<unsigned long long>(unsigned long long b) { return b; }
template <> __device__ __forceinline__ double from_bits64<double>(unsigned long long b) { return __longlong_as_double((long long)b); }

template <typename A, int BYTES = sizeof(A)> struct Desc;

template <typename A> struct Desc<A, 4> {
  using word = unsigned long long;
  static __device__ __forceinline__ void store(word* d, uint32_t status, A v) {
    const word w = ((word)bits_of(v) << 32) | status;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(d), "l"(w) : "memory");
  }
  static __device__ __forceinline__ void load(const word* d, uint32_t& status, A& v) {
    word w;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(d) : "memory");
    status = (uint32_t)w;
    v = from_bits32<A>((uint32_t)(w >> 32));
  }
};

struct alignas(16) desc128_t { unsigned long long status, value; };

template <typename A> struct Desc<A, 8> {
  using word = desc128_t;
  static __device__ __forceinline__ void store(word* d, uint32_t status, A v) {
    const unsigned long long s = status, b = bits_of(v);
    asm volatile("{.reg .b128 t; mov.b128 t, {%1,%2}; st.relaxed.gpu.global.b128 [%0], t;}" ::"l"(d), "l"(s), "l"(b) : "memory");
  }
  static __device__ __forceinline__ void load(const word* d, uint32_t& status, A& v) {
    unsigned long long s, b;
    asm volatile("{.reg .b128 t; ld.relaxed.gpu.global.b128 t, [%2]; mov.b128 {%0,%1}, t;}" : "=l"(s), "=l"(b) : "l"(d) : "memory");
    status = (s >> 32) ? 0u : (uint32_t)s;
    v = from_bits64<A>(b);
  }
};

// decoupled look-back by one full warp: sum the aggregates of the predecessors down to the first
// tile whose inclusive prefix is known.
template <typename A>
__device__ __forceinline__ A lookback(const typename Desc<A>::word* desc, long long tile, uint32_t epoch) {
  const int lane = sb::lane_id();
  A prefix = A(0);
  long long pred = tile - 1 - lane;
  for(;;) {
    uint32_t st = (epoch << 2) | ST_INCL;
    A val = A(0);
    if(pred >= 0) {
      do { Desc<A>::load(desc + pred, st, val); } while((st >> 2) != epoch);
    }
    const unsigned incl = __ballot_sync(SB_FULL_MASK, (st & 3u) == ST_INCL);
    if(incl) {
      const int first = __ffs(incl) - 1;
      prefix += sb::warp_sum<A>(lane <= first ? val : A(0));
      break;
    }
    prefix += sb::warp_sum<A>(val);
    pred -= 32;
  }
  return prefix;
}

template <typename A, int MODE, bool ALIGNED>
__global__ void __launch_bounds__(SCAN_NT)
scan_kernel(const A* in, A* out, size_t n, int exclusive, size_t seg_len, size_t seg_mask, const uint8_t* __restrict__ flags,
    const A* __restrict__ seed, typename Desc<A>::word* desc, unsigned long long* ticket, unsigned long long ticket_base,
    uint32_t epoch, int use_lookback) {
  using C = ScanCfg<A>;
  constexpr int VEC = C::VEC;
  constexpr int K = SCAN_K;
  __shared__ long long s_tile;
  __shared__ A s_wagg[SCAN_NW];
  __shared__ int s_wflag[SCAN_NW];
  __shared__ A s_prefix;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  long long tile;
  if(use_lookback) {
    if(threadIdx.x == 0) s_tile = (long long)(atomicAdd(ticket, 1ull) - ticket_base);
    __syncthreads();
    tile = s_tile;
  } else {
    tile = blockIdx.x;
  }
  const size_t base = (size_t)tile * C::TILE;
  const size_t wbase = base + (size_t)warp * C::WCHUNK + (size_t)lane * VEC;

  // ---- load -------------------------------------------------------------------------------------
  A v[K][VEC];
  unsigned hf[K];
  const bool full_tile = base + C::TILE <= n;
#pragma unroll
  for(int j = 0; j < K; ++j) {
    const size_t idx = wbase + (size_t)j * C::ROW;
    if(ALIGNED && (full_tile || idx + VEC <= n)) {
      sb::Vec<A> t = sb::ldg_stream<A>(in + idx);
#pragma unroll
      for(int c = 0; c < VEC; ++c) v[j][c] = t.v[c];
    } else {
#pragma unroll
      for(int c = 0; c < VEC; ++c) v[j][c] = (idx + c < n) ? in[idx + c] : A(0);
    }
    hf[j] = 0;
    if(MODE == MODE_IMPLICIT) {
      if(seg_mask != ~(size_t)0) {
#pragma unroll
        for(int c = 0; c < VEC; ++c) hf[j] |= (((idx + c) & seg_mask) == 0 ? 1u : 0u) << c;
      } else {
        size_t r = idx % seg_len;
#pragma unroll
        for(int c = 0; c < VEC; ++c) {
          hf[j] |= (r == 0 ? 1u : 0u) << c;
          r = (r + 1 == seg_len) ? 0 : r + 1;
        }
      }
    } else if(MODE == MODE_FLAGS) {
#pragma unroll
      for(int c = 0; c < VEC; ++c) hf[j] |= ((idx + c < n && flags[idx + c]) ? 1u : 0u) << c;
      if(idx == 0) hf[j] |= 1u;
    }
  }

  // ---- pass 1: thread partials -> warp segmented scan per row -> warp aggregate --------------------
  A carry_in[K];
  bool open[K];
  A row_carry = A(0);
  bool warp_open = true;
#pragma unroll
  for(int j = 0; j < K; ++j) {
    A s = A(0);
    bool f = false;
#pragma unroll
    for(int c = 0; c < VEC; ++c) {
      if(MODE != MODE_PLAIN && ((hf[j] >> c) & 1u)) {
        s = v[j][c];
        f = true;
      } else {
        s += v[j][c];
      }
    }
    unsigned hm = 0;
    int h = 0;
    if(MODE != MODE_PLAIN) {
      hm = __ballot_sync(SB_FULL_MASK, f);
      const unsigned m = hm & (0xffffffffu >> (31 - lane));
      h = m ? 31 - __clz(m) : 0;
    }
    if(lane == 0 && !f) s += row_carry;
    A inc = s;
#pragma unroll
    for(int d = 1; d < 32; d <<= 1) {
      const A t = sb::shfl_up(inc, d);
      if(lane >= h + d) inc += t;
    }
    A ex = sb::shfl_up(inc, 1);
    if(lane == 0) ex = row_carry;
    carry_in[j] = ex;
    open[j] = warp_open && (MODE == MODE_PLAIN || (hm & ((1u << lane) - 1u)) == 0u);
    row_carry = sb::shfl_idx(inc, 31);
    if(MODE != MODE_PLAIN && hm) warp_open = false;
  }
  if(lane == 0) {
    s_wagg[warp] = row_carry;
    s_wflag[warp] = warp_open ? 0 : 1;
  }
  __syncthreads();

  // ---- block level: prefix of the warps before mine, tile aggregate ---------------------------------
  A run = A(0), my_prefix = A(0);
  bool run_open = true, my_open = true;
#pragma unroll
  for(int u = 0; u < SCAN_NW; ++u) {
    if(u == warp) {
      my_prefix = run;
      my_open = run_open;
    }
    if(MODE != MODE_PLAIN && s_wflag[u]) {
      run = s_wagg[u];
      run_open = false;
    } else {
      run += s_wagg[u];
    }
  }
  const A tile_agg = run;
  const bool tile_has_head = !run_open;

  // ---- decoupled look-back (warp 0) ---------------------------------------------------------------
  A tile_prefix = A(0);
  if(use_lookback) {
    if(warp == 0) {
      A p = A(0);
      if(tile == 0) {
        if(MODE == MODE_PLAIN && seed) p = *seed;
        if(lane == 0) Desc<A>::store(desc, (epoch << 2) | ST_INCL, tile_has_head ? tile_agg : p + tile_agg);
      } else {
        if(lane == 0) Desc<A>::store(desc + tile, (epoch << 2) | (tile_has_head ? ST_INCL : ST_AGG), tile_agg);
        const bool starts_on_head = MODE != MODE_PLAIN && (sb::shfl_idx(hf[0], 0) & 1u);
        if(!starts_on_head) {
          p = lookback<A>(desc, tile, epoch);
          if(!tile_has_head && lane == 0) Desc<A>::store(desc + tile, (epoch << 2) | ST_INCL, p + tile_agg);
        }
      }
      if(lane == 0) s_prefix = p;
    }
    __syncthreads();
    tile_prefix = s_prefix;
  }
  const A warp_prefix = my_prefix + (my_open ? tile_prefix : A(0));

  // ---- pass 2: rescan with the known carry, store ---------------------------------------------------
#pragma unroll
  for(int j = 0; j < K; ++j) {
    const size_t idx = wbase + (size_t)j * C::ROW;
    A r = carry_in[j] + (open[j] ? warp_prefix : A(0));
    sb::Vec<A> o;
#pragma unroll
    for(int c = 0; c < VEC; ++c) {
      if(MODE != MODE_PLAIN && ((hf[j] >> c) & 1u)) r = A(0);
      const A e = r;
      r += v[j][c];
      o.v[c] = exclusive ? e : r;
    }
    if(ALIGNED && (full_tile || idx + VEC <= n)) {
      sb::stg_vec<A>(out + idx, o);
    } else {
#pragma unroll
      for(int c = 0; c < VEC; ++c)
        if(idx + c < n) out[idx + c] = o.v[c];
    }
  }
}

template <typename T>
int scan_launch(sb_ctx* ctx, const T* in, T* out, size_t n, int exclusive, int mode, size_t seg_len, const uint8_t* flags, const T* seed) {
  using A = typename sb_arith<T>::type;
  using C = ScanCfg<A>;
  using W = typename Desc<A>::word;
  if(!ctx || (n && (!in || !out))) return SB_ERR_BAD_ARG;
  if(!sb::aligned_to(in, sizeof(T)) || !sb::aligned_to(out, sizeof(T)) || !sb::aligned_to(seed, sizeof(T))) return SB_ERR_MISALIGNED;
  if(n == 0) return SB_OK;
  sb_device_guard guard(ctx->device);

  const size_t tiles = (n + C::TILE - 1) / C::TILE;
  if(tiles > 0x7fffffffull) return SB_ERR_BAD_ARG;
  size_t seg_mask = ~(size_t)0;
  int use_lookback = 1;
  if(mode == MODE_IMPLICIT) {
    if(sb::is_pow2(seg_len)) {
      seg_mask = seg_len - 1;
      if(seg_len <= (size_t)C::TILE) use_lookback = 0;  // every tile starts on a head
    }
  }
  if(tiles == 1 && !(mode == MODE_PLAIN && seed)) use_lookback = 0;

  W* desc = nullptr;
  uint32_t epoch = 0;
  unsigned long long tbase = 0;
  if(use_lookback) {
    int rc = sb_ctx_reserve_desc(ctx, sizeof(A) == 4 ? tiles : 0, sizeof(A) == 8 ? tiles : 0);
    if(rc != SB_OK) return rc;
    if(ctx->scan_epoch >= (1u << 30) - 1) {
      if(ctx->desc64) SB_CUDA_TRY(cudaMemsetAsync(ctx->desc64, 0, ctx->desc64_count * 8, ctx->stream));
      if(ctx->desc128) SB_CUDA_TRY(cudaMemsetAsync(ctx->desc128, 0, ctx->desc128_count * 16, ctx->stream));
      ctx->scan_epoch = 0;
    }
    epoch = ++ctx->scan_epoch;
    desc = reinterpret_cast<W*>(sizeof(A) == 4 ? ctx->desc64 : ctx->desc128);
    tbase = ctx->scan_ticket_base;
    ctx->scan_ticket_base += tiles;
  }
  const bool aligned = sb::aligned_to(in, 16) && sb::aligned_to(out, 16);
  const A* ain = reinterpret_cast<const A*>(in);
  A* aout = reinterpret_cast<A*>(out);
  const A* aseed = reinterpret_cast<const A*>(seed);
  const dim3 grid((unsigned)tiles), block(SCAN_NT);
#define SB_SCAN_LAUNCH(MODE_, AL_)                                                                                       \
  scan_kernel<A, MODE_, AL_><<<grid, block, 0, ctx->stream>>>(ain, aout, n, exclusive, seg_len, seg_mask, flags, aseed, \
      desc, ctx->scan_ticket, tbase, epoch, use_lookback)
  if(mode == MODE_PLAIN) {
    if(aligned) SB_SCAN_LAUNCH(MODE_PLAIN, true); else SB_SCAN_LAUNCH(MODE_PLAIN, false);
  } else if(mode == MODE_IMPLICIT) {
    if(aligned) SB_SCAN_LAUNCH(MODE_IMPLICIT, true); else SB_SCAN_LAUNCH(MODE_IMPLICIT, false);
  } else {
    if(aligned) SB_SCAN_LAUNCH(MODE_FLAGS, true); else SB_SCAN_LAUNCH(MODE_FLAGS, false);
  }
#undef SB_SCAN_LAUNCH
  SB_CHECK_LAUNCH(ctx);
  return SB_OK;
}

template <typename T>
int scan_impl(sb_ctx* ctx, const T* in, T* out, size_t n, int exclusive, size_t group_len, const T* seed) {
  if(group_len == 0 || group_len >= n) return scan_launch<T>(ctx, in, out, n, exclusive, MODE_PLAIN, 0, nullptr, seed);
  if(seed) return SB_ERR_BAD_ARG;  // a carry-in only makes sense for the global scan
  return scan_launch<T>(ctx, in, out, n, exclusive, MODE_IMPLICIT, group_len, nullptr, nullptr);
}

template <typename T>
int segscan_impl(sb_ctx* ctx, const T* in, T* out, size_t n, size_t seg_len, const uint8_t* flags, int exclusive) {
  if(flags) return scan_launch<T>(ctx, in, out, n, exclusive, MODE_FLAGS, 0, flags, nullptr);
  if(seg_len == 0) return SB_ERR_SEG_LEN;
  if(seg_len >= n) return scan_launch<T>(ctx, in, out, n, exclusive, MODE_PLAIN, 0, nullptr, nullptr);
  return scan_launch<T>(ctx, in, out, n, exclusive, MODE_IMPLICIT, seg_len, nullptr, nullptr);
}

}  // namespace

extern "C" {
int sb_scan_i32(sb_ctx* c, const int32_t* in, int32_t* out, size_t n, int ex, size_t g, const int32_t* seed) { return scan_impl<int32_t>(c, in, out, n, ex, g, seed); }
int sb_scan_i64(sb_ctx* c, const int64_t* in, int64_t* out, size_t n, int ex, size_t g, const int64_t* seed) { return scan_impl<int64_t>(c, in, out, n, ex, g, seed); }
int sb_scan_f32(sb_ctx* c, const float* in, float* out, size_t n, int ex, size_t g, const float* seed) { return scan_impl<float>(c, in, out, n, ex, g, seed); }
int sb_scan_f64(sb_ctx* c, const double* in, double* out, size_t n, int ex, size_t g, const double* seed) { return scan_impl<double>(c, in, out, n, ex, g, seed); }

int sb_segscan_i32(sb_ctx* c, const int32_t* in, int32_t* out, size_t n, size_t L, const uint8_t* f, int ex) { return segscan_impl<int32_t>(c, in, out, n, L, f, ex); }
int sb_segscan_i64(sb_ctx* c, const int64_t* in, int64_t* out, size_t n, size_t L, const uint8_t* f, int ex) { return segscan_impl<int64_t>(c, in, out, n, L, f, ex); }
int sb_segscan_f32(sb_ctx* c, const float* in, float* out, size_t n, size_t L, const uint8_t* f, int ex) { return segscan_impl<float>(c, in, out, n, L, f, ex); }
int sb_segscan_f64(sb_ctx* c, const double* in, double* out, size_t n, size_t L, const uint8_t* f, int ex) { return segscan_impl<double>(c, in, out, n, L, f, ex); }
}
