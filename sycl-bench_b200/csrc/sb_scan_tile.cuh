// Device building blocks shared by the scan kernels (sb_scan.cu, sb_scan_cyclic.cu): tile geometry, tile
// descriptors, mbarrier / TMA bulk-copy primitives and the per-warp row scans.  See sb_scan.cu for the design.
#pragma once
#include "sb_common.cuh"
#include "sb_peer.cuh"

namespace sbscan {

enum : uint32_t { ST_AGG = 1u, ST_INCL = 2u };
enum { MODE_PLAIN = 0, MODE_IMPLICIT = 1, MODE_FLAGS = 2 };

template <typename A, int NW_, int K> struct TileCfg {
  static constexpr int VEC = 16 / sizeof(A);
  static constexpr int NW = NW_;
  static constexpr int ROW = 32 * VEC;      // elements one warp covers with one 128-bit access per lane
  static constexpr int WCHUNK = K * ROW;    // consecutive elements owned by a warp
  static constexpr int TILE = NW * WCHUNK;  // elements per tile
  static constexpr int TILE_BYTES = NW * K * 512;
};

// ---- descriptors ---------------------------------------------------------------------------------------------
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

// ---- mbarrier / TMA bulk-copy primitives ------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while(!done) {
    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
}
// non-blocking probe: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{.reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}"
               : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void bulk_load_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, unsigned long long policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}

// ---- head flags as bit masks ---------------------------------------------------------------------------------------
// The segmented scans restart at a head: s = head ? v : s + v.  Written as a branch per element that was a bit test, a
// compare, a select and a move per element and pass (ncu, explicit-flag kernel: SEL 21 %, IMAD 18 %, ISETP 14 %, LOP3 9 % of
// the executed instructions, 30 of the 48 thread instructions per element).  With the flag expanded to an all-ones / all-zeros
// word it is one AND on the bits and one add: s = (s & ~m) + v  (0 + v == v exactly; a NaN in front of a head is cleared too).
__device__ __forceinline__ uint32_t head_mask(unsigned hf, int c) { return 0u - ((hf >> c) & 1u); }  // bit c of hf -> 0xffffffff / 0
__device__ __forceinline__ uint32_t clear_if(uint32_t s, uint32_t m) { return s & ~m; }
__device__ __forceinline__ float clear_if(float s, uint32_t m) { return __uint_as_float(__float_as_uint(s) & ~m); }
__device__ __forceinline__ unsigned long long clear_if(unsigned long long s, uint32_t m) {
  return s & ~(((unsigned long long)m << 32) | m);
}
__device__ __forceinline__ double clear_if(double s, uint32_t m) {
  return __longlong_as_double(__double_as_longlong(s) & (long long)~(((unsigned long long)m << 32) | m));
}

// VEC flag bytes (one per element, any non-zero value = head) -> one bit per element: byte-wise compare, keep bit c of byte c,
// gather the bytes in the top byte by a multiply (3 instructions instead of a test + select + shift + or per element)
__device__ __forceinline__ unsigned pack_flag_bytes(uint32_t w) { return ((__vcmpne4(w, 0u) & 0x08040201u) * 0x01010101u) >> 24; }

// ---- per-warp building blocks ---------------------------------------------------------------------------------
// head flags of the K x VEC elements a thread owns (bit c of hf[j] = element c of row j is a head)
template <typename A, int MODE, int K>
__device__ __forceinline__ void tile_head_flags(unsigned (&hf)[K], size_t wbase, size_t n, size_t seg_len, size_t seg_mask,
    const uint8_t* __restrict__ flags) {
  constexpr int VEC = 16 / sizeof(A);
  constexpr int ROW = 32 * VEC;
#pragma unroll
  for(int j = 0; j < K; ++j) {
    const size_t idx = wbase + (size_t)j * ROW;
    hf[j] = 0;
    if(MODE == MODE_IMPLICIT) {
      if(seg_mask != ~(size_t)0) {
        if(seg_mask >= (size_t)(VEC - 1)) {
          // power-of-two segment of at least one vector: only element 0 of an aligned vector can be a head
          hf[j] = (idx & seg_mask) == 0 ? 1u : 0u;
        } else {
#pragma unroll
          for(int c = 0; c < VEC; ++c) hf[j] |= (((idx + c) & seg_mask) == 0 ? 1u : 0u) << c;
        }
      } else {
        size_t r = idx % seg_len;
#pragma unroll
        for(int c = 0; c < VEC; ++c) {
          hf[j] |= (r == 0 ? 1u : 0u) << c;
          r = (r + 1 == seg_len) ? 0 : r + 1;
        }
      }
    } else if(MODE == MODE_FLAGS) {
      if(idx + VEC <= n && (reinterpret_cast<uintptr_t>(flags + idx) & (VEC - 1)) == 0) {
        uint32_t w;  // VEC flag bytes in one access
        if(VEC == 4) w = __ldg(reinterpret_cast<const uint32_t*>(flags + idx));
        else w = __ldg(reinterpret_cast<const uint16_t*>(flags + idx));
        hf[j] = pack_flag_bytes(w);
      } else {
#pragma unroll
        for(int c = 0; c < VEC; ++c) hf[j] |= ((idx + c < n && flags[idx + c]) ? 1u : 0u) << c;
      }
      if(idx == 0) hf[j] |= 1u;
    }
  }
}

// explicit head flags staged in the ring by the producer (one flag byte per element, warp chunk at `wflags`):
// VEC bytes per lane and row, bank-conflict free
template <typename A, int K>
__device__ __forceinline__ void tile_head_flags_ring(unsigned (&hf)[K], const unsigned char* wflags, size_t wbase) {
  constexpr int VEC = 16 / sizeof(A);
  constexpr int ROW = 32 * VEC;
  const unsigned char* fp = wflags + (threadIdx.x & 31) * VEC;
#pragma unroll
  for(int j = 0; j < K; ++j) {
    uint32_t w;
    if(VEC == 4) w = *reinterpret_cast<const uint32_t*>(fp + j * ROW);
    else w = *reinterpret_cast<const uint16_t*>(fp + j * ROW);
    hf[j] = pack_flag_bytes(w);
    if(wbase + (size_t)j * ROW == 0) hf[j] |= 1u;  // element 0 always starts a segment
  }
}

// guarded global -> register load of a (possibly partial, possibly unaligned) tile
template <typename A, bool ALIGNED, int K>
__device__ __forceinline__ void tile_load_global(A (&v)[K][16 / sizeof(A)], const A* in, size_t wbase, bool full_tile, size_t n) {
  constexpr int VEC = 16 / sizeof(A);
#pragma unroll
  for(int j = 0; j < K; ++j) {
    const size_t idx = wbase + (size_t)j * (32 * VEC);
    if(ALIGNED && (full_tile || idx + VEC <= n)) {
      sb::Vec<A> t = sb::ldg_stream<A>(in + idx);
#pragma unroll
      for(int c = 0; c < VEC; ++c) v[j][c] = t.v[c];
    } else {
#pragma unroll
      for(int c = 0; c < VEC; ++c) v[j][c] = (idx + c < n) ? in[idx + c] : A(0);
    }
  }
}

// shared-memory ring -> registers (the lane-consecutive 16-byte layout is bank-conflict free)
template <typename A, int K>
__device__ __forceinline__ void tile_load_smem(A (&v)[K][16 / sizeof(A)], const unsigned char* warp_rows) {
  constexpr int VEC = 16 / sizeof(A);
  const unsigned char* sp = warp_rows + (threadIdx.x & 31) * 16;
#pragma unroll
  for(int j = 0; j < K; ++j) {
    const sb::Vec<A> t = *reinterpret_cast<const sb::Vec<A>*>(sp + (size_t)j * 512);
#pragma unroll
    for(int c = 0; c < VEC; ++c) v[j][c] = t.v[c];
  }
}

// the warp's K rows reduced to (sum since the last head, contains a head)
template <typename A, int MODE, int K>
__device__ __forceinline__ void warp_reduce_rows(const A (&v)[K][16 / sizeof(A)], const unsigned (&hf)[K], A& agg, bool& has_head) {
  constexpr int VEC = 16 / sizeof(A);
  const int lane = threadIdx.x & 31;
  if(MODE == MODE_PLAIN) {
    A s = A(0);
#pragma unroll
    for(int j = 0; j < K; ++j) {
      A r = v[j][0];
#pragma unroll
      for(int c = 1; c < VEC; ++c) r += v[j][c];
      s += r;
    }
    agg = sb::warp_sum<A>(s);
    has_head = false;
    return;
  }
  {
    unsigned any = 0;
#pragma unroll
    for(int j = 0; j < K; ++j) any |= hf[j];
    if(!__any_sync(SB_FULL_MASK, any != 0u)) {  // no head in the warp chunk: one warp sum instead of K
      warp_reduce_rows<A, MODE_PLAIN, K>(v, hf, agg, has_head);
      return;
    }
  }
  // Two forms.  PER CHUNK (second branch), measured at 2^28 against PER ROW (profiles/r02_ops_v16_phasea_ab.jsonl): +13-18 % for
  // 64-bit types with explicit flags at random positions (int64 3.14 -> 3.66 TB/s), +6.5 % for 64-bit implicit heads at
  // L = 65536, 0 ... +10 % for 32-bit explicit flags — but -4.5 ... -6 % for 32-bit implicit heads (the K partials push the
  // 18-worker instantiation over its 96 registers), which therefore keep the per-row form.
  constexpr bool PER_ROW =
#ifdef SB_SCAN_PHASEA_PER_ROW
      true;
#else
#ifdef SB_SCAN_PHASEA_PER_CHUNK32
      false;
#else
      MODE == MODE_IMPLICIT && sizeof(A) == 4;
#endif
#endif
  if(PER_ROW) {
    // one warp sum per row (K x 5 shuffles), restarted at the last head of a row
    A run = A(0);
    bool runf = false;
#pragma unroll
    for(int j = 0; j < K; ++j) {
      A s = A(0);
      const bool f = hf[j] != 0u;
#pragma unroll
      for(int c = 0; c < VEC; ++c) s = clear_if(s, head_mask(hf[j], c)) + v[j][c];
      const unsigned hm = __ballot_sync(SB_FULL_MASK, f);
      if(hm) {
        const int hl = 31 - __clz(hm);
        run = sb::warp_sum<A>(lane >= hl ? s : A(0));
        runf = true;
      } else {
        run += sb::warp_sum<A>(s);
      }
    }
    agg = run;
    has_head = runf;
  } else {
    // Only what follows the LAST head of the chunk counts.  Its place (row jl, lane ll) comes out of K ballots; every lane
    // then adds up the rows behind it — of row jl only the lanes from ll on, whose partial `s` already starts at their own
    // last head — and ONE warp sum finishes: 5 shuffles per chunk instead of 5 per row.
    A s[K];
    int jl = -1, ll = 0;
#pragma unroll
    for(int j = 0; j < K; ++j) {
      s[j] = A(0);
#pragma unroll
      for(int c = 0; c < VEC; ++c) s[j] = clear_if(s[j], head_mask(hf[j], c)) + v[j][c];
      const unsigned hm = __ballot_sync(SB_FULL_MASK, hf[j] != 0u);
      if(hm) {
        jl = j;
        ll = 31 - __clz(hm);
      }
    }
    A acc = A(0);
#pragma unroll
    for(int j = 0; j < K; ++j)
      if(j > jl || (j == jl && lane >= ll)) acc += s[j];
    agg = sb::warp_sum<A>(acc);
    has_head = jl >= 0;
  }
}

// pass 1: thread partials -> per-row warp segmented scan.  The K row scans are independent shuffle
// chains (the compiler interleaves them: ILP K); the carries between rows are added afterwards.
// carry_in[j] = running sum entering the thread's part of row j, relative to the start of the warp
// chunk (or the last head inside it); open[j] = no head between the start of the chunk and there.
// Segmented variant of pass 1 with the row scans chained through the carry: fewer live registers
// (no per-row totals / head masks), which matters more than ILP once head flags are in play.
template <typename A, int MODE, int K>
__device__ __forceinline__ void tile_pass1_chained(const A (&v)[K][16 / sizeof(A)], const unsigned (&hf)[K], A (&carry_in)[K],
    bool (&open)[K], A& warp_agg, bool& warp_has_head) {
  constexpr int VEC = 16 / sizeof(A);
  const int lane = threadIdx.x & 31;
  A row_carry = A(0);
  bool warp_open = true;
#pragma unroll
  for(int j = 0; j < K; ++j) {
    A s = A(0);
    const bool f = hf[j] != 0u;
#pragma unroll
    for(int c = 0; c < VEC; ++c) s = clear_if(s, head_mask(hf[j], c)) + v[j][c];
    const unsigned hm = __ballot_sync(SB_FULL_MASK, f);
    const unsigned m = hm & (0xffffffffu >> (31 - lane));
    const int h = m ? 31 - __clz(m) : 0;
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
    open[j] = warp_open && (hm & ((1u << lane) - 1u)) == 0u;
    row_carry = sb::shfl_idx(inc, 31);
    if(hm) warp_open = false;
  }
  warp_agg = row_carry;
  warp_has_head = !warp_open;
}

// UNCHAINED: chunks that hold heads also take the K independent row scans below (ILP K instead of K dependent
// shuffle chains) at the price of the per-row totals / head masks in registers.  Measured at 2^28
// (profiles/r02_ops_v2*.jsonl): explicit flags +4..5 % (32-bit), 64-bit independent tiles +4.5..7 %, but the
// implicit-head look-back kernel loses 6-13 % (96-register cap: spills), so the callers choose.
template <typename A, int MODE, int K, bool UNCHAINED = false>
__device__ __forceinline__ void tile_pass1(const A (&v)[K][16 / sizeof(A)], const unsigned (&hf)[K], A (&carry_in)[K], bool (&open)[K],
    A& warp_agg, bool& warp_has_head) {
  if(MODE != MODE_PLAIN) {
    // a warp chunk without any head (the common case once segments are longer than K rows) is a plain scan:
    // take the unchained path, whose K row scans are independent shuffle chains
    unsigned any = 0;
#pragma unroll
    for(int j = 0; j < K; ++j) any |= hf[j];
    if(!__any_sync(SB_FULL_MASK, any != 0u)) {
      tile_pass1<A, MODE_PLAIN, K>(v, hf, carry_in, open, warp_agg, warp_has_head);
      return;
    }
#ifndef SB_SCAN_UNCHAINED
    if(!UNCHAINED) {
      tile_pass1_chained<A, MODE, K>(v, hf, carry_in, open, warp_agg, warp_has_head);
      return;
    }
#endif
  }
  constexpr int VEC = 16 / sizeof(A);
  const int lane = threadIdx.x & 31;
  A tot[K];          // row sum since the last head of the row (whole row if it has none)
  unsigned hm[K];    // lanes of row j that contain a head
#pragma unroll
  for(int j = 0; j < K; ++j) {
    A s = A(0);
    const bool f = MODE != MODE_PLAIN && hf[j] != 0u;
#pragma unroll
    for(int c = 0; c < VEC; ++c) s = (MODE != MODE_PLAIN ? clear_if(s, head_mask(hf[j], c)) : s) + v[j][c];
    hm[j] = 0;
    int h = 0;
    if(MODE != MODE_PLAIN) {
      hm[j] = __ballot_sync(SB_FULL_MASK, f);
      const unsigned m = hm[j] & (0xffffffffu >> (31 - lane));
      h = m ? 31 - __clz(m) : 0;
    }
    A inc = s;
    if(MODE == MODE_PLAIN) {
      sb::warp_inclusive_scan<A>(inc);  // predicated shuffle-adds
    } else {
#pragma unroll
      for(int d = 1; d < 32; d <<= 1) {
        const A t = sb::shfl_up(inc, d);
        if(lane >= h + d) inc += t;
      }
    }
    const A ex = sb::shfl_up(inc, 1);
    carry_in[j] = lane == 0 ? A(0) : ex;
    tot[j] = sb::shfl_idx(inc, 31);
  }
  A off = A(0);
  bool warp_open = true;
#pragma unroll
  for(int j = 0; j < K; ++j) {
    const bool row_open = MODE == MODE_PLAIN || (hm[j] & ((1u << lane) - 1u)) == 0u;  // no head in front of me in this row
    if(row_open) carry_in[j] += off;
    open[j] = warp_open && row_open;
    if(MODE != MODE_PLAIN && hm[j]) {
      off = tot[j];
      warp_open = false;
    } else {
      off += tot[j];
    }
  }
  warp_agg = off;
  warp_has_head = !warp_open;
}

// pass 2: rescan with the known carry, 128-bit coalesced stores from registers
// (HINT: the 128-bit stores carry an L2 cache policy, see sb_scan_cyclic.cu)
// (EXCL: -1 = the run-time `exclusive` flag decides per element — a 64-bit select costs two SEL per element, 16 % of the
// instructions of the 64-bit global scans in ncu —, 0 / 1 = decided at compile time by the caller's branch)
template <typename A, int MODE, bool ALIGNED, int K, bool HINT = false, int EXCL = -1>
__device__ __forceinline__ void tile_pass2_store(const A (&v)[K][16 / sizeof(A)], const unsigned (&hf)[K], const A (&carry_in)[K],
    const bool (&open)[K], A warp_prefix, size_t wbase, bool full_tile, size_t n, A* out, int exclusive, unsigned long long policy = 0) {
  constexpr int VEC = 16 / sizeof(A);
#pragma unroll
  for(int j = 0; j < K; ++j) {
    const size_t idx = wbase + (size_t)j * (32 * VEC);
    A r = carry_in[j] + (open[j] ? warp_prefix : A(0));
    sb::Vec<A> o;
#pragma unroll
    for(int c = 0; c < VEC; ++c) {
      if(MODE != MODE_PLAIN) r = clear_if(r, head_mask(hf[j], c));
      const A e = r;
      r += v[j][c];
      o.v[c] = (EXCL < 0 ? exclusive != 0 : EXCL != 0) ? e : r;
    }
    if(ALIGNED && (full_tile || idx + VEC <= n)) {
      if(HINT) {
        const uint4& w = *reinterpret_cast<const uint4*>(&o);
        asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(out + idx), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w), "l"(policy) : "memory");
      } else {
        sb::stg_vec<A>(out + idx, o);
      }
    } else {
#pragma unroll
      for(int c = 0; c < VEC; ++c)
        if(idx + c < n) out[idx + c] = o.v[c];
    }
  }
}

// prefix of the warps in front of `warp` inside a tile and the tile's own (aggregate, has-head)
template <typename A, int MODE, int NW>
__device__ __forceinline__ void combine_warps(const A* wagg, const int* wflag, int warp, A& my_prefix, bool& my_open, A& tile_agg,
    bool& tile_has_head) {
  A run = A(0);
  bool run_open = true;
  my_prefix = A(0);
  my_open = true;
#pragma unroll
  for(int u = 0; u < NW; ++u) {
    if(u == warp) {
      my_prefix = run;
      my_open = run_open;
    }
    if(MODE != MODE_PLAIN && wflag[u]) {
      run = wagg[u];
      run_open = false;
    } else {
      run += wagg[u];
    }
  }
  tile_agg = run;
  tile_has_head = !run_open;
}

// ---- polling helpers shared by the look-back warps -------------------------------------------------------------
// Warp-collective: sum of the aggregates desc[0 .. count) of one round, in a fixed order (lane-strided partial sums,
// then the xor tree), polling until every descriptor carries `epoch`.  All loads of a pass are issued back to back.
template <typename A>
__device__ __forceinline__ A sum_round_aggregates(const typename Desc<A>::word* row, int count, uint32_t epoch, Deadline& dl, unsigned* err) {
  constexpr int U = 5;  // 5 x 32 descriptors per window: one window covers a round of 148 tiles
  const int lane = sb::lane_id();
  A lane_sum = A(0);
  for(int lo = 0; lo < count; lo += 32 * U) {
    uint32_t st[U];
    A val[U];
#pragma unroll
    for(int u = 0; u < U; ++u) {
      const int q = lo + u * 32 + lane;
      st[u] = (epoch << 2) | ST_AGG;
      val[u] = A(0);
      if(q < count) Desc<A>::load(row + q, st[u], val[u]);
    }
    for(;;) {
      bool missing = false;
#pragma unroll
      for(int u = 0; u < U; ++u) {
        const int q = lo + u * 32 + lane;
        if(q < count && (st[u] >> 2) != epoch) {
          missing = true;
          Desc<A>::load(row + q, st[u], val[u]);
        }
      }
      if(!__any_sync(SB_FULL_MASK, missing)) break;
      if(__any_sync(SB_FULL_MASK, dl.expired())) {
        if(lane == 0) *err = 1u;
        break;
      }
    }
#pragma unroll
    for(int u = 0; u < U; ++u) lane_sum += val[u];
  }
  return sb::warp_sum<A>(lane_sum);
}

}  // namespace sbscan
