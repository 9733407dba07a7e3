// sb_scan_<T> / sb_segscan_<T>: single-pass prefix sums with decoupled look-back.
//
// Upstream pattern/prefixsum.cpp:1, pattern/scan.cpp:1 and pattern/segmentedscan.cpp:1 are
// one-line TODO stubs; the semantics implemented here are the written contract in DESIGN.md
// (§Scan semantics), derived from the two implemented siblings (same iota input,
// pattern/reduction.cpp:31-34; `--local` as the group/segment length,
// pattern/segmentedreduction.cpp:35-61).
//
// B200 design: HBM-bound copy-like stream (algorithmic bytes 2*N*s).
//
//  scan_ws_kernel (16-byte aligned buffers) — persistent, warp-specialised, mbarrier-only:
//    * grid = co-resident CTAs (sm_count x occupancy), tiles dealt round-robin: tile = r*G + pos.
//    * producer warp: keeps a ring of WS_STAGES cp.async.bulk (TMA) tile loads in flight
//      (full/empty mbarriers, complete_tx byte counting).
//    * worker warps, software-pipelined over the CTA's tiles with depth WS_DEPTH:
//        phase A (tile k+D): read the tile from the ring, reduce it to one (sum, has-head) pair per
//                            warp, hand the pairs to the look-back warp (agg mbarrier);
//        phase B (tile k)  : wait for the tile prefix (prefix mbarrier), read the tile again from
//                            the ring, scan it with the carry and write 128-bit coalesced stores
//                            straight from registers, then release the stage (empty mbarrier).
//    * look-back warp: combines the warp pairs, publishes the tile descriptor, runs the ROUND
//      look-back (below) and posts the tile prefix.  Workers never wait for a look-back that has not
//      had D tiles of time to finish; there is no __syncthreads after the prologue.
//    History (profiles/): v1 one CTA per 16 KB tile + LDG: 2.1 TB/s, 40 barrier-stall cycles per
//    issue, memory duty ~25 % (ncu).  v2 persistent + TMA ring, block barriers: 2.7-3.3 TB/s,
//    5.5 TB/s without look-back.  v3 + round look-back: 2.9-4.2 TB/s.
//
//  Round look-back: the G tiles of a round are in flight together and publish their aggregates at
//    about the same time, so a tile sums the aggregates of positions pos-1..0 of ITS OWN round
//    (SCAN_WINDOW x 32 descriptors per step, all loads issued before the first wait; a tile that
//    contains a segment head ends the walk) and adds the previous round's prefix, which the CTA
//    at position G-1 publishes once per round.  One dependent step per ROUND instead of one per
//    32 tiles, nobody waits for a neighbour's look-back, and the summation order is fixed
//    (deterministic floating point).
//
//  scan_kernel (unaligned or tiny inputs): one CTA per 16 KB tile, tile ids from an atomic ticket,
//    classic windowed look-back.
//
//  Common: a warp owns K consecutive rows of 32 x 16 B; per row one warp (segmented) scan by
//  shuffles.  Descriptors {status, value} are written/read with ONE relaxed 64-bit (4-byte T) or
//  128-bit (8-byte T, st/ld .b128) access, so no fence separates flag and value; the status
//  carries a per-launch epoch, so descriptors are never reset between launches.  Segmented
//  variants run the same code on (value, head-flag) pairs: (v1,f1)(+)(v2,f2) = (f2 ? v2 : v1+v2,
//  f1|f2); if every tile starts on a head (power-of-two segment <= tile) no descriptor is touched.
#include <cuda.h>

#include <atomic>
#include <cstdlib>

#include "sb_common.cuh"

namespace {

// ---- tuning knobs (make tune builds variants with -D...) -------------------------------------------------
#ifndef SB_SCAN_NT
#define SB_SCAN_NT 256  // ticketed kernel: threads
#endif
#ifndef SB_SCAN_K
#define SB_SCAN_K 4  // ticketed kernel: rows per warp
#endif
#ifndef SB_SCAN_WINDOW
#define SB_SCAN_WINDOW 5  // look-back: batches of 32 descriptors polled per step (5 x 32 covers a whole round of 148 tiles)
#endif
#ifndef SB_WS_NW
#define SB_WS_NW 16  // persistent kernel: worker warps
#endif
#ifndef SB_WS_K
#define SB_WS_K 8  // rows per worker warp and tile (tile = NW*K*512 bytes = 64 KB)
#endif
#ifndef SB_WS_STAGES
#define SB_WS_STAGES 3
#endif
#ifndef SB_WS_DEPTH
#define SB_WS_DEPTH 1  // phase A runs this many tiles ahead of phase B
#endif
#ifndef SB_WS_MINB
#define SB_WS_MINB 1
#endif
constexpr int SCAN_NT = SB_SCAN_NT;
constexpr int SCAN_K = SB_SCAN_K;
constexpr int SCAN_WINDOW = SB_SCAN_WINDOW;
constexpr int WS_NW = SB_WS_NW, WS_K = SB_WS_K, WS_STAGES = SB_WS_STAGES, WS_DEPTH = SB_WS_DEPTH;
#ifndef SB_WS_NW_WIDE
#define SB_WS_NW_WIDE 18  // worker warps of the 72 KB-tile instantiations (global scans; 32-bit independent tiles)
#endif
constexpr int WS_NW_WIDE = SB_WS_NW_WIDE;
#ifndef SB_WS_NW_NARROW
#define SB_WS_NW_NARROW 15  // 15 workers + producer = 16 warps: 128 registers per thread for 64-bit independent tiles
#endif
constexpr int WS_NW_NARROW = SB_WS_NW_NARROW;
#ifndef SB_WS_NW_FLAGS
#define SB_WS_NW_FLAGS SB_WS_NW  // worker warps of the explicit-flag instantiation
#endif
constexpr int WS_NW_FLAGS = SB_WS_NW_FLAGS;
constexpr int WS_NSLOT = 2 * WS_DEPTH + 2;  // a fast warp is at most DEPTH+1 tiles ahead of a slow one

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

// classic decoupled look-back by one full warp: sum the aggregates of the predecessors down to the
// first tile whose inclusive prefix is known (ticketed kernel)
template <typename A>
__device__ __forceinline__ A lookback(const typename Desc<A>::word* desc, long long tile, uint32_t epoch) {
  const int lane = sb::lane_id();
  A prefix = A(0);
  long long pred = tile - 1 - lane;
  for(;;) {
    uint32_t st[SCAN_WINDOW];
    A val[SCAN_WINDOW];
#pragma unroll
    for(int w = 0; w < SCAN_WINDOW; ++w) {
      st[w] = (epoch << 2) | ST_INCL;
      val[w] = A(0);
      if(pred - 32 * w >= 0) Desc<A>::load(desc + (pred - 32 * w), st[w], val[w]);
    }
#pragma unroll
    for(int w = 0; w < SCAN_WINDOW; ++w) {
      if(pred - 32 * w >= 0) {
        while((st[w] >> 2) != epoch) Desc<A>::load(desc + (pred - 32 * w), st[w], val[w]);
      }
      const unsigned incl = __ballot_sync(SB_FULL_MASK, (st[w] & 3u) == ST_INCL);
      if(incl) {
        const int first = __ffs(incl) - 1;
        prefix += sb::warp_sum<A>(lane <= first ? val[w] : A(0));
        return prefix;
      }
      prefix += sb::warp_sum<A>(val[w]);
    }
    pred -= 32 * SCAN_WINDOW;
  }
}

// round look-back (persistent kernel), see the header comment
template <typename A, int MODE>
__device__ __forceinline__ A lookback_round(const typename Desc<A>::word* desc, const typename Desc<A>::word* rdesc, long long tile,
    long long G, uint32_t epoch, const A* __restrict__ seed) {
  constexpr int U = SCAN_WINDOW;
  const int lane = sb::lane_id();
  const long long r = tile / G;
  const int pos = (int)(tile - r * G);
  const typename Desc<A>::word* row = desc + r * G;

  // previous round's prefix: start the load now, look at it last
  uint32_t rst = (epoch << 2) | ST_INCL;
  A rval = A(0);
  if(lane == 0) {
    if(r > 0) Desc<A>::load(rdesc + (r - 1), rst, rval);
    else if(MODE == MODE_PLAIN && seed) rval = *seed;
  }

  A lane_sum = A(0);
  bool head_found = false;
  for(int hi = pos; hi > 0 && !head_found; hi -= 32 * U) {
    uint32_t st[U];
    A val[U];
#pragma unroll
    for(int u = 0; u < U; ++u) {
      const int q = hi - 1 - (u * 32 + lane);
      st[u] = (epoch << 2) | ST_AGG;
      val[u] = A(0);
      if(q >= 0) Desc<A>::load(row + q, st[u], val[u]);
    }
    // Poll until every descriptor of the window is there, re-issuing ALL missing loads of a pass back to back: one
    // L2 round trip per pass.  (Waiting for them one after the other cost up to U + 1 dependent round trips of
    // ~1.1 us each under load — the phase trace showed look-backs of 4-8 us after the last aggregate was out.)
    for(;;) {
      bool missing = false;
#pragma unroll
      for(int u = 0; u < U; ++u) {
        const int q = hi - 1 - (u * 32 + lane);
        if(q >= 0 && (st[u] >> 2) != epoch) {
          missing = true;
          Desc<A>::load(row + q, st[u], val[u]);
        }
      }
      if(lane == 0 && r > 0 && (rst >> 2) != epoch) Desc<A>::load(rdesc + (r - 1), rst, rval);  // keep the round prefix polling too
      if(!__any_sync(SB_FULL_MASK, missing)) break;
    }
#pragma unroll
    for(int u = 0; u < U; ++u) {
      if(head_found) break;
      if(MODE != MODE_PLAIN) {
        const unsigned heads = __ballot_sync(SB_FULL_MASK, (st[u] & 3u) == ST_INCL);
        if(heads) {
          const int first = __ffs(heads) - 1;  // lowest lane = nearest predecessor with a head
          if(lane <= first) lane_sum += val[u];
          head_found = true;
          continue;
        }
      }
      lane_sum += val[u];
    }
  }
  A total = sb::warp_sum<A>(lane_sum);
  if(!head_found) {
    if(lane == 0 && r > 0) {
      while((rst >> 2) != epoch) Desc<A>::load(rdesc + (r - 1), rst, rval);
    }
    total += sb::shfl_idx(rval, 0);
  }
  return total;
}

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
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

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
#pragma unroll
        for(int c = 0; c < VEC; ++c) hf[j] |= (((w >> (8 * c)) & 0xffu) ? 1u : 0u) << c;
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
    hf[j] = 0;
#pragma unroll
    for(int c = 0; c < VEC; ++c) hf[j] |= (((w >> (8 * c)) & 0xffu) ? 1u : 0u) << c;
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
  A run = A(0);
  bool runf = false;
#pragma unroll
  for(int j = 0; j < K; ++j) {
    A s = A(0);
    bool f = false;
#pragma unroll
    for(int c = 0; c < VEC; ++c) {
      if((hf[j] >> c) & 1u) {
        s = v[j][c];
        f = true;
      } else {
        s += v[j][c];
      }
    }
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
    bool f = false;
#pragma unroll
    for(int c = 0; c < VEC; ++c) {
      if((hf[j] >> c) & 1u) {
        s = v[j][c];
        f = true;
      } else {
        s += v[j][c];
      }
    }
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

template <typename A, int MODE, int K>
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
    tile_pass1_chained<A, MODE, K>(v, hf, carry_in, open, warp_agg, warp_has_head);
    return;
  }
  constexpr int VEC = 16 / sizeof(A);
  const int lane = threadIdx.x & 31;
  A tot[K];          // row sum since the last head of the row (whole row if it has none)
  unsigned hm[K];    // lanes of row j that contain a head
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
template <typename A, int MODE, bool ALIGNED, int K>
__device__ __forceinline__ void tile_pass2_store(const A (&v)[K][16 / sizeof(A)], const unsigned (&hf)[K], const A (&carry_in)[K],
    const bool (&open)[K], A warp_prefix, size_t wbase, bool full_tile, size_t n, A* out, int exclusive) {
  constexpr int VEC = 16 / sizeof(A);
#pragma unroll
  for(int j = 0; j < K; ++j) {
    const size_t idx = wbase + (size_t)j * (32 * VEC);
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

// ---- kernel 1: one CTA per tile, tile id from an atomic ticket (unaligned buffers, small inputs) -------------
template <typename A, int MODE, bool ALIGNED>
__global__ void __launch_bounds__(SCAN_NT)
scan_kernel(const A* in, A* out, size_t n, int exclusive, size_t seg_len, size_t seg_mask, const uint8_t* __restrict__ flags,
    const A* __restrict__ seed, typename Desc<A>::word* desc, unsigned long long* ticket, unsigned long long ticket_base,
    uint32_t epoch, int use_lookback) {
  using C = TileCfg<A, SCAN_NT / 32, SCAN_K>;
  constexpr int VEC = C::VEC;
  constexpr int K = SCAN_K;
  __shared__ long long s_tile;
  __shared__ A s_wagg[C::NW];
  __shared__ int s_wflag[C::NW];
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
  const bool full_tile = base + C::TILE <= n;

  A v[K][VEC];
  unsigned hf[K];
  tile_load_global<A, ALIGNED, K>(v, in, wbase, full_tile, n);
  tile_head_flags<A, MODE, K>(hf, wbase, n, seg_len, seg_mask, flags);

  A carry_in[K];
  bool open[K];
  A warp_agg;
  bool warp_has_head;
  tile_pass1<A, MODE, K>(v, hf, carry_in, open, warp_agg, warp_has_head);
  if(lane == 0) {
    s_wagg[warp] = warp_agg;
    s_wflag[warp] = warp_has_head ? 1 : 0;
  }
  __syncthreads();
  A my_prefix, tile_agg;
  bool my_open, tile_has_head;
  combine_warps<A, MODE, C::NW>(s_wagg, s_wflag, warp, my_prefix, my_open, tile_agg, tile_has_head);

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
  tile_pass2_store<A, MODE, ALIGNED, K>(v, hf, carry_in, open, warp_prefix, wbase, full_tile, n, out, exclusive);
}

// Round look-back with every descriptor load of the round in flight at once (G <= 32 * MAXB): one L2 round
// trip when the aggregates are there, instead of one per SCAN_WINDOW x 32 descriptors.  The phase trace
// (profiles/tb_trace.py) showed 2-4.5 us per look-back with the windowed walk — several dependent L2 round
// trips at ~1 us each under streaming load — which capped the kernel at one tile per ~3.4 us and CTA.
template <typename A, int MAXB>
__device__ __forceinline__ A lookback_round_wide(const typename Desc<A>::word* desc, const typename Desc<A>::word* rdesc, long long tile,
    long long G, uint32_t epoch, const A* __restrict__ seed) {
  const int lane = sb::lane_id();
  const long long r = tile / G;
  const int pos = (int)(tile - r * G);
  const typename Desc<A>::word* row = desc + r * G;
  uint32_t rst = (epoch << 2) | ST_INCL;
  A rval = A(0);
  if(lane == 0) {
    if(r > 0) Desc<A>::load(rdesc + (r - 1), rst, rval);
    else if(seed) rval = *seed;
  }
  uint32_t st[MAXB];
  A val[MAXB];
#pragma unroll
  for(int u = 0; u < MAXB; ++u) {
    const int q = u * 32 + lane;
    st[u] = (epoch << 2) | ST_AGG;
    val[u] = A(0);
    if(q < pos) Desc<A>::load(row + q, st[u], val[u]);
  }
  // one L2 round trip per polling pass: every missing descriptor (and the round prefix) is re-requested at once
  for(;;) {
    bool missing = false;
#pragma unroll
    for(int u = 0; u < MAXB; ++u) {
      const int q = u * 32 + lane;
      if(q < pos && (st[u] >> 2) != epoch) {
        missing = true;
        Desc<A>::load(row + q, st[u], val[u]);
      }
    }
    if(lane == 0 && r > 0 && (rst >> 2) != epoch) {
      missing = true;
      Desc<A>::load(rdesc + (r - 1), rst, rval);
    }
    if(!__any_sync(SB_FULL_MASK, missing)) break;
  }
  A lane_sum = A(0);
#pragma unroll
  for(int u = 0; u < MAXB; ++u) lane_sum += val[u];
  const A total = sb::warp_sum<A>(lane_sum);
  return total + sb::shfl_idx(rval, 0);
}

// ---- kernel 2: persistent, warp-specialised, TMA ring (see header) ---------------------------------------------
// bytes of one ring stage: the data tile and, for explicit head flags, the tile's flag bytes behind it
template <typename A, int MODE, int K, int NW>
__host__ __device__ constexpr int ws_stage_bytes() {
  return TileCfg<A, NW, K>::TILE_BYTES + (MODE == MODE_FLAGS ? TileCfg<A, NW, K>::TILE : 0);
}

template <typename A, int MODE, int K, int NW>
__global__ void __launch_bounds__((NW + 2) * 32, SB_WS_MINB)
scan_ws_kernel(const A* in, A* out, size_t n, int exclusive, size_t seg_len, size_t seg_mask, const uint8_t* __restrict__ flags,
    const A* __restrict__ seed, typename Desc<A>::word* desc, uint32_t epoch, int use_lookback, long long num_tiles, int flags_in_ring) {
  using C = TileCfg<A, NW, K>;
  constexpr int VEC = C::VEC;
  constexpr int STAGE_BYTES = ws_stage_bytes<A, MODE, K, NW>();
  extern __shared__ __align__(128) unsigned char ring[];  // WS_STAGES x STAGE_BYTES
  __shared__ __align__(8) unsigned long long full_bar[WS_STAGES], empty_bar[WS_STAGES], agg_bar[WS_NSLOT], prefix_bar[WS_NSLOT];
  __shared__ A s_wagg[WS_NSLOT][NW];
  __shared__ int s_wflag[WS_NSLOT][NW];
  __shared__ A s_tprefix[WS_NSLOT];
  __shared__ int s_first_head[WS_NSLOT];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long G = gridDim.x;
  const long long my_tiles = (num_tiles - (long long)blockIdx.x + G - 1) / G;  // tiles blockIdx.x + i*G

  if(threadIdx.x == 0) {
#pragma unroll
    for(int s = 0; s < WS_STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), NW);
    }
#pragma unroll
    for(int s = 0; s < WS_NSLOT; ++s) {
      mbar_init(smem_u32(&agg_bar[s]), NW);
      mbar_init(smem_u32(&prefix_bar[s]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();  // the only block-wide barrier

  if(warp == NW) {
    // ===== producer: TMA bulk loads into the ring =====
    for(long long i = 0; i < my_tiles; ++i) {
      const int stage = (int)(i % WS_STAGES);
      if(i >= WS_STAGES) mbar_wait(smem_u32(&empty_bar[stage]), (uint32_t)((i / WS_STAGES) - 1) & 1u);
      const long long tile = (long long)blockIdx.x + i * G;
      const size_t base = (size_t)tile * C::TILE;
      const uint32_t bar = smem_u32(&full_bar[stage]);
      if(base + C::TILE <= n) {
        if(lane == 0) {
          const bool with_flags = MODE == MODE_FLAGS && flags_in_ring;
          mbar_expect_tx(bar, C::TILE_BYTES + (with_flags ? C::TILE : 0));
          bulk_load(smem_u32(ring) + stage * STAGE_BYTES, in + base, C::TILE_BYTES, bar);
          if(with_flags) bulk_load(smem_u32(ring) + stage * STAGE_BYTES + C::TILE_BYTES, flags + base, C::TILE, bar);
        }
      } else {
        // ragged last tile: nothing is staged (the workers read it from global memory with guards);
        // the arrive only completes the barrier phase the workers wait on
        if(lane == 0) mbar_arrive(bar);
      }
    }
  } else if(warp == NW + 1) {
    // ===== look-back warp =====
    if(!use_lookback) return;
    const typename Desc<A>::word* rdesc = desc + num_tiles;  // one prefix per round behind the tile descriptors
    for(long long i = 0; i < my_tiles; ++i) {
      const int slot = (int)(i % WS_NSLOT);
      const long long tile = (long long)blockIdx.x + i * G;
      mbar_wait(smem_u32(&agg_bar[slot]), (uint32_t)(i / WS_NSLOT) & 1u);
      A my_prefix, tile_agg;
      bool my_open, tile_has_head;
      combine_warps<A, MODE, NW>(s_wagg[slot], s_wflag[slot], -1, my_prefix, my_open, tile_agg, tile_has_head);
      if(lane == 0) Desc<A>::store(desc + tile, (epoch << 2) | (tile_has_head ? ST_INCL : ST_AGG), tile_agg);
      A p = A(0);
      const bool starts_on_head = MODE != MODE_PLAIN && s_first_head[slot];
      if(!starts_on_head) p = lookback_round<A, MODE>(desc, rdesc, tile, G, epoch, seed);
      if(lane == 0) {
        if((tile + 1) % G == 0)
          Desc<A>::store(const_cast<typename Desc<A>::word*>(rdesc) + tile / G, (epoch << 2) | ST_INCL, tile_has_head ? tile_agg : p + tile_agg);
        s_tprefix[slot] = p;
        mbar_arrive(smem_u32(&prefix_bar[slot]));
      }
    }
  } else {
    if(!use_lookback) {
      // ===== workers, single phase: every tile starts on a head, so a tile needs nothing from other tiles.
      // The agg mbarrier doubles as the workers' barrier for the warp pairs of the tile.
      for(long long it = 0; it < my_tiles; ++it) {
        const int stage = (int)(it % WS_STAGES), slot = (int)(it % WS_NSLOT);
        const long long tile = (long long)blockIdx.x + it * G;
        const size_t base = (size_t)tile * C::TILE;
        const size_t wbase = base + (size_t)warp * C::WCHUNK + (size_t)lane * VEC;
        const bool full_tile = base + C::TILE <= n;
        A v[K][VEC];
        unsigned hf[K];
        const bool ring_flags = MODE == MODE_FLAGS && flags_in_ring && full_tile;
        if(!ring_flags) tile_head_flags<A, MODE, K>(hf, wbase, n, seg_len, seg_mask, flags);  // flag loads fly while we wait
        mbar_wait(smem_u32(&full_bar[stage]), (uint32_t)(it / WS_STAGES) & 1u);
        if(ring_flags) tile_head_flags_ring<A, K>(hf, ring + (size_t)stage * STAGE_BYTES + C::TILE_BYTES + (size_t)warp * C::WCHUNK, wbase);
        if(full_tile) tile_load_smem<A, K>(v, ring + (size_t)stage * STAGE_BYTES + (size_t)warp * K * 512);
        else tile_load_global<A, true, K>(v, in, wbase, false, n);  // ragged last tile
        A carry_in[K];
        bool open[K];
        A warp_agg;
        bool warp_has_head;
        tile_pass1<A, MODE, K>(v, hf, carry_in, open, warp_agg, warp_has_head);
        if(lane == 0) {
          mbar_arrive(smem_u32(&empty_bar[stage]));  // the tile lives in registers now
          s_wagg[slot][warp] = warp_agg;
          s_wflag[slot][warp] = warp_has_head ? 1 : 0;
          mbar_arrive(smem_u32(&agg_bar[slot]));
        }
        A my_prefix = A(0), tile_agg;
        bool my_open = true, tile_has_head;
        if(!(MODE != MODE_PLAIN && (sb::shfl_idx(hf[0], 0) & 1u))) {  // a warp chunk that starts on a head needs no one
          mbar_wait(smem_u32(&agg_bar[slot]), (uint32_t)(it / WS_NSLOT) & 1u);
          combine_warps<A, MODE, NW>(s_wagg[slot], s_wflag[slot], warp, my_prefix, my_open, tile_agg, tile_has_head);
        }
        tile_pass2_store<A, MODE, true, K>(v, hf, carry_in, open, my_prefix, wbase, full_tile, n, out, exclusive);
      }
      return;
    }
    // ===== workers: phase A runs WS_DEPTH tiles ahead of phase B =====
    for(long long it = 0; it < my_tiles + WS_DEPTH; ++it) {
      if(it < my_tiles) {
        // ---- phase A: tile -> per-warp (sum, has-head) for the look-back warp
        const int stage = (int)(it % WS_STAGES), slot = (int)(it % WS_NSLOT);
        const long long tile = (long long)blockIdx.x + it * G;
        const size_t wbase = (size_t)tile * C::TILE + (size_t)warp * C::WCHUNK + (size_t)lane * VEC;
        const bool full_tile = (size_t)(tile + 1) * C::TILE <= n;
        A v[K][VEC];
        unsigned hf[K];
        const bool ring_flags = MODE == MODE_FLAGS && flags_in_ring && full_tile;
        if(!ring_flags) tile_head_flags<A, MODE, K>(hf, wbase, n, seg_len, seg_mask, flags);  // flag loads fly while we wait
        mbar_wait(smem_u32(&full_bar[stage]), (uint32_t)(it / WS_STAGES) & 1u);
        if(ring_flags) tile_head_flags_ring<A, K>(hf, ring + (size_t)stage * STAGE_BYTES + C::TILE_BYTES + (size_t)warp * C::WCHUNK, wbase);
        if(full_tile) tile_load_smem<A, K>(v, ring + (size_t)stage * STAGE_BYTES + (size_t)warp * K * 512);
        else tile_load_global<A, true, K>(v, in, wbase, false, n);  // ragged last tile
        A agg;
        bool has_head;
        warp_reduce_rows<A, MODE, K>(v, hf, agg, has_head);
        if(lane == 0) {
          s_wagg[slot][warp] = agg;
          s_wflag[slot][warp] = has_head ? 1 : 0;
          if(warp == 0) s_first_head[slot] = (int)(hf[0] & 1u);
          mbar_arrive(smem_u32(&agg_bar[slot]));
        }
      }
      if(it >= WS_DEPTH) {
        // ---- phase B: scan the tile with its prefix, store, release the stage
        const long long b = it - WS_DEPTH;
        const int stage = (int)(b % WS_STAGES), slot = (int)(b % WS_NSLOT);
        const long long tile = (long long)blockIdx.x + b * G;
        const size_t base = (size_t)tile * C::TILE;
        const size_t wbase = base + (size_t)warp * C::WCHUNK + (size_t)lane * VEC;
        const bool full_tile = base + C::TILE <= n;
        A v[K][VEC];
        unsigned hf[K];
        if(MODE == MODE_FLAGS && flags_in_ring && full_tile)
          tile_head_flags_ring<A, K>(hf, ring + (size_t)stage * STAGE_BYTES + C::TILE_BYTES + (size_t)warp * C::WCHUNK, wbase);
        else
          tile_head_flags<A, MODE, K>(hf, wbase, n, seg_len, seg_mask, flags);
        if(full_tile) tile_load_smem<A, K>(v, ring + (size_t)stage * STAGE_BYTES + (size_t)warp * K * 512);
        else tile_load_global<A, true, K>(v, in, wbase, false, n);  // ragged last tile
        A carry_in[K];
        bool open[K];
        A warp_agg;
        bool warp_has_head;
        tile_pass1<A, MODE, K>(v, hf, carry_in, open, warp_agg, warp_has_head);
        // The tile lives in registers now (pass 1 consumed every ring read): hand the stage back BEFORE
        // waiting for the prefix, so that the producer can refill it during the look-back wait.  ncu showed
        // ~45 % of the warp samples in mbarrier wait loops with one 64 KB stage in flight per SM at most.
        __syncwarp();
        if(lane == 0) mbar_arrive(smem_u32(&empty_bar[stage]));
        // all warps' phase-A pairs of this tile (and, with look-back, the tile prefix) are complete
        const uint32_t par = (uint32_t)(b / WS_NSLOT) & 1u;
        A tile_prefix = A(0);
        if(use_lookback) {
          mbar_wait(smem_u32(&prefix_bar[slot]), par);
          tile_prefix = s_tprefix[slot];
        } else {
          mbar_wait(smem_u32(&agg_bar[slot]), par);
        }
        A my_prefix, tile_agg;
        bool my_open, tile_has_head;
        combine_warps<A, MODE, NW>(s_wagg[slot], s_wflag[slot], warp, my_prefix, my_open, tile_agg, tile_has_head);
        const A warp_prefix = my_prefix + (my_open ? tile_prefix : A(0));
        tile_pass2_store<A, MODE, true, K>(v, hf, carry_in, open, warp_prefix, wbase, full_tile, n, out, exclusive);
      }
    }
  }
}

// ---- kernel 3: independent tiles (every tile starts on a segment head): no descriptors, no look-back ------
// Same ring and worker code as the single-phase path above, but tiles are handed out IN ORDER by an
// atomic ticket, taken by the producer whenever a ring stage frees up, which keeps the DRAM access
// front compact and ordered.  A persistent copy kernel with static grid-stride assignment measures
// 5.97 TB/s on B200, the same kernel with ticketed chunks 6.95 TB/s (profiles/r01_stream_probe.txt);
// this path gained 4-10 % (f32 L=256: 6.1 -> 6.76 TB/s).  With a look-back the ticketed assignment
// LOSES (2.3-5.0 instead of 5.6-5.8 TB/s): a CTA gets a random position in every round, whereas the
// static round-robin settles into a stable stagger in which the aggregates a tile needs are almost
// always published before it asks — so the look-back kernel above keeps static assignment.
constexpr int WS_TILE_IDS = 8;  // ring of tile ids handed from the producer to the workers

template <typename A, int MODE, int K, int NW>
__global__ void __launch_bounds__((NW + 1) * 32, SB_WS_MINB)
scan_independent_kernel(const A* in, A* out, size_t n, int exclusive, size_t seg_len, size_t seg_mask, const uint8_t* __restrict__ flags,
    long long num_tiles, int flags_in_ring, unsigned long long* ticket, unsigned long long ticket_base) {
  using C = TileCfg<A, NW, K>;
  constexpr int VEC = C::VEC;
  constexpr int STAGE_BYTES = ws_stage_bytes<A, MODE, K, NW>();
  extern __shared__ __align__(128) unsigned char ring[];  // WS_STAGES x STAGE_BYTES
  __shared__ __align__(8) unsigned long long full_bar[WS_STAGES], empty_bar[WS_STAGES], agg_bar[WS_NSLOT];
  __shared__ A s_wagg[WS_NSLOT][NW];
  __shared__ int s_wflag[WS_NSLOT][NW];
  __shared__ int s_tile_id[WS_TILE_IDS];  // tile of the CTA's i-th iteration, -1 = no more tiles

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if(threadIdx.x == 0) {
#pragma unroll
    for(int s = 0; s < WS_STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), NW);
    }
#pragma unroll
    for(int s = 0; s < WS_NSLOT; ++s) mbar_init(smem_u32(&agg_bar[s]), NW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if(warp == NW) {
    // ===== producer: draws tile tickets, TMA bulk loads into the ring =====
    if(lane != 0) return;
    for(int i = 0;; ++i) {
      const int stage = i % WS_STAGES;
      if(i >= WS_STAGES) mbar_wait(smem_u32(&empty_bar[stage]), (uint32_t)((i / WS_STAGES) - 1) & 1u);
      const long long tile = (long long)(atomicAdd(ticket, 1ull) - ticket_base);
      const uint32_t bar = smem_u32(&full_bar[stage]);
      if(tile >= num_tiles) {
        s_tile_id[i % WS_TILE_IDS] = -1;
        mbar_arrive(bar);  // releases the workers; they read the -1 and leave
        break;
      }
      s_tile_id[i % WS_TILE_IDS] = (int)tile;
      const size_t base = (size_t)tile * C::TILE;
      if(base + C::TILE <= n) {
        const bool with_flags = MODE == MODE_FLAGS && flags_in_ring;
        mbar_expect_tx(bar, C::TILE_BYTES + (with_flags ? C::TILE : 0));
        bulk_load(smem_u32(ring) + stage * STAGE_BYTES, in + base, C::TILE_BYTES, bar);
        if(with_flags) bulk_load(smem_u32(ring) + stage * STAGE_BYTES + C::TILE_BYTES, flags + base, C::TILE, bar);
      } else {
        mbar_arrive(bar);  // ragged last tile: read from global memory by the workers
      }
    }
    return;
  }
  // ===== workers =====
  for(int it = 0;; ++it) {
    const int stage = it % WS_STAGES, slot = it % WS_NSLOT;
    mbar_wait(smem_u32(&full_bar[stage]), (uint32_t)(it / WS_STAGES) & 1u);
    const int tile = s_tile_id[it % WS_TILE_IDS];
    if(tile < 0) break;
    const size_t base = (size_t)tile * C::TILE;
    const size_t wbase = base + (size_t)warp * C::WCHUNK + (size_t)lane * VEC;
    const bool full_tile = base + C::TILE <= n;
    A v[K][VEC];
    unsigned hf[K];
    if(MODE == MODE_FLAGS && flags_in_ring && full_tile)
      tile_head_flags_ring<A, K>(hf, ring + (size_t)stage * STAGE_BYTES + C::TILE_BYTES + (size_t)warp * C::WCHUNK, wbase);
    else
      tile_head_flags<A, MODE, K>(hf, wbase, n, seg_len, seg_mask, flags);
    if(full_tile) tile_load_smem<A, K>(v, ring + (size_t)stage * STAGE_BYTES + (size_t)warp * K * 512);
    else tile_load_global<A, true, K>(v, in, wbase, false, n);
    A carry_in[K];
    bool open[K];
    A warp_agg;
    bool warp_has_head;
    tile_pass1<A, MODE, K>(v, hf, carry_in, open, warp_agg, warp_has_head);
    if(lane == 0) {
      mbar_arrive(smem_u32(&empty_bar[stage]));  // the tile lives in registers now
      s_wagg[slot][warp] = warp_agg;
      s_wflag[slot][warp] = warp_has_head ? 1 : 0;
      mbar_arrive(smem_u32(&agg_bar[slot]));
    }
    A my_prefix = A(0), tile_agg;
    bool my_open = true, tile_has_head;
    if(!(MODE != MODE_PLAIN && (sb::shfl_idx(hf[0], 0) & 1u))) {  // a warp chunk that starts on a head needs no one
      mbar_wait(smem_u32(&agg_bar[slot]), (uint32_t)(it / WS_NSLOT) & 1u);
      combine_warps<A, MODE, NW>(s_wagg[slot], s_wflag[slot], warp, my_prefix, my_open, tile_agg, tile_has_head);
    }
    tile_pass2_store<A, MODE, true, K>(v, hf, carry_in, open, my_prefix, wbase, full_tile, n, out, exclusive);
  }
}

// ---- kernel 4: thread-blocked scan through swizzled TMA tensor tiles (global scans, 16-byte aligned) --------
// The array is viewed as a 2-D tensor of 128-byte rows.  A tile is TB_NW*32 rows; TMA (cp.async.bulk.tensor,
// SWIZZLE_128B) drops it into the ring so that chunk j of row r sits at chunk j ^ (r & 7): every worker thread
// owns ONE row = 128 consecutive bytes of the array and reads / rewrites it with eight conflict-free
// LDS.128 / STS.128.  Per tile a thread does one serial scan of its 32 (16) elements, ONE warp scan of the
// thread totals and one add per element — about 6x fewer instructions than the row-per-warp kernels above
// (ncu: 51 warp instructions per 64-bit element there, issue slots 50 % busy).  Results are written back into
// the same ring stage and leave through a TMA tensor store issued by a dedicated warp, so worker warps issue
// no global memory instruction at all.  Rows beyond the end of the array are zero-filled on load and clipped
// on store by the TMA unit: no ragged-tile code.  Warp roles: TB_NW workers, loader, storer, two look-back warps (even / odd tiles of the CTA).
#ifndef SB_TB_NW
#define SB_TB_NW 12
#endif
#ifndef SB_TB_STAGES
#define SB_TB_STAGES 3
#endif
constexpr int TB_NW = SB_TB_NW, TB_STAGES = SB_TB_STAGES;
constexpr int TB_ROWS = TB_NW * 32;          // rows (= worker threads) per tile
constexpr int TB_STAGE_BYTES = TB_ROWS * 128;
constexpr int TB_BOXES = (TB_ROWS + 255) / 256;  // TMA boxes per tile (box height is limited to 256 rows)
constexpr int TB_BOX_ROWS = TB_ROWS / TB_BOXES;
static_assert(TB_BOX_ROWS * TB_BOXES == TB_ROWS && (TB_BOX_ROWS * 128) % 1024 == 0, "tile does not split into swizzle-aligned boxes");
constexpr int TB_THREADS = (TB_NW + 4) * 32;  // workers, loader, storer, two look-back warps

// optional phase trace (-DSB_TB_TRACE): %globaltimer stamps of the first TB_TRACE_TILES tiles of every CTA
#ifdef SB_TB_TRACE
constexpr int TB_TRACE_TILES = 48, TB_TRACE_EVENTS = 16;
__device__ unsigned long long tb_trace[160][TB_TRACE_TILES][TB_TRACE_EVENTS];
__device__ __forceinline__ void tb_stamp(long long tile_i, int ev) {
  if(tile_i < TB_TRACE_TILES && blockIdx.x < 160) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    tb_trace[blockIdx.x][tile_i][ev] = t;
  }
}
#define TB_STAMP(i, ev) tb_stamp(i, ev)
#else
#define TB_STAMP(i, ev)
#endif

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tm), "r"(src), "r"(c0), "r"(c1) : "memory");
}

template <typename A>
__global__ void __launch_bounds__(TB_THREADS, 1)
scan_tb_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out, long long total_rows, int exclusive,
    const A* __restrict__ seed, typename Desc<A>::word* desc, typename Desc<A>::word* total_slot, uint32_t epoch, long long num_tiles) {
  constexpr int E = 128 / sizeof(A);   // elements per row / thread
  constexpr int VEC = 16 / sizeof(A);  // elements per 16-byte chunk
  extern __shared__ unsigned char tb_raw[];
  __shared__ __align__(8) unsigned long long full_bar[TB_STAGES], empty_bar[TB_STAGES], stored_bar, out_free_bar, agg_bar[2], prefix_bar[2];
  __shared__ A s_wagg[2][TB_NW], s_wprefix[2][TB_NW];

  const uint32_t ring = (smem_u32(tb_raw) + 1023u) & ~1023u;  // SWIZZLE_128B needs 1024-byte aligned tiles
  const uint32_t out_tile = ring + TB_STAGES * TB_STAGE_BYTES;  // staging tile of the TMA store
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long G = gridDim.x;
  const long long my_tiles = (num_tiles - (long long)blockIdx.x + G - 1) / G;

  if(threadIdx.x == 0) {
#pragma unroll
    for(int s = 0; s < TB_STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), TB_NW);
    }
    mbar_init(smem_u32(&stored_bar), TB_NW);
    mbar_init(smem_u32(&out_free_bar), 1);
#pragma unroll
    for(int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&agg_bar[s]), TB_NW);
      mbar_init(smem_u32(&prefix_bar[s]), TB_NW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();  // the only block-wide barrier

  if(warp == TB_NW) {
    // ===== loader: TMA tensor loads into the ring; a stage is free again as soon as the workers hold it in registers =====
    if(lane != 0) return;
    for(long long i = 0; i < my_tiles; ++i) {
      const int stage = (int)(i % TB_STAGES);
      if(i >= TB_STAGES) mbar_wait(smem_u32(&empty_bar[stage]), (uint32_t)((i / TB_STAGES) - 1) & 1u);
      const long long row0 = ((long long)blockIdx.x + i * G) * TB_ROWS;
      const uint32_t bar = smem_u32(&full_bar[stage]);
      int boxes = 0;
#pragma unroll
      for(int h = 0; h < TB_BOXES; ++h) boxes += (row0 + h * TB_BOX_ROWS < total_rows) ? 1 : 0;
      TB_STAMP(i, 0);  // load issue
      mbar_expect_tx(bar, (uint32_t)boxes * TB_BOX_ROWS * 128u);
#pragma unroll
      for(int h = 0; h < TB_BOXES; ++h)
        if(row0 + h * TB_BOX_ROWS < total_rows) tma_load_2d(ring + stage * TB_STAGE_BYTES + h * TB_BOX_ROWS * 128, &tm_in, 0, (int)(row0 + h * TB_BOX_ROWS), bar);
    }
  } else if(warp == TB_NW + 1 || warp == TB_NW + 3) {
    // ===== look-back warps (one for the even, one for the odd tiles of this CTA, so that consecutive look-backs
    // overlap): tile aggregate -> descriptor, round look-back, per-warp prefixes for the workers =====
    const typename Desc<A>::word* rdesc = desc + num_tiles;
    for(long long i = (warp == TB_NW + 1) ? 0 : 1; i < my_tiles; i += 2) {
      const int slot = (int)(i & 1);
      const long long tile = (long long)blockIdx.x + i * G;
      mbar_wait(smem_u32(&agg_bar[slot]), (uint32_t)(i >> 1) & 1u);
      if(lane == 0) TB_STAMP(i, 1);  // aggregates of all warps in
      const A mine = lane < TB_NW ? s_wagg[slot][lane] : A(0);
      A incl = mine;
#pragma unroll
      for(int d = 1; d < 32; d <<= 1) {
        const A up = sb::shfl_up(incl, d);
        if(lane >= d) incl += up;
      }
      const A tile_agg = sb::shfl_idx(incl, 31);
      if(lane == 0) Desc<A>::store(desc + tile, (epoch << 2) | ST_AGG, tile_agg);
      const A p = G <= 160 ? lookback_round_wide<A, 5>(desc, rdesc, tile, G, epoch, seed) : lookback_round<A, MODE_PLAIN>(desc, rdesc, tile, G, epoch, seed);
      if(lane == 0 && (tile + 1) % G == 0)
        Desc<A>::store(const_cast<typename Desc<A>::word*>(rdesc) + tile / G, (epoch << 2) | ST_INCL, p + tile_agg);
      if(lane == 0 && tile == num_tiles - 1) Desc<A>::store(total_slot, (epoch << 2) | ST_INCL, p + tile_agg);  // for the tail kernel
      if(lane == 0) TB_STAMP(i, 2);  // look-back done
      if(lane < TB_NW) {  // every writer releases its own value: the barrier counts TB_NW arrivals
        s_wprefix[slot][lane] = p + (incl - mine);
        mbar_arrive(smem_u32(&prefix_bar[slot]));
      }
    }
  } else if(warp == TB_NW + 2) {
    // ===== storer: TMA tensor store of the staging tile =====
    if(lane != 0) return;
    for(long long i = 0; i < my_tiles; ++i) {
      mbar_wait(smem_u32(&stored_bar), (uint32_t)i & 1u);
      TB_STAMP(i, 3);  // all rows written to the staging tile
      const long long row0 = ((long long)blockIdx.x + i * G) * TB_ROWS;
#pragma unroll
      for(int h = 0; h < TB_BOXES; ++h)
        if(row0 + h * TB_BOX_ROWS < total_rows) tma_store_2d(&tm_out, out_tile + h * TB_BOX_ROWS * 128, 0, (int)(row0 + h * TB_BOX_ROWS));
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the staging tile has been read
      TB_STAMP(i, 4);  // staging tile free
      mbar_arrive(smem_u32(&out_free_bar));
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all stores have landed before the CTA retires
  } else {
    // ===== workers: thread t owns row t of every tile.  Software pipeline over two register sets:
    //   phase A(k+1): ring -> registers (the stage goes straight back to the loader), serial row scan, warp scan
    //                 of the row totals, warp aggregate to the look-back warp
    //   phase B(k)  : wait for the tile prefix (it has had a whole tile time to arrive), add, write the row to
    //                 the staging tile, hand it to the storer.
    const uint32_t row_off = (uint32_t)threadIdx.x * 128u;
    const uint32_t sw = (uint32_t)(threadIdx.x & 7);
    auto phase_a = [&](long long k, A (&x)[E], A& thread_excl) {
      const int stage = (int)(k % TB_STAGES), slot = (int)(k & 1);
      const uint32_t row = ring + stage * TB_STAGE_BYTES + row_off;
      const bool valid = ((long long)blockIdx.x + k * G) * TB_ROWS + threadIdx.x < total_rows;
      if(threadIdx.x == 0) TB_STAMP(k, 5);  // phase A: start waiting for the data
      mbar_wait(smem_u32(&full_bar[stage]), (uint32_t)(k / TB_STAGES) & 1u);
      if(threadIdx.x == 0) TB_STAMP(k, 6);  // data there
#pragma unroll
      for(int j = 0; j < 8; ++j) {
        uint4 q;
        const uint32_t addr = row + (((uint32_t)j ^ sw) << 4);
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(addr) : "memory");
        const sb::Vec<A>& t = *reinterpret_cast<const sb::Vec<A>*>(&q);
#pragma unroll
        for(int c = 0; c < VEC; ++c) x[j * VEC + c] = valid ? t.v[c] : A(0);  // rows behind the end of the array count as zeros
      }
      // serial inclusive scan of the row (left to right: the order of a sequential scan)
#pragma unroll
      for(int e = 1; e < E; ++e) x[e] += x[e - 1];
      __syncwarp();
      if(lane == 0) mbar_arrive(smem_u32(&empty_bar[stage]));  // the warp's rows live in registers now
      A incl = x[E - 1];
#pragma unroll
      for(int d = 1; d < 32; d <<= 1) {
        const A up = sb::shfl_up(incl, d);
        if(lane >= d) incl += up;
      }
      thread_excl = incl - x[E - 1];
      __syncwarp();  // every lane is past its read of s_wprefix[slot] (phase B of tile k-2) before the slot is re-armed
      if(lane == 31) {
        s_wagg[slot][warp] = incl;
        mbar_arrive(smem_u32(&agg_bar[slot]));
      }
      if(threadIdx.x == 0) TB_STAMP(k, 7);  // phase A done
    };
    auto phase_b = [&](long long k, A (&x)[E], A thread_excl) {
      const int slot = (int)(k & 1);
      if(threadIdx.x == 0) TB_STAMP(k, 8);  // phase B: waiting for the prefix
      mbar_wait(smem_u32(&prefix_bar[slot]), (uint32_t)(k >> 1) & 1u);
      if(threadIdx.x == 0) TB_STAMP(k, 9);  // prefix there
      const A add = s_wprefix[slot][warp] + thread_excl;
      if(exclusive) {
#pragma unroll
        for(int e = E - 1; e > 0; --e) x[e] = x[e - 1] + add;
        x[0] = add;
      } else {
#pragma unroll
        for(int e = 0; e < E; ++e) x[e] += add;
      }
      if(k > 0) mbar_wait(smem_u32(&out_free_bar), (uint32_t)(k - 1) & 1u);  // the store of tile k-1 has read the staging tile
#pragma unroll
      for(int j = 0; j < 8; ++j) {
        sb::Vec<A> t;
#pragma unroll
        for(int c = 0; c < VEC; ++c) t.v[c] = x[j * VEC + c];
        const uint32_t addr = out_tile + row_off + (((uint32_t)j ^ sw) << 4);
        const uint4& q = *reinterpret_cast<const uint4*>(&t);
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(q.x), "r"(q.y), "r"(q.z), "r"(q.w) : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // make the rows visible to the TMA store
      __syncwarp();
      if(threadIdx.x == 0) TB_STAMP(k, 10);  // phase B done
      if(lane == 0) mbar_arrive(smem_u32(&stored_bar));
    };
    A xa[E], xb[E];
    A ea = A(0), eb = A(0);
    if(my_tiles > 0) phase_a(0, xa, ea);
    for(long long k = 0; k < my_tiles; k += 2) {
      if(k + 1 < my_tiles) phase_a(k + 1, xb, eb);
      phase_b(k, xa, ea);
      if(k + 1 < my_tiles) {
        if(k + 2 < my_tiles) phase_a(k + 2, xa, ea);
        phase_b(k + 1, xb, eb);
      }
    }
  }
}

// elements behind the last full 128-byte row (fewer than 128 / sizeof(A)); runs after scan_tb_kernel on the stream
template <typename A>
__global__ void scan_tb_tail_kernel(const A* in, A* out, size_t n_main, size_t n, int exclusive, const typename Desc<A>::word* total_slot) {
  if(threadIdx.x != 0 || blockIdx.x != 0) return;
  uint32_t st;
  A run;
  Desc<A>::load(total_slot, st, run);  // seed + sum of the first n_main elements (in and out may be the same buffer)
  for(size_t i = n_main; i < n; ++i) {
    const A x = in[i];
    if(exclusive) out[i] = run;
    run += x;
    if(!exclusive) out[i] = run;
  }
}

// co-resident CTAs per SM of a persistent kernel instantiation.  The shared-memory opt-in is a per-device
// function attribute, so the query (and the opt-in) is made once per device, on the current device.
constexpr int SB_MAX_DEVICES = 64;
struct OccupancyCache {
  std::atomic<int> per_sm[SB_MAX_DEVICES];  // 0: not queried yet, -1: does not fit
};
template <typename Kern>
int resident_ctas_per_sm(OccupancyCache& cache, int device, Kern kern, int threads, size_t smem) {
  if(device < 0 || device >= SB_MAX_DEVICES) return 0;
  int nb = cache.per_sm[device].load(std::memory_order_acquire);
  if(nb != 0) return nb;
  if(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, smem) != cudaSuccess || nb <= 0) {
    (void)cudaGetLastError();
    nb = -1;
  }
  cache.per_sm[device].store(nb, std::memory_order_release);
  return nb;
}

// rows per worker warp: 8 (64 KB tiles); 6 with explicit flags so that 3 stages of data + flag bytes fit in 227 KB
template <int MODE> struct ws_rows { static constexpr int K = MODE == MODE_FLAGS ? (WS_K * 3) / 4 : WS_K; };

template <typename A, int MODE, int NW>
int launch_ws(sb_ctx* ctx, const A* in, A* out, size_t n, int exclusive, size_t seg_len, size_t seg_mask, const uint8_t* flags,
    const A* seed, typename Desc<A>::word* desc, uint32_t epoch, int use_lookback, size_t tiles) {
  constexpr int K = ws_rows<MODE>::K;
  constexpr size_t smem = (size_t)WS_STAGES * ws_stage_bytes<A, MODE, K, NW>();
  constexpr int WS_THREADS = (NW + 2) * 32;
  static OccupancyCache ws_cache, ind_cache;
  const int per_sm = resident_ctas_per_sm(ws_cache, ctx->device, scan_ws_kernel<A, MODE, K, NW>, WS_THREADS, smem);
  if(per_sm <= 0) return SB_ERR_WORKSPACE;
  size_t grid = (size_t)ctx->sm_count * per_sm;
  if(grid > tiles) grid = tiles;
  int flags_in_ring = (MODE == MODE_FLAGS && sb::aligned_to(flags, 16)) ? 1 : 0;  // TMA needs 16-byte aligned sources
  long long num_tiles = (long long)tiles;
  if(!use_lookback) {
    const int ind_per_sm = resident_ctas_per_sm(ind_cache, ctx->device, scan_independent_kernel<A, MODE, K, NW>, (NW + 1) * 32, smem);
    if(ind_per_sm <= 0) return SB_ERR_WORKSPACE;
    size_t g = (size_t)ctx->sm_count * ind_per_sm;
    if(g > tiles) g = tiles;
    // every CTA's producer draws tickets until one is past the last tile: tiles + grid tickets per launch
    const unsigned long long tbase = ctx->scan_ticket_base;
    ctx->scan_ticket_base += tiles + g;
    scan_independent_kernel<A, MODE, K, NW><<<(unsigned)g, (NW + 1) * 32, smem, ctx->stream>>>(in, out, n, exclusive, seg_len, seg_mask, flags,
        num_tiles, flags_in_ring, ctx->scan_ticket, tbase);
    return SB_OK;
  }
  // Cooperative launch: the runtime guarantees that all CTAs of the grid are co-resident (or refuses the
  // launch) — the forward-progress condition of the static round-robin look-back.  No grid barrier is used.
  void* args[] = {(void*)&in, (void*)&out, (void*)&n, (void*)&exclusive, (void*)&seg_len, (void*)&seg_mask, (void*)&flags, (void*)&seed,
      (void*)&desc, (void*)&epoch, (void*)&use_lookback, (void*)&num_tiles, (void*)&flags_in_ring};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)scan_ws_kernel<A, MODE, K, NW>, dim3((unsigned)grid), dim3(WS_THREADS), args, smem, ctx->stream);
  if(e != cudaSuccess) return (int)e;
  return SB_OK;
}

// ---- host side of kernel 4: tensor maps + cooperative launch ---------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    (void)cudaGetLastError();
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
// [rows][128 bytes] view of a 16-byte aligned array, box = 256 rows, 128-byte swizzle
bool make_row_map(CUtensorMap* tm, const void* base, unsigned long long rows) {
  EncodeTiledFn enc = tensor_map_encoder();
  if(!enc) return false;
  const cuuint64_t dims[2] = {32, rows};          // 32 x u32 = 128 bytes per row
  const cuuint64_t strides[1] = {128};            // bytes between rows
  const cuuint32_t box[2] = {32, TB_BOX_ROWS};
  const cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// returns SB_ERR_UNSUPPORTED when the tensor-map path is not available (caller falls back to scan_ws_kernel)
template <typename A>
int launch_tb(sb_ctx* ctx, const A* in, A* out, size_t n, int exclusive, const A* seed, typename Desc<A>::word* desc, uint32_t epoch, size_t tiles) {
  constexpr size_t E = 128 / sizeof(A);
  const unsigned long long rows = n / E;
  if(rows == 0 || rows >= 0xffffffffull) return SB_ERR_UNSUPPORTED;
  alignas(64) CUtensorMap tm_in, tm_out;
  if(!make_row_map(&tm_in, in, rows) || !make_row_map(&tm_out, out, rows)) return SB_ERR_UNSUPPORTED;
  constexpr size_t smem = (size_t)(TB_STAGES + 1) * TB_STAGE_BYTES + 1024;  // ring + staging tile + alignment slack
  static OccupancyCache cache;
  const int per_sm = resident_ctas_per_sm(cache, ctx->device, scan_tb_kernel<A>, TB_THREADS, smem);
  if(per_sm <= 0) return SB_ERR_UNSUPPORTED;
  const long long num_tiles = (long long)((rows + TB_ROWS - 1) / TB_ROWS);
  size_t grid = (size_t)ctx->sm_count;  // one CTA per SM (launch bounds)
  if((long long)grid > num_tiles) grid = (size_t)num_tiles;
  long long total_rows = (long long)rows;
  typename Desc<A>::word* total_slot = desc + 2 * tiles + 32;  // scratch behind the tile and round descriptors
  void* args[] = {(void*)&tm_in, (void*)&tm_out, (void*)&total_rows, (void*)&exclusive, (void*)&seed, (void*)&desc, (void*)&total_slot, (void*)&epoch, (void*)&num_tiles};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)scan_tb_kernel<A>, dim3((unsigned)grid), dim3(TB_THREADS), args, smem, ctx->stream);
  if(e != cudaSuccess) return (int)e;
  const size_t n_main = (size_t)rows * E;
  if(n_main < n) {
    ctx->launches++;
    scan_tb_tail_kernel<A><<<1, 32, 0, ctx->stream>>>(in, out, n_main, n, exclusive, total_slot);
  }
  return SB_OK;
}

template <typename T>
int scan_launch(sb_ctx* ctx, const T* in, T* out, size_t n, int exclusive, int mode, size_t seg_len, const uint8_t* flags, const T* seed) {
  using A = typename sb_arith<T>::type;
  using W = typename Desc<A>::word;
  if(!ctx || (n && (!in || !out))) return SB_ERR_BAD_ARG;
  if(!sb::aligned_to(in, sizeof(T)) || !sb::aligned_to(out, sizeof(T)) || !sb::aligned_to(seed, sizeof(T))) return SB_ERR_MISALIGNED;
  if(n == 0) return SB_OK;
  sb_device_guard guard(ctx->device);

  // persistent kernel for 16-byte aligned buffers with at least a few tiles; ticketed kernel otherwise
  const bool aligned = sb::aligned_to(in, 16) && sb::aligned_to(out, 16);
  // 18 worker warps (72 KB tiles) where they measured faster: global scans, whose rate is tile bytes per look-back
  // round, and 32-bit independent tiles; 16 warps (64 KB) everywhere else
  constexpr size_t wide_tile = TileCfg<A, WS_NW_WIDE, WS_K>::TILE;
  const bool wide = mode == MODE_PLAIN || (mode == MODE_IMPLICIT && sizeof(A) == 4 && sb::is_pow2(seg_len) && wide_tile % seg_len == 0);
  // 64-bit independent tiles: 15 workers, so that the CTA is 16 warps and a thread may use 128 registers — at 17
  // warps ptxas has 96 and spills (fp64 group scan 5.5 -> 6.4 TB/s)
  constexpr size_t narrow_tile = TileCfg<A, WS_NW_NARROW, WS_K>::TILE;
  const bool narrow = mode == MODE_IMPLICIT && sizeof(A) == 8 && sb::is_pow2(seg_len) && narrow_tile % seg_len == 0;
  const size_t ws_tile = mode == MODE_FLAGS ? TileCfg<A, WS_NW_FLAGS, ws_rows<MODE_FLAGS>::K>::TILE
                         : wide ? wide_tile : narrow ? narrow_tile : TileCfg<A, WS_NW, WS_K>::TILE;
  const size_t simple_tile = TileCfg<A, SCAN_NT / 32, SCAN_K>::TILE;
  const bool use_ws = aligned && n >= 4 * ws_tile;
  const size_t tile = use_ws ? ws_tile : simple_tile;
  const size_t tiles = (n + tile - 1) / tile;
  if(tiles > 0x7fffffffull) return SB_ERR_BAD_ARG;

  size_t seg_mask = ~(size_t)0;
  int use_lookback = 1;
  if(mode == MODE_IMPLICIT && sb::is_pow2(seg_len)) {
    seg_mask = seg_len - 1;
    if(tile % seg_len == 0) use_lookback = 0;  // every tile starts on a head
  }
  if(tiles == 1 && !(mode == MODE_PLAIN && seed)) use_lookback = 0;

  W* desc = nullptr;
  uint32_t epoch = 0;
  unsigned long long tbase = 0;
  if(use_lookback) {
    const size_t ndesc = 2 * tiles + 64;  // tile descriptors + (persistent kernel) one prefix per round
    int rc = sb_ctx_reserve_desc(ctx, sizeof(A) == 4 ? ndesc : 0, sizeof(A) == 8 ? ndesc : 0);
    if(rc != SB_OK) return rc;
    if(ctx->scan_epoch >= (1u << 30) - 1) {
      if(ctx->desc64) SB_CUDA_TRY(cudaMemsetAsync(ctx->desc64, 0, ctx->desc64_count * 8, ctx->stream));
      if(ctx->desc128) SB_CUDA_TRY(cudaMemsetAsync(ctx->desc128, 0, ctx->desc128_count * 16, ctx->stream));
      ctx->scan_epoch = 0;
    }
    epoch = ++ctx->scan_epoch;
    desc = reinterpret_cast<W*>(sizeof(A) == 4 ? ctx->desc64 : ctx->desc128);
    if(!use_ws) {
      tbase = ctx->scan_ticket_base;
      ctx->scan_ticket_base += tiles;
    }
  }
  const A* ain = reinterpret_cast<const A*>(in);
  A* aout = reinterpret_cast<A*>(out);
  const A* aseed = reinterpret_cast<const A*>(seed);
  if(use_ws) {
    int rc;
    // global scans: SB_SCAN_TB=1 selects the experimental thread-blocked kernel on swizzled TMA tensor tiles
    // (parity-green, currently slower than the row-per-warp kernel: DESIGN.md section 4.4)
    static const bool tb_enabled = getenv("SB_SCAN_TB") && getenv("SB_SCAN_TB")[0] == '1';
    if(mode == MODE_PLAIN && use_lookback && tb_enabled) {
      rc = launch_tb<A>(ctx, ain, aout, n, exclusive, aseed, desc, epoch, tiles);
      if(rc == SB_OK) {
        SB_CHECK_LAUNCH(ctx);
        return SB_OK;
      }
      if(rc != SB_ERR_UNSUPPORTED) return rc;
    }
    if(mode == MODE_PLAIN) rc = launch_ws<A, MODE_PLAIN, WS_NW_WIDE>(ctx, ain, aout, n, exclusive, seg_len, seg_mask, flags, aseed, desc, epoch, use_lookback, tiles);
    else if(mode == MODE_IMPLICIT && wide) rc = launch_ws<A, MODE_IMPLICIT, WS_NW_WIDE>(ctx, ain, aout, n, exclusive, seg_len, seg_mask, flags, aseed, desc, epoch, use_lookback, tiles);
    else if(mode == MODE_IMPLICIT && narrow) rc = launch_ws<A, MODE_IMPLICIT, WS_NW_NARROW>(ctx, ain, aout, n, exclusive, seg_len, seg_mask, flags, aseed, desc, epoch, use_lookback, tiles);
    else if(mode == MODE_IMPLICIT) rc = launch_ws<A, MODE_IMPLICIT, WS_NW>(ctx, ain, aout, n, exclusive, seg_len, seg_mask, flags, aseed, desc, epoch, use_lookback, tiles);
    else rc = launch_ws<A, MODE_FLAGS, WS_NW_FLAGS>(ctx, ain, aout, n, exclusive, seg_len, seg_mask, flags, aseed, desc, epoch, use_lookback, tiles);
    if(rc != SB_OK) return rc;
    SB_CHECK_LAUNCH(ctx);
    return SB_OK;
  }
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

#ifdef SB_TB_TRACE
extern "C" __attribute__((visibility("default"))) int sb_tb_trace_dump(unsigned long long* dst) {
  return (int)cudaMemcpyFromSymbol(dst, tb_trace, sizeof(tb_trace));
}
#endif

size_t sb_scan_tile_bytes() { return (size_t)WS_NW * WS_K * 512; }

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
