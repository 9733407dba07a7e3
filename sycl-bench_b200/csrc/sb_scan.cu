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
//  scan_independent_kernel: the same ring and worker code without descriptors and look-back warp, tiles handed out in
//    order by an atomic ticket — whenever every tile starts on a segment head (implicit heads with a power-of-two
//    segment that divides the tile; explicit flags after flags_tile_heads_kernel has said so on the device).
//
//  Tile shapes (workers x rows of 512 B, chosen per case from measurements; see the knobs below): plain scans 16 x 9
//    (32-bit) / 14 x 10 (64-bit); implicit heads 18 x 8 or 15 / 16 x 8 (independent tiles), 18 x 8 / 14 x 8 (look-back);
//    explicit flags 14 x 8 with the tile's flag bytes in the same ring stage.
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
#include <atomic>
#include <cstdlib>

#include "sb_scan_tile.cuh"

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
// Global (plain) scans: the tile stays ~70 KB — their rate is tile bytes per look-back round — but its shape is chosen per element
// size from a sweep at 2^28 / 2^30 (profiles/r02_ops_v21_plain_tile_shapes.jsonl; default before: 18 workers x 8 rows for all):
//   32-bit: 16 workers x 9 rows (72 KB, 18 warps, 96 registers):  fp32 6.15 -> 6.37, int32 6.36 -> 6.34 (2^30: 6.52 -> 6.62) TB/s
//   64-bit: 14 workers x 10 rows (70 KB, 16 warps, 128 registers): int64 6.23 -> 6.38, fp64 6.24 -> 6.32 (2^30 int64: 6.34 -> 6.42)
//   (16 x 9 loses 7 % for 64-bit types — spills at 96 registers —, 14 x 10 loses 1-4 % for 32-bit, 12 x 12 loses everywhere)
#ifndef SB_WS_LB32_SHAPE
#define SB_WS_LB32_SHAPE 0  // 1: 32-bit implicit heads on the look-back path take 14 workers x 10 rows (128 registers) instead of 18 x 8:
                            // measured 15-23 % SLOWER (L = 65536 int32 5.56 -> 4.28 TB/s), with either form of phase A
                            // (profiles/r02_ops_v24_lb32_14x10_ab.jsonl) — what helps 64-bit types and the two-phase kernel does not carry over
#endif
#ifndef SB_WS_PLAIN_SHAPE
#define SB_WS_PLAIN_SHAPE 1  // 0: 18 x 8 for every type
#endif
template <int BYTES> struct plain_shape {
  static constexpr int NW = !SB_WS_PLAIN_SHAPE ? SB_WS_NW_WIDE : BYTES == 8 ? 14 : 16;
  static constexpr int K = !SB_WS_PLAIN_SHAPE ? SB_WS_K : BYTES == 8 ? 10 : 9;
};
#ifndef SB_WS_NW_NARROW
#define SB_WS_NW_NARROW 15  // 15 workers + producer = 16 warps: 128 registers per thread for 64-bit independent tiles
#endif
constexpr int WS_NW_NARROW = SB_WS_NW_NARROW;
#ifndef SB_WS_NW_FLAGS
#define SB_WS_NW_FLAGS 14  // worker warps of the explicit-flag instantiations: 14 + producer + look-back = 16 warps, 128 registers
#endif
constexpr int WS_NW_FLAGS = SB_WS_NW_FLAGS;
// Measured and NOT kept for this kernel (profiles/r02_ops_v6_wprefix_ab.jsonl, r02_ops_v8_totals_ab.jsonl):
//  * per-warp prefixes computed once by the look-back warp instead of every worker lane walking the NW warp sums: 15 % fewer
//    worker instructions for 64-bit types, but the two extra warp scans sit on the inter-CTA hand-over (aggregate ->
//    descriptor, prefix -> workers) and the kernel LOSES 10-17 % (i32 6.30 -> 5.17, i64 6.18 -> 5.56 TB/s);
//  * the two-phase kernel's independent round totals (leader sums a round, every look-back warp keeps a running sum) instead
//    of the chained round prefix: 5.7-5.9 against 6.26-6.33 TB/s — with the tile waiting in shared memory phase A is only ONE
//    tile ahead, and the extra hop (aggregate -> leader -> total -> look-back) does not fit into one round of look-ahead.
// Both say the same: the hand-over chain, not the issue slots, bounds this kernel.
//  * tile descriptors published ahead of the warp's own look-back (what lifted the two-phase kernel by 10 %): neutral here
//    (2^28: i32 6.34 / 6.34, i64 6.08 / 6.24; 2^30: 6.59 / 6.48 TB/s with / without, profiles/r02_ops_v9_publish_ahead_ab.jsonl) —
//    phase A is one tile ahead, there is next to nothing to publish early.  Off; -DSB_WS_PUBLISH_AHEAD=1 builds it.
#ifndef SB_WS_PUBLISH_AHEAD
#define SB_WS_PUBLISH_AHEAD 0
#endif
#ifndef SB_WS_NW_LB64
#define SB_WS_NW_LB64 14  // 64-bit types, segmented modes with look-back: 14 + 2 = 16 warps -> 128 registers per thread
#endif
constexpr int WS_NW_LB64 = SB_WS_NW_LB64;
#ifndef SB_WS_K_LB64
#define SB_WS_K_LB64 8  // rows per warp of that instantiation with implicit heads.  10 rows (70 KB tiles, 3 stages = 210 KB) measured
                        // -1 ... -7 % except L = 2^20 (+3 %): not the tile size (profiles/r02_ops_v20_lb64_k10_ab.jsonl)
#endif
constexpr int WS_K_LB64 = SB_WS_K_LB64;
#ifndef SB_SCAN_SEG_EARLY
#define SB_SCAN_SEG_EARLY 1  // explicit flags, 32-bit: the look-back waits only for the descriptors up to the nearest head
#endif
constexpr int WS_NSLOT = 2 * WS_DEPTH + 2;  // a fast warp is at most DEPTH+1 tiles ahead of a slow one

using namespace sbscan;

// classic decoupled look-back by one full warp: sum the aggregates of the predecessors down to the
// first tile whose inclusive prefix is known (ticketed kernel)
// (every polling loop carries a deadline: a CTA that waits for a descriptor that never comes raises the ctx error flag and
// carries on with what it has — sb_ctx_sync reports SB_ERR_TIMEOUT — instead of hanging the device)
template <typename A>
__device__ __forceinline__ A lookback(const typename Desc<A>::word* desc, long long tile, uint32_t epoch, Deadline& dl, unsigned* err) {
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
      for(;;) {
        const bool missing = pred - 32 * w >= 0 && (st[w] >> 2) != epoch;
        if(missing) Desc<A>::load(desc + (pred - 32 * w), st[w], val[w]);
        if(!__any_sync(SB_FULL_MASK, missing && (st[w] >> 2) != epoch)) break;
        if(__any_sync(SB_FULL_MASK, dl.expired())) {
          if(lane == 0) *err = 1u;
          return prefix;
        }
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
template <typename A, int MODE, typename Idle>
__device__ __forceinline__ A lookback_round(const typename Desc<A>::word* desc, const typename Desc<A>::word* rdesc, long long tile,
    long long G, uint32_t epoch, const A* __restrict__ seed, Deadline& dl, unsigned* err, Idle on_idle) {
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
    // Explicit flags, 32-bit types (SB_SCAN_SEG_EARLY): the walk ends at the NEAREST predecessor that holds a head, so only
    // the descriptors from there to this tile are waited for — with heads in (almost) every tile a tile depends on one
    // neighbour instead of on every tile in front of it in the round.  Measured at 2^28 with heads at random positions,
    // mean distance 256 (profiles/r02_ops_v12_seg_early_ab.jsonl): int32 4.20 -> 4.36, fp32 3.85 -> 4.15 TB/s; but 64-bit
    // types LOSE 8-11 % (3.11 -> 2.85) and implicit heads (L = 65536 ... 2^20) 0-6 %: those kernels are bound by their
    // instruction count, not by the descriptor chain, and the two extra ballots per window sit in the look-back warp's
    // critical path.  Hence only where it measured faster.
    for(;;) {
      bool missing = false;
      if(MODE == MODE_FLAGS && sizeof(A) == 4 && SB_SCAN_SEG_EARLY) {
        int stop_u = U;  // window of the nearest head seen so far (warp-uniform)
#pragma unroll
        for(int u = 0; u < U; ++u) {
          if(u > stop_u) continue;
          const int q = hi - 1 - (u * 32 + lane);
          const bool present = q < 0 || (st[u] >> 2) == epoch;
          const unsigned heads = __ballot_sync(SB_FULL_MASK, q >= 0 && present && (st[u] & 3u) == ST_INCL);
          unsigned need = 0xffffffffu;  // lanes of this window whose descriptors the walk needs (lane 0 = nearest)
          if(heads) {
            need = 0xffffffffu >> (32 - __ffs(heads));
            stop_u = u;
          }
          if(!present && ((need >> lane) & 1u)) {
            missing = true;
            Desc<A>::load(row + q, st[u], val[u]);
          }
        }
      } else {
#pragma unroll
        for(int u = 0; u < U; ++u) {
          const int q = hi - 1 - (u * 32 + lane);
          if(q >= 0 && (st[u] >> 2) != epoch) {
            missing = true;
            Desc<A>::load(row + q, st[u], val[u]);
          }
        }
      }
      if(lane == 0 && r > 0 && (rst >> 2) != epoch) Desc<A>::load(rdesc + (r - 1), rst, rval);  // keep the round prefix polling too
      if(!__any_sync(SB_FULL_MASK, missing)) break;
      on_idle();  // (publishes this CTA's next aggregate if its phase A has finished meanwhile)
      if(__any_sync(SB_FULL_MASK, dl.expired())) {
        if(lane == 0) *err = 1u;
        break;
      }
    }
#pragma unroll
    for(int u = 0; u < U; ++u) {
      if(head_found) break;
      if(MODE != MODE_PLAIN) {
        // (the epoch test matters with SB_SCAN_SEG_EARLY: descriptors behind the nearest head may still be stale)
        const unsigned heads = __ballot_sync(SB_FULL_MASK, (st[u] >> 2) == epoch && (st[u] & 3u) == ST_INCL);
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
    for(;;) {
      const bool missing = lane == 0 && r > 0 && (rst >> 2) != epoch;
      if(missing) Desc<A>::load(rdesc + (r - 1), rst, rval);
      if(!__any_sync(SB_FULL_MASK, missing && (rst >> 2) != epoch)) break;
      on_idle();
      if(__any_sync(SB_FULL_MASK, dl.expired())) {
        if(lane == 0) *err = 1u;
        break;
      }
    }
    total += sb::shfl_idx(rval, 0);
  }
  return total;
}

// ---- kernel 1: one CTA per tile, tile id from an atomic ticket (unaligned buffers, small inputs) -------------
template <typename A, int MODE, bool ALIGNED>
__global__ void __launch_bounds__(SCAN_NT)
scan_kernel(const A* in, A* out, size_t n, int exclusive, size_t seg_len, size_t seg_mask, const uint8_t* __restrict__ flags,
    const A* __restrict__ seed, typename Desc<A>::word* desc, unsigned long long* ticket, unsigned long long ticket_base,
    uint32_t epoch, int use_lookback, unsigned* err, unsigned long long timeout_ns) {
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
          Deadline dl(timeout_ns);
          p = lookback<A>(desc, tile, epoch, dl, err);
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

// ---- kernel 2: persistent, warp-specialised, TMA ring (see header) ---------------------------------------------
// bytes of one ring stage: the data tile and, for explicit head flags, the tile's flag bytes behind it
template <typename A, int MODE, int K, int NW>
__host__ __device__ constexpr int ws_stage_bytes() {
  return TileCfg<A, NW, K>::TILE_BYTES + (MODE == MODE_FLAGS ? TileCfg<A, NW, K>::TILE : 0);
}

template <typename A, int MODE, int K, int NW>
__global__ void __launch_bounds__((NW + 2) * 32, SB_WS_MINB)
scan_ws_kernel(const A* in, A* out, size_t n, int exclusive, size_t seg_len, size_t seg_mask, const uint8_t* __restrict__ flags,
    const A* __restrict__ seed, typename Desc<A>::word* desc, uint32_t epoch, int use_lookback, long long num_tiles, int flags_in_ring,
    unsigned* err, unsigned long long timeout_ns, const unsigned long long* gate, unsigned long long gate_seq, int gate_want) {
  using C = TileCfg<A, NW, K>;
  constexpr int VEC = C::VEC;
  constexpr int STAGE_BYTES = ws_stage_bytes<A, MODE, K, NW>();
  extern __shared__ __align__(128) unsigned char ring[];  // WS_STAGES x STAGE_BYTES
  __shared__ __align__(8) unsigned long long full_bar[WS_STAGES], empty_bar[WS_STAGES], agg_bar[WS_NSLOT], prefix_bar[WS_NSLOT];
  if(gate && (int)(*gate == gate_seq) != gate_want) return;  // explicit flags: the other kernel of the pair runs (flags_tile_heads_kernel)
  // (Measured and removed: SINGLE-PASS tiles for dense explicit flags — no phase A, the warp sums come out of pass 1 of phase B
  // and the tile waits in registers for a look-back that ends at the neighbouring tile.  One pass over the ring instead of
  // two, yet 18-25 % SLOWER with random heads at mean distance 32 ... 1024 (int32 4.34 -> 3.25, int64 3.08 -> 2.53 TB/s,
  // profiles/r02_ops_v15_single_pass_ab.jsonl): pass 1 -> publish -> look-back -> pass 2 is then a serial chain per tile with
  // the workers idle in between, whereas phase A of the next tile overlaps it in the two-phase form.)
  __shared__ A s_wagg[WS_NSLOT][NW];
  __shared__ int s_wflag[WS_NSLOT][NW];
  __shared__ A s_tprefix[WS_NSLOT];
  __shared__ A s_tagg[WS_NSLOT];    // look-back warp, lane 0: aggregate / has-head of the tiles it has published
  __shared__ int s_thead[WS_NSLOT];
  __shared__ __align__(8) unsigned long long xchg_bar;  // workers' barrier of "lazy" tiles (below)

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long G = gridDim.x;
  const long long my_tiles = (num_tiles - (long long)blockIdx.x + G - 1) / G;  // tiles blockIdx.x + i*G

  // Segmented modes: does tile t start on a segment head?  (one flag byte, or arithmetic on the tile base)
  auto starts_on_head = [&](long long t) -> bool {
    if(MODE == MODE_PLAIN) return false;
    if(t == 0) return true;  // element 0 always starts a segment
    const size_t first = (size_t)t * C::TILE;
    if(MODE == MODE_IMPLICIT) return seg_mask != ~(size_t)0 ? (first & seg_mask) == 0 : first % seg_len == 0;
    return __ldg(flags + first) != 0;
  };
  // A tile is LAZY when the tile behind it starts on a head (or there is none): no later tile can need its aggregate
  // — a walk back from tile j stops at the nearest tile holding a head, and tile t+1 holds one — so phase A is
  // skipped (a dummy descriptor keeps the pollers going) and the warps exchange their sums inside phase B instead.
  // With heads every 256 elements every tile is lazy and the kernel does one pass over the ring per tile instead
  // of two: ncu had the explicit-flag kernel at 64 % issue-slot utilisation with 4.5 warps per scheduler, i.e.
  // instruction-bound (profiles/r02_ncu_scan_flags.txt).
  // Explicit flags only: there phase A is as expensive as pass 1 (chained per-row reductions).  With implicit heads
  // phase A is cheap and the extra barrier in phase B costs more than it saves (L=65536: 4.5 -> 3.8 TB/s measured).
  auto is_lazy = [&](long long t) -> bool { return MODE == MODE_FLAGS && (t + 1 >= num_tiles || starts_on_head(t + 1)); };

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
    mbar_init(smem_u32(&xchg_bar), NW);
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
    Deadline dl_seg(timeout_ns);
    // publish(must): tile aggregate -> descriptor, for the tiles of this CTA up to `must` (with SB_WS_PUBLISH_AHEAD also for a
    // tile whose phase A has already finished, from inside the look-back's polling loop; see the note at the macro).
    long long pa = 0, cur = 0;
    auto publish = [&](long long must) {  // blocks until tile `must` of this CTA is out (-1: only what is ready)
      while(pa < my_tiles) {
        const int ls = (int)(pa % WS_NSLOT);
        const uint32_t lp = (uint32_t)(pa / WS_NSLOT) & 1u;
        if(pa <= must) {
          mbar_wait(smem_u32(&agg_bar[ls]), lp);
        } else {
          if(!SB_WS_PUBLISH_AHEAD || pa > cur + WS_DEPTH) break;
          if(!__all_sync(SB_FULL_MASK, mbar_test(smem_u32(&agg_bar[ls]), lp))) break;
        }
        const long long t = (long long)blockIdx.x + pa * G;
        A tile_agg = A(0);
        bool tile_has_head = false;
        if(!is_lazy(t)) {
          A my_prefix;
          bool my_open;
          combine_warps<A, MODE, NW>(s_wagg[ls], s_wflag[ls], -1, my_prefix, my_open, tile_agg, tile_has_head);
        }
        // a lazy tile publishes a placeholder: its value is never used (see is_lazy), its presence is what pollers wait for
        if(lane == 0) {
          Desc<A>::store(desc + t, (epoch << 2) | (tile_has_head ? ST_INCL : ST_AGG), tile_agg);
          s_tagg[ls] = tile_agg;
          s_thead[ls] = tile_has_head ? 1 : 0;
        }
        ++pa;
      }
    };
    for(long long i = 0; i < my_tiles; ++i) {
      const int slot = (int)(i % WS_NSLOT);
      const long long tile = (long long)blockIdx.x + i * G;
      const bool on_head = starts_on_head(tile);  // (the flag-byte load flies during the wait)
      cur = i;
      publish(i);
      A p = A(0);
      if(!on_head) p = lookback_round<A, MODE>(desc, rdesc, tile, G, epoch, seed, dl_seg, err, [&] { publish(-1); });
      if(lane == 0) {
        if((tile + 1) % G == 0) {
          const A tile_agg = s_tagg[slot];  // written by this lane in publish()
          Desc<A>::store(const_cast<typename Desc<A>::word*>(rdesc) + tile / G, (epoch << 2) | ST_INCL, s_thead[slot] ? tile_agg : p + tile_agg);
        }
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
    unsigned lazy_hist = 0;   // bit d: tile (it - d) of this CTA is lazy
    unsigned lazy_done = 0;   // lazy tiles finished so far: phase of xchg_bar
    bool lazy_nxt = my_tiles > 0 && is_lazy((long long)blockIdx.x);
    for(long long it = 0; it < my_tiles + WS_DEPTH; ++it) {
      lazy_hist <<= 1;
      // the decision for the NEXT tile is requested one iteration ahead: with explicit flags it is a global load
      const bool lazy_cur = lazy_nxt;
      lazy_nxt = it + 1 < my_tiles && is_lazy((long long)blockIdx.x + (it + 1) * G);
      if(it < my_tiles) {
        // ---- phase A: tile -> per-warp (sum, has-head) for the look-back warp
        const int stage = (int)(it % WS_STAGES), slot = (int)(it % WS_NSLOT);
        const long long tile = (long long)blockIdx.x + it * G;
        if(lazy_cur) {
          lazy_hist |= 1u;
          if(lane == 0) mbar_arrive(smem_u32(&agg_bar[slot]));  // nothing to contribute: wakes the look-back warp
        } else {
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
            mbar_arrive(smem_u32(&agg_bar[slot]));
          }
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
        const bool lazy = (lazy_hist >> WS_DEPTH) & 1u;
        A v[K][VEC];
        unsigned hf[K];
        if(lazy) mbar_wait(smem_u32(&full_bar[stage]), (uint32_t)(b / WS_STAGES) & 1u);  // phase A did not wait for the data
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
        tile_pass1<A, MODE, K, MODE == MODE_FLAGS>(v, hf, carry_in, open, warp_agg, warp_has_head);
        // The tile lives in registers now (pass 1 consumed every ring read): hand the stage back BEFORE
        // waiting for the prefix, so that the producer can refill it during the look-back wait.  ncu showed
        // ~45 % of the warp samples in mbarrier wait loops with one 64 KB stage in flight per SM at most.
        __syncwarp();
        if(lane == 0) mbar_arrive(smem_u32(&empty_bar[stage]));
        if(lazy) {
          // the warp sums phase A did not produce: exchange them now (one barrier among the workers)
          if(lane == 0) {
            s_wagg[slot][warp] = warp_agg;
            s_wflag[slot][warp] = warp_has_head ? 1 : 0;
            mbar_arrive(smem_u32(&xchg_bar));
          }
          mbar_wait(smem_u32(&xchg_bar), lazy_done & 1u);
          ++lazy_done;
        }
        // all warps' pairs of this tile (and, with look-back, the tile prefix) are complete
        const uint32_t par = (uint32_t)(b / WS_NSLOT) & 1u;
        A tile_prefix = A(0);
        if(use_lookback) {
          mbar_wait(smem_u32(&prefix_bar[slot]), par);
          tile_prefix = s_tprefix[slot];
        } else {
          mbar_wait(smem_u32(&agg_bar[slot]), par);
        }
        if(MODE == MODE_PLAIN && use_lookback) {
          A my_prefix, tile_agg;
          bool my_open, tile_has_head;
          combine_warps<A, MODE, NW>(s_wagg[slot], s_wflag[slot], warp, my_prefix, my_open, tile_agg, tile_has_head);
          const A warp_prefix = my_prefix + tile_prefix;
          // exclusive flag resolved at compile time: no 64-bit select per element (+1-2 %)
          if(exclusive) tile_pass2_store<A, MODE, true, K, false, 1>(v, hf, carry_in, open, warp_prefix, wbase, full_tile, n, out, 1);
          else tile_pass2_store<A, MODE, true, K, false, 0>(v, hf, carry_in, open, warp_prefix, wbase, full_tile, n, out, 0);
        } else {
          A my_prefix, tile_agg;
          bool my_open, tile_has_head;
          combine_warps<A, MODE, NW>(s_wagg[slot], s_wflag[slot], warp, my_prefix, my_open, tile_agg, tile_has_head);
          const A warp_prefix = my_prefix + (my_open ? tile_prefix : A(0));
          tile_pass2_store<A, MODE, true, K>(v, hf, carry_in, open, warp_prefix, wbase, full_tile, n, out, exclusive);
        }
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
    long long num_tiles, int flags_in_ring, unsigned long long* ticket, unsigned long long ticket_base, const unsigned long long* gate,
    unsigned long long gate_seq, int gate_want) {
  using C = TileCfg<A, NW, K>;
  constexpr int VEC = C::VEC;
  constexpr int STAGE_BYTES = ws_stage_bytes<A, MODE, K, NW>();
  extern __shared__ __align__(128) unsigned char ring[];  // WS_STAGES x STAGE_BYTES
  __shared__ __align__(8) unsigned long long full_bar[WS_STAGES], empty_bar[WS_STAGES], agg_bar[WS_NSLOT];
  if(gate && (int)(*gate == gate_seq) != gate_want) {
    // not this kernel's case.  The host has already counted this launch's tickets (tiles + grid): draw them
    if(threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(ticket, (unsigned long long)num_tiles + gridDim.x);
    return;
  }
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
    tile_pass1<A, MODE, K, (MODE == MODE_FLAGS || sizeof(A) == 8)>(v, hf, carry_in, open, warp_agg, warp_has_head);
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

// ---- explicit head flags: does every tile start on a head? ----------------------------------------------------------
// If so (the benchmark's own flags: a head every `--local` elements, and the tile is a multiple of that) no tile needs
// anything from another tile and the independent-tile kernel does the job without descriptors, look-back warp and phase A
// (2^28 elements: int32 5.04 -> 6.58, fp32 5.24 -> 6.59, int64 4.08 -> 6.65, fp64 4.24 -> 6.77 TB/s).
// The host cannot know without a round trip, so it launches this check (one byte per tile), then BOTH scan kernels, each
// of which leaves at once unless the verdict names it.  Verdict: *gate == seq  <=>  some tile of check number `seq` does NOT
// start on a head (sequence numbers only grow, so atomicMax never needs a reset).
__global__ void __launch_bounds__(256) flags_tile_heads_kernel(const uint8_t* __restrict__ flags, long long num_tiles, size_t tile,
    unsigned long long* gate, unsigned long long seq) {
  const long long t = 1 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int bad = t < num_tiles && __ldg(flags + (size_t)t * tile) == 0 ? 1 : 0;
  if(__syncthreads_or(bad) && threadIdx.x == 0) atomicMax(gate, seq);
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

// rows per worker warp: 8.  Explicit flags ran 6 rows (16 workers) so that 3 stages of data + flag bytes fit in 227 KB; with
// 14 workers 8 rows fit for every type:
// (64-bit types carry half as many flag bytes per tile byte: 14 workers x 8 rows = 56 KB + 7 KB of flags per stage fit three times,
//  so they keep 8 rows: the benchmark's flags int64 5.17 -> 6.65, fp64 5.84 -> 6.77 TB/s, random heads +4-8 %
//  (profiles/r02_ops_v18_flags_k8_64bit_ab.jsonl; SB_WS_FLAGS_K64=6 builds the previous 6-row tiles for A/B)
#ifndef SB_WS_FLAGS_K64
#define SB_WS_FLAGS_K64 SB_WS_K
#endif
// (32-bit types: 14 workers x 8 rows = 56 KB + 14 KB of flags per stage fit three times as well; against 16 workers x 6 rows
//  random heads gain 8-9 % (int32 4.42 -> 4.77 TB/s at mean distance 256), the benchmark's flags are unchanged (6.55-6.9):
//  profiles/r02_ops_v19_flags32_14x8_ab.jsonl.  -DSB_WS_NW_FLAGS=16 -DSB_WS_FLAGS_K32=6 builds the previous shape.)
#ifndef SB_WS_FLAGS_K32
#define SB_WS_FLAGS_K32 SB_WS_K
#endif
template <int MODE, int BYTES = 4> struct ws_rows {
  static constexpr int K = MODE == MODE_FLAGS ? (BYTES == 8 ? SB_WS_FLAGS_K64 : SB_WS_FLAGS_K32) : WS_K;
};

template <typename A, int MODE, int NW, int KROWS = 0>
int launch_ws(sb_ctx* ctx, const A* in, A* out, size_t n, int exclusive, size_t seg_len, size_t seg_mask, const uint8_t* flags,
    const A* seed, typename Desc<A>::word* desc, uint32_t epoch, int use_lookback, size_t tiles, const unsigned long long* gate = nullptr,
    unsigned long long gate_seq = 0, int gate_want = 0) {
  constexpr int K = KROWS > 0 ? KROWS : ws_rows<MODE, (int)sizeof(A)>::K;
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
    // (the host mirror of the ticket advances only once the launch is known to have been accepted)
    const unsigned long long tbase = ctx->scan_ticket_base;
    scan_independent_kernel<A, MODE, K, NW><<<(unsigned)g, (NW + 1) * 32, smem, ctx->stream>>>(in, out, n, exclusive, seg_len, seg_mask, flags,
        num_tiles, flags_in_ring, ctx->scan_ticket, tbase, gate, gate_seq, gate_want);
    if(cudaError_t e = cudaGetLastError(); e != cudaSuccess) return (int)e;
    ctx->scan_ticket_base += tiles + g;
    return SB_OK;
  }
  // Cooperative launch: the runtime guarantees that all CTAs of the grid are co-resident (or refuses the
  // launch) — the forward-progress condition of the static round-robin look-back.  No grid barrier is used.
  unsigned* err = ctx->err_dev;
  unsigned long long timeout_ns = sb_peer_timeout_ns();
  void* args[] = {(void*)&in, (void*)&out, (void*)&n, (void*)&exclusive, (void*)&seg_len, (void*)&seg_mask, (void*)&flags, (void*)&seed,
      (void*)&desc, (void*)&epoch, (void*)&use_lookback, (void*)&num_tiles, (void*)&flags_in_ring, (void*)&err, (void*)&timeout_ns,
      (void*)&gate, (void*)&gate_seq, (void*)&gate_want};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)scan_ws_kernel<A, MODE, K, NW>, dim3((unsigned)grid), dim3(WS_THREADS), args, smem, ctx->stream);
  if(e != cudaSuccess) return (int)e;
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
  // 32-bit implicit heads that need the look-back (no tile size is a multiple of the segment: L > 16384 or not a power of
  // two) take the widest tile as well: in the look-back regime the rate is tile bytes per round.  Measured at 2^28
  // (profiles/r02_ops_v14_lb32_wide_ab.jsonl; SB_SCAN_LB32_NW=16 keeps 16 workers): L = 65536 int32 4.56 -> 5.54, fp32
  // 4.06 -> 5.13 TB/s; L = 2^20 +1-8 %, L = 50000 +4-7 %.
  static const bool lb32_wide = !(getenv("SB_SCAN_LB32_NW") && atoi(getenv("SB_SCAN_LB32_NW")) == 16);
  const bool implicit32_lookback = mode == MODE_IMPLICIT && sizeof(A) == 4 && lb32_wide &&
                                   !(sb::is_pow2(seg_len) && (wide_tile % seg_len == 0 || TileCfg<A, WS_NW, WS_K>::TILE % seg_len == 0));
  const bool wide = mode == MODE_PLAIN || (mode == MODE_IMPLICIT && sizeof(A) == 4 && sb::is_pow2(seg_len) && wide_tile % seg_len == 0) ||
                    implicit32_lookback;
  // 64-bit independent tiles: 15 workers, so that the CTA is 16 warps and a thread may use 128 registers — at 17
  // warps ptxas has 96 and spills (fp64 group scan 5.5 -> 6.4 TB/s)
  constexpr size_t narrow_tile = TileCfg<A, WS_NW_NARROW, WS_K>::TILE;
  const bool narrow = mode == MODE_IMPLICIT && sizeof(A) == 8 && sb::is_pow2(seg_len) && narrow_tile % seg_len == 0;
  // 64-bit types on the look-back kernel with implicit heads: 14 workers + producer + look-back = 16 warps, so that a
  // thread may use 128 registers (at 18 warps ptxas has 96 and spills 100-270 bytes, profiles/r02_ptxas_spills.txt).
  // SB_SCAN_LB64_NW=16 keeps the 16-worker instantiation (A/B measurements).
  static const bool lb64_enabled = !(getenv("SB_SCAN_LB64_NW") && atoi(getenv("SB_SCAN_LB64_NW")) == 16);
  // 64-bit segments of 1024 ... 8192 elements: neither 15 x 8 rows (7680 elements) nor 14 x 8 (7168) is a multiple of them, but
  // 16 x 8 rows (8192) is — independent tiles with 16 workers (96 registers, a few spills) still beat the look-back kernel
  // by a wide margin (L = 16384 on the look-back path measured 3.0 TB/s against 5.5-6.6 for independent tiles)
  const bool mid = mode == MODE_IMPLICIT && sizeof(A) == 8 && !narrow && sb::is_pow2(seg_len) && TileCfg<A, WS_NW, WS_K>::TILE % seg_len == 0;
  // (explicit flags run 14 workers for every type, see SB_WS_NW_FLAGS)
  const bool lb64 = lb64_enabled && sizeof(A) == 8 && mode == MODE_IMPLICIT && !narrow && !mid;
  const size_t ws_tile = mode == MODE_FLAGS ? TileCfg<A, WS_NW_FLAGS, ws_rows<MODE_FLAGS, (int)sizeof(A)>::K>::TILE
                         : mode == MODE_PLAIN ? TileCfg<A, plain_shape<sizeof(A)>::NW, plain_shape<sizeof(A)>::K>::TILE
                         : (SB_WS_LB32_SHAPE && implicit32_lookback) ? TileCfg<A, 14, 10>::TILE : wide ? wide_tile : narrow ? narrow_tile : lb64 ? TileCfg<A, WS_NW_LB64, WS_K_LB64>::TILE : TileCfg<A, WS_NW, WS_K>::TILE;
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
    if(!use_ws) tbase = ctx->scan_ticket_base;  // advanced after the launch succeeded
  }
  const A* ain = reinterpret_cast<const A*>(in);
  A* aout = reinterpret_cast<A*>(out);
  const A* aseed = reinterpret_cast<const A*>(seed);
  if(use_ws) {
    int rc;
    if(mode == MODE_PLAIN) rc = launch_ws<A, MODE_PLAIN, plain_shape<sizeof(A)>::NW, plain_shape<sizeof(A)>::K>(ctx, ain, aout, n, exclusive, seg_len, seg_mask, flags, aseed, desc, epoch, use_lookback, tiles);
#if SB_WS_LB32_SHAPE
    else if(implicit32_lookback) rc = launch_ws<A, MODE_IMPLICIT, 14, 10>(ctx, ain, aout, n, exclusive, seg_len, seg_mask, flags, aseed, desc, epoch, use_lookback, tiles);
#endif
    else if(mode == MODE_IMPLICIT && wide) rc = launch_ws<A, MODE_IMPLICIT, WS_NW_WIDE>(ctx, ain, aout, n, exclusive, seg_len, seg_mask, flags, aseed, desc, epoch, use_lookback, tiles);
    else if(mode == MODE_IMPLICIT && narrow) rc = launch_ws<A, MODE_IMPLICIT, WS_NW_NARROW>(ctx, ain, aout, n, exclusive, seg_len, seg_mask, flags, aseed, desc, epoch, use_lookback, tiles);
    else if(mode == MODE_IMPLICIT && lb64) rc = launch_ws<A, MODE_IMPLICIT, WS_NW_LB64, WS_K_LB64>(ctx, ain, aout, n, exclusive, seg_len, seg_mask, flags, aseed, desc, epoch, use_lookback, tiles);
    else if(mode == MODE_IMPLICIT) rc = launch_ws<A, MODE_IMPLICIT, WS_NW>(ctx, ain, aout, n, exclusive, seg_len, seg_mask, flags, aseed, desc, epoch, use_lookback, tiles);
    else {
      // explicit flags: check kernel + both scan kernels, gated by its verdict (see flags_tile_heads_kernel).
      // SB_SCAN_FLAGS_GATE=0 launches the look-back kernel alone (A/B measurements).
      static const bool gate_enabled = !(getenv("SB_SCAN_FLAGS_GATE") && getenv("SB_SCAN_FLAGS_GATE")[0] == '0');
      const unsigned long long* gate = nullptr;
      unsigned long long seq = 0;
      rc = SB_OK;
      if(gate_enabled && use_lookback && tiles >= 64) {
        unsigned long long* g = ctx->scan_ticket + 16;  // byte 128 of the ctx's 256-byte ticket block (zeroed at creation)
        seq = ++ctx->flags_gate_seq;
        flags_tile_heads_kernel<<<(unsigned)((tiles + 254) / 256), 256, 0, ctx->stream>>>(flags, (long long)tiles, tile, g, seq);
        gate = g;
        ctx->launches += 2;  // the check and the kernel of the pair that leaves at once are launches too (sb_ctx_launch_count)
        // (both kernels of the pair share one tile size: 14 workers x 8 rows, a multiple of 256 elements for every type)
        rc = launch_ws<A, MODE_FLAGS, WS_NW_FLAGS>(ctx, ain, aout, n, exclusive, seg_len, seg_mask, flags, aseed, desc, epoch, 0, tiles, gate, seq, 0);
      }
      if(rc == SB_OK) rc = launch_ws<A, MODE_FLAGS, WS_NW_FLAGS>(ctx, ain, aout, n, exclusive, seg_len, seg_mask, flags, aseed, desc, epoch, use_lookback, tiles, gate, seq, 1);
    }
    if(rc != SB_OK) return rc;
    SB_CHECK_LAUNCH(ctx);
    return SB_OK;
  }
  const dim3 grid((unsigned)tiles), block(SCAN_NT);
#define SB_SCAN_LAUNCH(MODE_, AL_)                                                                                       \
  scan_kernel<A, MODE_, AL_><<<grid, block, 0, ctx->stream>>>(ain, aout, n, exclusive, seg_len, seg_mask, flags, aseed, \
      desc, ctx->scan_ticket, tbase, epoch, use_lookback, ctx->err_dev, sb_peer_timeout_ns())
  if(mode == MODE_PLAIN) {
    if(aligned) SB_SCAN_LAUNCH(MODE_PLAIN, true); else SB_SCAN_LAUNCH(MODE_PLAIN, false);
  } else if(mode == MODE_IMPLICIT) {
    if(aligned) SB_SCAN_LAUNCH(MODE_IMPLICIT, true); else SB_SCAN_LAUNCH(MODE_IMPLICIT, false);
  } else {
    if(aligned) SB_SCAN_LAUNCH(MODE_FLAGS, true); else SB_SCAN_LAUNCH(MODE_FLAGS, false);
  }
#undef SB_SCAN_LAUNCH
  SB_CHECK_LAUNCH(ctx);
  if(use_lookback) ctx->scan_ticket_base += tiles;
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

size_t sb_scan_min_tile_bytes() { return (size_t)(SCAN_NT / 32) * SCAN_K * 512; }  // the ticketed kernel's 16 KB tile

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
