// sb_scan_cyclic_<T> + sb_peer_*: global prefix sum that reads every element from HBM ONCE, on one GPU or on the
// P GPUs of a box, with the exchange of partial sums done by the kernel itself over NVLink peer memory.
//
// Upstream has no scan code (pattern/prefixsum.cpp:1 is a TODO stub) and no multi-device code; the semantics are
// DESIGN.md's scan contract (global inclusive / exclusive sum of in[i] = (T)i, pattern/reduction.cpp:31-34).
//
// Why: with CONTIGUOUS shards no output of shard r can be written before every shard < r has been read completely,
// so reduce-then-scan (sb_comm_scan / sharded.py) reads a shard twice: 3*N*s/P bytes per GPU, 0.735 weak-scaling
// efficiency at 8 GPUs in round 1.  Here the global array is dealt to the GPUs BLOCK-CYCLICALLY in blocks of one
// kernel round (G tiles = 148 x 70 KB = 10.1 MB): round r of rank g is block r*P + g.  Every rank runs the same
// persistent kernel; per tile it does
//   phase A (tile i + D): TMA bulk load (L2 policy evict_last) -> per-warp sums -> tile aggregate -> descriptor;
//   phase B (tile i)    : TMA bulk load of the same tile again — it was read D rounds ago and is still in the
//                         126 MB L2 (policy evict_first: last use) — scan with the tile prefix, 128-bit stores
//                         (policy evict_first).  The two kinds of load alternate in one smem ring, so the distance
//                         between the two reads of a tile is exactly D rounds (D x 10.1 MB of L2).
// The prefix of tile (round r, rank g, position pos) is
//      E[r]                      everything in rounds < r, on all ranks        (running sum kept by every look-back warp)
//    + sum_{g' < g} T[r][g']     the round totals of the lower ranks           (mailbox, written by the peers)
//    + sum_{q < pos} agg[r][q]   the tile aggregates in front of it            (local descriptors, as in scan_ws_kernel)
// T[r][g] = total of round r on rank g is summed by the leader (CTA 0's look-back warp) as soon as the G aggregates are out and stored into
// the mailbox of EVERY rank (st.relaxed.sys over NVLink; 16 bytes per round and peer).  Nothing in that chain waits
// for a scan to finish: all totals are independent, so the only latency is aggregate -> leader -> NVLink -> look-back
// (a few microseconds), which the D rounds of look-ahead cover.  HBM traffic is 2*N*s/P per GPU, the same as on one
// GPU; NCCL is not involved.  On ONE GPU the same kernel takes the inter-CTA prefix chain off the critical path.
//
// Mailbox protocol: entry (round, writer) = two 64-bit words {tag:32 | value bits:32} (low, high half), each
// self-validating, tag = sequence number of the operation in the peer group (all ranks call in the same order, like
// any collective).  Two mailbox planes alternate between operations; before its first write of operation k the leader
// posts started[rank] = k to every peer and waits until every peer has started operation k-1, i.e. has finished
// k-2, the last user of the plane.  Every polling loop gives up after SB_CYCLIC_TIMEOUT_MS (8 s) and raises the ctx error
// flag (sb_ctx_sync returns SB_ERR_TIMEOUT) instead of hanging the device when a peer never shows up.
#include <cstdlib>
#include <cstring>
#include <new>

#include "sb_peer.cuh"
#include "sb_scan_tile.cuh"

// Tile shape: 14 workers x 10 rows x 512 B = 70 KB (16 warps with producer and look-back warp -> 128 registers per thread).
// The first shape, 18 workers x 8 rows (72 KB, 20 warps -> 96 registers), spilled 44 bytes in the fp64 instantiation; measured
// on one GPU at 2^28 (profiles/r02_ops_v23_cyclic_tile_shapes.jsonl): fp64 4.96 -> 6.17, int64 6.00 -> 6.26, int32 6.25 -> 6.34,
// fp32 6.23 -> 6.26 TB/s; 16 x 9 loses 8 % for 64-bit types.
#ifndef SB_CY_NW
#define SB_CY_NW 14
#endif
#ifndef SB_CY_K
#define SB_CY_K 10
#endif
#ifndef SB_CY_STAGES
#define SB_CY_STAGES 3
#endif
// phase B source: 0 = second TMA load of the tile through the smem ring, interleaved with the phase-A loads (the
// distance between the two reads of a tile is exactly D rounds); 1 = 128-bit loads straight into registers, the ring
// carrying phase-A loads only.  Measured (profiles/r02_cyclic_variants.txt): variant 1 keeps more HBM requests in
// flight but the ring's own look-ahead (3 tiles) adds to the distance — at D=2 40 % of the re-reads miss L2
// (dram read 3.0 GB for 2.15) and the kernel drops from 5.9 to 4.8 TB/s.  Default 0.
#ifndef SB_CY_BLOAD
#define SB_CY_BLOAD 0
#endif

namespace {
using namespace sbscan;

constexpr int CY_NW = SB_CY_NW, CY_K = SB_CY_K, CY_STAGES = SB_CY_STAGES;
constexpr int CY_NSLOT = 8;      // ring of per-tile slots (warp sums, tile prefix); look-ahead D <= CY_NSLOT - 2
constexpr int CY_MAX_DEPTH = CY_NSLOT - 2;
struct PeerTable {
  unsigned long long* mbox[CY_MAX_WORLD];     // plane of the current operation in peer d's mailbox
  unsigned long long* started[CY_MAX_WORLD];  // started[world] array of peer d
};

// phase B: the warp's K rows of a FULL tile, global (L2-resident) -> registers, one 128-bit load per lane and row
template <typename A, int K>
__device__ __forceinline__ void tile_load_l2(A (&v)[K][16 / sizeof(A)], const A* lane_ptr, int hint, unsigned long long policy) {
  constexpr int VEC = 16 / sizeof(A);
#pragma unroll
  for(int j = 0; j < K; ++j) {
    uint4 r;
    const A* p = lane_ptr + (size_t)j * (32 * VEC);
    if(hint) asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(policy));
    else asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    const sb::Vec<A>& t = *reinterpret_cast<const sb::Vec<A>*>(&r);
#pragma unroll
    for(int c = 0; c < VEC; ++c) v[j][c] = t.v[c];
  }
}
// phase A: lane-partial sum of the warp's K rows, read from the ring (two accumulators; fixed order)
template <typename A, int K>
__device__ __forceinline__ A tile_sum_smem(const unsigned char* warp_rows) {
  constexpr int VEC = 16 / sizeof(A);
  const unsigned char* sp = warp_rows + (threadIdx.x & 31) * 16;
  A s0 = A(0), s1 = A(0);
#pragma unroll
  for(int j = 0; j < K; ++j) {
    const sb::Vec<A> t = *reinterpret_cast<const sb::Vec<A>*>(sp + (size_t)j * 512);
    A r = t.v[0];
#pragma unroll
    for(int c = 1; c < VEC; ++c) r += t.v[c];
    if(j & 1) s1 += r; else s0 += r;
  }
  return s0 + s1;
}
// the same for the ragged last tile, from global memory with guards
template <typename A, int K>
__device__ __forceinline__ A tile_sum_global(const A* in, size_t wbase, size_t n) {
  constexpr int VEC = 16 / sizeof(A);
  A s0 = A(0), s1 = A(0);
#pragma unroll
  for(int j = 0; j < K; ++j) {
    const size_t idx = wbase + (size_t)j * (32 * VEC);
    A r = A(0);
#pragma unroll
    for(int c = 0; c < VEC; ++c) r += (idx + c < n) ? in[idx + c] : A(0);
    if(j & 1) s1 += r; else s0 += r;
  }
  return s0 + s1;
}

// ---- the kernel ------------------------------------------------------------------------------------------------------
// warps: [0, NW) workers, NW producer, NW+1 look-back (in CTA 0 also the leader that publishes the round totals; a
// separate leader warp would make the CTA 21 warps, which ptxas rounds to 24: 80 instead of 96 registers and spills)
template <typename A, int K, int NW>
__global__ void __launch_bounds__((NW + 2) * 32, 1)
scan_cyclic_kernel(const A* in, A* out, size_t n, int exclusive, const A* __restrict__ seed, typename Desc<A>::word* desc, uint32_t epoch,
    long long num_tiles, long long G, int D, int hint, const __grid_constant__ PeerTable peers, int world, int rank, uint32_t tag, unsigned* err,
    unsigned long long timeout_ns) {
  using C = TileCfg<A, NW, K>;
  constexpr int VEC = C::VEC;
  constexpr int STAGE_BYTES = C::TILE_BYTES;
  extern __shared__ __align__(128) unsigned char ring[];  // CY_STAGES x STAGE_BYTES
  __shared__ __align__(8) unsigned long long full_bar[CY_STAGES], empty_bar[CY_STAGES], agg_bar[CY_NSLOT], prefix_bar[CY_NSLOT];
  __shared__ A s_wagg[CY_NSLOT][NW];  // phase A: per-warp sums of a tile
  __shared__ A s_tprefix[CY_NSLOT];   // look-back warp: prefix of the tile

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long pos = blockIdx.x;
  const long long my_tiles = pos < num_tiles ? (num_tiles - pos + G - 1) / G : 0;  // tiles pos + i*G: one per round

  if(threadIdx.x == 0) {
#pragma unroll
    for(int s = 0; s < CY_STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), NW);
    }
#pragma unroll
    for(int s = 0; s < CY_NSLOT; ++s) {
      mbar_init(smem_u32(&agg_bar[s]), NW);
      mbar_init(smem_u32(&prefix_bar[s]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();  // the only block-wide barrier

  if(warp == NW) {
    // ===== producer: phase-A loads (HBM, L2 policy evict_last) and, interleaved D rounds behind them, the phase-B
    // loads of the same tiles (L2 hits, policy evict_first: last use) =====
    if(lane != 0) return;
    unsigned long long keep = 0, stream = 0;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(stream));
    long long q = 0;  // load sequence number: stage = q % CY_STAGES
    auto issue = [&](long long i, unsigned long long policy) {
      const int stage = (int)(q % CY_STAGES);
      if(q >= CY_STAGES) mbar_wait(smem_u32(&empty_bar[stage]), (uint32_t)((q / CY_STAGES) - 1) & 1u);
      ++q;
      const size_t base = (size_t)(pos + i * G) * C::TILE;
      const uint32_t bar = smem_u32(&full_bar[stage]);
      if(base + C::TILE <= n) {
        mbar_expect_tx(bar, C::TILE_BYTES);
        if(hint) bulk_load_hint(smem_u32(ring) + stage * STAGE_BYTES, in + base, C::TILE_BYTES, bar, policy);
        else bulk_load(smem_u32(ring) + stage * STAGE_BYTES, in + base, C::TILE_BYTES, bar);
      } else {
        mbar_arrive(bar);  // ragged last tile: the workers read it from global memory with guards
      }
    };
    for(long long it = 0; it < my_tiles + D; ++it) {
      if(it < my_tiles) issue(it, keep);
      if(SB_CY_BLOAD == 0 && it >= D) issue(it - D, stream);
    }
  } else if(warp == NW + 1) {
    // ===== look-back warp: tile aggregate -> descriptor; prefix = E[r] + lower ranks' totals of round r + local aggregates =====
    Deadline dl(timeout_ns);
    const unsigned long long* mine = peers.mbox[rank];
    const bool leader = blockIdx.x == 0;  // position 0 has a tile in every round
    if(leader) {
      // plane hand-over: tell everybody that this rank is in operation `tag`, wait until everybody is at least in
      // tag-1 (has finished tag-2, the last user of this mailbox plane)
      if(lane < world) {
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(peers.started[lane] + rank), "l"((unsigned long long)tag) : "memory");
      }
      for(;;) {
        bool behind = false;
        if(lane < world) {
          unsigned long long v;
          asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(peers.started[rank] + lane) : "memory");
          behind = (int32_t)((uint32_t)v - (tag - 1u)) < 0;
        }
        if(!__any_sync(SB_FULL_MASK, behind)) break;
        if(__any_sync(SB_FULL_MASK, dl.expired())) {
          if(lane == 0) *err = 1u;
          break;
        }
      }
    }
    A E = A(0);
    if(seed) E = *seed;
    // Publishing runs AHEAD of looking back.  The aggregate of round pa (and, in the leader, the round total T[pa]) goes
    // out as soon as phase A of that round is complete — up to D rounds ahead of the look-back, also from inside the
    // look-back's polling loop.  If a warp published round i+1 only after its look-back of round i, every total would wait
    // for the lower ranks' totals of the round before: a chain across the ranks (rank g's T[i] after rank 0's T[i-g], rank
    // 0's after rank P-1's T[i-2]) whose latency grows with P — the 64-bit leg went 0.84 / 0.88 / 0.93 ms at 2 / 4 / 8
    // GPUs with the sequential loop (profiles/r02_scale_n*.json).
    long long pa = 0, cur = 0;
    auto publish = [&](long long must) {  // blocks until round `must` is out (-1: only what is ready)
      while(pa < my_tiles) {
        const int ls = (int)(pa % CY_NSLOT);
        const uint32_t lp = (uint32_t)(pa / CY_NSLOT) & 1u;
        if(pa <= must) {
          mbar_wait(smem_u32(&agg_bar[ls]), lp);
        } else {
          if(pa > cur + D) break;  // phase A is at most D rounds ahead of the tile this warp owes a prefix for
          if(!__all_sync(SB_FULL_MASK, mbar_test(smem_u32(&agg_bar[ls]), lp))) break;
        }
        A my_prefix, tile_agg;
        bool my_open, tile_has_head;
        combine_warps<A, MODE_PLAIN, NW>(s_wagg[ls], nullptr, -1, my_prefix, my_open, tile_agg, tile_has_head);
        if(lane == 0) Desc<A>::store(desc + pos + pa * G, (epoch << 2) | ST_AGG, tile_agg);
        if(leader) {
          // round total -> every rank's mailbox; depends on nothing but this GPU's phase A
          __syncwarp();
          const long long left = num_tiles - pa * G;
          const int count = (int)(left < G ? left : G);
          const A total = sum_round_aggregates<A>(desc + pa * G, count, epoch, dl, err);
          if(lane < world) {
            unsigned long long w0, w1;
            mbox_pack<A>(total, tag, w0, w1);
            mbox_store(peers.mbox[lane] + (pa * world + rank) * 2, w0, w1);
          }
        }
        ++pa;
      }
    };
    for(long long i = 0; i < my_tiles; ++i) {  // i is the round index of this CTA's i-th tile
      const int slot = (int)(i % CY_NSLOT);
      cur = i;
      publish(i);
      // mailbox: lane l reads T[i-1][l] (all ranks, to advance E) and, if l < rank, T[i][l]
      unsigned long long p0 = 0, p1 = 0, c0 = 0, c1 = 0;
      const bool want_prev = i > 0 && lane < world, want_cur = lane < rank;
      if(want_prev) mbox_load(mine + ((i - 1) * world + lane) * 2, p0, p1);
      if(want_cur) mbox_load(mine + (i * world + lane) * 2, c0, c1);
      const A local = sum_round_aggregates<A>(desc + i * G, (int)pos, epoch, dl, err);
      for(;;) {
        bool missing = false;
        if(want_prev && !mbox_ready(p0, p1, tag)) {
          missing = true;
          mbox_load(mine + ((i - 1) * world + lane) * 2, p0, p1);
        }
        if(want_cur && !mbox_ready(c0, c1, tag)) {
          missing = true;
          mbox_load(mine + (i * world + lane) * 2, c0, c1);
        }
        if(!__any_sync(SB_FULL_MASK, missing)) break;
        publish(-1);  // do not sit on finished aggregates while waiting for other ranks
        if(__any_sync(SB_FULL_MASK, dl.expired())) {
          if(lane == 0) *err = 1u;
          break;
        }
      }
      E += sb::warp_sum<A>(want_prev ? mbox_unpack<A>(p0, p1) : A(0));
      const A lower = sb::warp_sum<A>(want_cur ? mbox_unpack<A>(c0, c1) : A(0));
      if(lane == 0) {
        s_tprefix[slot] = (E + lower) + local;
        mbar_arrive(smem_u32(&prefix_bar[slot]));
      }
    }
  } else {
    // ===== workers: per iteration phase A of tile `it` (from the ring), then phase B of tile it-D =====
    unsigned long long stream = 0;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(stream));
    long long q = 0;  // ring sequence number, in step with the producer
    for(long long it = 0; it < my_tiles + D; ++it) {
      const bool do_b = it >= D;
      const long long b = it - D;
      const size_t bbase = do_b ? (size_t)(pos + b * G) * C::TILE : 0;
      const size_t bw = bbase + (size_t)warp * C::WCHUNK + (size_t)lane * VEC;
      const bool bfull = bbase + C::TILE <= n;
      A v[K][VEC];
      unsigned hf[K];
      if(SB_CY_BLOAD == 1 && do_b) {
        // variant 1: the K 128-bit loads per lane are in flight while phase A below runs (volatile asm: issued here)
        if(bfull) tile_load_l2<A, K>(v, in + bw, hint, stream);
        else tile_load_global<A, true, K>(v, in, bw, false, n);
      }
      if(it < my_tiles) {
        // ---- phase A: tile -> per-warp sums, streamed out of the ring with a handful of registers
        const int stage = (int)(q % CY_STAGES), slot = (int)(it % CY_NSLOT);
        const uint32_t par = (uint32_t)(q / CY_STAGES) & 1u;
        ++q;
        const size_t base = (size_t)(pos + it * G) * C::TILE;
        mbar_wait(smem_u32(&full_bar[stage]), par);
        A s;
        if(base + C::TILE <= n) {
          s = tile_sum_smem<A, K>(ring + (size_t)stage * STAGE_BYTES + (size_t)warp * K * 512);
        } else {
          s = tile_sum_global<A, K>(in, base + (size_t)warp * C::WCHUNK + (size_t)lane * VEC, n);
        }
        const A agg = sb::warp_sum<A>(s);
        if(lane == 0) {
          mbar_arrive(smem_u32(&empty_bar[stage]));  // warp_sum consumed every lane's smem reads
          s_wagg[slot][warp] = agg;
          mbar_arrive(smem_u32(&agg_bar[slot]));
        }
      }
      if(do_b) {
        // ---- phase B: the same tile again, scan with its prefix, store
        const int slot = (int)(b % CY_NSLOT);
        if(SB_CY_BLOAD == 0) {
          const int stage = (int)(q % CY_STAGES);
          const uint32_t par = (uint32_t)(q / CY_STAGES) & 1u;
          ++q;
          mbar_wait(smem_u32(&full_bar[stage]), par);
          if(bfull) tile_load_smem<A, K>(v, ring + (size_t)stage * STAGE_BYTES + (size_t)warp * K * 512);
          else tile_load_global<A, true, K>(v, in, bw, false, n);
        }
        A carry_in[K];
        bool open[K];
        A warp_agg;
        bool warp_has_head;
        tile_pass1<A, MODE_PLAIN, K>(v, hf, carry_in, open, warp_agg, warp_has_head);
        if(SB_CY_BLOAD == 0) {
          __syncwarp();
          if(lane == 0) mbar_arrive(smem_u32(&empty_bar[(int)((q - 1) % CY_STAGES)]));  // the tile lives in registers now
        }
        mbar_wait(smem_u32(&prefix_bar[slot]), (uint32_t)(b / CY_NSLOT) & 1u);
        // (the per-warp prefixes are NOT precomputed by the look-back warp: that was measured 10-17 % slower on the
        // look-back kernel, sb_scan.cu SB_WPREFIX — the hand-over is the critical path, not the workers' issue slots)
        const A tile_prefix = s_tprefix[slot];
        A my_prefix, tile_agg;
        bool my_open, tile_has_head;
        combine_warps<A, MODE_PLAIN, NW>(s_wagg[slot], nullptr, warp, my_prefix, my_open, tile_agg, tile_has_head);
        const A warp_prefix = my_prefix + tile_prefix;
        if(hint) {
          if(exclusive) tile_pass2_store<A, MODE_PLAIN, true, K, true, 1>(v, hf, carry_in, open, warp_prefix, bw, bfull, n, out, 1, stream);
          else tile_pass2_store<A, MODE_PLAIN, true, K, true, 0>(v, hf, carry_in, open, warp_prefix, bw, bfull, n, out, 0, stream);
        } else {
          tile_pass2_store<A, MODE_PLAIN, true, K, false>(v, hf, carry_in, open, warp_prefix, bw, bfull, n, out, exclusive);
        }
      }
    }
  }
}

// ---- host side -------------------------------------------------------------------------------------------------------
constexpr int CY_THREADS = (CY_NW + 2) * 32;
template <typename A> constexpr size_t cy_smem() { return (size_t)CY_STAGES * TileCfg<A, CY_NW, CY_K>::TILE_BYTES; }

// tiles per round = co-resident CTAs of the kernel (sm_count x 1); 0 if the kernel does not fit
template <typename A>
long long tiles_per_round(sb_ctx* ctx) {
  static int per_sm_cache[64] = {};
  if(ctx->device < 0 || ctx->device >= 64) return 0;
  int nb = per_sm_cache[ctx->device];
  if(nb == 0) {
    if(cudaFuncSetAttribute(scan_cyclic_kernel<A, CY_K, CY_NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cy_smem<A>()) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, scan_cyclic_kernel<A, CY_K, CY_NW>, CY_THREADS, cy_smem<A>()) != cudaSuccess || nb <= 0) {
      (void)cudaGetLastError();
      nb = -1;
    }
    per_sm_cache[ctx->device] = nb;
  }
  return nb > 0 ? (long long)ctx->sm_count * nb : 0;
}

int env_int(const char* name, int dflt, int lo, int hi) {
  const char* v = getenv(name);
  if(!v || !*v) return dflt;
  const int x = atoi(v);
  return x < lo ? lo : (x > hi ? hi : x);
}

int private_peer(sb_ctx* ctx, sb_peer** out) {
  if(!ctx->self_peer) {
    sb_peer* p = nullptr;
    int rc = sb_peer_create(ctx, 0, 1, &p);
    if(rc != SB_OK) return rc;
    ctx->self_peer = p;
  }
  *out = ctx->self_peer;
  return SB_OK;
}

template <typename T>
int scan_cyclic_impl(sb_ctx* ctx, sb_peer* peer, const T* in, T* out, size_t n, int exclusive, const T* seed) {
  using A = typename sb_arith<T>::type;
  using W = typename Desc<A>::word;
  if(!ctx || (n && (!in || !out))) return SB_ERR_BAD_ARG;
  if(peer && peer->ctx != ctx) return SB_ERR_BAD_ARG;
  if(!sb::aligned_to(in, 16) || !sb::aligned_to(out, 16) || !sb::aligned_to(seed, sizeof(T))) return SB_ERR_MISALIGNED;
  sb_device_guard guard(ctx->device);
  if(!peer) {
    int rc = private_peer(ctx, &peer);
    if(rc != SB_OK) return rc;
  }
  for(int d = 0; d < peer->world; ++d)
    if(!peer->connected[d]) return SB_ERR_PEER;
  const uint32_t tag = ++peer->op_seq;  // advances on every rank, also when there is nothing to do
  if(n == 0) return SB_OK;
  const long long G = tiles_per_round<A>(ctx);
  if(G <= 0) return SB_ERR_WORKSPACE;
  constexpr size_t tile = TileCfg<A, CY_NW, CY_K>::TILE;
  const size_t tiles = (n + tile - 1) / tile;
  const size_t rounds = (tiles + (size_t)G - 1) / (size_t)G;
  if(rounds > (size_t)CY_MAX_ROUNDS) return SB_ERR_BAD_ARG;
  int rc = sb_ctx_reserve_desc(ctx, sizeof(A) == 4 ? tiles + 64 : 0, sizeof(A) == 8 ? tiles + 64 : 0);
  if(rc != SB_OK) return rc;
  if(ctx->scan_epoch >= (1u << 30) - 1) {
    if(ctx->desc64) SB_CUDA_TRY(cudaMemsetAsync(ctx->desc64, 0, ctx->desc64_count * 8, ctx->stream));
    if(ctx->desc128) SB_CUDA_TRY(cudaMemsetAsync(ctx->desc128, 0, ctx->desc128_count * 16, ctx->stream));
    ctx->scan_epoch = 0;
  }
  uint32_t epoch = ++ctx->scan_epoch;
  W* desc = reinterpret_cast<W*>(sizeof(A) == 4 ? ctx->desc64 : ctx->desc128);
  PeerTable pt;
  memset(&pt, 0, sizeof(pt));
  for(int d = 0; d < peer->world; ++d) {
    char* b = static_cast<char*>(peer->peer_base[d]);
    pt.started[d] = reinterpret_cast<unsigned long long*>(b);
    pt.mbox[d] = reinterpret_cast<unsigned long long*>(b + CY_HEADER_BYTES + (tag & 1u) * plane_bytes(peer->world));
  }
  // look-ahead in rounds and L2 policies: defaults from the measurements in profiles/ (r02); the env knobs exist for tuning runs
  static const int env_depth = env_int("SB_CYCLIC_DEPTH", 0, 0, CY_MAX_DEPTH), env_hint = env_int("SB_CYCLIC_HINT", -1, -1, 1);
  int D = env_depth > 0 ? env_depth : 2;  // rounds of 10.1 MB: L2 keeps ~3 of them (profiles/r02_cyclic_variants.txt); 8 GPUs: D=1/2/3 -> 7.00/7.00/6.87x on config 5
  int hint = env_hint >= 0 ? env_hint : 1;
  size_t grid = (size_t)G < tiles ? (size_t)G : tiles;
  const A* ain = reinterpret_cast<const A*>(in);
  A* aout = reinterpret_cast<A*>(out);
  const A* aseed = reinterpret_cast<const A*>(seed);
  long long num_tiles = (long long)tiles, g = G;
  int world = peer->world, rank = peer->rank;
  uint32_t tg = tag;
  unsigned* err = ctx->err_dev;
  unsigned long long timeout_ns = sb_peer_timeout_ns();
  void* args[] = {(void*)&ain, (void*)&aout, (void*)&n, (void*)&exclusive, (void*)&aseed, (void*)&desc, (void*)&epoch, (void*)&num_tiles, (void*)&g,
      (void*)&D, (void*)&hint, (void*)&pt, (void*)&world, (void*)&rank, (void*)&tg, (void*)&err, (void*)&timeout_ns};
  // cooperative launch: all CTAs co-resident (the look-back and leader warps wait for other CTAs of the grid)
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)scan_cyclic_kernel<A, CY_K, CY_NW>, dim3((unsigned)grid), dim3(CY_THREADS), args, cy_smem<A>(), ctx->stream);
  if(e != cudaSuccess) return (int)e;
  SB_CHECK_LAUNCH(ctx);
  return SB_OK;
}

size_t dtype_size(int dtype) {
  switch(dtype) {
  case SB_I32: case SB_F32: return 4;
  case SB_I64: case SB_F64: return 8;
  default: return 0;
  }
}

}  // namespace

unsigned long long sb_peer_timeout_ns() {
  static const int ms = env_int("SB_CYCLIC_TIMEOUT_MS", CY_TIMEOUT_MS, 1, 600000);
  return (unsigned long long)ms * 1000000ull;
}

void sb_ctx_release_self_peer(sb_ctx* ctx) {
  if(ctx->self_peer) {
    sb_peer_destroy(ctx->self_peer);
    ctx->self_peer = nullptr;
  }
}

extern "C" {

size_t sb_cyclic_block_elems(sb_ctx* ctx, int dtype) {
  if(!ctx) return 0;
  sb_device_guard guard(ctx->device);
  switch(dtype) {
  case SB_I32: return (size_t)tiles_per_round<uint32_t>(ctx) * TileCfg<uint32_t, CY_NW, CY_K>::TILE;
  case SB_F32: return (size_t)tiles_per_round<float>(ctx) * TileCfg<float, CY_NW, CY_K>::TILE;
  case SB_I64: return (size_t)tiles_per_round<unsigned long long>(ctx) * TileCfg<unsigned long long, CY_NW, CY_K>::TILE;
  case SB_F64: return (size_t)tiles_per_round<double>(ctx) * TileCfg<double, CY_NW, CY_K>::TILE;
  default: return 0;
  }
}

int sb_peer_create(sb_ctx* ctx, int rank, int world, sb_peer** out) {
  if(!ctx || !out || world < 1 || world > CY_MAX_WORLD || rank < 0 || rank >= world) return SB_ERR_BAD_ARG;
  *out = nullptr;
  sb_device_guard guard(ctx->device);
  sb_peer* p = new(std::nothrow) sb_peer();
  if(!p) return SB_ERR_WORKSPACE;
  p->ctx = ctx;
  p->rank = rank;
  p->world = world;
  p->bytes = mailbox_bytes(world);
  if(cudaMalloc(&p->base, p->bytes) != cudaSuccess) {
    (void)cudaGetLastError();
    delete p;
    return SB_ERR_WORKSPACE;
  }
  // zero = "no operation yet" (tags start at 1); synchronous, so that the memory is clean before any peer can map it
  if(cudaMemset(p->base, 0, p->bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
    cudaFree(p->base);
    (void)cudaGetLastError();
    delete p;
    return SB_ERR_WORKSPACE;
  }
  p->peer_base[rank] = p->base;
  p->connected[rank] = true;
  *out = p;
  return SB_OK;
}

int sb_peer_destroy(sb_peer* p) {
  if(!p) return SB_OK;
  sb_device_guard guard(p->ctx->device);
  cudaStreamSynchronize(p->ctx->stream);
  for(int d = 0; d < p->world; ++d)
    if(p->ipc_opened[d] && p->peer_base[d]) cudaIpcCloseMemHandle(p->peer_base[d]);
  cudaFree(p->base);
  (void)cudaGetLastError();
  delete p;
  return SB_OK;
}

int sb_peer_rank(sb_peer* p) { return p ? p->rank : -1; }
int sb_peer_world(sb_peer* p) { return p ? p->world : 0; }

int sb_peer_handle(sb_peer* p, void* handle, size_t handle_bytes) {
  if(!p || !handle || handle_bytes < SB_PEER_HANDLE_BYTES) return SB_ERR_BAD_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) <= SB_PEER_HANDLE_BYTES, "handle does not fit");
  sb_device_guard guard(p->ctx->device);
  cudaIpcMemHandle_t h;
  if(cudaIpcGetMemHandle(&h, p->base) != cudaSuccess) {
    (void)cudaGetLastError();
    return SB_ERR_PEER;
  }
  memset(handle, 0, handle_bytes);
  memcpy(handle, &h, sizeof(h));
  return SB_OK;
}

int sb_peer_connect(sb_peer* p, int peer_rank, const void* handle, size_t handle_bytes) {
  if(!p || !handle || handle_bytes < sizeof(cudaIpcMemHandle_t) || peer_rank < 0 || peer_rank >= p->world) return SB_ERR_BAD_ARG;
  if(peer_rank == p->rank || p->connected[peer_rank]) return SB_OK;
  sb_device_guard guard(p->ctx->device);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* ptr = nullptr;
  if(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
    (void)cudaGetLastError();
    return SB_ERR_PEER;
  }
  p->peer_base[peer_rank] = ptr;
  p->ipc_opened[peer_rank] = true;
  p->connected[peer_rank] = true;
  return SB_OK;
}

int sb_peer_connect_local(sb_peer* p, sb_peer* other) {
  if(!p || !other || p->world != other->world || other->rank == p->rank) return SB_ERR_BAD_ARG;
  if(p->connected[other->rank]) return SB_OK;
  sb_device_guard guard(p->ctx->device);
  if(other->ctx->device != p->ctx->device) {
    int can = 0;
    if(cudaDeviceCanAccessPeer(&can, p->ctx->device, other->ctx->device) != cudaSuccess || !can) {
      (void)cudaGetLastError();
      return SB_ERR_PEER;
    }
    cudaError_t e = cudaDeviceEnablePeerAccess(other->ctx->device, 0);
    if(e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
      (void)cudaGetLastError();
      return SB_ERR_PEER;
    }
    (void)cudaGetLastError();
  }
  p->peer_base[other->rank] = other->base;
  p->connected[other->rank] = true;
  return SB_OK;
}

int sb_scan_cyclic_i32(sb_ctx* c, sb_peer* p, const int32_t* in, int32_t* out, size_t n, int ex, const int32_t* seed) { return scan_cyclic_impl<int32_t>(c, p, in, out, n, ex, seed); }
int sb_scan_cyclic_i64(sb_ctx* c, sb_peer* p, const int64_t* in, int64_t* out, size_t n, int ex, const int64_t* seed) { return scan_cyclic_impl<int64_t>(c, p, in, out, n, ex, seed); }
int sb_scan_cyclic_f32(sb_ctx* c, sb_peer* p, const float* in, float* out, size_t n, int ex, const float* seed) { return scan_cyclic_impl<float>(c, p, in, out, n, ex, seed); }
int sb_scan_cyclic_f64(sb_ctx* c, sb_peer* p, const double* in, double* out, size_t n, int ex, const double* seed) { return scan_cyclic_impl<double>(c, p, in, out, n, ex, seed); }

int sb_scan_cyclic(sb_ctx* c, sb_peer* p, int dtype, const void* in, void* out, size_t n, int exclusive) {
  switch(dtype) {
  case SB_I32: return sb_scan_cyclic_i32(c, p, (const int32_t*)in, (int32_t*)out, n, exclusive, nullptr);
  case SB_I64: return sb_scan_cyclic_i64(c, p, (const int64_t*)in, (int64_t*)out, n, exclusive, nullptr);
  case SB_F32: return sb_scan_cyclic_f32(c, p, (const float*)in, (float*)out, n, exclusive, nullptr);
  case SB_F64: return sb_scan_cyclic_f64(c, p, (const double*)in, (double*)out, n, exclusive, nullptr);
  default: return SB_ERR_BAD_ARG;
  }
}

}  // extern "C"
