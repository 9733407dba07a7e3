// sb_segreduce_<T>: per-segment sum, in place, result at the segment head.
//
// Replaces SegmentedReduction<T>::submit_ndrange / submit_hierarchical
// (/root/reference/pattern/segmentedreduction.cpp:35-61, :63-96): one work-group of `--local`
// work-items per segment, one element per work-item, log2(L) barrier tree in local memory, then
// `acc[gid] = scratch[0]` for lid == 0.  Layout contract: :112-119 (head = sum, rest untouched).
//
// B200 design: HBM-bound stream, algorithmic bytes N*s + (N/L)*s.
//  - no shared memory and no barriers: a warp streams "super-rows" of 4 rows x 32 lanes x 16 B
//    (2 KB, four 128-bit loads in flight per lane, 128 KB per resident SM);
//  - the segment SHAPE is a template parameter (lanes per segment 1..32, two rows, four rows), so
//    every instantiation is a straight-line loop body without guards; only the last, partial
//    super-row takes the guarded path.  The first version decided the shape at run time inside
//    the loop and sat at 69 % issue-slot utilisation for L=64 (profiles/r01_ncu_summary.txt);
//  - a segment is summed by xor-shuffles over exactly the lanes that hold it; when a segment is one
//    row, 64-bit / fp types sum the four rows of a super-row with 6 shuffles instead of 20
//    (warp_sum4), 32-bit integers use redux.sync;
//  - the head is rewritten by the lane that owns it (one scattered store per segment);
//  - segments longer than a super-row accumulate in registers over their super-rows first;
//  - grid = sm_count x 8 CTAs of 256 threads, warp-granular grid-stride: a CTA reads 16 KB
//    contiguous per iteration.
//  - large inputs (>= 4 tiles of 64 KB) with segments of at most a super-row go through segreduce_ring_kernel: the
//    same per-segment arithmetic, but the rows arrive through a persistent cp.async.bulk (TMA) ring, tiles handed out
//    in order by an atomic ticket: 192 KB in flight per SM regardless of registers and occupancy.  ncu on the
//    register-fed kernel at L=64 (profiles/r02_ncu_segreduce_L64.txt): 47-54 % warps active, 22-26 % issue active,
//    28 long-scoreboard cycles per instruction — latency-bound with too few bytes in flight.
// A generic warp-per-segment kernel covers everything else (L not a power of two, L < 16 B,
// unaligned buffers); it sums exactly the subset of a segment the reference's halving tree
// reaches (oracle/sb_oracle.c: ref_tree_member).
#include <cstdlib>

#include "sb_scan_tile.cuh"

namespace {

constexpr int SR_NT = 256;
constexpr int SR_CTAS_PER_SM = 8;

template <typename A> struct acc_of { using type = A; };
template <> struct acc_of<uint16_t> { using type = uint32_t; };

// sum over aligned groups of LANES lanes (compile-time), result in every lane of the group
template <typename Acc, int LANES>
__device__ __forceinline__ Acc group_sum(Acc s) {
  if(LANES == 32) return sb::warp_sum<Acc>(s);
#pragma unroll
  for(int d = LANES >> 1; d > 0; d >>= 1) s += __shfl_xor_sync(SB_FULL_MASK, s, d);
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

// SHAPE: 1..32 = lanes per segment (segment <= one row), 64 = segment is two rows, 128 = segment is a super-row
// `row0` points at lane 0 of row 0 of the super-row; valid_rows (GUARDED only) = rows that exist.
template <typename A, bool GUARDED>
__device__ __forceinline__ void load_superrow(sb::Vec<A> (&t)[4], const A* row0, int lane, int valid_rows) {
  constexpr int VEC = 16 / sizeof(A);
  constexpr int ROW = 32 * VEC;
  const A* p = row0 + lane * VEC;
#pragma unroll
  for(int u = 0; u < 4; ++u)
    if(!GUARDED || u < valid_rows) t[u] = sb::ldg_stream<A>(p + u * ROW);
}

template <typename A, int SHAPE, bool GUARDED>
__device__ __forceinline__ void reduce_superrow(const sb::Vec<A> (&t)[4], A* row0, int lane, int valid_rows) {
  using Acc = typename acc_of<A>::type;
  constexpr int VEC = 16 / sizeof(A);
  constexpr int ROW = 32 * VEC;
  A* p = row0 + lane * VEC;
  Acc s[4];
#pragma unroll
  for(int u = 0; u < 4; ++u) {
    s[u] = Acc(0);
    if(!GUARDED || u < valid_rows) {
#pragma unroll
      for(int c = 0; c < VEC; ++c) s[u] += t[u].v[c];
    }
  }
  if(SHAPE == 32 && !sb::has_redux<Acc>::value) {
    const Acc r = warp_sum4<Acc>(s[0], s[1], s[2], s[3], lane);
    const int u = lane >> 3;
    if((lane & 7) == 0 && (!GUARDED || u < valid_rows)) row0[u * ROW] = (A)r;
  } else if(SHAPE <= 32) {
#pragma unroll
    for(int u = 0; u < 4; ++u) {
      const Acc r = group_sum<Acc, SHAPE>(s[u]);
      if((lane & (SHAPE - 1)) == 0 && (!GUARDED || u < valid_rows)) p[u * ROW] = (A)r;
    }
  } else if(SHAPE == 64) {
    if(!sb::has_redux<Acc>::value) {
      const Acc r = warp_sum2<Acc>(s[0] + s[1], s[2] + s[3], lane);
      const int u = lane >> 4;
      if((lane & 15) == 0 && (!GUARDED || 2 * u < valid_rows)) row0[u * 2 * ROW] = (A)r;
    } else {
      const Acc a = sb::warp_sum<Acc>(s[0] + s[1]), b = sb::warp_sum<Acc>(s[2] + s[3]);
      if(lane == 0) {
        row0[0] = (A)a;
        if(!GUARDED || valid_rows > 2) row0[2 * ROW] = (A)b;
      }
    }
  } else {
    const Acc a = sb::warp_sum<Acc>((s[0] + s[1]) + (s[2] + s[3]));
    if(lane == 0) row0[0] = (A)a;
  }
}

template <typename A, int NT, int SHAPE>
__global__ void __launch_bounds__(NT) segreduce_rows_kernel(A* buf, size_t n) {
  constexpr int VEC = 16 / sizeof(A);
  constexpr size_t ROW = 32 * VEC;
  constexpr size_t SUPER = 4 * ROW;
  const int lane = threadIdx.x & 31;
  const size_t gw = (size_t)blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
  const size_t nw = (size_t)gridDim.x * (NT / 32);
  const size_t full = n / SUPER;
  A* row0 = buf + gw * SUPER;
  const size_t step = nw * SUPER;
  sb::Vec<A> cur[4];
  // 64-bit types with short segments spend long in the shuffle chains: keep the loads of the warp's
  // next super-row in flight meanwhile (measured: f64 L=64 4.35 -> 5.08 TB/s; the extra registers
  // cost occupancy and lose 5-12 % for 32-bit types and for L >= 256, so those stay unpipelined)
  constexpr bool PREFETCH = sizeof(A) == 8 && SHAPE <= 64;
  if(PREFETCH) {
    sb::Vec<A> nxt[4];
    if(gw < full) load_superrow<A, false>(cur, row0, lane, 4);
    for(size_t sr = gw; sr < full; sr += nw, row0 += step) {
      const bool more = sr + nw < full;
      if(more) load_superrow<A, false>(nxt, row0 + step, lane, 4);
      reduce_superrow<A, SHAPE, false>(cur, row0, lane, 4);
      if(more) {
#pragma unroll
        for(int u = 0; u < 4; ++u) cur[u] = nxt[u];
      }
    }
  } else {
    for(size_t sr = gw; sr < full; sr += nw, row0 += step) {
      load_superrow<A, false>(cur, row0, lane, 4);
      reduce_superrow<A, SHAPE, false>(cur, row0, lane, 4);
    }
  }
  // partial last super-row (n is a multiple of the row length here, so whole rows remain)
  const size_t rest = n - full * SUPER;
  if(rest && gw == full % nw) {
    const int valid_rows = (int)(rest / ROW);
    load_superrow<A, true>(cur, buf + full * SUPER, lane, valid_rows);
    reduce_superrow<A, SHAPE, true>(cur, buf + full * SUPER, lane, valid_rows);
  }
}

// ---- the same arithmetic fed by a TMA ring (large inputs) ----------------------------------------------------------
// warps [0, SRR_NW): workers, each owns 8 consecutive rows (two super-rows) of a tile; warp SRR_NW: producer.
constexpr int SRR_NW = 16, SRR_STAGES = 3, SRR_TILE_BYTES = SRR_NW * 8 * 512, SRR_IDS = 8;

// LANE (segments of 128 / 256 / 512 bytes = SHAPE 8 / 16 / 32): a lane owns 128 CONTIGUOUS bytes of the tile instead of 16 bytes
// of each of 8 rows, i.e. a whole segment or a half / quarter of one: 8 LDS.128 (chunk order rotated by the lane number, so the
// eight lanes of a quarter-warp hit eight different 16-byte bank groups), a serial sum in registers, log2(SHAPE / 8) xor-shuffles
// instead of log2(SHAPE) per row, and one head store per segment for the whole 4 KB of the warp.
template <typename A, int SHAPE, bool LANE>
__global__ void __launch_bounds__((SRR_NW + 1) * 32, 1)
segreduce_ring_kernel(A* buf, size_t n, long long num_tiles, unsigned long long* ticket, unsigned long long ticket_base) {
  using namespace sbscan;
  constexpr int VEC = 16 / sizeof(A);
  constexpr size_t ROW = 32 * VEC, SUPER = 4 * ROW, TILE = SRR_TILE_BYTES / sizeof(A);
  extern __shared__ __align__(128) unsigned char ring[];  // SRR_STAGES x SRR_TILE_BYTES
  __shared__ __align__(8) unsigned long long full_bar[SRR_STAGES], empty_bar[SRR_STAGES];
  __shared__ long long s_tile_id[SRR_IDS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if(threadIdx.x == 0) {
#pragma unroll
    for(int s = 0; s < SRR_STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), SRR_NW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if(warp == SRR_NW) {
    if(lane != 0) return;
    for(int i = 0;; ++i) {
      const int stage = i % SRR_STAGES;
      if(i >= SRR_STAGES) mbar_wait(smem_u32(&empty_bar[stage]), (uint32_t)((i / SRR_STAGES) - 1) & 1u);
      const long long tile = (long long)(atomicAdd(ticket, 1ull) - ticket_base);
      const uint32_t bar = smem_u32(&full_bar[stage]);
      if(tile >= num_tiles) {
        s_tile_id[i % SRR_IDS] = -1;
        mbar_arrive(bar);  // releases the workers; they read the -1 and leave
        break;
      }
      s_tile_id[i % SRR_IDS] = tile;
      const size_t base = (size_t)tile * TILE;
      if(base + TILE <= n) {
        mbar_expect_tx(bar, SRR_TILE_BYTES);
        bulk_load(smem_u32(ring) + stage * SRR_TILE_BYTES, buf + base, SRR_TILE_BYTES, bar);
      } else {
        mbar_arrive(bar);  // ragged last tile: read from global memory by the workers
      }
    }
    return;
  }
  for(int it = 0;; ++it) {
    const int stage = it % SRR_STAGES;
    mbar_wait(smem_u32(&full_bar[stage]), (uint32_t)(it / SRR_STAGES) & 1u);
    const long long tile = s_tile_id[it % SRR_IDS];
    if(tile < 0) break;
    const size_t base = (size_t)tile * TILE + (size_t)warp * 2 * SUPER;  // first element of the warp's two super-rows
    if(LANE && (size_t)(tile + 1) * TILE <= n) {
      using Acc = typename acc_of<A>::type;
      constexpr int LPS = SHAPE / 8;  // lanes per segment
      const unsigned char* sp = ring + (size_t)stage * SRR_TILE_BYTES + (size_t)threadIdx.x * 128;
      sb::Vec<A> t[8];
#pragma unroll
      for(int k = 0; k < 8; ++k) t[k] = *reinterpret_cast<const sb::Vec<A>*>(sp + ((k + lane) & 7) * 16);
      __syncwarp();
      if(lane == 0) mbar_arrive(smem_u32(&empty_bar[stage]));  // the bytes live in registers now
      Acc acc[VEC];
#pragma unroll
      for(int c = 0; c < VEC; ++c) acc[c] = (Acc)t[0].v[c];
#pragma unroll
      for(int k = 1; k < 8; ++k) {
#pragma unroll
        for(int c = 0; c < VEC; ++c) acc[c] += (Acc)t[k].v[c];
      }
#pragma unroll
      for(int w = VEC / 2; w > 0; w >>= 1) {
#pragma unroll
        for(int c = 0; c < w; ++c) acc[c] += acc[c + w];
      }
      Acc r = acc[0];
#pragma unroll
      for(int d = LPS >> 1; d > 0; d >>= 1) r += __shfl_xor_sync(SB_FULL_MASK, r, d);
      if((lane & (LPS - 1)) == 0) buf[(size_t)tile * TILE + (size_t)threadIdx.x * (128 / sizeof(A))] = (A)r;
    } else if((size_t)(tile + 1) * TILE <= n) {
      sb::Vec<A> t0[4], t1[4];
      const unsigned char* sp = ring + (size_t)stage * SRR_TILE_BYTES + (size_t)warp * 8 * 512 + lane * 16;
#pragma unroll
      for(int u = 0; u < 4; ++u) {
        t0[u] = *reinterpret_cast<const sb::Vec<A>*>(sp + u * 512);
        t1[u] = *reinterpret_cast<const sb::Vec<A>*>(sp + (4 + u) * 512);
      }
      __syncwarp();
      if(lane == 0) mbar_arrive(smem_u32(&empty_bar[stage]));  // the rows live in registers now
      reduce_superrow<A, SHAPE, false>(t0, buf + base, lane, 4);
      reduce_superrow<A, SHAPE, false>(t1, buf + base + SUPER, lane, 4);
    } else {
      if(lane == 0) mbar_arrive(smem_u32(&empty_bar[stage]));
      // ragged last tile (n is a multiple of the row length): whole rows remain, read with guards from global memory
#pragma unroll
      for(int h = 0; h < 2; ++h) {
        const size_t b = base + h * SUPER;
        if(b < n) {
          const size_t rows = (n - b) / ROW;
          const int valid_rows = rows < 4 ? (int)rows : 4;
          sb::Vec<A> t[4];
          load_superrow<A, true>(t, buf + b, lane, valid_rows);
          reduce_superrow<A, SHAPE, true>(t, buf + b, lane, valid_rows);
        }
      }
    }
  }
}

// rows of a partial super-row may themselves be partial when a segment is shorter than a row:
// such tails are left to the generic kernel (see segreduce_impl: the row kernel gets a multiple of ROW)

// segments longer than a super-row: one warp per segment, register accumulation
template <typename A, int NT>
__global__ void __launch_bounds__(NT) segreduce_long_kernel(A* buf, size_t n, size_t seg_len) {
  using Acc = typename acc_of<A>::type;
  constexpr int VEC = 16 / sizeof(A);
  constexpr size_t ROW = 32 * VEC;
  constexpr size_t SUPER = 4 * ROW;
  const int lane = threadIdx.x & 31;
  const size_t gw = (size_t)blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
  const size_t nw = (size_t)gridDim.x * (NT / 32);
  const size_t nseg = n / seg_len;
  const size_t per = seg_len / SUPER;
  for(size_t seg = gw; seg < nseg; seg += nw) {
    A* p = buf + seg * seg_len + (size_t)lane * VEC;
    Acc a[4];
#pragma unroll
    for(int u = 0; u < 4; ++u) a[u] = Acc(0);
    for(size_t k = 0; k < per; ++k, p += SUPER) {
      sb::Vec<A> t[4];
#pragma unroll
      for(int u = 0; u < 4; ++u) t[u] = sb::ldg_stream<A>(p + u * ROW);
#pragma unroll
      for(int u = 0; u < 4; ++u) {
#pragma unroll
        for(int c = 0; c < VEC; ++c) a[u] += t[u].v[c];
      }
    }
    const Acc total = sb::warp_sum<Acc>((a[0] + a[1]) + (a[2] + a[3]));
    if(lane == 0) buf[seg * seg_len] = (A)total;
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

// SB_SEGREDUCE_RING=0 keeps the register-fed kernel for every size (A/B measurements)
inline bool ring_enabled() {
  static const bool on = !(getenv("SB_SEGREDUCE_RING") && getenv("SB_SEGREDUCE_RING")[0] == '0');
  return on;
}

// SB_SEGREDUCE_LANE=0 keeps the row-per-warp arithmetic inside the ring kernel (A/B measurements)
inline bool lane_enabled() {
  static const bool on = !(getenv("SB_SEGREDUCE_LANE") && getenv("SB_SEGREDUCE_LANE")[0] == '0');
  return on;
}

template <typename A, int SHAPE, bool LANE>
bool launch_ring(sb_ctx* ctx, A* buf, size_t n, size_t tiles) {
  constexpr size_t smem = (size_t)SRR_STAGES * SRR_TILE_BYTES;
  static bool configured[64] = {};
  if(ctx->device < 0 || ctx->device >= 64) return false;
  if(!configured[ctx->device]) {
    if(cudaFuncSetAttribute(segreduce_ring_kernel<A, SHAPE, LANE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess) configured[ctx->device] = true;
    (void)cudaGetLastError();
    if(!configured[ctx->device]) return false;
  }
  size_t grid = (size_t)ctx->sm_count;
  if(grid > tiles) grid = tiles;
  // every CTA's producer draws tickets until one is past the last tile: tiles + grid tickets per launch
  const unsigned long long tbase = ctx->scan_ticket_base;
  segreduce_ring_kernel<A, SHAPE, LANE><<<(unsigned)grid, (SRR_NW + 1) * 32, smem, ctx->stream>>>(buf, n, (long long)tiles, ctx->scan_ticket, tbase);
  if(cudaPeekAtLastError() == cudaSuccess) ctx->scan_ticket_base += tiles + grid;
  return true;
}

template <typename A, int SHAPE>
void launch_rows(sb_ctx* ctx, A* buf, size_t n, size_t blocks) {
  constexpr size_t TILE = SRR_TILE_BYTES / sizeof(A);
  const size_t tiles = (n + TILE - 1) / TILE;
  if(ring_enabled() && tiles >= 4) {
    constexpr bool CAN_LANE = SHAPE == 8 || SHAPE == 16 || SHAPE == 32;
    if(CAN_LANE && lane_enabled()) {
      if(launch_ring<A, SHAPE, CAN_LANE>(ctx, buf, n, tiles)) return;
    } else if(launch_ring<A, SHAPE, false>(ctx, buf, n, tiles)) {
      return;
    }
  }
  segreduce_rows_kernel<A, SR_NT, SHAPE><<<(unsigned)blocks, SR_NT, 0, ctx->stream>>>(buf, n);
}

template <typename T>
int segreduce_impl(sb_ctx* ctx, T* inout, size_t n, size_t seg_len) {
  using A = typename sb_arith<T>::type;
  if(!ctx || (n && !inout)) return SB_ERR_BAD_ARG;
  if(seg_len == 0 || n % seg_len != 0) return SB_ERR_SEG_LEN;
  if(!sb::aligned_to(inout, sizeof(T))) return SB_ERR_MISALIGNED;
  if(n == 0) return SB_OK;
  sb_device_guard guard(ctx->device);
  constexpr size_t VEC = 16 / sizeof(T);
  constexpr size_t ROW = 32 * VEC, SUPER = 4 * ROW;
  constexpr size_t WPB = SR_NT / 32;
  A* buf = reinterpret_cast<A*>(inout);
  const bool pow2 = sb::is_pow2(seg_len);
  const size_t max_blocks = (size_t)ctx->sm_count * SR_CTAS_PER_SM;
  auto blocks_for = [&](size_t tasks) {
    size_t b = (tasks + WPB - 1) / WPB;
    return b > max_blocks ? max_blocks : (b ? b : 1);
  };
  if(pow2 && seg_len >= VEC && sb::aligned_to(inout, 16)) {
    if(seg_len > SUPER) {
      segreduce_long_kernel<A, SR_NT><<<(unsigned)blocks_for(n / seg_len), SR_NT, 0, ctx->stream>>>(buf, n, seg_len);
      SB_CHECK_LAUNCH(ctx);
      return SB_OK;
    }
    // the row kernel takes whole rows; a tail shorter than a row (only possible when seg_len < ROW) goes generic
    const size_t n_rows = (n / ROW) * ROW;
    if(n_rows) {
      // FLAT grid (one super-row per warp, CTAs dispatched in order): in-order dispatch keeps the DRAM
      // access front compact (profiles/r01_stream_probe.txt; L=256: 5.2 -> 5.5 TB/s i32, 5.9 -> 6.4 f64).
      // Exception: 64-bit types with segments of at most two rows are bound by their shuffle chains and
      // do better with the persistent, software-pipelined loop (f64 L=64: 4.78 flat, 4.98 persistent).
      const size_t lps = seg_len / VEC;  // lanes per segment if <= 32, else 64 / 128 = two / four rows
      const bool persistent = sizeof(T) == 8 && lps <= 64;
      size_t blocks = ((n_rows + SUPER - 1) / SUPER + WPB - 1) / WPB;
      if(persistent) blocks = blocks_for((n_rows + SUPER - 1) / SUPER);
      if(blocks > (size_t(1) << 24)) blocks = size_t(1) << 24;
      switch(lps) {
      case 1: launch_rows<A, 1>(ctx, buf, n_rows, blocks); break;
      case 2: launch_rows<A, 2>(ctx, buf, n_rows, blocks); break;
      case 4: launch_rows<A, 4>(ctx, buf, n_rows, blocks); break;
      case 8: launch_rows<A, 8>(ctx, buf, n_rows, blocks); break;
      case 16: launch_rows<A, 16>(ctx, buf, n_rows, blocks); break;
      case 32: launch_rows<A, 32>(ctx, buf, n_rows, blocks); break;
      case 64: launch_rows<A, 64>(ctx, buf, n_rows, blocks); break;
      default: launch_rows<A, 128>(ctx, buf, n_rows, blocks); break;
      }
      SB_CHECK_LAUNCH(ctx);
    }
    if(n_rows < n) {
      segreduce_generic_kernel<A, SR_NT><<<(unsigned)blocks_for((n - n_rows) / seg_len), SR_NT, 0, ctx->stream>>>(buf + n_rows, n - n_rows, seg_len, 1);
      SB_CHECK_LAUNCH(ctx);
    }
    return SB_OK;
  }
  segreduce_generic_kernel<A, SR_NT><<<(unsigned)blocks_for(n / seg_len), SR_NT, 0, ctx->stream>>>(buf, n, seg_len, pow2 ? 1 : 0);
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
