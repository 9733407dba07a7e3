// Peer-memory mailboxes shared by the kernels that exchange values between GPUs themselves (sb_scan_cyclic.cu,
// sb_reduce.cu): layout of the mailbox allocation, self-validating entries, polling with a deadline.
//
// Mailbox of one rank (cudaMalloc, mapped into every peer by CUDA IPC or peer access):
//   [0, 256)            started[world]      u64: sequence number of the scan this rank has entered (plane hand-over)
//   [256, 768)          reduce slots        2 planes x CY_MAX_WORLD entries (all-reduce of sb_reduce_all)
//   [1024, ...)         scan planes         2 planes x CY_MAX_ROUNDS x world entries (round totals of sb_scan_cyclic)
// An entry is two 64-bit words {tag:32 | value bits:32} (low, high half), each self-validating: a reader accepts an
// entry when both words carry the tag of the current operation, so a 16-byte store may arrive torn.
#pragma once
#include "sb_common.cuh"

constexpr int CY_MAX_WORLD = 16;
constexpr int CY_MAX_ROUNDS = 16384;  // 16384 x 10.4 MB = 170 GB per GPU
constexpr size_t CY_STARTED_BYTES = 256;
constexpr size_t CY_REDUCE_OFFSET = 256;   // 2 x CY_MAX_WORLD x 16 B
constexpr size_t CY_HEADER_BYTES = 1024;
constexpr int CY_TIMEOUT_MS = 8000;  // default; SB_CYCLIC_TIMEOUT_MS overrides (tests)

struct sb_peer {
  sb_ctx* ctx = nullptr;
  int rank = 0, world = 1;
  void* base = nullptr;  // own mailbox
  size_t bytes = 0;
  void* peer_base[CY_MAX_WORLD] = {};
  bool ipc_opened[CY_MAX_WORLD] = {};
  bool connected[CY_MAX_WORLD] = {};
  uint32_t op_seq = 0;   // scans issued on this group
  uint32_t red_seq = 0;  // all-reduces issued on this group
};

inline size_t plane_bytes(int world) { return (size_t)CY_MAX_ROUNDS * world * 16; }
inline size_t mailbox_bytes(int world) { return CY_HEADER_BYTES + 2 * plane_bytes(world); }
unsigned long long sb_peer_timeout_ns();  // sb_scan_cyclic.cu

// what a kernel needs for an all-reduce of one value: slot[d] = this rank's entry in peer d's current reduce plane,
// mine = base of the own current reduce plane (entry r = rank r's value)
struct ReducePeers {
  unsigned long long* slot[CY_MAX_WORLD];
  const unsigned long long* mine;
  int world, rank;
  uint32_t tag;
  unsigned long long timeout_ns;
  unsigned* err;
};

// ---- polling with a deadline -------------------------------------------------------------------------------------
struct Deadline {
  unsigned spins = 0;
  unsigned long long t0 = 0, limit_ns;
  __device__ explicit Deadline(unsigned long long ns) : limit_ns(ns) {}
  // true once the caller has been polling for longer than limit_ns in total (looked at every 256 passes)
  __device__ __forceinline__ bool expired() {
    if((++spins & 255u) != 0) return false;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if(t0 == 0) {
      t0 = t;
      return false;
    }
    return t - t0 > limit_ns;
  }
};

__device__ __forceinline__ void mbox_load(const unsigned long long* p, unsigned long long& w0, unsigned long long& w1) {
  asm volatile("ld.relaxed.sys.global.v2.u64 {%0,%1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p) : "memory");
}
__device__ __forceinline__ void mbox_store(unsigned long long* p, unsigned long long w0, unsigned long long w1) {
  asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ unsigned long long mbox_bits(uint32_t v) { return v; }
__device__ __forceinline__ unsigned long long mbox_bits(float v) { return __float_as_uint(v); }
__device__ __forceinline__ unsigned long long mbox_bits(unsigned long long v) { return v; }
__device__ __forceinline__ unsigned long long mbox_bits(double v) { return (unsigned long long)__double_as_longlong(v); }
template <typename A> __device__ __forceinline__ void mbox_pack(A v, uint32_t tag, unsigned long long& w0, unsigned long long& w1) {
  const unsigned long long b = mbox_bits(v);  // 4-byte types: high half is zero
  w0 = ((unsigned long long)tag << 32) | (b & 0xffffffffull);
  w1 = ((unsigned long long)tag << 32) | (b >> 32);
}
template <typename A> __device__ __forceinline__ A mbox_unpack(unsigned long long w0, unsigned long long w1);
template <> __device__ __forceinline__ uint32_t mbox_unpack<uint32_t>(unsigned long long w0, unsigned long long) { return (uint32_t)w0; }
template <> __device__ __forceinline__ float mbox_unpack<float>(unsigned long long w0, unsigned long long) { return __uint_as_float((uint32_t)w0); }
template <> __device__ __forceinline__ unsigned long long mbox_unpack<unsigned long long>(unsigned long long w0, unsigned long long w1) {
  return (w0 & 0xffffffffull) | (w1 << 32);
}
template <> __device__ __forceinline__ double mbox_unpack<double>(unsigned long long w0, unsigned long long w1) {
  return __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
}
__device__ __forceinline__ bool mbox_ready(unsigned long long w0, unsigned long long w1, uint32_t tag) {
  return (uint32_t)(w0 >> 32) == tag && (uint32_t)(w1 >> 32) == tag;
}

// Warp-collective all-reduce (sum) of one value per rank through the mailboxes: every rank stores its value into
// slot[rank] of every peer's plane, polls its own plane until all `world` entries carry `tag`, and adds them in rank
// order (the same order on every rank: floating-point results are identical everywhere).
template <typename A>
__device__ __forceinline__ A mailbox_allreduce(A v, const ReducePeers& rp) {
  const int lane = threadIdx.x & 31;
  if(lane < rp.world) {
    unsigned long long w0, w1;
    mbox_pack<A>(v, rp.tag, w0, w1);
    mbox_store(rp.slot[lane], w0, w1);
  }
  unsigned long long w0 = 0, w1 = 0;
  Deadline dl(rp.timeout_ns);
  for(;;) {
    bool missing = false;
    if(lane < rp.world && !mbox_ready(w0, w1, rp.tag)) {
      mbox_load(rp.mine + lane * 2, w0, w1);
      missing = !mbox_ready(w0, w1, rp.tag);
    }
    if(!__any_sync(SB_FULL_MASK, missing)) break;
    if(__any_sync(SB_FULL_MASK, dl.expired())) {
      if(lane == 0) *rp.err = 1u;
      break;
    }
  }
  const A mine = lane < rp.world ? mbox_unpack<A>(w0, w1) : A(0);
  A total = A(0);
  for(int r = 0; r < rp.world; ++r) total += __shfl_sync(SB_FULL_MASK, mine, r);
  return total;
}
