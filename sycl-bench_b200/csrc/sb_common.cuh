// libsb_cuda internals shared by all translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/sb_cuda.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libsb_cuda is written for sm_100a (B200) only"
#endif

#define SB_FULL_MASK 0xffffffffu

struct sb_peer;

struct sb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 0;
  uint64_t launches = 0;
  bool deterministic = false;  // sb_ctx_set_deterministic: bit-reproducible floating-point reductions

  // reduce: per-block partials (8 B each) + completion ticket
  void* partials = nullptr;
  size_t partials_slots = 0;
  unsigned int* reduce_ticket = nullptr;

  // scan: tile descriptors (separate arrays for 64-bit and 128-bit descriptors), dynamic tile ticket
  void* desc64 = nullptr;
  size_t desc64_count = 0;
  void* desc128 = nullptr;
  size_t desc128_count = 0;
  unsigned long long* scan_ticket = nullptr;  // monotonically increasing, never reset
  unsigned long long scan_ticket_base = 0;    // host mirror of the value the next launch starts at
  uint32_t scan_epoch = 0;                    // descriptor generation, bumped per launch
  unsigned long long flags_gate_seq = 0;      // explicit-flag scans: number of the last tile-head check (sb_scan.cu)

  // sb_scan_cyclic on one GPU: a private peer group of world size 1 (created on first use)
  sb_peer* self_peer = nullptr;
  // device-side error flag (pinned, mapped host memory): a kernel that gave up waiting for another CTA / GPU sets it
  unsigned* err_host = nullptr;
  unsigned* err_dev = nullptr;

  // timing
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
  bool timing = false;

  // host-buffer entry points: cached device staging
  void* stage[2] = {nullptr, nullptr};
  size_t stage_bytes[2] = {0, 0};
  void* stage_small = nullptr;  // 64 B result slot
  // device -> host copies of finished chunks run on a second stream so that they overlap the
  // host -> device copies of later chunks (PCIe is full duplex)
  cudaStream_t d2h_stream = nullptr;
  static constexpr int kChunkEvents = 8;
  cudaEvent_t chunk_ev[kChunkEvents] = {};
};

#define SB_CUDA_TRY(expr)                      \
  do {                                         \
    cudaError_t _e = (expr);                   \
    if(_e != cudaSuccess) return (int)_e;      \
  } while(0)

#define SB_CHECK_LAUNCH(ctx)                   \
  do {                                         \
    (ctx)->launches++;                         \
    cudaError_t _e = cudaGetLastError();       \
    if(_e != cudaSuccess) return (int)_e;      \
  } while(0)

struct sb_device_guard {
  int prev = -1;
  explicit sb_device_guard(int dev) {
    cudaGetDevice(&prev);
    if(prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~sb_device_guard() {
    if(prev >= 0) cudaSetDevice(prev);
  }
};

// ---- arithmetic type mapping: integers are summed as unsigned (wrap mod 2^w, no UB) -------------
template <typename T> struct sb_arith { using type = T; };
template <> struct sb_arith<int16_t> { using type = uint16_t; };
template <> struct sb_arith<int32_t> { using type = uint32_t; };
template <> struct sb_arith<int64_t> { using type = unsigned long long; };
template <> struct sb_arith<long long> { using type = unsigned long long; };

namespace sb {

template <typename T> struct vec_of { static constexpr int N = 16 / sizeof(T); };

// 16-byte register tile of a thread
template <typename T>
struct alignas(16) Vec {
  static constexpr int N = 16 / sizeof(T);
  T v[N];
};

// streaming 128-bit load, read-only path (input is never written by the kernel)
template <typename T>
__device__ __forceinline__ Vec<T> ldg_nc(const T* p) {
  uint4 r;
  asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return *reinterpret_cast<Vec<T>*>(&r);
}
// same, but `volatile`: the compiler keeps these loads in program order, i.e. a batch of them written
// back to back is ISSUED back to back.  Without it nvcc sinks some loads next to their first use for
// some element types (measured on the reduce kernel: int32 5.85 and fp64 5.6 TB/s against 6.9-7.05 for
// fp32 / int64 with identical source), which halves the bytes in flight per thread.
template <typename T>
__device__ __forceinline__ Vec<T> ldg_nc_ordered(const T* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return *reinterpret_cast<Vec<T>*>(&r);
}
// streaming 128-bit load through the coherent path (buffer may be written by this kernel)
template <typename T>
__device__ __forceinline__ Vec<T> ldg_stream(const T* p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return *reinterpret_cast<Vec<T>*>(&r);
}
template <typename T>
__device__ __forceinline__ void stg_vec(T* p, const Vec<T>& v) {
  const uint4& r = *reinterpret_cast<const uint4*>(&v);
  asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w) : "memory");
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// warp sum: redux.sync for 32-bit integers, xor-shuffle tree otherwise
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for(int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(SB_FULL_MASK, v, d);
  return v;
}
template <>
__device__ __forceinline__ uint32_t warp_sum<uint32_t>(uint32_t v) {
  return __reduce_add_sync(SB_FULL_MASK, v);
}
template <>
__device__ __forceinline__ uint16_t warp_sum<uint16_t>(uint16_t v) {
  return (uint16_t)__reduce_add_sync(SB_FULL_MASK, (uint32_t)v);
}

// types whose warp sum is a single redux.sync
template <typename T> struct has_redux { static constexpr bool value = false; };
template <> struct has_redux<uint32_t> { static constexpr bool value = true; };
template <> struct has_redux<uint16_t> { static constexpr bool value = true; };

template <typename T>
__device__ __forceinline__ T shfl_up(T v, int d) { return __shfl_up_sync(SB_FULL_MASK, v, d); }
template <typename T>
__device__ __forceinline__ T shfl_idx(T v, int l) { return __shfl_sync(SB_FULL_MASK, v, l); }

// One step of an inclusive warp scan: x += (value of lane - d), only in lanes >= d.  shfl.sync.up returns a
// predicate that says whether the source lane exists, so the add is predicated directly: no compare, no select
// (ncu on the scan kernel: SEL 18 % + ISETP 8 % of all executed instructions with the `if(lane >= d)` form).
template <int D> __device__ __forceinline__ void scan_step_up(uint32_t& x) {
  asm volatile("{.reg .pred p; .reg .u32 t; shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff; @p add.u32 %0, %0, t;}" : "+r"(x) : "n"(D));
}
template <int D> __device__ __forceinline__ void scan_step_up(float& x) {
  asm volatile("{.reg .pred p; .reg .f32 t; shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff; @p add.rn.f32 %0, %0, t;}" : "+f"(x) : "n"(D));
}
template <int D> __device__ __forceinline__ void scan_step_up(unsigned long long& x) {
  asm volatile(
      "{.reg .pred p; .reg .u32 lo, hi, tlo, thi; .reg .u64 t; mov.b64 {lo, hi}, %0;"
      " shfl.sync.up.b32 tlo|p, lo, %1, 0, 0xffffffff; shfl.sync.up.b32 thi, hi, %1, 0, 0xffffffff;"
      " mov.b64 t, {tlo, thi}; @p add.u64 %0, %0, t;}" : "+l"(x) : "n"(D));
}
template <int D> __device__ __forceinline__ void scan_step_up(double& x) {
  asm volatile(
      "{.reg .pred p; .reg .u32 lo, hi, tlo, thi; .reg .f64 t; mov.b64 {lo, hi}, %0;"
      " shfl.sync.up.b32 tlo|p, lo, %1, 0, 0xffffffff; shfl.sync.up.b32 thi, hi, %1, 0, 0xffffffff;"
      " mov.b64 t, {tlo, thi}; @p add.rn.f64 %0, %0, t;}" : "+d"(x) : "n"(D));
}
// inclusive warp scan (plain sum) built from the five steps
template <typename T> __device__ __forceinline__ void warp_inclusive_scan(T& x) {
  scan_step_up<1>(x);
  scan_step_up<2>(x);
  scan_step_up<4>(x);
  scan_step_up<8>(x);
  scan_step_up<16>(x);
}

inline bool is_pow2(size_t x) { return x && !(x & (x - 1)); }
inline bool aligned_to(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

}  // namespace sb

// workspace growth helpers (sb_ctx.cu)
int sb_ctx_reserve_desc(sb_ctx* ctx, size_t count64, size_t count128);
int sb_ctx_reserve_stage(sb_ctx* ctx, int which, size_t bytes);
int sb_ctx_reserve_partials(sb_ctx* ctx, size_t slots);
size_t sb_scan_min_tile_bytes();                      // sb_scan.cu: smallest tile any scan kernel uses (most descriptors)
void sb_ctx_release_self_peer(sb_ctx* ctx);           // sb_scan_cyclic.cu
int sb_ctx_take_error(sb_ctx* ctx);                   // sb_ctx.cu: SB_ERR_TIMEOUT once if a kernel raised the error flag
