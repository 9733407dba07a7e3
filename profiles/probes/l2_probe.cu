// L2-residency probe (NOT product code).  Question: can a persistent kernel read a tile from HBM in a
// "phase A" (reduce), and re-read the same tile D rounds later in a "phase B" (scan + store) from the
// 126 MB L2 instead of HBM?  If yes, a scan whose tile prefix is not known for several microseconds
// (cross-GPU exchange) still moves only 2*N*s bytes through HBM.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o l2_probe l2_probe.cu
// Run  : ./l2_probe                 -> timing table (CUDA events, 2 GiB in + 2 GiB out)
//        ./l2_probe D HINT SYNC     -> one variant, three launches (for ncu --metrics dram__bytes_read.sum,...)
//
// Round r = G tiles dealt round-robin (tile = r*G + blockIdx.x), as in scan_ws_kernel.  Per iteration `it` a CTA
//   phase A: loads tile (it, pos) and xors it into a register (HINT: L2 evict_last policy), bumps round_cnt[it];
//   phase B: waits until round_cnt[it-D] == tiles of that round (SYNC; stands in for "all aggregates of the round
//            are published"), loads tile (it-D, pos) again (HINT: evict_first) and stores it (HINT: evict_first).
// D = 0 skips phase A altogether: the plain persistent copy this design has to match.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if(e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while(0)

__device__ __forceinline__ uint4 ld_plain(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ld_hint(const uint4* p, unsigned long long pol) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ void st_plain(uint4* p, uint4 r) {
  asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w) : "memory");
}
__device__ __forceinline__ void st_hint(uint4* p, uint4 r, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w), "l"(pol) : "memory");
}

template <int U, int NT>
__global__ void __launch_bounds__(NT) two_phase(const uint4* __restrict__ in, uint4* __restrict__ out, size_t nvec, int D, int hint, int sync,
    unsigned* round_cnt, unsigned* sink) {
  const long long G = gridDim.x;
  const size_t tile_vec = (size_t)NT * U;
  const long long ntiles = (long long)(nvec / tile_vec);
  const long long R = (ntiles + G - 1) / G;
  unsigned long long keep, stream;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(stream));
  unsigned acc = 0;
  for(long long it = 0; it < R + D; ++it) {
    if(D > 0 && it < R) {
      const long long tile = it * G + blockIdx.x;
      if(tile < ntiles) {
        const uint4* p = in + (size_t)tile * tile_vec + threadIdx.x;
        uint4 t[U];
#pragma unroll
        for(int u = 0; u < U; ++u) t[u] = hint ? ld_hint(p + (size_t)u * NT, keep) : ld_plain(p + (size_t)u * NT);
#pragma unroll
        for(int u = 0; u < U; ++u) acc ^= t[u].x ^ t[u].y ^ t[u].z ^ t[u].w;
        if(sync) {
          __syncthreads();
          if(threadIdx.x == 0) atomicAdd(round_cnt + it, 1u);
        }
      }
    }
    if(it >= D) {
      const long long r = it - D;
      const long long tile = r * G + blockIdx.x;
      if(tile < ntiles) {
        if(sync && D > 0) {
          if(threadIdx.x == 0) {
            const long long left = ntiles - r * G;
            const unsigned want = (unsigned)(left < G ? left : G);
            unsigned v;
            long long spins = 0;  // bail out instead of hanging the box if the grid is not co-resident
            do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(round_cnt + r) : "memory"); } while(v < want && ++spins < (1ll << 22));
            if(v < want) *sink = 0xdeadu;
          }
          __syncthreads();
        }
        const uint4* p = in + (size_t)tile * tile_vec + threadIdx.x;
        uint4* q = out + (size_t)tile * tile_vec + threadIdx.x;
        uint4 t[U];
#pragma unroll
        for(int u = 0; u < U; ++u) t[u] = hint ? ld_hint(p + (size_t)u * NT, stream) : ld_plain(p + (size_t)u * NT);
#pragma unroll
        for(int u = 0; u < U; ++u) {
          t[u].x += 1;
          if(hint) st_hint(q + (size_t)u * NT, t[u], stream); else st_plain(q + (size_t)u * NT, t[u]);
        }
      }
    }
  }
  if(acc == 0x12345u) *sink = acc;
}

int main(int argc, char** argv) {
  const size_t bytes = size_t(2) << 30;
  uint4 *a, *b;
  CK(cudaMalloc(&a, bytes));
  CK(cudaMalloc(&b, bytes));
  CK(cudaMemset(a, 1, bytes));
  unsigned *cnt, *sink;
  CK(cudaMalloc(&cnt, 1 << 20));
  CK(cudaMalloc(&sink, 4));
  const size_t nvec = bytes / 16;
  int sms;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  int l2 = 0, persist_max = 0;
  CK(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, 0));
  CK(cudaDeviceGetAttribute(&persist_max, cudaDevAttrMaxPersistingL2CacheSize, 0));
  printf("SMs %d, L2 %d bytes, max persisting L2 %d bytes\n", sms, l2, persist_max);

  auto launch = [&](int U, int ctas, int D, int hint, int sync) {
    CK(cudaMemsetAsync(cnt, 0, 1 << 20));
    if(U == 8) two_phase<8, 512><<<sms * ctas, 512>>>(a, b, nvec, D, hint, sync, cnt, sink);
    else if(U == 4) two_phase<4, 512><<<sms * ctas, 512>>>(a, b, nvec, D, hint, sync, cnt, sink);
    else two_phase<2, 512><<<sms * ctas, 512>>>(a, b, nvec, D, hint, sync, cnt, sink);
  };
  if(argc >= 4) {
    const int D = atoi(argv[1]), hint = atoi(argv[2]), sync = atoi(argv[3]);
    const int U = argc > 4 ? atoi(argv[4]) : 8, ctas = argc > 5 ? atoi(argv[5]) : 1;
    for(int i = 0; i < 3; ++i) launch(U, ctas, D, hint, sync);
    CK(cudaDeviceSynchronize());
    return 0;
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  auto timeit = [&](int U, int ctas, int D, int hint, int sync) {
    for(int i = 0; i < 2; ++i) launch(U, ctas, D, hint, sync);
    std::vector<float> ms;
    for(int i = 0; i < 7; ++i) {
      CK(cudaMemsetAsync(cnt, 0, 1 << 20));
      CK(cudaEventRecord(e0));
      if(U == 8) two_phase<8, 512><<<sms * ctas, 512>>>(a, b, nvec, D, hint, sync, cnt, sink);
      else if(U == 4) two_phase<4, 512><<<sms * ctas, 512>>>(a, b, nvec, D, hint, sync, cnt, sink);
      else two_phase<2, 512><<<sms * ctas, 512>>>(a, b, nvec, D, hint, sync, cnt, sink);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float t;
      CK(cudaEventElapsedTime(&t, e0, e1));
      ms.push_back(t);
    }
    std::sort(ms.begin(), ms.end());
    const double round_mb = (double)sms * ctas * 512 * U * 16 / 1e6;
    printf("tile %3d KB x %d CTA/SM (round %5.1f MB)  D=%d hint=%d sync=%d   %8.4f ms  %8.1f GB/s (med %8.1f)   [2*N*s accounting]\n", 512 * U * 16 / 1024, ctas,
        round_mb, D, hint, sync, ms[0], 2.0 * bytes / ms[0] / 1e6, 2.0 * bytes / ms[3] / 1e6);
  };
  for(int pass = 0; pass < 2; ++pass) {
    if(pass == 1) {
      if(persist_max <= 0) break;
      CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)persist_max));
      printf("---- with cudaLimitPersistingL2CacheSize = %d\n", persist_max);
    }
    for(int U : {8, 4}) {
      for(int ctas : {1, 2}) {
        if(pass == 1 && !(U == 8 && ctas == 1)) continue;
        timeit(U, ctas, 0, 0, 0);  // plain persistent copy
        for(int D : {1, 2, 3, 4, 6, 8}) {
          timeit(U, ctas, D, 0, 1);
          timeit(U, ctas, D, 1, 1);
        }
        timeit(U, ctas, 2, 0, 0);  // no round synchronisation: CTAs drift
        timeit(U, ctas, 2, 1, 0);
      }
    }
  }
  return 0;
}
