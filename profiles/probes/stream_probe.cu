// Streaming-copy probe (NOT product code): which load/store flavour gets closest to the measured
// copy peak on B200?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_probe stream_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if(e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while(0)

template <int VARIANT>
__device__ __forceinline__ uint4 ld(const uint4* p) {
  uint4 r;
  if(VARIANT == 0) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  else if(VARIANT == 1) asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  else if(VARIANT == 2) asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  else asm volatile("ld.global.lu.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
template <int VARIANT>
__device__ __forceinline__ void st(uint4* p, uint4 r) {
  if(VARIANT == 0) asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w) : "memory");
  else if(VARIANT == 1) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w) : "memory");
  else if(VARIANT == 2) asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w) : "memory");
  else asm volatile("st.global.wt.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w) : "memory");
}

template <int LV, int SV, int U, int NT>
__global__ void __launch_bounds__(NT) copy_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t nvec) {
  const size_t stride = (size_t)gridDim.x * NT * U;
  for(size_t b = (size_t)blockIdx.x * NT * U + threadIdx.x; b < nvec; b += stride) {
    uint4 t[U];
#pragma unroll
    for(int u = 0; u < U; ++u) if(b + (size_t)u * NT < nvec) t[u] = ld<LV>(in + b + (size_t)u * NT);
#pragma unroll
    for(int u = 0; u < U; ++u) if(b + (size_t)u * NT < nvec) { t[u].x += 1; st<SV>(out + b + (size_t)u * NT, t[u]); }
  }
}

// 256-bit accesses (sm_100): ld.global.v8 / st.global.v8, optional L2::evict_first on the load
struct alignas(32) u8x { unsigned v[8]; };
template <int EVICT, int U, int NT>
__global__ void __launch_bounds__(NT) copy256_kernel(const u8x* __restrict__ in, u8x* __restrict__ out, size_t nvec) {
  const size_t stride = (size_t)gridDim.x * NT * U;
  for(size_t b = (size_t)blockIdx.x * NT * U + threadIdx.x; b < nvec; b += stride) {
    u8x t[U];
#pragma unroll
    for(int u = 0; u < U; ++u) if(b + (size_t)u * NT < nvec) {
      const u8x* p = in + b + (size_t)u * NT;
      if(EVICT) asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(t[u].v[0]), "=r"(t[u].v[1]), "=r"(t[u].v[2]), "=r"(t[u].v[3]), "=r"(t[u].v[4]), "=r"(t[u].v[5]), "=r"(t[u].v[6]), "=r"(t[u].v[7]) : "l"(p));
      else asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(t[u].v[0]), "=r"(t[u].v[1]), "=r"(t[u].v[2]), "=r"(t[u].v[3]), "=r"(t[u].v[4]), "=r"(t[u].v[5]), "=r"(t[u].v[6]), "=r"(t[u].v[7]) : "l"(p));
    }
#pragma unroll
    for(int u = 0; u < U; ++u) if(b + (size_t)u * NT < nvec) {
      t[u].v[0] += 1;
      u8x* p = out + b + (size_t)u * NT;
      asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(t[u].v[0]), "r"(t[u].v[1]), "r"(t[u].v[2]), "r"(t[u].v[3]), "r"(t[u].v[4]), "r"(t[u].v[5]), "r"(t[u].v[6]), "r"(t[u].v[7]) : "memory");
    }
  }
}

// persistent CTAs, chunks handed out in order by an atomic ticket (what the hardware scheduler does for flat grids)
template <int U, int NT>
__global__ void __launch_bounds__(NT) copy_dynamic(const uint4* __restrict__ in, uint4* __restrict__ out, size_t nvec, unsigned long long* ticket) {
  __shared__ unsigned long long s_chunk;
  const size_t nchunks = (nvec + (size_t)NT * U - 1) / ((size_t)NT * U);
  for(;;) {
    __syncthreads();
    if(threadIdx.x == 0) s_chunk = atomicAdd(ticket, 1ull);
    __syncthreads();
    const size_t c = s_chunk;
    if(c >= nchunks) break;
    const size_t b = c * NT * U + threadIdx.x;
    uint4 t[U];
#pragma unroll
    for(int u = 0; u < U; ++u) if(b + (size_t)u * NT < nvec) t[u] = in[b + (size_t)u * NT];
#pragma unroll
    for(int u = 0; u < U; ++u) if(b + (size_t)u * NT < nvec) { t[u].x += 1; out[b + (size_t)u * NT] = t[u]; }
  }
}

// persistent, static assignment, but CONTIGUOUS per-CTA ranges instead of grid-stride
template <int U, int NT>
__global__ void __launch_bounds__(NT) copy_blocked(const uint4* __restrict__ in, uint4* __restrict__ out, size_t nvec) {
  const size_t per = (nvec + gridDim.x - 1) / gridDim.x;
  const size_t lo = (size_t)blockIdx.x * per, hi = lo + per < nvec ? lo + per : nvec;
  for(size_t b = lo + threadIdx.x; b < hi; b += (size_t)NT * U) {
    uint4 t[U];
#pragma unroll
    for(int u = 0; u < U; ++u) if(b + (size_t)u * NT < hi) t[u] = in[b + (size_t)u * NT];
#pragma unroll
    for(int u = 0; u < U; ++u) if(b + (size_t)u * NT < hi) { t[u].x += 1; out[b + (size_t)u * NT] = t[u]; }
  }
}

// non-persistent: one chunk per CTA
template <int U, int NT>
__global__ void __launch_bounds__(NT) copy_flat(const uint4* __restrict__ in, uint4* __restrict__ out, size_t nvec) {
  const size_t b = (size_t)blockIdx.x * NT * U + threadIdx.x;
  uint4 t[U];
#pragma unroll
  for(int u = 0; u < U; ++u) if(b + (size_t)u * NT < nvec) t[u] = in[b + (size_t)u * NT];
#pragma unroll
  for(int u = 0; u < U; ++u) if(b + (size_t)u * NT < nvec) { t[u].x += 1; out[b + (size_t)u * NT] = t[u]; }
}

template <typename F>
void timeit(const char* name, size_t bytes, F f) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for(int i = 0; i < 3; ++i) f();
  std::vector<float> ms;
  for(int i = 0; i < 10; ++i) { CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float t; CK(cudaEventElapsedTime(&t, e0, e1)); ms.push_back(t); }
  std::sort(ms.begin(), ms.end());
  printf("%-44s %8.4f ms  %8.1f GB/s (med %8.1f)\n", name, ms[0], 2.0 * bytes / ms[0] / 1e6, 2.0 * bytes / ms[5] / 1e6);
}

int main() {
  const size_t bytes = size_t(2) << 30;  // 2 GiB in + 2 GiB out
  uint4 *a, *b; CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMemset(a, 1, bytes));
  const size_t nvec = bytes / 16;
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  timeit("cudaMemcpyAsync D2D", bytes, [&] { CK(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice)); });
#define RUN(LV, SV, U, NT, CTAS) timeit("persist ld" #LV " st" #SV " U" #U " NT" #NT " x" #CTAS, bytes, [&] { copy_kernel<LV, SV, U, NT><<<sms * CTAS, NT>>>(a, b, nvec); });
  RUN(0, 0, 4, 256, 8) RUN(1, 0, 4, 256, 8) RUN(2, 0, 4, 256, 8) RUN(3, 0, 4, 256, 8)
  RUN(0, 1, 4, 256, 8) RUN(0, 2, 4, 256, 8) RUN(0, 3, 4, 256, 8) RUN(3, 1, 4, 256, 8)
  RUN(0, 0, 2, 256, 8) RUN(0, 0, 8, 256, 4) RUN(0, 0, 4, 512, 4) RUN(0, 0, 4, 1024, 2) RUN(0, 0, 4, 128, 16) RUN(0, 0, 1, 256, 8) RUN(0, 1, 8, 256, 4)
  timeit("persist 256-bit U2 NT256 x8", bytes, [&] { copy256_kernel<0, 2, 256><<<sms * 8, 256>>>((const u8x*)a, (u8x*)b, nvec / 2); });
  timeit("persist 256-bit evict_first U2 NT256 x8", bytes, [&] { copy256_kernel<1, 2, 256><<<sms * 8, 256>>>((const u8x*)a, (u8x*)b, nvec / 2); });
  timeit("persist 256-bit U4 NT256 x4", bytes, [&] { copy256_kernel<0, 4, 256><<<sms * 4, 256>>>((const u8x*)a, (u8x*)b, nvec / 2); });
  unsigned long long* tk; CK(cudaMalloc(&tk, 8));
  timeit("persist dynamic ticket U4 NT256 x8", bytes, [&] { CK(cudaMemsetAsync(tk, 0, 8)); copy_dynamic<4, 256><<<sms * 8, 256>>>(a, b, nvec, tk); });
  timeit("persist dynamic ticket U16 NT256 x4 (64KB)", bytes, [&] { CK(cudaMemsetAsync(tk, 0, 8)); copy_dynamic<16, 256><<<sms * 4, 256>>>(a, b, nvec, tk); });
  timeit("persist blocked (contiguous per CTA) x8", bytes, [&] { copy_blocked<4, 256><<<sms * 8, 256>>>(a, b, nvec); });
  timeit("persist grid-stride x1 NT1024 U4", bytes, [&] { copy_kernel<0, 0, 4, 1024><<<sms, 1024>>>(a, b, nvec); });
  timeit("flat U16 NT256 (64KB/CTA)", bytes, [&] { copy_flat<16, 256><<<(unsigned)((nvec + 4095) / 4096), 256>>>(a, b, nvec); });
  timeit("flat U4 NT256", bytes, [&] { copy_flat<4, 256><<<(unsigned)((nvec + 1023) / 1024), 256>>>(a, b, nvec); });
  timeit("flat U1 NT256", bytes, [&] { copy_flat<1, 256><<<(unsigned)((nvec + 255) / 256), 256>>>(a, b, nvec); });
  timeit("flat U2 NT512", bytes, [&] { copy_flat<2, 512><<<(unsigned)((nvec + 1023) / 1024), 512>>>(a, b, nvec); });
  timeit("flat U8 NT128", bytes, [&] { copy_flat<8, 128><<<(unsigned)((nvec + 1023) / 1024), 128>>>(a, b, nvec); });
  return 0;
}
