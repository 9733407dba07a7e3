// Context, memory helpers, CUDA-event timer.  The ctx is the counterpart of the reference's
// BenchmarkArgs::device_queue (include/command_line.h:129-140): one device + one in-order stream.
#include <cstdio>
#include <cstring>
#include <new>

#include "sb_common.cuh"

extern "C" {

int sb_version(void) { return SB_VERSION; }

const char* sb_strerror(int status) {
  switch(status) {
  case SB_OK: return "ok";
  case SB_ERR_BAD_ARG: return "sb: bad argument";
  case SB_ERR_MISALIGNED: return "sb: pointer not aligned to the element size";
  case SB_ERR_WORKSPACE: return "sb: workspace allocation failed";
  case SB_ERR_SEG_LEN: return "sb: invalid segment length (zero, or problem size not a multiple of it)";
  case SB_ERR_NO_DEVICE: return "sb: no CUDA device (libsb_cuda has no CPU fallback)";
  case SB_ERR_UNSUPPORTED: return "sb: device is not compute capability 10.x (library is sm_100a only)";
  case SB_ERR_TIMER: return "sb: sb_timer_end without sb_timer_begin";
  case SB_ERR_NCCL: return "sb: NCCL unavailable (libnccl.so.2 not loadable) or an NCCL call failed";
  case SB_ERR_TIMEOUT: return "sb: a kernel gave up waiting for another CTA or GPU (results of that call are invalid)";
  case SB_ERR_PEER: return "sb: peer memory unavailable (CUDA IPC / peer access failed, or a peer is not connected)";
  default: break;
  }
  if(status > 0) return cudaGetErrorString((cudaError_t)status);
  return "sb: unknown error";
}

int sb_device_count(int* count) {
  if(!count) return SB_ERR_BAD_ARG;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if(e != cudaSuccess) {
    (void)cudaGetLastError();
    *count = 0;
    return SB_ERR_NO_DEVICE;
  }
  *count = n;
  return SB_OK;
}

int sb_ctx_create(int device, void* stream, sb_ctx** out) {
  if(!out) return SB_ERR_BAD_ARG;
  *out = nullptr;
  int n = 0;
  if(cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    (void)cudaGetLastError();
    return SB_ERR_NO_DEVICE;
  }
  if(device < 0 || device >= n) return SB_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  SB_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if(prop.major != 10) return SB_ERR_UNSUPPORTED;

  sb_device_guard guard(device);
  sb_ctx* ctx = new(std::nothrow) sb_ctx();
  if(!ctx) return SB_ERR_WORKSPACE;
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  if(stream != SB_STREAM_OWNED) {
    ctx->stream = (cudaStream_t)stream;
  } else {
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if(e != cudaSuccess) { delete ctx; return (int)e; }
    ctx->own_stream = true;
  }
  // fixed-size scratch: reduce partials for up to 32 CTAs/SM, tickets
  ctx->partials_slots = (size_t)ctx->sm_count * 32;
  cudaError_t e = cudaMalloc(&ctx->partials, ctx->partials_slots * 8);
  if(e == cudaSuccess) e = cudaMalloc((void**)&ctx->reduce_ticket, 256);
  if(e == cudaSuccess) e = cudaMalloc((void**)&ctx->scan_ticket, 256);
  if(e == cudaSuccess) e = cudaMalloc(&ctx->stage_small, 256);
  if(e == cudaSuccess) e = cudaMemset(ctx->reduce_ticket, 0, 256);
  if(e == cudaSuccess) e = cudaMemset(ctx->scan_ticket, 0, 256);
  if(e == cudaSuccess) e = cudaHostAlloc((void**)&ctx->err_host, 64, cudaHostAllocMapped | cudaHostAllocPortable);
  if(e == cudaSuccess) {
    *ctx->err_host = 0u;
    e = cudaHostGetDevicePointer((void**)&ctx->err_dev, ctx->err_host, 0);
  }
  if(e == cudaSuccess) e = cudaEventCreate(&ctx->ev_begin);
  if(e == cudaSuccess) e = cudaEventCreate(&ctx->ev_end);
  if(e == cudaSuccess) e = cudaDeviceSynchronize();
  if(e != cudaSuccess) {
    sb_ctx_destroy(ctx);
    return e == cudaErrorMemoryAllocation ? SB_ERR_WORKSPACE : (int)e;
  }
  *out = ctx;
  return SB_OK;
}

int sb_ctx_destroy(sb_ctx* ctx) {
  if(!ctx) return SB_OK;
  sb_device_guard guard(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  sb_ctx_release_self_peer(ctx);
  if(ctx->err_host) cudaFreeHost(ctx->err_host);
  cudaFree(ctx->partials);
  cudaFree(ctx->reduce_ticket);
  cudaFree(ctx->scan_ticket);
  cudaFree(ctx->desc64);
  cudaFree(ctx->desc128);
  cudaFree(ctx->stage[0]);
  cudaFree(ctx->stage[1]);
  cudaFree(ctx->stage_small);
  if(ctx->d2h_stream) {
    cudaStreamSynchronize(ctx->d2h_stream);
    cudaStreamDestroy(ctx->d2h_stream);
  }
  for(cudaEvent_t e : ctx->chunk_ev)
    if(e) cudaEventDestroy(e);
  if(ctx->ev_begin) cudaEventDestroy(ctx->ev_begin);
  if(ctx->ev_end) cudaEventDestroy(ctx->ev_end);
  if(ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  (void)cudaGetLastError();
  delete ctx;
  return SB_OK;
}

int sb_ctx_set_stream(sb_ctx* ctx, void* stream) {
  if(!ctx) return SB_ERR_BAD_ARG;
  sb_device_guard guard(ctx->device);
  // tickets, partials and descriptors of the ctx may still be in use on the old stream, owned or borrowed
  SB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if(ctx->own_stream && ctx->stream) {
    cudaStreamDestroy(ctx->stream);
    ctx->own_stream = false;
  }
  if(stream != SB_STREAM_OWNED) {
    ctx->stream = (cudaStream_t)stream;
  } else {
    SB_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
  }
  return SB_OK;
}

void* sb_ctx_stream(sb_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int sb_ctx_sync(sb_ctx* ctx) {
  if(!ctx) return SB_ERR_BAD_ARG;
  sb_device_guard guard(ctx->device);
  SB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return sb_ctx_take_error(ctx);
}

int sb_ctx_device_name(sb_ctx* ctx, char* buf, size_t buflen) {
  if(!ctx || !buf || buflen == 0) return SB_ERR_BAD_ARG;
  cudaDeviceProp prop;
  SB_CUDA_TRY(cudaGetDeviceProperties(&prop, ctx->device));
  std::snprintf(buf, buflen, "%s", prop.name);
  return SB_OK;
}

int sb_ctx_pci_bus_id(sb_ctx* ctx, char* buf, size_t buflen) {
  if(!ctx || !buf || buflen < 16) return SB_ERR_BAD_ARG;
  SB_CUDA_TRY(cudaDeviceGetPCIBusId(buf, (int)buflen, ctx->device));
  return SB_OK;
}

int sb_ctx_set_deterministic(sb_ctx* ctx, int on) {
  if(!ctx) return SB_ERR_BAD_ARG;
  ctx->deterministic = on != 0;
  return SB_OK;
}

int sb_ctx_sm_count(sb_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
uint64_t sb_ctx_launch_count(sb_ctx* ctx) { return ctx ? ctx->launches : 0; }

int sb_timer_begin(sb_ctx* ctx) {
  if(!ctx) return SB_ERR_BAD_ARG;
  sb_device_guard guard(ctx->device);
  SB_CUDA_TRY(cudaEventRecord(ctx->ev_begin, ctx->stream));
  ctx->timing = true;
  return SB_OK;
}

int sb_timer_end(sb_ctx* ctx, float* elapsed_ms) {
  if(!ctx || !elapsed_ms) return SB_ERR_BAD_ARG;
  if(!ctx->timing) return SB_ERR_TIMER;
  sb_device_guard guard(ctx->device);
  SB_CUDA_TRY(cudaEventRecord(ctx->ev_end, ctx->stream));
  SB_CUDA_TRY(cudaEventSynchronize(ctx->ev_end));
  SB_CUDA_TRY(cudaEventElapsedTime(elapsed_ms, ctx->ev_begin, ctx->ev_end));
  ctx->timing = false;
  return sb_ctx_take_error(ctx);
}

int sb_malloc(sb_ctx* ctx, void** dptr, size_t bytes) {
  if(!ctx || !dptr) return SB_ERR_BAD_ARG;
  sb_device_guard guard(ctx->device);
  *dptr = nullptr;
  if(bytes == 0) bytes = 16;
  SB_CUDA_TRY(cudaMalloc(dptr, bytes));
  return SB_OK;
}

int sb_free(sb_ctx* ctx, void* dptr) {
  if(!ctx) return SB_ERR_BAD_ARG;
  sb_device_guard guard(ctx->device);
  SB_CUDA_TRY(cudaFree(dptr));
  return SB_OK;
}

int sb_host_alloc(sb_ctx* ctx, void** hptr, size_t bytes) {
  if(!ctx || !hptr) return SB_ERR_BAD_ARG;
  sb_device_guard guard(ctx->device);
  *hptr = nullptr;
  if(bytes == 0) bytes = 16;
  SB_CUDA_TRY(cudaHostAlloc(hptr, bytes, cudaHostAllocMapped | cudaHostAllocPortable));
  return SB_OK;
}

int sb_host_free(sb_ctx* ctx, void* hptr) {
  if(!ctx) return SB_ERR_BAD_ARG;
  sb_device_guard guard(ctx->device);
  SB_CUDA_TRY(cudaFreeHost(hptr));
  return SB_OK;
}

int sb_memcpy_h2d(sb_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if(!ctx || (bytes && (!dst || !src))) return SB_ERR_BAD_ARG;
  sb_device_guard guard(ctx->device);
  SB_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return SB_OK;
}

int sb_memcpy_d2h(sb_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if(!ctx || (bytes && (!dst || !src))) return SB_ERR_BAD_ARG;
  sb_device_guard guard(ctx->device);
  SB_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return SB_OK;
}

int sb_memset(sb_ctx* ctx, void* dst, int byte, size_t bytes) {
  if(!ctx || (bytes && !dst)) return SB_ERR_BAD_ARG;
  sb_device_guard guard(ctx->device);
  SB_CUDA_TRY(cudaMemsetAsync(dst, byte, bytes, ctx->stream));
  return SB_OK;
}

}  // extern "C"

int sb_ctx_take_error(sb_ctx* ctx) {
  if(ctx->err_host && *(volatile unsigned*)ctx->err_host) {
    *(volatile unsigned*)ctx->err_host = 0u;
    return SB_ERR_TIMEOUT;
  }
  return SB_OK;
}

// ---- workspace growth (amortised; never shrinks) -------------------------------------------------

static int grow(sb_ctx* ctx, void** ptr, size_t* have, size_t want, size_t elem_bytes) {
  if(want <= *have) return SB_OK;
  // growth implies the old array is idle only after the stream drains
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if(e != cudaSuccess) return (int)e;
  if(*ptr) cudaFree(*ptr);
  *ptr = nullptr;
  *have = 0;
  size_t cap = want + want / 4 + 1024;
  e = cudaMalloc(ptr, cap * elem_bytes);
  if(e != cudaSuccess) {
    (void)cudaGetLastError();
    return SB_ERR_WORKSPACE;
  }
  e = cudaMemsetAsync(*ptr, 0, cap * elem_bytes, ctx->stream);
  if(e != cudaSuccess) return (int)e;
  *have = cap;
  return SB_OK;
}

int sb_ctx_reserve_desc(sb_ctx* ctx, size_t count64, size_t count128) {
  int rc = grow(ctx, &ctx->desc64, &ctx->desc64_count, count64, 8);
  if(rc != SB_OK) return rc;
  return grow(ctx, &ctx->desc128, &ctx->desc128_count, count128, 16);
}

int sb_ctx_reserve_partials(sb_ctx* ctx, size_t slots) {
  if(slots <= ctx->partials_slots) return SB_OK;
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if(e != cudaSuccess) return (int)e;
  cudaFree(ctx->partials);
  ctx->partials = nullptr;
  ctx->partials_slots = 0;
  const size_t cap = slots + slots / 4;
  e = cudaMalloc(&ctx->partials, cap * 8);
  if(e != cudaSuccess) {
    (void)cudaGetLastError();
    return SB_ERR_WORKSPACE;
  }
  ctx->partials_slots = cap;
  return SB_OK;
}

int sb_ctx_reserve_stage(sb_ctx* ctx, int which, size_t bytes) {
  if(bytes <= ctx->stage_bytes[which]) return SB_OK;
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if(e != cudaSuccess) return (int)e;
  if(ctx->stage[which]) cudaFree(ctx->stage[which]);
  ctx->stage[which] = nullptr;
  ctx->stage_bytes[which] = 0;
  e = cudaMalloc(&ctx->stage[which], bytes);
  if(e != cudaSuccess) {
    (void)cudaGetLastError();
    return SB_ERR_WORKSPACE;
  }
  ctx->stage_bytes[which] = bytes;
  return SB_OK;
}
