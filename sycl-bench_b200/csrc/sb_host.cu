// Host-buffer entry points: the call shape of the reference plugin, where setup() uploads a host
// std::vector (/root/reference/pattern/reduction.cpp:36-42, include/memory_wrappers.h:47-57) and
// verify() reads the result back through a host accessor (pattern/reduction.cpp:60,
// pattern/segmentedreduction.cpp:102).  The call returns after the result is in host memory.
// Staging buffers are cached in the ctx.
//
// These calls are PCIe-bound (~55 GB/s per direction against ~6 TB/s in HBM), so the only thing
// that matters is keeping both directions of the link busy: operations that return an array are
// chunked (64 MiB); chunk k is uploaded and processed on the ctx stream while chunk k-1 is
// downloaded on a second stream.  A global scan carries its running total from chunk to chunk as a
// device-resident seed (sb_scan's `seed`), restarting scans and the segmented operations cut the
// chunks on segment boundaries.
#include "sb_common.cuh"

namespace {

constexpr size_t kChunkBytes = size_t(64) << 20;

size_t dtype_size(int dtype) {
  switch(dtype) {
  case SB_I16: return 2;
  case SB_I32: case SB_F32: return 4;
  case SB_I64: case SB_F64: return 8;
  default: return 0;
  }
}

int ensure_d2h_stream(sb_ctx* ctx) {
  if(!ctx->d2h_stream) SB_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
  for(cudaEvent_t& e : ctx->chunk_ev)
    if(!e) SB_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return SB_OK;
}

// after the kernels of chunk k were enqueued on the ctx stream: download [dst, dst+bytes) on the d2h stream
int download_chunk(sb_ctx* ctx, int k, void* h_dst, const void* d_src, size_t bytes) {
  cudaEvent_t ev = ctx->chunk_ev[k % sb_ctx::kChunkEvents];
  SB_CUDA_TRY(cudaEventRecord(ev, ctx->stream));
  SB_CUDA_TRY(cudaStreamWaitEvent(ctx->d2h_stream, ev, 0));
  SB_CUDA_TRY(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->d2h_stream));
  return SB_OK;
}

int finish(sb_ctx* ctx) {
  SB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if(ctx->d2h_stream) SB_CUDA_TRY(cudaStreamSynchronize(ctx->d2h_stream));
  return sb_ctx_take_error(ctx);
}

// error exit of a chunked call: copies already queued still write into the caller's host buffer, so drain both
// streams before handing the buffer back
int fail(sb_ctx* ctx, int rc) {
  cudaStreamSynchronize(ctx->stream);
  if(ctx->d2h_stream) cudaStreamSynchronize(ctx->d2h_stream);
  (void)cudaGetLastError();
  return rc;
}

template <typename A>
__global__ void next_seed_kernel(const A* __restrict__ last_in, const A* __restrict__ last_out, int exclusive, A* __restrict__ seed) {
  // running total after a chunk: the last inclusive value (an exclusive result lacks the last input)
  *seed = exclusive ? (A)(*last_out + *last_in) : *last_out;
}

int scan_call(sb_ctx* ctx, int dtype, const void* in, void* out, size_t n, int exclusive, size_t group_len, const void* seed) {
  switch(dtype) {
  case SB_I32: return sb_scan_i32(ctx, (const int32_t*)in, (int32_t*)out, n, exclusive, group_len, (const int32_t*)seed);
  case SB_I64: return sb_scan_i64(ctx, (const int64_t*)in, (int64_t*)out, n, exclusive, group_len, (const int64_t*)seed);
  case SB_F32: return sb_scan_f32(ctx, (const float*)in, (float*)out, n, exclusive, group_len, (const float*)seed);
  default: return sb_scan_f64(ctx, (const double*)in, (double*)out, n, exclusive, group_len, (const double*)seed);
  }
}

int segscan_call(sb_ctx* ctx, int dtype, const void* in, void* out, size_t n, size_t seg_len, const uint8_t* flags, int exclusive) {
  switch(dtype) {
  case SB_I32: return sb_segscan_i32(ctx, (const int32_t*)in, (int32_t*)out, n, seg_len, flags, exclusive);
  case SB_I64: return sb_segscan_i64(ctx, (const int64_t*)in, (int64_t*)out, n, seg_len, flags, exclusive);
  case SB_F32: return sb_segscan_f32(ctx, (const float*)in, (float*)out, n, seg_len, flags, exclusive);
  default: return sb_segscan_f64(ctx, (const double*)in, (double*)out, n, seg_len, flags, exclusive);
  }
}

int segreduce_call(sb_ctx* ctx, int dtype, void* buf, size_t n, size_t seg_len) {
  switch(dtype) {
  case SB_I16: return sb_segreduce_i16(ctx, (int16_t*)buf, n, seg_len);
  case SB_I32: return sb_segreduce_i32(ctx, (int32_t*)buf, n, seg_len);
  case SB_I64: return sb_segreduce_i64(ctx, (int64_t*)buf, n, seg_len);
  case SB_F32: return sb_segreduce_f32(ctx, (float*)buf, n, seg_len);
  default: return sb_segreduce_f64(ctx, (double*)buf, n, seg_len);
  }
}

void launch_next_seed(sb_ctx* ctx, int dtype, const void* last_in, const void* last_out, int exclusive, void* seed) {
  switch(dtype) {
  case SB_I32: next_seed_kernel<uint32_t><<<1, 1, 0, ctx->stream>>>((const uint32_t*)last_in, (const uint32_t*)last_out, exclusive, (uint32_t*)seed); break;
  case SB_I64: next_seed_kernel<unsigned long long><<<1, 1, 0, ctx->stream>>>((const unsigned long long*)last_in, (const unsigned long long*)last_out, exclusive, (unsigned long long*)seed); break;
  case SB_F32: next_seed_kernel<float><<<1, 1, 0, ctx->stream>>>((const float*)last_in, (const float*)last_out, exclusive, (float*)seed); break;
  default: next_seed_kernel<double><<<1, 1, 0, ctx->stream>>>((const double*)last_in, (const double*)last_out, exclusive, (double*)seed); break;
  }
  ctx->launches++;
}

// elements per chunk: ~64 MiB, a multiple of `align` elements (segment length) and of 16 bytes
size_t chunk_elems(size_t s, size_t align) {
  size_t a = align ? align : 1, b = 16;
  while(b) {  // gcd(align, 16)
    const size_t t = a % b;
    a = b;
    b = t;
  }
  const size_t unit = (align ? align : 1) / a * 16;  // lcm(align, 16)
  return (kChunkBytes / s) / unit * unit;            // 0 when one unit does not fit: caller takes one piece
}

}  // namespace

extern "C" {

int sb_host_reduce(sb_ctx* ctx, int dtype, const void* h_in, size_t n, void* h_out) {
  const size_t s = dtype_size(dtype);
  if(!ctx || !h_out || (n && !h_in) || s == 0 || dtype == SB_I16) return SB_ERR_BAD_ARG;
  sb_device_guard guard(ctx->device);
  int rc = sb_ctx_reserve_stage(ctx, 0, n * s + 16);
  if(rc != SB_OK) return rc;
  void* d_in = ctx->stage[0];
  void* d_out = ctx->stage_small;
  if(n) SB_CUDA_TRY(cudaMemcpyAsync(d_in, h_in, n * s, cudaMemcpyHostToDevice, ctx->stream));
  switch(dtype) {
  case SB_I32: rc = sb_reduce_i32(ctx, (const int32_t*)d_in, n, (int32_t*)d_out); break;
  case SB_I64: rc = sb_reduce_i64(ctx, (const int64_t*)d_in, n, (int64_t*)d_out); break;
  case SB_F32: rc = sb_reduce_f32(ctx, (const float*)d_in, n, (float*)d_out); break;
  default: rc = sb_reduce_f64(ctx, (const double*)d_in, n, (double*)d_out); break;
  }
  if(rc != SB_OK) return fail(ctx, rc);
  SB_CUDA_TRY(cudaMemcpyAsync(h_out, d_out, s, cudaMemcpyDeviceToHost, ctx->stream));
  return finish(ctx);
}

int sb_host_scan(sb_ctx* ctx, int dtype, const void* h_in, void* h_out, size_t n, int exclusive, size_t group_len) {
  const size_t s = dtype_size(dtype);
  if(!ctx || (n && (!h_in || !h_out)) || s == 0 || dtype == SB_I16) return SB_ERR_BAD_ARG;
  if(n == 0) return SB_OK;
  sb_device_guard guard(ctx->device);
  int rc = sb_ctx_reserve_stage(ctx, 0, n * s + 16);
  if(rc == SB_OK) rc = sb_ctx_reserve_stage(ctx, 1, n * s + 16);
  if(rc == SB_OK) rc = ensure_d2h_stream(ctx);
  if(rc != SB_OK) return rc;
  char* d_in = (char*)ctx->stage[0];
  char* d_out = (char*)ctx->stage[1];
  const char* hi = (const char*)h_in;
  char* ho = (char*)h_out;
  const bool global = group_len == 0 || group_len >= n;
  size_t chunk = chunk_elems(s, global ? 1 : group_len);
  if(chunk == 0) chunk = n;  // a restart interval above the chunk size: one piece
  char* seed = (char*)ctx->stage_small;
  int k = 0;
  for(size_t b = 0; b < n; b += chunk, ++k) {
    const size_t m = (n - b < chunk) ? n - b : chunk;
    if(k >= sb_ctx::kChunkEvents) SB_CUDA_TRY(cudaEventSynchronize(ctx->chunk_ev[k % sb_ctx::kChunkEvents]));
    SB_CUDA_TRY(cudaMemcpyAsync(d_in + b * s, hi + b * s, m * s, cudaMemcpyHostToDevice, ctx->stream));
    rc = scan_call(ctx, dtype, d_in + b * s, d_out + b * s, m, exclusive, global ? 0 : group_len, (global && k > 0) ? seed : nullptr);
    if(rc != SB_OK) return fail(ctx, rc);
    if(global && b + m < n) launch_next_seed(ctx, dtype, d_in + (b + m - 1) * s, d_out + (b + m - 1) * s, exclusive, seed);
    rc = download_chunk(ctx, k, ho + b * s, d_out + b * s, m * s);
    if(rc != SB_OK) return fail(ctx, rc);
  }
  return finish(ctx);
}

int sb_host_segreduce(sb_ctx* ctx, int dtype, void* h_inout, size_t n, size_t seg_len) {
  const size_t s = dtype_size(dtype);
  if(!ctx || (n && !h_inout) || s == 0) return SB_ERR_BAD_ARG;
  if(seg_len == 0 || n % seg_len != 0) return SB_ERR_SEG_LEN;
  if(n == 0) return SB_OK;
  sb_device_guard guard(ctx->device);
  int rc = sb_ctx_reserve_stage(ctx, 0, n * s + 16);
  if(rc == SB_OK) rc = ensure_d2h_stream(ctx);
  if(rc != SB_OK) return rc;
  char* d = (char*)ctx->stage[0];
  char* h = (char*)h_inout;
  size_t chunk = chunk_elems(s, seg_len);
  if(chunk == 0) chunk = n;
  int k = 0;
  for(size_t b = 0; b < n; b += chunk, ++k) {
    const size_t m = (n - b < chunk) ? n - b : chunk;
    if(k >= sb_ctx::kChunkEvents) SB_CUDA_TRY(cudaEventSynchronize(ctx->chunk_ev[k % sb_ctx::kChunkEvents]));
    SB_CUDA_TRY(cudaMemcpyAsync(d + b * s, h + b * s, m * s, cudaMemcpyHostToDevice, ctx->stream));
    rc = segreduce_call(ctx, dtype, d + b * s, m, seg_len);
    if(rc != SB_OK) return fail(ctx, rc);
    // only the heads changed, but the reference reads the whole buffer back (verify(), :102)
    rc = download_chunk(ctx, k, h + b * s, d + b * s, m * s);
    if(rc != SB_OK) return fail(ctx, rc);
  }
  return finish(ctx);
}

int sb_host_segscan(sb_ctx* ctx, int dtype, const void* h_in, void* h_out, size_t n, size_t seg_len, const uint8_t* h_flags, int exclusive) {
  const size_t s = dtype_size(dtype);
  if(!ctx || (n && (!h_in || !h_out)) || s == 0 || dtype == SB_I16) return SB_ERR_BAD_ARG;
  if(!h_flags && seg_len == 0) return SB_ERR_SEG_LEN;
  if(n == 0) return SB_OK;
  sb_device_guard guard(ctx->device);
  int rc = sb_ctx_reserve_stage(ctx, 0, n * s + 16);
  if(rc == SB_OK) rc = ensure_d2h_stream(ctx);
  if(rc != SB_OK) return rc;
  char* d = (char*)ctx->stage[0];
  const char* hi = (const char*)h_in;
  char* ho = (char*)h_out;
  if(h_flags || seg_len >= n) {
    // explicit flags (segments may straddle any cut) or a single segment: one piece, in place
    uint8_t* d_flags = nullptr;
    if(h_flags) {
      rc = sb_ctx_reserve_stage(ctx, 1, n + 16);
      if(rc != SB_OK) return rc;
      d_flags = (uint8_t*)ctx->stage[1];
      SB_CUDA_TRY(cudaMemcpyAsync(d_flags, h_flags, n, cudaMemcpyHostToDevice, ctx->stream));
    }
    SB_CUDA_TRY(cudaMemcpyAsync(d, h_in, n * s, cudaMemcpyHostToDevice, ctx->stream));
    rc = segscan_call(ctx, dtype, d, d, n, seg_len, d_flags, exclusive);
    if(rc != SB_OK) return fail(ctx, rc);
    SB_CUDA_TRY(cudaMemcpyAsync(h_out, d, n * s, cudaMemcpyDeviceToHost, ctx->stream));
    return finish(ctx);
  }
  size_t chunk = chunk_elems(s, seg_len);
  if(chunk == 0) chunk = n;
  int k = 0;
  for(size_t b = 0; b < n; b += chunk, ++k) {
    const size_t m = (n - b < chunk) ? n - b : chunk;
    if(k >= sb_ctx::kChunkEvents) SB_CUDA_TRY(cudaEventSynchronize(ctx->chunk_ev[k % sb_ctx::kChunkEvents]));
    SB_CUDA_TRY(cudaMemcpyAsync(d + b * s, hi + b * s, m * s, cudaMemcpyHostToDevice, ctx->stream));
    rc = segscan_call(ctx, dtype, d + b * s, d + b * s, m, seg_len, nullptr, exclusive);
    if(rc != SB_OK) return fail(ctx, rc);
    rc = download_chunk(ctx, k, ho + b * s, d + b * s, m * s);
    if(rc != SB_OK) return fail(ctx, rc);
  }
  return finish(ctx);
}

// descriptors an op on n elements needs at most: the smallest tile any scan kernel uses gives the largest count
static size_t scan_desc_upper_bound(size_t n, size_t s) {
  const size_t tile = sb_scan_min_tile_bytes() / s;
  return 2 * ((n + tile - 1) / tile) + 64;  // tile descriptors + one prefix per round + slack
}

size_t sb_workspace_bytes(int op, int dtype, size_t n) {
  const size_t s = dtype_size(dtype);
  if(s == 0) return 0;
  switch(op) {
  case SB_OP_REDUCE: return 256 * 32 * 8 + 256;  // block partials for up to 256 SMs + tickets
  case SB_OP_SCAN:
  case SB_OP_SEGSCAN: return scan_desc_upper_bound(n, s) * (s == 4 ? 8 : 16) + 256;
  case SB_OP_SEGREDUCE: return 0;
  default: return 0;
  }
}

int sb_ctx_reserve(sb_ctx* ctx, int op, int dtype, size_t n) {
  const size_t s = dtype_size(dtype);
  if(!ctx || s == 0) return SB_ERR_BAD_ARG;
  if(op != SB_OP_SCAN && op != SB_OP_SEGSCAN) return SB_OK;  // everything else uses the fixed-size scratch of sb_ctx_create
  sb_device_guard guard(ctx->device);
  const size_t count = scan_desc_upper_bound(n, s);
  return sb_ctx_reserve_desc(ctx, s == 4 ? count : 0, s == 8 ? count : 0);
}

}  // extern "C"
