// Host-buffer entry points: the call shape of the reference plugin, where setup() uploads a host
// std::vector (/root/reference/pattern/reduction.cpp:36-42, include/memory_wrappers.h:47-57) and
// verify() reads the result back through a host accessor (pattern/reduction.cpp:60,
// pattern/segmentedreduction.cpp:102).  H2D, the kernel and D2H all run on the ctx stream; the
// call returns after the result is in host memory.  Staging buffers are cached in the ctx.
#include "sb_common.cuh"

namespace {

size_t dtype_size(int dtype) {
  switch(dtype) {
  case SB_I16: return 2;
  case SB_I32: case SB_F32: return 4;
  case SB_I64: case SB_F64: return 8;
  default: return 0;
  }
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
  if(rc != SB_OK) return rc;
  SB_CUDA_TRY(cudaMemcpyAsync(h_out, d_out, s, cudaMemcpyDeviceToHost, ctx->stream));
  SB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return SB_OK;
}

int sb_host_scan(sb_ctx* ctx, int dtype, const void* h_in, void* h_out, size_t n, int exclusive, size_t group_len) {
  const size_t s = dtype_size(dtype);
  if(!ctx || (n && (!h_in || !h_out)) || s == 0 || dtype == SB_I16) return SB_ERR_BAD_ARG;
  if(n == 0) return SB_OK;
  sb_device_guard guard(ctx->device);
  int rc = sb_ctx_reserve_stage(ctx, 0, n * s + 16);
  if(rc != SB_OK) return rc;
  void* d = ctx->stage[0];
  SB_CUDA_TRY(cudaMemcpyAsync(d, h_in, n * s, cudaMemcpyHostToDevice, ctx->stream));
  switch(dtype) {
  case SB_I32: rc = sb_scan_i32(ctx, (const int32_t*)d, (int32_t*)d, n, exclusive, group_len, nullptr); break;
  case SB_I64: rc = sb_scan_i64(ctx, (const int64_t*)d, (int64_t*)d, n, exclusive, group_len, nullptr); break;
  case SB_F32: rc = sb_scan_f32(ctx, (const float*)d, (float*)d, n, exclusive, group_len, nullptr); break;
  default: rc = sb_scan_f64(ctx, (const double*)d, (double*)d, n, exclusive, group_len, nullptr); break;
  }
  if(rc != SB_OK) return rc;
  SB_CUDA_TRY(cudaMemcpyAsync(h_out, d, n * s, cudaMemcpyDeviceToHost, ctx->stream));
  SB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return SB_OK;
}

int sb_host_segreduce(sb_ctx* ctx, int dtype, void* h_inout, size_t n, size_t seg_len) {
  const size_t s = dtype_size(dtype);
  if(!ctx || (n && !h_inout) || s == 0) return SB_ERR_BAD_ARG;
  if(seg_len == 0 || n % seg_len != 0) return SB_ERR_SEG_LEN;
  if(n == 0) return SB_OK;
  sb_device_guard guard(ctx->device);
  int rc = sb_ctx_reserve_stage(ctx, 0, n * s + 16);
  if(rc != SB_OK) return rc;
  void* d = ctx->stage[0];
  SB_CUDA_TRY(cudaMemcpyAsync(d, h_inout, n * s, cudaMemcpyHostToDevice, ctx->stream));
  switch(dtype) {
  case SB_I16: rc = sb_segreduce_i16(ctx, (int16_t*)d, n, seg_len); break;
  case SB_I32: rc = sb_segreduce_i32(ctx, (int32_t*)d, n, seg_len); break;
  case SB_I64: rc = sb_segreduce_i64(ctx, (int64_t*)d, n, seg_len); break;
  case SB_F32: rc = sb_segreduce_f32(ctx, (float*)d, n, seg_len); break;
  default: rc = sb_segreduce_f64(ctx, (double*)d, n, seg_len); break;
  }
  if(rc != SB_OK) return rc;
  // only the heads changed, but the reference reads the whole buffer back (verify(), :102)
  SB_CUDA_TRY(cudaMemcpyAsync(h_inout, d, n * s, cudaMemcpyDeviceToHost, ctx->stream));
  SB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return SB_OK;
}

int sb_host_segscan(sb_ctx* ctx, int dtype, const void* h_in, void* h_out, size_t n, size_t seg_len, const uint8_t* h_flags, int exclusive) {
  const size_t s = dtype_size(dtype);
  if(!ctx || (n && (!h_in || !h_out)) || s == 0 || dtype == SB_I16) return SB_ERR_BAD_ARG;
  if(!h_flags && seg_len == 0) return SB_ERR_SEG_LEN;
  if(n == 0) return SB_OK;
  sb_device_guard guard(ctx->device);
  int rc = sb_ctx_reserve_stage(ctx, 0, n * s + 16);
  if(rc != SB_OK) return rc;
  uint8_t* d_flags = nullptr;
  if(h_flags) {
    rc = sb_ctx_reserve_stage(ctx, 1, n + 16);
    if(rc != SB_OK) return rc;
    d_flags = (uint8_t*)ctx->stage[1];
    SB_CUDA_TRY(cudaMemcpyAsync(d_flags, h_flags, n, cudaMemcpyHostToDevice, ctx->stream));
  }
  void* d = ctx->stage[0];
  SB_CUDA_TRY(cudaMemcpyAsync(d, h_in, n * s, cudaMemcpyHostToDevice, ctx->stream));
  switch(dtype) {
  case SB_I32: rc = sb_segscan_i32(ctx, (const int32_t*)d, (int32_t*)d, n, seg_len, d_flags, exclusive); break;
  case SB_I64: rc = sb_segscan_i64(ctx, (const int64_t*)d, (int64_t*)d, n, seg_len, d_flags, exclusive); break;
  case SB_F32: rc = sb_segscan_f32(ctx, (const float*)d, (float*)d, n, seg_len, d_flags, exclusive); break;
  default: rc = sb_segscan_f64(ctx, (const double*)d, (double*)d, n, seg_len, d_flags, exclusive); break;
  }
  if(rc != SB_OK) return rc;
  SB_CUDA_TRY(cudaMemcpyAsync(h_out, d, n * s, cudaMemcpyDeviceToHost, ctx->stream));
  SB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return SB_OK;
}

size_t sb_workspace_bytes(int op, int dtype, size_t n) {
  const size_t s = dtype_size(dtype);
  if(s == 0) return 0;
  switch(op) {
  case SB_OP_REDUCE: return 148 * 32 * 8 + 256;  // block partials + ticket
  case SB_OP_SCAN:
  case SB_OP_SEGSCAN: {
    const size_t tile = sb_scan_tile_bytes() / s;
    const size_t tiles = (n + tile - 1) / tile;
    return (2 * tiles + 64) * (s == 4 ? 8 : 16) + 256;  // tile descriptors + per-round prefixes + ticket
  }
  case SB_OP_SEGREDUCE: return 0;
  default: return 0;
  }
}

}  // extern "C"
