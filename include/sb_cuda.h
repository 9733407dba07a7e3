/* libsb_cuda — C ABI of the B200-native pattern/ reduce/scan library.
 *
 * This header is the lower drop-in boundary.  Every entry point states the
 * reference interface it replaces as `file:line` inside /root/reference (unisa-hpc/sycl-bench
 * 2.0.4).  The reference has no library layer: its device code is inline SYCL lambdas inside the
 * benchmark classes, submitted through `sycl::queue::submit`.  libsb_cuda is what a maintainer
 * binds behind those `submit` call sites (INTEGRATION.md shows the exact stubs).
 *
 * Conventions
 *  - plain C, `extern "C"`, pointers + sizes only; no C++/torch types.
 *  - every data pointer of an `sb_<op>_<T>` call is a DEVICE-accessible pointer owned by the
 *    caller (cudaMalloc, or pinned+mapped host memory); the library never allocates per call.
 *  - calls are ASYNCHRONOUS on the ctx stream (like `queue.submit`, include/common.h:70-75 waits
 *    afterwards); results stay in device memory (pattern/reduction.cpp:60,96).
 *  - return value: 0 = ok; negative = sb_status below; positive = cudaError_t passed through.
 *    No C++ exception crosses this ABI.
 *  - a ctx is single-threaded; distinct ctxs are independent.
 *  - there is NO CPU fallback: without a CUDA device sb_ctx_create fails with SB_ERR_NO_DEVICE.
 *
 * Type suffixes follow include/type_traits.h:14-23 of the reference:
 *    i16 = short ("int16"), i32 = int ("int32"), i64 = long long ("int64"),
 *    f32 = float ("fp32"),  f64 = double ("fp64").
 */
#ifndef SB_CUDA_H
#define SB_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define SB_VERSION 100 /* 0.1.0 */

typedef struct sb_ctx sb_ctx;
#define SB_STREAM_OWNED ((void*)(intptr_t)-1)

enum sb_status {
  SB_OK = 0,
  SB_ERR_BAD_ARG = -1,     /* null pointer, n too large, bad enum                              */
  SB_ERR_MISALIGNED = -2,  /* pointer not aligned to sizeof(T)                                 */
  SB_ERR_WORKSPACE = -3,   /* internal workspace allocation failed                             */
  SB_ERR_SEG_LEN = -4,     /* seg_len == 0, or n % seg_len != 0 where the reference asserts it */
  SB_ERR_NO_DEVICE = -5,   /* no CUDA device / device index out of range                       */
  SB_ERR_UNSUPPORTED = -6, /* device is not sm_100 (library is built for sm_100a only)         */
  SB_ERR_TIMER = -7,       /* sb_timer_end without sb_timer_begin                              */
  SB_ERR_NCCL = -8,        /* libnccl.so.2 could not be loaded, or an NCCL call failed         */
  SB_ERR_TIMEOUT = -9,     /* a kernel gave up waiting for another CTA / GPU; reported once by the next
                              sb_ctx_sync / sb_timer_end / sb_host_* call; results of that op are invalid */
  SB_ERR_PEER = -10        /* CUDA IPC / peer access failed, or a peer of the group is not connected   */
};

enum sb_dtype { SB_I16 = 0, SB_I32 = 1, SB_I64 = 2, SB_F32 = 3, SB_F64 = 4 };
enum sb_op { SB_OP_REDUCE = 0, SB_OP_SCAN = 1, SB_OP_SEGREDUCE = 2, SB_OP_SEGSCAN = 3 };

/* ---- library / context ---------------------------------------------------------------------- */

int sb_version(void);
const char* sb_strerror(int status);

/* Number of CUDA devices (replaces the device selectors of include/command_line.h:182-202). */
int sb_device_count(int* count);

/* One ctx per (device, stream): the counterpart of `BenchmarkArgs::device_queue`
 * (include/command_line.h:133,152-154).  `stream` is a cudaStream_t: NULL is the CUDA default
 * stream (as in every CUDA API), SB_STREAM_OWNED asks for a new non-blocking stream owned by the
 * ctx (the in-order queue of include/command_line.h:154).  The ctx owns the scratch the kernels need: block partials
 * and tickets have a fixed size; the tile descriptors of the scans grow on demand — the first scan of a given size
 * on a ctx synchronises the stream and reallocates them (never under stream capture, never shrinking).  Call
 * sb_ctx_reserve from setup() to keep that out of a timed or captured region. */
int sb_ctx_create(int device, void* stream, sb_ctx** out);
int sb_ctx_destroy(sb_ctx* ctx);
int sb_ctx_set_stream(sb_ctx* ctx, void* stream);
void* sb_ctx_stream(sb_ctx* ctx);
/* queue.wait_and_throw(): include/common.h:61,75. */
int sb_ctx_sync(sb_ctx* ctx);
/* device-name CSV column: include/common.h:42-43. */
int sb_ctx_device_name(sb_ctx* ctx, char* buf, size_t buflen);
int sb_ctx_sm_count(sb_ctx* ctx);
/* PCI bus id of the ctx device ("0000:1b:00.0"): lets a caller find the same device in NVML, whose indices
 * ignore CUDA_VISIBLE_DEVICES (used by the harness's energy / clock hook, the reference's empty
 * include/nv_energy_meas.h). */
int sb_ctx_pci_bus_id(sb_ctx* ctx, char* buf, size_t buflen);
/* Number of kernels this ctx has launched so far (bench.py's gpu_launches). */
uint64_t sb_ctx_launch_count(sb_ctx* ctx);
/* on != 0: floating-point sb_reduce_* uses a fixed summation order (bit-reproducible run to run,
 * about 3 % slower).  Default off: chunks are handed to CTAs in order of arrival.  Integer results
 * and all scans are reproducible either way. */
int sb_ctx_set_deterministic(sb_ctx* ctx, int on);
/* Upper bound of the device scratch an op on n elements of `dtype` touches on any device this library supports
 * (the scans pick their tile size by alignment and size; the bound assumes the smallest tile). */
size_t sb_workspace_bytes(int op, int dtype, size_t n);
/* Grow the ctx scratch now to what `op` on n elements can need (sb_workspace_bytes), so that the op itself never
 * allocates: the place for it is setup() (include/common.h:55-59), outside the timed run(). */
int sb_ctx_reserve(sb_ctx* ctx, int op, int dtype, size_t n);

/* CUDA-event timing on the ctx stream; replaces the SYCL event profiling of
 * include/common.h:83-101 (kernel-time).  begin/end may bracket any number of sb_* calls. */
int sb_timer_begin(sb_ctx* ctx);
int sb_timer_end(sb_ctx* ctx, float* elapsed_ms); /* synchronises on the end event */

/* ---- memory helpers (PrefetchedBuffer, include/memory_wrappers.h:39-92) --------------------- */

int sb_malloc(sb_ctx* ctx, void** dptr, size_t bytes);       /* initialize(q, range)  :42-45 */
int sb_free(sb_ctx* ctx, void* dptr);
int sb_host_alloc(sb_ctx* ctx, void** hptr, size_t bytes);   /* pinned + mapped host memory   */
int sb_host_free(sb_ctx* ctx, void* hptr);
int sb_memcpy_h2d(sb_ctx* ctx, void* dst, const void* src, size_t bytes); /* :47-57, async */
int sb_memcpy_d2h(sb_ctx* ctx, void* dst, const void* src, size_t bytes); /* get_host_access :81 */
int sb_memset(sb_ctx* ctx, void* dst, int byte, size_t bytes);

/* ---- generate_input on the device ----------------------------------------------------------- */
/* out[i] = static_cast<T>(offset + i)   (pattern/reduction.cpp:31-34,
 * pattern/segmentedreduction.cpp:25-28).  `offset` is the global index of the shard's first
 * element (multi-GPU).  Integer types wrap, fp32/fp64 round to nearest even like x86 cvtsi2ss. */
int sb_iota_i16(sb_ctx*, int16_t* out, size_t n, uint64_t offset);
int sb_iota_i32(sb_ctx*, int32_t* out, size_t n, uint64_t offset);
int sb_iota_i64(sb_ctx*, int64_t* out, size_t n, uint64_t offset);
int sb_iota_f32(sb_ctx*, float* out, size_t n, uint64_t offset);
int sb_iota_f64(sb_ctx*, double* out, size_t n, uint64_t offset);

/* out[i] = value: the device-side form of `std::fill(in_vec.begin(), in_vec.end(), 1)` + upload in the
 * sycl2020 reduction trio (sycl2020/atomics/atomic_reduction.cpp:19-27,
 * sycl2020/group_algorithms/reduce_over_group.cpp:21-29, sycl2020/kernel_reduction/kernel_reduction.cpp:19-27). */
int sb_fill_i32(sb_ctx*, int32_t* out, size_t n, int32_t value);
int sb_fill_i64(sb_ctx*, int64_t* out, size_t n, int64_t value);
int sb_fill_f32(sb_ctx*, float* out, size_t n, float value);
int sb_fill_f64(sb_ctx*, double* out, size_t n, double value);

/* dst[0..bytes) = src[0..bytes), device to device, one launch for 16-byte congruent pointers.  The copy
 * kernel of micro/DRAM.cpp:54-62 (`out[gid] = in[gid]`, one work-item per element, in a 1-, 2- or 3-D
 * index space over the same contiguous buffer): the in-suite calibrator for the HBM roofline.
 * Algorithmic bytes 2 * bytes.  The ranges must not overlap. */
int sb_copy(sb_ctx*, void* dst, const void* src, size_t bytes);

/* ---- reduction ------------------------------------------------------------------------------ */
/* out[0] = sum(in[0..n)), one launch.  Replaces the whole multi-pass driver
 * Reduction<T>::submit (pattern/reduction.cpp:71-97) and both kernels
 * local_reduce_ndrange (:99-126) / local_reduce_hierarchical (:128-167).
 * Integers wrap mod 2^w (bit-exact with the reference for any order).  Floating point sums in an order that differs
 * from the reference's tree (tolerance) and, by default, from run to run: chunks are handed to CTAs in
 * order of arrival.  sb_ctx_set_deterministic(ctx, 1) fixes the order (bit-reproducible, about 3 % slower).
 * n == 0 writes 0.  `out` must not alias `in`. */
int sb_reduce_i32(sb_ctx*, const int32_t* in, size_t n, int32_t* out);
int sb_reduce_i64(sb_ctx*, const int64_t* in, size_t n, int64_t* out);
int sb_reduce_f32(sb_ctx*, const float* in, size_t n, float* out);
int sb_reduce_f64(sb_ctx*, const double* in, size_t n, double* out);

/* ---- fused two-input reductions ------------------------------------------ */
/* out[0] = sum a[i] * b[i], one launch, a and b read once, nothing else written.  Replaces the whole
 * run() of ScalarProdBench (single-kernel/scalar_prod.cpp:54-175): the product kernel (:55-83), and per
 * pass the work-group tree kernel (:96-170) and the gather kernel (:172-178) over the product buffer.
 * Every product is rounded to T before it is added (the reference stores it first); integers wrap. */
int sb_dot_i32(sb_ctx*, const int32_t* a, const int32_t* b, size_t n, int32_t* out);
int sb_dot_i64(sb_ctx*, const int64_t* a, const int64_t* b, size_t n, int64_t* out);
int sb_dot_f32(sb_ctx*, const float* a, const float* b, size_t n, float* out);
int sb_dot_f64(sb_ctx*, const double* a, const double* b, size_t n, double* out);
/* Simple linear regression y ~ b0 + b1 x in one launch.  out6 = { sum xy, sum xx, sum x, sum y, b1, b0 }.
 * Replaces LinearRegressionCoeffBench::run (single-kernel/lin_reg_coeff.cpp:126-146): two vec_product
 * launches (:52-69), four in-place multi-pass reductions (:71-124) and the host arithmetic (:134-142),
 * which is reproduced operation for operation in T: mean = sum / n, ss_xy = sum xy - mean_x mean_y,
 * ss_xx = sum xx - mean_x mean_x, b1 = ss_xy / ss_xx, b0 = mean_y - b1 mean_x.  x and y are NOT modified
 * (the reference reduces them in place).  n == 0: the sums are 0 and the coefficients NaN. */
int sb_linreg_f32(sb_ctx*, const float* x, const float* y, size_t n, float* out6);
int sb_linreg_f64(sb_ctx*, const double* x, const double* y, size_t n, double* out6);
/* Squared error of N candidate regression lines over N points:
 *   out[g] = sum_i ((alpha[g] * x[i] + beta[g]) - y[i])^2,   g < n.
 * Replaces LinearRegressionBench::run (single-kernel/lin_reg_error.cpp:54-79; in1 = x, in2 = y): one work-item per
 * candidate, each looping over all N points (N^2 multiply-adds).  Here: one fused streaming reduction of the five sums the
 * quadratic form needs (x and y read once, double accumulators) and one map over alpha / beta — 5*N*s bytes, two launches.
 * The result is the correctly rounded value to within a few ulp of double; the reference's sequential sum in T differs
 * from it by its own rounding error, inside the tolerance the tests state (1e-5*log2 N fp32, 1e-12*log2 N fp64).
 * Conditioning: a candidate that fits the points almost exactly has an error far below the O(N) sums it is computed
 * from; such values are accurate to a few N * 2^-52 in absolute terms (fp64, N = 2^22: 1e-10), not relatively. */
int sb_linreg_error_f32(sb_ctx*, const float* x, const float* y, const float* alpha, const float* beta, size_t n, float* out);
int sb_linreg_error_f64(sb_ctx*, const double* x, const double* y, const double* alpha, const double* beta, size_t n, double* out);
/* Nearest-cluster pass of k-means as the reference computes it (single-kernel/kmeans.cpp:49-82): features are stored
 * feature-major (features[l*n + g]), clusters[k*nfeatures + l]; distances are summed in T.  The reference records the
 * POINT's index, not the cluster's (`index = gid`, :74): membership[g] = (int)g if some cluster's squared distance is below
 * the running minimum — which starts at FLT_MAX (:66) —, else 0.  Bit-exact. */
int sb_kmeans_f32(sb_ctx*, const float* features, const float* clusters, size_t n, int nfeatures, int nclusters, int32_t* membership);
int sb_kmeans_f64(sb_ctx*, const double* features, const double* clusters, size_t n, int nfeatures, int nclusters, int32_t* membership);

/* ---- segmented reduction -------------------------------------------------------------------- */
/* In place, reference layout (pattern/segmentedreduction.cpp:35-61,63-96; contract :112-119):
 *   inout[g*seg_len] = tree_sum(inout[g*seg_len .. g*seg_len+seg_len-1]),   other elements untouched.
 * n % seg_len must be 0 (the reference asserts it, :23).  For a seg_len that is not a power of
 * two the reference's halving tree silently skips elements; the same subset is summed here
 * (oracle/sb_oracle.c: ref_tree_member).  */
int sb_segreduce_i16(sb_ctx*, int16_t* inout, size_t n, size_t seg_len);
int sb_segreduce_i32(sb_ctx*, int32_t* inout, size_t n, size_t seg_len);
int sb_segreduce_i64(sb_ctx*, int64_t* inout, size_t n, size_t seg_len);
int sb_segreduce_f32(sb_ctx*, float* inout, size_t n, size_t seg_len);
int sb_segreduce_f64(sb_ctx*, double* inout, size_t n, size_t seg_len);

/* ---- scan / prefix sum ---------------------------------------------------------------------- */
/* pattern/prefixsum.cpp:1 and pattern/scan.cpp:1 are TODO stubs upstream; semantics are the ones
 * written down in DESIGN.md §Scan semantics (parity unpinned upstream).
 *   group_len == 0 : global scan, single pass, decoupled look-back.
 *   group_len  > 0 : the scan restarts every group_len elements ("per-work-group granularity").
 *   exclusive != 0 : out[i] = sum of the elements strictly before i (0 at a restart).
 * `seed` (device pointer, may be NULL) is added to every output of a global scan: the carry-in of
 * a shard in the multi-GPU path.  in == out (in place) is allowed. */
int sb_scan_i32(sb_ctx*, const int32_t* in, int32_t* out, size_t n, int exclusive, size_t group_len, const int32_t* seed);
int sb_scan_i64(sb_ctx*, const int64_t* in, int64_t* out, size_t n, int exclusive, size_t group_len, const int64_t* seed);
int sb_scan_f32(sb_ctx*, const float* in, float* out, size_t n, int exclusive, size_t group_len, const float* seed);
int sb_scan_f64(sb_ctx*, const double* in, double* out, size_t n, int exclusive, size_t group_len, const double* seed);

/* ---- segmented scan ------------------------------------------------------------------------- */
/* pattern/segmentedscan.cpp:1 is a TODO stub upstream (DESIGN.md §Scan semantics).
 * Sum-scan that restarts at segment heads.  head_flags == NULL: heads are implicit at
 * i % seg_len == 0.  head_flags != NULL: one uint8 per element, non-zero marks a head
 * (seg_len is ignored; element 0 is always a head).  Operator on (value, flag) pairs:
 * (v1,f1) (+) (v2,f2) = (f2 ? v2 : v1+v2, f1|f2). */
int sb_segscan_i32(sb_ctx*, const int32_t* in, int32_t* out, size_t n, size_t seg_len, const uint8_t* head_flags, int exclusive);
int sb_segscan_i64(sb_ctx*, const int64_t* in, int64_t* out, size_t n, size_t seg_len, const uint8_t* head_flags, int exclusive);
int sb_segscan_f32(sb_ctx*, const float* in, float* out, size_t n, size_t seg_len, const uint8_t* head_flags, int exclusive);
int sb_segscan_f64(sb_ctx*, const double* in, double* out, size_t n, size_t seg_len, const uint8_t* head_flags, int exclusive);

/* ---- multi-GPU carry ------------------------------------------------------------------------ */
/* inout[i] += carry[0]  (the "fused carry-add pass" of BASELINE.json north_star). */
int sb_carry_add_i32(sb_ctx*, int32_t* inout, size_t n, const int32_t* carry);
int sb_carry_add_i64(sb_ctx*, int64_t* inout, size_t n, const int64_t* carry);
int sb_carry_add_f32(sb_ctx*, float* inout, size_t n, const float* carry);
int sb_carry_add_f64(sb_ctx*, double* inout, size_t n, const double* carry);
/* carry_out[0] = sum(totals[0..rank))  — turns the all-gathered shard totals into this rank's
 * exclusive carry on the device (no host round trip). */
int sb_exclusive_carry_i32(sb_ctx*, const int32_t* totals, int rank, int32_t* carry_out);
int sb_exclusive_carry_i64(sb_ctx*, const int64_t* totals, int rank, int64_t* carry_out);
int sb_exclusive_carry_f32(sb_ctx*, const float* totals, int rank, float* carry_out);
int sb_exclusive_carry_f64(sb_ctx*, const double* totals, int rank, double* carry_out);

/* ---- on-device verification against the closed forms of the iota input ------ */
/* Counts elements of `data` (a result computed from in[i] = (T)(offset+i)) that differ from the
 * closed form; integer types only (mod 2^w).  mismatches is a DEVICE pointer to one uint64.
 *   kind 0: inclusive scan   kind 1: exclusive scan   (seg_len == 0: global; else restart every seg_len)
 *   kind 2: segmented-reduction layout (head = segment sum, others = input). */
int sb_check_iota_i32(sb_ctx*, const int32_t* data, size_t n, uint64_t offset, int kind, size_t seg_len, uint64_t* mismatches);
int sb_check_iota_i64(sb_ctx*, const int64_t* data, size_t n, uint64_t offset, int kind, size_t seg_len, uint64_t* mismatches);
/* Floating point: counts elements with |data - closed form| > rel_tol * max(|closed form|, 1) (NaN counts); the closed
 * form is evaluated in double.  The untouched elements of kind 2 must equal the input (T)(offset+i) exactly.  rel_tol is
 * the caller's statement of BASELINE.json's tolerance (1e-5 log2 N for fp32, 1e-12 log2 N for fp64). */
int sb_check_iota_f32(sb_ctx*, const float* data, size_t n, uint64_t offset, int kind, size_t seg_len, double rel_tol, uint64_t* mismatches);
int sb_check_iota_f64(sb_ctx*, const double* data, size_t n, uint64_t offset, int kind, size_t seg_len, double rel_tol, uint64_t* mismatches);

/* ---- host-buffer (end-to-end) entry points -------------------------------------------------- */
/* Same operations on HOST buffers: H2D of the input, the kernel, D2H of the result, all on the ctx
 * stream, synchronous on return.  This is the call shape of the reference plugin where setup()
 * uploads a host std::vector (pattern/reduction.cpp:36-42) and verify() reads the result back
 * (:60).  Device staging buffers are cached in the ctx and grow on demand. */
int sb_host_reduce(sb_ctx*, int dtype, const void* h_in, size_t n, void* h_out);
int sb_host_scan(sb_ctx*, int dtype, const void* h_in, void* h_out, size_t n, int exclusive, size_t group_len);
int sb_host_segreduce(sb_ctx*, int dtype, void* h_inout, size_t n, size_t seg_len);
int sb_host_segscan(sb_ctx*, int dtype, const void* h_in, void* h_out, size_t n, size_t seg_len, const uint8_t* h_head_flags, int exclusive);

/* ---- sharded operations over the GPUs of one box (single process, NCCL over NVLink) ---------- */
/* The reference is single-device; this is the multi-GPU path of BASELINE.json north_star: the input
 * is sharded contiguously over ngpus devices, every rank reduces / scans its shard with the kernels
 * above, and NCCL exchanges ONE element per rank: all-reduce(sum) for the reduction, all-gather of
 * the shard totals for the scan's exclusive carry.  libnccl.so.2 is loaded lazily (dlopen), so the
 * library has no NCCL dependency unless these calls are used.
 *   rank 0 is `rank0` (its device, its stream: the harness's timer ctx); ranks 1.. get new ctxs on
 *   the following devices.  in[r] / out[r] are device pointers on rank r's device, n[r] its shard size.
 *   All work is asynchronous; sb_comm_join makes rank 0's stream wait for every other rank, so that
 *   sb_timer_end / sb_ctx_sync on rank 0 cover the whole operation. */
typedef struct sb_comm sb_comm;
enum sb_scan_strategy {
  SB_REDUCE_THEN_SCAN = 0, /* sb_reduce -> all-gather -> seeded sb_scan: 3 passes over a shard (default) */
  SB_SCAN_THEN_CARRY = 1,  /* sb_scan -> all-gather -> sb_carry_add: the 4-pass recipe, inclusive only */
  SB_SCAN_CYCLIC = 2       /* block-cyclic shards, sb_scan_cyclic: 2 passes, totals exchanged by the kernel over NVLink
                              peer memory, no NCCL call; every n[r] must be the same */
};
int sb_comm_create(sb_ctx* rank0, int ngpus, sb_comm** out);
int sb_comm_destroy(sb_comm* comm);
int sb_comm_size(sb_comm* comm);
sb_ctx* sb_comm_ctx(sb_comm* comm, int rank);
int sb_comm_join(sb_comm* comm);
/* out[r][0] = sum over ALL shards, on every rank */
int sb_comm_reduce(sb_comm* comm, int dtype, const void* const* in, const size_t* n, void* const* out);
/* out[r] = rank r's shard of the GLOBAL inclusive / exclusive scan */
int sb_comm_scan(sb_comm* comm, int dtype, const void* const* in, void* const* out, const size_t* n, int exclusive, int strategy);

/* Independent per-shard operations ("replicas"): segmented reduction, segmented scan and the
 * group-scope scan shard on segment boundaries and need no exchange at all.  One call enqueues the op on every rank's
 * stream and joins them into rank 0's.  op = SB_OP_SEGREDUCE (inout = out, in ignored), SB_OP_SEGSCAN (implicit heads
 * every seg_len), SB_OP_SCAN (group scope: restart every seg_len).  Every n[r] must be a multiple of seg_len. */
int sb_comm_foreach(sb_comm* comm, int op, int dtype, const void* const* in, void* const* out, const size_t* n, size_t seg_len, int exclusive);

/* ---- block-cyclic prefix sum: one read of the input per GPU, exchange inside the kernel ------------------------ */
/* The global array is dealt to the `world` ranks in blocks of sb_cyclic_block_elems() elements (one round of the
 * persistent kernel: sm_count tiles of 70 KB): block k lives on rank k % world, a rank stores its blocks back to
 * back, every rank holds the same number n of elements (the blocks of the last round may be short: they all have
 * n - floor(n / block) * block elements).  Global index of local element i of rank g, with b = i / block:
 *     b < n / block :  (b * world + g) * block + i % block
 *     otherwise     :  (n / block) * world * block + g * (n % block) + i % block
 * sb_scan_cyclic_<T> writes out[i] = inclusive / exclusive prefix sum at that global index (+ *seed if given).  The
 * kernel publishes one total per round and rank into every rank's mailbox (sb_peer) with system-scope stores over
 * NVLink and keeps the tile it is waiting for in the 126 MB L2, so every element is read from HBM once and written
 * once.  peer == NULL: one GPU (world 1) — the same two-phase kernel, an alternative to sb_scan_<T> for global scans.
 * All ranks of a peer group must issue their sb_scan_cyclic calls in the same order with the same n (it is a
 * collective).  in/out must be 16-byte aligned; in == out is allowed. */
typedef struct sb_peer sb_peer;
#define SB_PEER_HANDLE_BYTES 64
size_t sb_cyclic_block_elems(sb_ctx* ctx, int dtype);
/* one process per GPU: create, exchange the 64-byte handles (any transport), connect every other rank */
int sb_peer_create(sb_ctx* ctx, int rank, int world, sb_peer** out);
int sb_peer_handle(sb_peer* peer, void* handle, size_t handle_bytes);                              /* cudaIpcGetMemHandle */
int sb_peer_connect(sb_peer* peer, int peer_rank, const void* handle, size_t handle_bytes);        /* cudaIpcOpenMemHandle */
/* one process, several GPUs: connect two peers of the same group directly (enables peer access) */
int sb_peer_connect_local(sb_peer* peer, sb_peer* other);
int sb_peer_destroy(sb_peer* peer);
int sb_peer_rank(sb_peer* peer);
int sb_peer_world(sb_peer* peer);
int sb_scan_cyclic_i32(sb_ctx*, sb_peer* peer, const int32_t* in, int32_t* out, size_t n, int exclusive, const int32_t* seed);
int sb_scan_cyclic_i64(sb_ctx*, sb_peer* peer, const int64_t* in, int64_t* out, size_t n, int exclusive, const int64_t* seed);
int sb_scan_cyclic_f32(sb_ctx*, sb_peer* peer, const float* in, float* out, size_t n, int exclusive, const float* seed);
int sb_scan_cyclic_f64(sb_ctx*, sb_peer* peer, const double* in, double* out, size_t n, int exclusive, const double* seed);
int sb_scan_cyclic(sb_ctx*, sb_peer* peer, int dtype, const void* in, void* out, size_t n, int exclusive);
/* Reduction with the all-reduce fused into the kernel: out[0] = sum over the inputs of ALL ranks of the peer group, on
 * every rank.  The last block of sb_reduce's kernel stores this GPU's sum into every peer's mailbox over NVLink,
 * waits for the other ranks' sums and adds them in rank order (identical floating-point result on every rank).
 * Replaces sb_reduce + ncclAllReduce of one element (BASELINE.json north_star) by one launch per GPU.  A collective:
 * every rank calls it, in the same order.  peer == NULL or a group of one: plain sb_reduce. */
int sb_reduce_all_i32(sb_ctx*, sb_peer* peer, const int32_t* in, size_t n, int32_t* out);
int sb_reduce_all_i64(sb_ctx*, sb_peer* peer, const int64_t* in, size_t n, int64_t* out);
int sb_reduce_all_f32(sb_ctx*, sb_peer* peer, const float* in, size_t n, float* out);
int sb_reduce_all_f64(sb_ctx*, sb_peer* peer, const double* in, size_t n, double* out);
int sb_reduce_all(sb_ctx*, sb_peer* peer, int dtype, const void* in, size_t n, void* out);
/* generate_input / verify() for that layout (n = elements of this rank, the same on every rank) */
int sb_iota_cyclic_i32(sb_ctx*, int32_t* out, size_t n, size_t block, int world, int rank);
int sb_iota_cyclic_i64(sb_ctx*, int64_t* out, size_t n, size_t block, int world, int rank);
int sb_iota_cyclic_f32(sb_ctx*, float* out, size_t n, size_t block, int world, int rank);
int sb_iota_cyclic_f64(sb_ctx*, double* out, size_t n, size_t block, int world, int rank);
int sb_check_iota_cyclic_i32(sb_ctx*, const int32_t* data, size_t n, size_t block, int world, int rank, int kind, double rel_tol, uint64_t* mismatches);
int sb_check_iota_cyclic_i64(sb_ctx*, const int64_t* data, size_t n, size_t block, int world, int rank, int kind, double rel_tol, uint64_t* mismatches);
int sb_check_iota_cyclic_f32(sb_ctx*, const float* data, size_t n, size_t block, int world, int rank, int kind, double rel_tol, uint64_t* mismatches);
int sb_check_iota_cyclic_f64(sb_ctx*, const double* data, size_t n, size_t block, int world, int rank, int kind, double rel_tol, uint64_t* mismatches);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* SB_CUDA_H */
