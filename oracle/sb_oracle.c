/* sb_oracle.c — CPU restatement of the reference's pattern/ algorithms.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (libsb_cuda, the harness, the pattern/
 * binaries) may call, link or load this file.  Users: tests/, __graft_entry__.smoke(), and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * Parity status
 *   reduction, segmentedreduction : PINNED.  Restated from the reference sources cited below and
 *       cross-checked (tests/test_oracle.py) against (a) the closed forms / known-answer values
 *       of SURVEY.md A.4 and (b) the UNMODIFIED reference translation units compiled against the
 *       host-only SYCL shim (oracle/_ref, built by oracle/Makefile from /root/reference).
 *   prefixsum, scan, segmentedscan : PARITY UNPINNED.  Upstream pattern/prefixsum.cpp:1,
 *       pattern/scan.cpp:1 and pattern/segmentedscan.cpp:1 are one-line TODO stubs; the functions
 *       below restate the contract written in DESIGN.md (§Scan semantics) and are pinned only by
 *       closed forms.
 *
 * All `file:line` citations are relative to /root/reference (unisa-hpc/sycl-bench 2.0.4).
 * Integer arithmetic is done on the unsigned counterpart (wrap mod 2^w, what the device does).
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* Which offsets j of a work-group of size L are folded into scratch[0] by
 *   for(int i = group_size / 2; i > 0; i /= 2) { barrier; if(lid < i) scratch[lid] += scratch[lid + i]; }
 * (pattern/reduction.cpp:117-121, pattern/segmentedreduction.cpp:52-56).  For a power-of-two L:
 * all of them.  Otherwise only the subset sums of the halving sequence L/2, L/4, ..., 1. */
ORACLE_API int ref_tree_member(size_t j, size_t L) {
  for(size_t i = L / 2; i > 0; i /= 2)
    if(j >= i) j -= i;
  return j == 0;
}

ORACLE_API int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

static int oracle_thread_id(void);

#define DEFINE_ORACLE(SUF, T, U, WIDE)                                                                                  \
  /* generate_input: pattern/reduction.cpp:31-34, pattern/segmentedreduction.cpp:25-28 */                              \
  ORACLE_API void ref_iota_##SUF(T* out, size_t n, uint64_t offset) {                                                   \
    for(size_t i = 0; i < n; ++i) out[i] = (T)(uint64_t)(offset + i);                                                   \
  }                                                                                                                     \
                                                                                                                        \
  /* one work-group of the tree: scratch[lid] = value; halving loop; returns scratch[0].                                \
   * pattern/reduction.cpp:115-123 (identical in :146-154 and segmentedreduction.cpp:50-56, :75-86) */                  \
  static T tree_##SUF(U* scratch, size_t L) {                                                                           \
    for(size_t i = L / 2; i > 0; i /= 2)                                                                                \
      for(size_t lid = 0; lid < i; ++lid) scratch[lid] = (U)(scratch[lid] + scratch[lid + i]);                          \
    return (T)scratch[0];                                                                                               \
  }                                                                                                                     \
                                                                                                                        \
  /* Reduction<T>::submit, pattern/reduction.cpp:71-97: n <- N, G <- N / L (floor: the tail is                          \
   * dropped when N % L != 0 because the assert at :29 is compiled out, CMakeLists.txt:17);                             \
   * loop { kernel; n <- G; G <- G > 1 ? ceil(n / L) : 0 }.  Result = out[0] (:60,:96).                                  \
   * N < L gives G = 0: the kernel is still submitted once with an empty range and out[0] is                            \
   * never written (uninitialised upstream); reported here as 0 with return value 1. */                                 \
  ORACLE_API int ref_reduce_##SUF(const T* in, size_t N, size_t L, T* result) {                                         \
    size_t n = N, G = N / L;                                                                                            \
    if(L == 0) return -1;                                                                                               \
    if(G == 0) { *result = (T)0; return 1; }                                                                            \
    U* cur = (U*)malloc((N ? N : 1) * sizeof(U));                                                                       \
    U* nxt = (U*)malloc(G * sizeof(U));                                                                                 \
    if(!cur || !nxt) { free(cur); free(nxt); return -2; }                                                               \
    memcpy(cur, in, N * sizeof(U));                                                                                     \
    int threads = oracle_num_threads();                                                                                 \
    (void)threads;                                                                                                      \
    do {                                                                                                                \
      _Pragma("omp parallel") {                                                                                         \
        U* scratch = (U*)malloc(L * sizeof(U));                                                                         \
        _Pragma("omp for schedule(static)") for(size_t g = 0; g < G; ++g) {                                             \
          for(size_t lid = 0; lid < L; ++lid) {                                                                         \
            const size_t gid = g * L + lid;                                                                             \
            scratch[lid] = (gid < n) ? cur[gid] : (U)0; /* :115 */                                                      \
          }                                                                                                             \
          nxt[g] = (U)tree_##SUF(scratch, L); /* :117-123 */                                                            \
        }                                                                                                               \
        free(scratch);                                                                                                  \
      }                                                                                                                 \
      n = G;                                                                                                            \
      G = (G > 1) ? (n + L - 1) / L : 0; /* :83-88 */                                                                   \
      if(G > 0) { U* t = cur; cur = nxt; nxt = t; } /* :90-92; buffers are large enough: G_k+1 <= G_k */                \
    } while(G > 0);                                                                                                     \
    *result = (T)nxt[0];                                                                                                \
    free(cur); free(nxt);                                                                                               \
    return 0;                                                                                                           \
  }                                                                                                                     \
                                                                                                                        \
  /* SegmentedReduction<T>::submit_*, pattern/segmentedreduction.cpp:35-61,63-96: in place,                             \
   * buf[g*L] = tree(buf[g*L .. g*L+L)), other elements untouched (:112-119). */                                        \
  ORACLE_API int ref_segreduce_##SUF(T* buf, size_t N, size_t L) {                                                      \
    if(L == 0 || N % L != 0) return -1;                                                                                 \
    const size_t G = N / L;                                                                                             \
    _Pragma("omp parallel") {                                                                                           \
      U* scratch = (U*)malloc(L * sizeof(U));                                                                           \
      _Pragma("omp for schedule(static)") for(size_t g = 0; g < G; ++g) {                                               \
        memcpy(scratch, buf + g * L, L * sizeof(U)); /* :50 */                                                          \
        buf[g * L] = tree_##SUF(scratch, L);         /* :52-58 */                                                       \
      }                                                                                                                 \
      free(scratch);                                                                                                    \
    }                                                                                                                   \
    return 0;                                                                                                           \
  }                                                                                                                     \
                                                                                                                        \
  /* The reference's own host check for segmentedreduction (sequential `T sum`),                                        \
   * pattern/segmentedreduction.cpp:107-111 — used to reproduce finding F4. */                                          \
  ORACLE_API T ref_segment_sequential_sum_##SUF(const T* seg, size_t L) {                                               \
    U sum = 0;                                                                                                          \
    for(size_t i = 0; i < L; ++i) sum = (U)(sum + ((const U*)seg)[i]);                                                  \
    return (T)sum;                                                                                                      \
  }                                                                                                                     \
                                                                                                                        \
  /* prefixsum / scan / segmentedscan — DESIGN.md §Scan semantics (upstream: TODO stubs).                               \
   * Sequential sum in the wide type WIDE (the integer itself; double for fp32; long double for                         \
   * fp64), rounded to T per element.  A head restarts the running sum.                                                 \
   *   group_len == 0 and flags == NULL : one segment (global scan).                                                    \
   *   flags != NULL : non-zero byte marks a head (element 0 always is one); group_len ignored. */                      \
  ORACLE_API int ref_scan_##SUF(const T* in, T* out, size_t n, int exclusive, size_t group_len, const uint8_t* flags) { \
    WIDE run = 0;                                                                                                       \
    for(size_t i = 0; i < n; ++i) {                                                                                     \
      const int head = (i == 0) || (flags ? flags[i] != 0 : (group_len && (i % group_len == 0)));                       \
      if(head) run = 0;                                                                                                 \
      const WIDE x = (WIDE)((const U*)in)[i]; /* read before write: in == out is allowed */                             \
      if(exclusive) out[i] = (T)(U)run;                                                                                 \
      run = (WIDE)(run + x);                                                                                            \
      if(!exclusive) out[i] = (T)(U)run;                                                                                \
    }                                                                                                                   \
    return 0;                                                                                                           \
  }                                                                                                                     \
                                                                                                                        \
  /* Multi-threaded sum / scan in T arithmetic: the CPU baseline legs of bench.py ("port").                             \
   * Two-pass chunked scan (per-thread totals, then rescan with the carry). */                                          \
  ORACLE_API T cpu_sum_##SUF(const T* in, size_t n) {                                                                   \
    U total = 0;                                                                                                        \
    _Pragma("omp parallel") {                                                                                           \
      U local = 0;                                                                                                      \
      _Pragma("omp for schedule(static) nowait") for(size_t i = 0; i < n; ++i) local = (U)(local + ((const U*)in)[i]);  \
      _Pragma("omp critical") total = (U)(total + local);                                                               \
    }                                                                                                                   \
    return (T)total;                                                                                                    \
  }                                                                                                                     \
  ORACLE_API int cpu_scan_##SUF(const T* in, T* out, size_t n, int exclusive) {                                         \
    int nt = oracle_num_threads();                                                                                      \
    if(nt < 1) nt = 1;                                                                                                  \
    U* totals = (U*)calloc((size_t)nt + 1, sizeof(U));                                                                  \
    if(!totals) return -2;                                                                                              \
    const size_t chunk = (n + (size_t)nt - 1) / (size_t)nt;                                                             \
    _Pragma("omp parallel num_threads(nt)") {                                                                           \
      const int t = oracle_thread_id();                                                                                 \
      const size_t b = (size_t)t * chunk, e = (b + chunk < n) ? b + chunk : n;                                          \
      U s = 0;                                                                                                          \
      for(size_t i = b; i < e; ++i) s = (U)(s + ((const U*)in)[i]);                                                     \
      totals[t + 1] = s;                                                                                                \
      _Pragma("omp barrier")                                                                                            \
      _Pragma("omp single") { for(int k = 1; k <= nt; ++k) totals[k] = (U)(totals[k] + totals[k - 1]); }                \
      U run = totals[t];                                                                                                \
      for(size_t i = b; i < e; ++i) {                                                                                   \
        const U x = ((const U*)in)[i];                                                                                  \
        if(exclusive) ((U*)out)[i] = run;                                                                               \
        run = (U)(run + x);                                                                                             \
        if(!exclusive) ((U*)out)[i] = run;                                                                              \
      }                                                                                                                 \
    }                                                                                                                   \
    free(totals);                                                                                                       \
    return 0;                                                                                                           \
  }

static int oracle_thread_id(void) {
#ifdef _OPENMP
  return omp_get_thread_num();
#else
  return 0;
#endif
}

DEFINE_ORACLE(i16, int16_t, uint16_t, uint16_t)
DEFINE_ORACLE(i32, int32_t, uint32_t, uint32_t)
DEFINE_ORACLE(i64, int64_t, uint64_t, uint64_t)
DEFINE_ORACLE(f32, float, float, double)
DEFINE_ORACLE(f64, double, double, long double)

/* ---- closed forms for the iota input, mod 2^64 ---------------------------------- */
static uint64_t tri(uint64_t g) { return (g & 1) ? g * ((g + 1) >> 1) : (g >> 1) * (g + 1); }

/* sum_{i<N} i */
ORACLE_API uint64_t kat_reduce(uint64_t N) { return N == 0 ? 0 : tri(N - 1); }
/* inclusive (exclusive) scan value at global index g; L == 0: global, else restart every L */
ORACLE_API uint64_t kat_scan(uint64_t g, int exclusive, uint64_t L) {
  if(L == 0) return exclusive ? (g == 0 ? 0 : tri(g - 1)) : tri(g);
  const uint64_t q = g / L, j = g % L;
  return exclusive ? q * L * j + (j == 0 ? 0 : tri(j - 1)) : q * L * (j + 1) + tri(j);
}
/* head of segment q of length L (power of two) */
ORACLE_API uint64_t kat_segreduce_head(uint64_t q, uint64_t L) { return q * L * L + tri(L - 1); }
