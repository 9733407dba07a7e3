/* sb_oracle_single_kernel.c — CPU restatement of single-kernel/scalar_prod.cpp and
 * single-kernel/lin_reg_coeff.cpp of the reference, and — second half of the file, with its own
 * parity note — of single-kernel/lin_reg_error.cpp and single-kernel/kmeans.cpp.
 *
 * TEST INFRASTRUCTURE ONLY (same rule as sb_oracle.c): never called, linked or loaded by the product.
 *
 * Parity status: PINNED.  tests/test_single_kernel.py checks every function below bit for bit against
 *   tests/golden/single_kernel_golden.json, produced by the UNMODIFIED translation units compiled against
 *   the host-only shim (oracle/_ref/{scalar_prod,lin_reg_coeff}, one OpenMP thread), and live against those
 *   binaries when present.
 *
 * Both reference benchmarks reduce IN PLACE over several launches and contain read/write overlaps between
 * work-items of one launch (scalar_prod.cpp:172-178 gather; lin_reg_coeff.cpp:92-114 group g writes
 * element g while another group still reads it).  The restatement fixes the schedule the single-threaded
 * shim executes: work-groups in ascending order, work-items of the gather kernel in ascending order — the
 * one under which every read sees the value of the previous pass, i.e. the intended result.
 *
 * `file:line` are relative to /root/reference.  Integers wrap (unsigned arithmetic).
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

#define DEFINE_SCALAR_PROD(SUF, T, U)                                                                                   \
  /* ScalarProdBench::run, scalar_prod.cpp:54-180.  `L` = --local.  Returns 0, or -1 when out of memory. */          \
  ORACLE_API int ref_scalar_prod_##SUF(const T* a, const T* b, size_t n, size_t L, T* out) {                            \
    if(n == 0 || L == 0) return -1;                                                                                     \
    U* g = (U*)malloc(n * sizeof(U));                                                                                   \
    U* local = (U*)malloc(L * sizeof(U));                                                                               \
    if(!g || !local) { free(g); free(local); return -1; }                                                               \
    for(size_t i = 0; i < n; ++i) g[i] = (U)((U)a[i] * (U)b[i]);              /* :66-69 product kernel */               \
    size_t array_size = n;                                                                                              \
    const size_t ept = 2;                                                     /* elements_per_thread, :89 */            \
    while(array_size != 1) {                                                  /* :91 */                                 \
      const size_t n_wg = (array_size + L * ept - 1) / (L * ept);             /* :92-93 */                              \
      for(size_t grp = 0; grp < n_wg; ++grp) {                                /* reduction kernel :104-134 */           \
        for(size_t lid = 0; lid < L; ++lid) {                                                                           \
          const size_t gid = grp * L + lid;                                                                             \
          U acc = 0;                                                          /* :109 */                                \
          for(size_t i = 0; i < ept; ++i) {                                                                             \
            const size_t e = gid + i * n_wg * L;                              /* :112 (an `int` upstream) */            \
            if(e < array_size) acc = (U)(acc + g[e]);                         /* :114-115 */                            \
          }                                                                                                             \
          local[lid] = acc;                                                                                             \
        }                                                                                                               \
        for(size_t stride = L / ept; stride >= 1; stride /= ept)              /* :120-127 */                            \
          for(size_t lid = 0; lid < stride; ++lid) local[lid] = (U)(local[lid] + local[lid + stride]);                  \
        g[grp * L] = local[0];                                                /* :130-132 */                            \
      }                                                                                                                 \
      for(size_t idx = 0; idx < n_wg; ++idx) g[idx] = g[idx * L];             /* gather kernel :172-178 */              \
      array_size = n_wg;                                                      /* :179 */                                \
    }                                                                                                                   \
    *(U*)out = g[0];                                                                                                    \
    free(g);                                                                                                            \
    free(local);                                                                                                        \
    return 0;                                                                                                           \
  }

DEFINE_SCALAR_PROD(i32, int32_t, uint32_t)
DEFINE_SCALAR_PROD(i64, int64_t, uint64_t)
DEFINE_SCALAR_PROD(f32, float, float)
DEFINE_SCALAR_PROD(f64, double, double)

#define DEFINE_LINREG(SUF, T)                                                                                           \
  /* LinearRegressionCoeffBench::reduce, lin_reg_coeff.cpp:71-124: in place on `g` (n elements, followed by           \
   * one readable pad element: the kernel reads g[2*gid + 1] even when that is == array_size, :98). */                 \
  static T linreg_reduce_##SUF(T* g, size_t n, size_t L, T* local) {                                                    \
    size_t array_size = n;                                                                                              \
    const size_t ept = 2;                                                                                               \
    while(array_size != 1) {                                                  /* :77 */                                 \
      const size_t n_wg = (array_size + L * ept - 1) / (L * ept);             /* :78-79 */                              \
      for(size_t grp = 0; grp < n_wg; ++grp) {                                                                          \
        for(size_t lid = 0; lid < L; ++lid) {                                                                           \
          const size_t gid = grp * L + lid;                                                                             \
          local[lid] = 0;                                                     /* :93 */                                 \
          if(ept * gid < array_size) local[lid] = g[ept * gid] + g[ept * gid + 1]; /* :95-97 */                         \
        }                                                                                                               \
        for(size_t stride = 1; stride < L; stride *= ept)                     /* :102-109 */                            \
          for(size_t lid = 0; lid < L; ++lid) {                                                                         \
            const size_t idx = ept * stride * lid;                                                                      \
            if(idx < L) local[idx] = local[idx] + local[idx + stride];                                                  \
          }                                                                                                             \
        g[grp] = local[0];                                                    /* :112-114 */                            \
      }                                                                                                                 \
      array_size = n_wg;                                                      /* :117 */                                \
    }                                                                                                                   \
    return g[0];                                                              /* :119-120 */                            \
  }                                                                                                                     \
  /* LinearRegressionCoeffBench::run, lin_reg_coeff.cpp:126-146.  out6 = {ss_xy sum, ss_xx sum, sum x, sum y,         \
   * coeff_b1, coeff_b0}; the first four are the values reduce() returns (before the mean correction). */             \
  ORACLE_API int ref_linreg_##SUF(const T* x, const T* y, size_t n, size_t L, T* out6) {                                \
    if(n == 0 || L < 2) return -1;                                                                                      \
    T* in1 = (T*)calloc(n + 2, sizeof(T));                                                                              \
    T* in2 = (T*)calloc(n + 2, sizeof(T));                                                                              \
    T* prod = (T*)calloc(n + 2, sizeof(T));                                                                             \
    T* local = (T*)calloc(2 * L + 2, sizeof(T));                                                                        \
    if(!in1 || !in2 || !prod || !local) { free(in1); free(in2); free(prod); free(local); return -1; }                   \
    memcpy(in1, x, n * sizeof(T));                                                                                      \
    memcpy(in2, y, n * sizeof(T));                                                                                      \
    for(size_t i = 0; i < n; ++i) prod[i] = in1[i] * in2[i];                  /* vec_product :52-69 */                  \
    const T sxy = linreg_reduce_##SUF(prod, n, L, local);                     /* :129 */                                \
    for(size_t i = 0; i < n; ++i) prod[i] = in1[i] * in1[i];                  /* :131 */                                \
    const T sxx = linreg_reduce_##SUF(prod, n, L, local);                     /* :133 */                                \
    const T sx = linreg_reduce_##SUF(in1, n, L, local);                       /* :135 */                                \
    const T sy = linreg_reduce_##SUF(in2, n, L, local);                       /* :136 */                                \
    volatile T mean_x = sx / (T)n, mean_y = sy / (T)n;                                                                  \
    volatile T mxy = mean_x * mean_y, mxx = mean_x * mean_x;                                                            \
    volatile T ss_xy = sxy - mxy, ss_xx = sxx - mxx;                          /* :138-139 */                            \
    volatile T b1 = ss_xy / ss_xx;                                            /* :141 */                                \
    volatile T b1mx = b1 * mean_x;                                                                                      \
    out6[0] = sxy; out6[1] = sxx; out6[2] = sx; out6[3] = sy;                                                           \
    out6[4] = b1;                                                                                                       \
    out6[5] = mean_y - b1mx;                                                  /* :142 */                                \
    free(in1); free(in2); free(prod); free(local);                                                                      \
    return 0;                                                                                                           \
  }

DEFINE_LINREG(f32, float)
DEFINE_LINREG(f64, double)

/* ---- single-kernel/lin_reg_error.cpp and single-kernel/kmeans.cpp (the rows the path widens into next) -------------
 *
 * Parity status of ref_lin_reg_error_*: PINNED WITHIN TOLERANCE, not bit for bit.  The reference is built with
 * -O3 -ffast-math (CMakeLists.txt:17), which lets the compiler reassociate and vectorise the inner sum
 * `error += e * e` (lin_reg_error.cpp:71-74): the order of its additions is whatever gcc chose, not something a
 * restatement can follow.  What is pinned (tests/test_single_kernel_widened.py against
 * tests/golden/single_kernel_widened_golden.json, made from oracle/_ref/lin_reg_error): the inputs, regenerated with
 * the reference's own rand() sequence (:38-43), bit for bit where the golden file records them, and every output
 * element within the north-star tolerance plus the reference's own criterion (relative L2 error < 1e-6, :80-103).
 * ref_kmeans_*: PINNED bit for bit (integer output) to oracle/_ref/kmeans.
 */
#include <float.h>

#define DEFINE_LINREG_ERROR(SUF, T)                                                                                     \
  /* LinearRegressionBench::setup, lin_reg_error.cpp:30-43: four rand() draws per element, in1, in2, alpha, beta.    \
   * `skip` draws are discarded first: the binary runs <float> and then <double> on one rand() stream (:129-133),    \
   * so the fp64 inputs of --size=N start 4*N draws in. */                                                            \
  ORACLE_API void ref_lin_reg_error_inputs_##SUF(size_t n, size_t skip, T* in1, T* in2, T* alpha, T* beta) {            \
    srand(1); /* the C default seed: the reference never calls srand */                                                \
    for(size_t i = 0; i < skip; ++i) (void)rand();                                                                      \
    for(size_t i = 0; i < n; ++i) {                                                                                     \
      in1[i] = (T)rand() / (T)RAND_MAX;                                                                                 \
      in2[i] = (T)rand() / (T)RAND_MAX;                                                                                 \
      alpha[i] = (T)rand() / (T)RAND_MAX;                                                                               \
      beta[i] = (T)rand() / (T)RAND_MAX;                                                                                \
    }                                                                                                                   \
  }                                                                                                                     \
  /* the kernel, lin_reg_error.cpp:64-77: out[g] = sum_i ((alpha[g] * in1[i] + beta[g]) - in2[i])^2, added in index  \
   * order in T (volatile: one rounding per operation, no contraction). */                                            \
  ORACLE_API void ref_lin_reg_error_##SUF(const T* in1, const T* in2, const T* alpha, const T* beta, size_t n, T* out) { \
    _Pragma("omp parallel for schedule(static)")                                                                        \
    for(size_t g = 0; g < n; ++g) {                                                                                     \
      const T a = alpha[g], b = beta[g];                                                                                \
      T error = 0;                                                                                                      \
      for(size_t i = 0; i < n; ++i) {                                                                                   \
        volatile T ax = a * in1[i];                                                                                     \
        volatile T axb = ax + b;                                                                                        \
        volatile T e = axb - in2[i];                                                                                    \
        volatile T ee = e * e;                                                                                          \
        error += ee;                                                                                                    \
      }                                                                                                                 \
      out[g] = error;                                                                                                   \
    }                                                                                                                   \
  }

DEFINE_LINREG_ERROR(f32, float)
DEFINE_LINREG_ERROR(f64, double)

#define DEFINE_KMEANS(SUF, T)                                                                                           \
  /* the kernel, kmeans.cpp:61-79.  features: nfeatures x n (feature-major, :70), clusters: read at [i*nfeatures+l]  \
   * (:70; the buffer upstream allocates is 3*n long, only nclusters*nfeatures elements are read).  The reference    \
   * stores the point's own index, not the cluster's (`index = gid`, :74): membership[g] = g as soon as any distance  \
   * is below the running minimum, which starts at FLT_MAX (<cfloat>'s, reached through the SYCL headers; the #ifndef \
   * fallback of :4-6 is dead).  A NaN distance compares false. */                                                    \
  ORACLE_API void ref_kmeans_##SUF(const T* features, const T* clusters, size_t n, int nfeatures, int nclusters,        \
      int* membership) {                                                                                                \
    for(size_t g = 0; g < n; ++g) {                                                                                     \
      int index = 0;                                                                                                    \
      T min_dist = (T)FLT_MAX;                                                                                          \
      for(int i = 0; i < nclusters; ++i) {                                                                              \
        T dist = 0;                                                                                                     \
        for(int l = 0; l < nfeatures; ++l) {                                                                            \
          volatile T d = features[(size_t)l * n + g] - clusters[i * nfeatures + l];                                     \
          volatile T dd = d * d;                                                                                        \
          dist += dd;                                                                                                   \
        }                                                                                                               \
        if(dist < min_dist) {                                                                                           \
          min_dist = dist;                                                                                              \
          index = (int)g;                                                                                               \
        }                                                                                                               \
      }                                                                                                                 \
      membership[g] = index;                                                                                            \
    }                                                                                                                   \
  }

DEFINE_KMEANS(f32, float)
DEFINE_KMEANS(f64, double)
