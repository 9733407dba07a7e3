/* sb_oracle_sycl2020.c — CPU restatement of the reference's sycl2020/ reduction trio.
 *
 * TEST INFRASTRUCTURE ONLY (same rule as sb_oracle.c): users are tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py — never the product path.
 *
 * Parity status: PINNED.  Cross-checked (tests/test_oracle.py) against the UNMODIFIED translation units
 *   sycl2020/atomics/atomic_reduction.cpp, sycl2020/group_algorithms/reduce_over_group.cpp and
 *   sycl2020/kernel_reduction/kernel_reduction.cpp compiled against the host-only shim (oracle/_ref) and
 *   against the committed golden vectors tests/golden/sycl2020_golden.json generated from them.
 *
 * What the three benchmarks compute (file:line relative to /root/reference):
 *   input  in[i] = 1 for every i                      atomic_reduction.cpp:23, reduce_over_group.cpp:25,
 *                                                     kernel_reduction.cpp:23
 *   ReductionAtomic<T>        out[0] += in[gid] atomically, one fetch_add per work-item (:33-40)
 *   ReduceGroupAlgorithm<T>   partial = reduce_over_group(in[gid]); work-item 0 of each group does
 *                             out[0] += partial atomically (:40-47); needs N % L == 0 (nd_range)
 *   KernelReduction<T,+,cf>   sycl::reduction over N / cf work-items, each combining cf consecutive
 *                             elements (:36-38): the tail N % cf is never read
 * Every addend is the same, so the result does not depend on the (unspecified) order of the atomics:
 * it is the value a T accumulator reaches after the same number of identical additions.  For fp32 that
 * is NOT N once N > 2^24: 16777216.0f + 1.0f == 16777216.0f, so ReductionAtomic_fp32 saturates at 2^24
 * (and upstream's verify(), which accumulates the expected value the same way (:46-47), agrees with it).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#define ORACLE_API __attribute__((visibility("default")))

/* polybench/common/polybenchUtilFuncts.h:14,27-42: float absVal(float), float percentDiff(double, double) */
static float abs_val(float a) { return a < 0 ? -a : a; }
ORACLE_API float ref2020_percent_diff(double val1, double val2) {
  if(abs_val((float)val1) < 0.01 && abs_val((float)val2) < 0.01) return 0.0f;
  return 100.0f * abs_val(abs_val((float)(val1 - val2)) / abs_val((float)(val1 + 0.00000001f)));
}

/* `count` additions of `step` to a T accumulator that starts at 0.  Integers wrap; floating point rounds
 * at every step, so a closed form is used only while every intermediate value is exact. */
#define DEFINE_REPEAT_INT(SUF, T, U)                                                                                    \
  static T repeat_add_##SUF(uint64_t count, uint64_t step) { return (T)(U)((U)count * (U)step); }
#define DEFINE_REPEAT_FP(SUF, T, MANT)                                                                                  \
  static T repeat_add_##SUF(uint64_t count, uint64_t step) {                                                            \
    const T s = (T)step;                                                                                                \
    /* exact while the running value stays below 2^MANT */                                                             \
    const uint64_t limit = (uint64_t)1 << MANT;                                                                         \
    uint64_t exact_steps = step ? limit / step : count;                                                                 \
    if(exact_steps > count) exact_steps = count;                                                                        \
    volatile T acc = (T)(exact_steps * step);                                                                           \
    for(uint64_t k = exact_steps; k < count; ++k) {                                                                     \
      const T next = acc + s;                                                                                           \
      if(next == acc) break; /* saturated: every further addition is absorbed */                                       \
      acc = next;                                                                                                       \
    }                                                                                                                   \
    return acc;                                                                                                         \
  }

DEFINE_REPEAT_INT(i32, int32_t, uint32_t)
DEFINE_REPEAT_INT(i64, int64_t, uint64_t)
DEFINE_REPEAT_FP(f32, float, 24)
DEFINE_REPEAT_FP(f64, double, 53)

#define DEFINE_TRIO(SUF, T)                                                                                             \
  /* atomic_reduction.cpp:33-40: N fetch_add(1) */                                                                      \
  ORACLE_API T ref2020_atomic_##SUF(uint64_t n) { return repeat_add_##SUF(n, 1); }                                      \
  /* reduce_over_group.cpp:40-47: every group folds L ones (exact for L < 2^24), N / L fetch_add(L) */                 \
  ORACLE_API T ref2020_group_##SUF(uint64_t n, uint64_t local) {                                                        \
    if(local == 0 || n % local != 0) return (T)-1; /* the nd_range is invalid: upstream throws */                      \
    return repeat_add_##SUF(n / local, local);                                                                          \
  }                                                                                                                     \
  /* kernel_reduction.cpp:36-38: N / cf work-items x cf elements; partial sums are combined pairwise by the           \
   * implementation, every partial is an exact integer, so the result is the exact count rounded once to T */        \
  ORACLE_API T ref2020_kernel_##SUF(uint64_t n, uint64_t cf) { return (T)((n / cf) * cf); }                             \
  /* verify(): atomic_reduction.cpp:42-52 accumulates the expected value in T; the other two compare with              \
   * `T verified_results = problem_size` where problem_size is an `int` copy of --size                                 \
   * (reduce_over_group.cpp:8,22,55; kernel_reduction.cpp:9,20,46).  Returns 1 for PASS. */                            \
  ORACLE_API int ref2020_verify_atomic_##SUF(T result, uint64_t n) {                                                    \
    return !(ref2020_percent_diff((double)result, (double)repeat_add_##SUF(n, 1)) > 0.05f);                             \
  }                                                                                                                     \
  ORACLE_API int ref2020_verify_size_##SUF(T result, uint64_t n) {                                                      \
    const T expected = (T)(int)(uint32_t)n;                                                                             \
    return !(ref2020_percent_diff((double)result, (double)expected) > 0.05);                                            \
  }

DEFINE_TRIO(i32, int32_t)
DEFINE_TRIO(i64, int64_t)
DEFINE_TRIO(f32, float)
DEFINE_TRIO(f64, double)
