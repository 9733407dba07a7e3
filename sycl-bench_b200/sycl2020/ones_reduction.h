// Shared body of the three sycl2020/ reduction benchmarks.  Upstream they are three
// ways to sum an array of ones with SYCL 2020 features:
//   sycl2020/atomics/atomic_reduction.cpp:28-41               one atomic_ref::fetch_add per work-item
//   sycl2020/group_algorithms/reduce_over_group.cpp:30-50     reduce_over_group + one atomic per work-group
//   sycl2020/kernel_reduction/kernel_reduction.cpp:28-41      sycl::reduction, cf elements per work-item
// Here all three are ONE launch of sb_reduce_<T> (the pattern/reduction kernel) over a device-filled
// array of ones; what differs per benchmark is which prefix of the input takes part (the `effective`
// element count) and the row name.
//
// verify(): upstream compares with `problem_size` through percentDiff <= 0.05 %
// (polybench/common/polybenchUtilFuncts.h:36-42, kernel_reduction.cpp:43-53).  The check here is stricter:
// integers bit-exact (count mod 2^w), floating point within the north-star tolerance — and, so that a
// PASS here implies a PASS upstream, the 0.05 % criterion as well.  Known upstream quirk, not reproduced:
// ReductionAtomic_fp32 saturates at 2^24 for N > 2^24 because 16777216.0f + 1.0f == 16777216.0f
// (tests/golden/sycl2020_golden.json, size 2^25); sb_reduce returns the count.
#pragma once

#include "../pattern/pattern_common.h"

namespace sycl2020 {

template <class T>
struct Fill;
template <> struct Fill<int> { static int ones(sb_ctx* c, int* p, size_t n) { return sb_fill_i32(c, reinterpret_cast<int32_t*>(p), n, 1); } };
template <> struct Fill<long long> { static int ones(sb_ctx* c, long long* p, size_t n) { return sb_fill_i64(c, reinterpret_cast<int64_t*>(p), n, 1); } };
template <> struct Fill<float> { static int ones(sb_ctx* c, float* p, size_t n) { return sb_fill_f32(c, p, n, 1.0f); } };
template <> struct Fill<double> { static int ones(sb_ctx* c, double* p, size_t n) { return sb_fill_f64(c, p, n, 1.0); } };

// 100 * |a - b| / |a|, 0 when both are (almost) 0 — the criterion of upstream's percentDiff, in double
inline double percent_diff(double a, double b) {
  if(std::fabs(a) < 0.01 && std::fabs(b) < 0.01) return 0.0;
  return 100.0 * std::fabs(a - b) / std::fabs(a + 1e-8);
}

template <class T>
class OnesReduction {
protected:
  BenchmarkArgs _args;
  PrefetchedBuffer<T> _in, _out;
  std::size_t _effective = 0;  // elements that take part

  // nd_range kernels upstream: an nd_range whose global size is not a multiple of the local size is invalid
  void require_nd_range() const {
    if(_args.local_size == 0 || _args.problem_size % _args.local_size != 0)
      throw std::invalid_argument("nd_range: --size must be a multiple of --local");
  }

public:
  explicit OnesReduction(const BenchmarkArgs& args) : _args{args}, _effective{args.problem_size} {}

  void setup() {
    _in.initialize(_args.ctx(), std::max<std::size_t>(_args.problem_size, 1));
    sb::check(Fill<T>::ones(_args.ctx(), _in.get(), _args.problem_size), "sb_fill");
    _out.initialize(_args.ctx(), 1);  // upstream uploads reduction_results = 0 and accumulates into it
  }

  void run(std::vector<sb::event>& events) {
    sb::check(pattern::Ops<T>::reduce(_args.ctx(), _in.get(), _effective, _out.get()), "sb_reduce");
    events.push_back({"sb_reduce"});
  }

  bool verify(VerificationSetting&) {
    const T result = _out.get_host_access()[0];
    if(!pattern::value_matches<T>(result, _effective, (long double)_effective, _effective)) return false;
    // upstream's criterion compares with --size even when a coarsening tail was dropped (size 7, cf 4: FAIL)
    return percent_diff((double)result, (double)_args.problem_size) <= 0.05;
  }

  static ThroughputMetric throughput(std::size_t effective) {
    return {((double)effective * sizeof(T) + sizeof(T)) / 1073741824.0, "GiB"};  // algorithmic bytes: n*s + s
  }
};

} // namespace sycl2020
