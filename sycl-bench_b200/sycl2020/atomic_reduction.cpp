// sycl2020/atomic_reduction — drop-in for /root/reference/sycl2020/atomics/atomic_reduction.cpp.
// Rows: ReductionAtomic_{int32,int64,fp32,fp64} (:55-61,:66-73).  Upstream issues one
// atomic_ref::fetch_add per work-item on out[0] (:33-40); here the same sum is one sb_reduce_<T> launch
// (see ones_reduction.h).  Like upstream's nd_range launch, --size must be a multiple of --local.
#include "ones_reduction.h"

template <class T>
class ReductionAtomic : public sycl2020::OnesReduction<T> {
public:
  ReductionAtomic(const BenchmarkArgs& args) : sycl2020::OnesReduction<T>{args} {}
  void setup() {
    this->require_nd_range();
    sycl2020::OnesReduction<T>::setup();
  }
  static ThroughputMetric getThroughputMetric(const BenchmarkArgs& args) { return sycl2020::OnesReduction<T>::throughput(args.problem_size); }
  static std::string getBenchmarkName(BenchmarkArgs&) { return std::string("ReductionAtomic_") + ReadableTypename<T>::name; }
};

int main(int argc, char** argv) {
  BenchmarkApp app(argc, argv);
  app.run<ReductionAtomic<int>>();
  app.run<ReductionAtomic<long long>>();
  app.run<ReductionAtomic<float>>();
  app.run<ReductionAtomic<double>>();
  return app.exitCode();
}
