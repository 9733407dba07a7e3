// sycl2020/reduce_over_group — drop-in for /root/reference/sycl2020/group_algorithms/reduce_over_group.cpp.
// Rows: ReduceGroupAlgorithm_{int32,int64,fp32,fp64} (:66-72,:77-84).  Upstream: reduce_over_group per
// work-group, then one atomic fetch_add of the partial per group (:40-47); here one sb_reduce_<T> launch
// (see ones_reduction.h).  --size must be a multiple of --local (nd_range, :36).
#include "ones_reduction.h"

template <class T>
class ReduceGroupAlgorithm : public sycl2020::OnesReduction<T> {
public:
  ReduceGroupAlgorithm(const BenchmarkArgs& args) : sycl2020::OnesReduction<T>{args} {}
  void setup() {
    this->require_nd_range();
    sycl2020::OnesReduction<T>::setup();
  }
  static ThroughputMetric getThroughputMetric(const BenchmarkArgs& args) { return sycl2020::OnesReduction<T>::throughput(args.problem_size); }
  static std::string getBenchmarkName(BenchmarkArgs&) { return std::string("ReduceGroupAlgorithm_") + ReadableTypename<T>::name; }
};

int main(int argc, char** argv) {
  BenchmarkApp app(argc, argv);
  app.run<ReduceGroupAlgorithm<int>>();
  app.run<ReduceGroupAlgorithm<long long>>();
  app.run<ReduceGroupAlgorithm<float>>();
  app.run<ReduceGroupAlgorithm<double>>();
  return app.exitCode();
}
