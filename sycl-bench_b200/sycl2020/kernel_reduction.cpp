// sycl2020/kernel_reduction — drop-in for /root/reference/sycl2020/kernel_reduction/kernel_reduction.cpp.
// Rows: KernelReduction_{int32,int64,fp32,fp64}_plus_cf{1,4,8} (:56-66,:71-90).  Upstream: sycl::reduction
// over --size / cf work-items, each combining cf consecutive elements (:36-38), so the last --size % cf
// elements are never read; here one sb_reduce_<T> launch over the first (--size / cf) * cf elements.
// The coarsening factor is a tuning knob of the SYCL kernel; sb_reduce has its own fixed per-thread batch,
// so the three rows per type differ only in that element count.
#include "ones_reduction.h"

template <class T, int coarse_factor>
class KernelReductionBench : public sycl2020::OnesReduction<T> {
public:
  KernelReductionBench(const BenchmarkArgs& args) : sycl2020::OnesReduction<T>{args} {
    this->_effective = (args.problem_size / coarse_factor) * coarse_factor;
  }
  static ThroughputMetric getThroughputMetric(const BenchmarkArgs& args) {
    return sycl2020::OnesReduction<T>::throughput((args.problem_size / coarse_factor) * coarse_factor);
  }
  static std::string getBenchmarkName(BenchmarkArgs&) {
    return std::string("KernelReduction_") + ReadableTypename<T>::name + "_plus_cf" + std::to_string(coarse_factor);
  }
};

template <class T>
void runCoarsening(BenchmarkApp& app) {
  app.run<KernelReductionBench<T, 1>>();
  app.run<KernelReductionBench<T, 4>>();
  app.run<KernelReductionBench<T, 8>>();
}

int main(int argc, char** argv) {
  BenchmarkApp app(argc, argv);
  runCoarsening<int>(app);
  runCoarsening<long long>(app);
  runCoarsening<float>(app);
  runCoarsening<double>(app);
  return app.exitCode();
}
