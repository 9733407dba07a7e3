// pattern/prefixsum — global INCLUSIVE prefix sum of in[i] = (T)i, out of place, single pass with
// decoupled look-back (sb_scan_<T>, group_len = 0).  Upstream pattern/prefixsum.cpp:1 is a TODO stub:
// DESIGN.md §Scan semantics is the contract.  Rows: Pattern_PrefixSum_NDRange_{int32,int64,fp32,fp64}.
#include "scan_common.h"

SB_SCAN_BENCH(PrefixSumNDRange, PrefixSum, "Pattern_PrefixSum_NDRange_")

int main(int argc, char** argv) {
  BenchmarkApp app(argc, argv);
  if(app.shouldRunNDRangeKernels()) {
    app.run<PrefixSumNDRange<int>>();
    app.run<PrefixSumNDRange<long long>>();
    app.run<PrefixSumNDRange<float>>();
    app.run<PrefixSumNDRange<double>>();
  } else {
    std::cerr << "prefixsum: --no-ndrange-kernels leaves nothing to run (see DESIGN.md)\n";
  }
  pattern::release_comms();
  return app.exitCode();
}
