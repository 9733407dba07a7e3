// pattern/segmentedscan — INCLUSIVE sum scan that restarts at segment heads.
//   Pattern_SegmentedScan_NDRange_<T>       heads implicit at i % --local == 0
//   Pattern_SegmentedScan_Flags_NDRange_<T> the same segmentation given as an explicit uint8 head-flag
//                                           array: exercises the (value, flag) operator and the
//                                           look-back over pairs (--no-flags skips these rows)
// Upstream pattern/segmentedscan.cpp:1 is a TODO stub: DESIGN.md §Scan semantics is the contract.
#include "scan_common.h"

SB_SCAN_BENCH(SegmentedScanNDRange, SegScanImplicit, "Pattern_SegmentedScan_NDRange_")
SB_SCAN_BENCH(SegmentedScanFlagsNDRange, SegScanFlags, "Pattern_SegmentedScan_Flags_NDRange_")

int main(int argc, char** argv) {
  BenchmarkApp app(argc, argv);
  if(!app.shouldRunNDRangeKernels()) {
    std::cerr << "segmentedscan: --no-ndrange-kernels leaves nothing to run (see DESIGN.md)\n";
    return app.exitCode();
  }
  app.run<SegmentedScanNDRange<int>>();
  app.run<SegmentedScanNDRange<long long>>();
  app.run<SegmentedScanNDRange<float>>();
  app.run<SegmentedScanNDRange<double>>();
  if(!app.getArgs().cli.isFlagSet("--no-flags")) {
    app.run<SegmentedScanFlagsNDRange<int>>();
    app.run<SegmentedScanFlagsNDRange<long long>>();
    app.run<SegmentedScanFlagsNDRange<float>>();
    app.run<SegmentedScanFlagsNDRange<double>>();
  }
  pattern::release_comms();
  return app.exitCode();
}
