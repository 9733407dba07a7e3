// pattern/scan — EXCLUSIVE sum scan; scope selected by --scan-scope=global|group|both (default both):
//   global: one scan over the whole array (decoupled look-back);
//   group : the scan restarts every --local elements (the per-work-group granularity of the
//           reference's implemented siblings, e.g. pattern/segmentedreduction.cpp:35-61).
// Upstream pattern/scan.cpp:1 is a TODO stub: DESIGN.md §Scan semantics is the contract.
// Rows: Pattern_Scan_Global_NDRange_<T>, Pattern_Scan_Group_NDRange_<T>.
#include "scan_common.h"

SB_SCAN_BENCH(ScanGlobalNDRange, ScanGlobal, "Pattern_Scan_Global_NDRange_")
SB_SCAN_BENCH(ScanGroupNDRange, ScanGroup, "Pattern_Scan_Group_NDRange_")

int main(int argc, char** argv) {
  BenchmarkApp app(argc, argv);
  const std::string scope = app.getArgs().cli.getOrDefault<std::string>("--scan-scope", "both");
  if(scope != "global" && scope != "group" && scope != "both") {
    std::cerr << "scan: --scan-scope must be global, group or both\n";
    return 2;
  }
  if(!app.shouldRunNDRangeKernels()) {
    std::cerr << "scan: --no-ndrange-kernels leaves nothing to run (see DESIGN.md)\n";
    return app.exitCode();
  }
  if(scope != "group") {
    app.run<ScanGlobalNDRange<int>>();
    app.run<ScanGlobalNDRange<long long>>();
    app.run<ScanGlobalNDRange<float>>();
    app.run<ScanGlobalNDRange<double>>();
  }
  if(scope != "global") {
    app.run<ScanGroupNDRange<int>>();
    app.run<ScanGroupNDRange<long long>>();
    app.run<ScanGroupNDRange<float>>();
    app.run<ScanGroupNDRange<double>>();
  }
  pattern::release_comms();
  return app.exitCode();
}
