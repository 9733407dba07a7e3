// micro/DRAM — drop-in for /root/reference/micro/DRAM.cpp: device copy bandwidth,
// the measured-peak denominator of the HBM roofline from inside the same harness.
// Rows: MicroBench_DRAM_{fp32,fp64}_{1,2,3} (:97-104,:107-120).
//
// Upstream copies a buffer of --size^3 BYTES (getBufferSize, :11-22), every element 33 (:42), with one
// work-item per element over a 1-, 2- or 3-dimensional index space (:54-62).  The buffer is contiguous
// row-major memory for every dimensionality, so the three rows of a type are the same memory operation;
// here each is ONE sb_copy launch over the whole buffer (128-bit accesses, 16 KB per CTA) and the rows are
// kept so that result tables line up with upstream's.  Throughput metric as upstream: 2 x buffer GiB (:45-50).
// verify(): every element of the output equals 33 (:64-93) — the whole buffer when it is at most 1 GiB,
// else the first and last 2^20 elements plus 4096 evenly spaced windows of 256 elements.
#include "../pattern/pattern_common.h"

template <class T> struct Fill33;
template <> struct Fill33<float> { static int run(sb_ctx* c, float* p, size_t n) { return sb_fill_f32(c, p, n, 33.f); } };
template <> struct Fill33<double> { static int run(sb_ctx* c, double* p, size_t n) { return sb_fill_f64(c, p, n, 33.0); } };

template <class T, int Dims>
class MicroBenchDRAM {
protected:
  BenchmarkArgs _args;
  PrefetchedBuffer<T> _in, _out;

  // elements of the buffer, computed like upstream's getBufferSize per dimensionality (:11-22): the integer division
  // by sizeof(T) happens at a different place for 1, 2 and 3 dimensions, which matters when --size is not a
  // multiple of sizeof(T) (--size=10, fp32: 250 / 250 / 200 elements)
  static std::size_t count(const BenchmarkArgs& a) {
    const std::size_t p = a.problem_size;
    if(Dims == 1) return p * p * p / sizeof(T);
    if(Dims == 2) return (p * p / sizeof(T)) * p;
    return (p / sizeof(T)) * p * p;
  }

public:
  MicroBenchDRAM(const BenchmarkArgs& args) : _args{args} {}

  void setup() {
    const std::size_t n = count(_args);
    _in.initialize(_args.ctx(), std::max<std::size_t>(n, 1));
    _out.initialize(_args.ctx(), std::max<std::size_t>(n, 1));
    sb::check(Fill33<T>::run(_args.ctx(), _in.get(), n), "sb_fill");
    sb::check(sb_memset(_args.ctx(), _out.get(), 0, n * sizeof(T)), "sb_memset");
  }

  void run(std::vector<sb::event>& events) {
    sb::check(sb_copy(_args.ctx(), _out.get(), _in.get(), count(_args) * sizeof(T)), "sb_copy");
    events.push_back({"sb_copy"});
  }

  bool verify(VerificationSetting&) {
    const std::size_t n = count(_args);
    auto window_ok = [&](std::size_t begin, std::size_t cnt) {
      const auto h = _out.get_host_access(begin, cnt);
      for(const T v : h)
        if(v != T(33)) return false;
      return true;
    };
    if(n == 0) return true;
    if(n * sizeof(T) <= (std::size_t(1) << 30)) return window_ok(0, n);
    const std::size_t edge = std::size_t(1) << 20;
    if(!window_ok(0, edge) || !window_ok(n - edge, edge)) return false;
    const std::size_t step = n / 4096;
    for(std::size_t w = 0; w < 4096; ++w)
      if(!window_ok(w * step + step / 2, 256)) return false;
    return true;
  }

  static ThroughputMetric getThroughputMetric(const BenchmarkArgs& args) {
    return {2.0 * count(args) * sizeof(T) / 1024.0 / 1024.0 / 1024.0, "GiB"};
  }
  static std::string getBenchmarkName(BenchmarkArgs&) {
    return std::string("MicroBench_DRAM_") + ReadableTypename<T>::name + "_" + std::to_string(Dims);
  }
};

int main(int argc, char** argv) {
  BenchmarkApp app(argc, argv);
  app.run<MicroBenchDRAM<float, 1>>();
  app.run<MicroBenchDRAM<float, 2>>();
  app.run<MicroBenchDRAM<float, 3>>();
  app.run<MicroBenchDRAM<double, 1>>();
  app.run<MicroBenchDRAM<double, 2>>();
  app.run<MicroBenchDRAM<double, 3>>();
  return app.exitCode();
}
