// pattern/segmentedreduction — per-segment sum in place, drop-in for
// /root/reference/pattern/segmentedreduction.cpp (segment length = --local, result at the segment
// head, every other element untouched; :35-61, contract :112-119).
// Rows: Pattern_SegmentedReduction_NDRange_{int16,int32,int64,fp32,fp64} (:134-139,:168-175).
// verify(): heads against the closed form g*L^2 + L(L-1)/2 (integers and fp64 exact — all partial
// sums are integers below 2^53 — fp32 within 1e-5*log2 N), non-heads bit-identical to the input,
// like upstream :107-119 but with a tolerance where upstream's exact `!=` cannot hold.
#include "pattern_common.h"

template <typename T>
class SegmentedReduction {
protected:
  BenchmarkArgs _args;
  PrefetchedBuffer<T> _buff;

public:
  SegmentedReduction(const BenchmarkArgs& args) : _args{args} {}

  void setup() {
    if(_args.local_size == 0 || _args.problem_size % _args.local_size != 0)
      throw std::invalid_argument("segmentedreduction: --size must be a multiple of --local (asserted upstream, segmentedreduction.cpp:23)");
    _buff.initialize(_args.ctx(), _args.problem_size);
    sb::check(pattern::Ops<T>::iota(_args.ctx(), _buff.get(), _args.problem_size, 0), "sb_iota");
  }

  void run(std::vector<sb::event>& events) {
    sb::check(pattern::Ops<T>::segreduce(_args.ctx(), _buff.get(), _args.problem_size, _args.local_size), "sb_segreduce");
    events.push_back({"sb_segreduce"});
  }

  bool verify(VerificationSetting&) {
    const std::size_t n = _args.problem_size, L = _args.local_size;
    if((L & (L - 1)) != 0) return verify_on_host(true);  // the reference tree skips elements for such L
    if constexpr(std::is_same_v<T, int> || std::is_same_v<T, long long>) {
      PrefetchedBuffer<unsigned long long> bad;
      bad.initialize(_args.ctx(), 1);
      int rc;
      if constexpr(std::is_same_v<T, int>) rc = sb_check_iota_i32(_args.ctx(), reinterpret_cast<const int32_t*>(_buff.get()), n, 0, 2, L, reinterpret_cast<uint64_t*>(bad.get()));
      else rc = sb_check_iota_i64(_args.ctx(), reinterpret_cast<const int64_t*>(_buff.get()), n, 0, 2, L, reinterpret_cast<uint64_t*>(bad.get()));
      sb::check(rc, "sb_check_iota");
      return bad.get_host_access()[0] == 0;
    } else {
      return verify_on_host(false);
    }
  }

  static ThroughputMetric getThroughputMetric(const BenchmarkArgs& args) {
    const double n = static_cast<double>(args.problem_size), L = static_cast<double>(std::max<std::size_t>(args.local_size, 1));
    return {(n * sizeof(T) + n / L * sizeof(T)) / 1073741824.0, "GiB"};  // N*s read + (N/L)*s written
  }

private:
  static bool tree_member(std::size_t j, std::size_t L) {
    for(std::size_t i = L / 2; i > 0; i /= 2)
      if(j >= i) j -= i;
    return j == 0;
  }
  bool verify_on_host(bool skipping_tree) {
    const std::size_t n = _args.problem_size, L = _args.local_size;
    const std::size_t chunk = std::max<std::size_t>(L, (std::size_t(1) << 24) / L * L);
    for(std::size_t begin = 0; begin < n; begin += chunk) {
      const std::size_t count = std::min(chunk, n - begin);
      const auto h = _buff.get_host_access(begin, count);
      for(std::size_t g = 0; g < count / L; ++g) {
        const std::size_t base = begin + g * L;
        long double real = 0.0L;
        uint64_t exact = 0;
        for(std::size_t j = 0; j < L; ++j)
          if(!skipping_tree || tree_member(j, L)) {
            exact += static_cast<uint64_t>(static_cast<T>(base + j));
            real += static_cast<long double>(static_cast<T>(base + j));
          }
        if(!pattern::value_matches<T>(h[g * L], exact, real, n)) return false;
        for(std::size_t j = 1; j < L; ++j)
          if(h[g * L + j] != static_cast<T>(base + j)) return false;
      }
    }
    return true;
  }
};

template <class T>
class SegmentedReductionNDRange : public SegmentedReduction<T> {
public:
  SegmentedReductionNDRange(const BenchmarkArgs& args) : SegmentedReduction<T>{args} {}
  static std::string getBenchmarkName(BenchmarkArgs&) { return std::string("Pattern_SegmentedReduction_NDRange_") + ReadableTypename<T>::name; }
};

int main(int argc, char** argv) {
  BenchmarkApp app(argc, argv);
  if(app.shouldRunNDRangeKernels()) {
    app.run<SegmentedReductionNDRange<short>>();
    app.run<SegmentedReductionNDRange<int>>();
    app.run<SegmentedReductionNDRange<long long>>();
    app.run<SegmentedReductionNDRange<float>>();
    app.run<SegmentedReductionNDRange<double>>();
  } else {
    std::cerr << "segmentedreduction: --no-ndrange-kernels leaves nothing to run (see DESIGN.md)\n";
  }
  return app.exitCode();
}
