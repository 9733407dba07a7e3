// pattern/segmentedreduction — per-segment sum in place, drop-in for
// /root/reference/pattern/segmentedreduction.cpp (segment length = --local, result at the segment
// head, every other element untouched; :35-61, contract :112-119).
// Rows: Pattern_SegmentedReduction_NDRange_{int16,int32,int64,fp32,fp64} (:134-139,:168-175).
// verify(): heads against the closed form g*L^2 + L(L-1)/2 (integers and fp64 exact — all partial
// sums are integers below 2^53 — fp32 within 1e-5*log2 N), non-heads bit-identical to the input,
// like upstream :107-119 but with a tolerance where upstream's exact `!=` cannot hold.
#include "sharded.h"

template <typename T>
class SegmentedReduction {
protected:
  BenchmarkArgs _args;
  PrefetchedBuffer<T> _buff;
  // --num-gpus=N: shards cut on segment boundaries, independent replicas (sb_comm_foreach); not for int16
  int _gpus = 1;
  sb_comm* _comm = nullptr;
  pattern::ShardedArray<T> _shards;

public:
  SegmentedReduction(const BenchmarkArgs& args) : _args{args} {}

  void setup() {
    if(_args.local_size == 0 || _args.problem_size % _args.local_size != 0)
      throw std::invalid_argument("segmentedreduction: --size must be a multiple of --local (asserted upstream, segmentedreduction.cpp:23)");
    _gpus = pattern::requested_gpus(_args);
    if(_gpus > 1 && (std::is_same_v<T, short> || (_args.local_size & (_args.local_size - 1)) != 0)) {
      // the int16 row and non-power-of-two --local (reference-tree subset, host-side check) stay on one GPU
      _gpus = 1;
    }
    if(_gpus > 1) {
      _comm = pattern::comm_for(_args, _gpus);
      _shards.initialize(_comm, _args.problem_size, pattern::lcm16(_args.local_size));
      _shards.fill_iota();
      return;
    }
    _buff.initialize(_args.ctx(), _args.problem_size);
    sb::check(pattern::Ops<T>::iota(_args.ctx(), _buff.get(), _args.problem_size, 0), "sb_iota");
  }

  void run(std::vector<sb::event>& events) {
    if(_gpus > 1) {
      sb::check(sb_comm_foreach(_comm, SB_OP_SEGREDUCE, dtype(), nullptr, _shards.out_ptrs(), _shards.counts(), _args.local_size, 0), "sb_comm_foreach");
      events.push_back({"sb_comm_foreach"});
      return;
    }
    sb::check(pattern::Ops<T>::segreduce(_args.ctx(), _buff.get(), _args.problem_size, _args.local_size), "sb_segreduce");
    events.push_back({"sb_segreduce"});
  }

  bool verify(VerificationSetting&) {
    const std::size_t n = _args.problem_size, L = _args.local_size;
    if constexpr(!std::is_same_v<T, short>) {
      if(_gpus > 1) {  // every shard, every element, on its own device (power-of-two --local only, see setup())
        for(int r = 0; r < _gpus; ++r)
          if(!pattern::verify_closed_form<T>(sb_comm_ctx(_comm, r), _shards.shard(r).get(), _shards.count(r), 2, L, _shards.begin(r), n)) return false;
        return true;
      }
      // every element on the device: heads against the closed form (integers exact, floating point within the
      // tolerance), the rest bit-identical to the input
      if((L & (L - 1)) == 0) return pattern::verify_closed_form<T>(_args.ctx(), _buff.get(), n, 2, L);
    }
    return verify_on_host((L & (L - 1)) != 0);  // int16, and --local values for which the reference tree skips elements
  }

  static ThroughputMetric getThroughputMetric(const BenchmarkArgs& args) {
    const double n = static_cast<double>(args.problem_size), L = static_cast<double>(std::max<std::size_t>(args.local_size, 1));
    return {(n * sizeof(T) + n / L * sizeof(T)) / 1073741824.0, "GiB"};  // N*s read + (N/L)*s written
  }

private:
  static constexpr int dtype() {
    if constexpr(std::is_same_v<T, short>) return SB_I16;
    else return pattern::DtypeOf<T>::value;
  }
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
  pattern::release_comms();
  return app.exitCode();
}
