// prefixsum / scan / segmentedscan share one benchmark class: upstream's three files are TODO
// stubs (/root/reference/pattern/prefixsum.cpp:1, scan.cpp:1, segmentedscan.cpp:1), so the
// semantics are the written contract of DESIGN.md §Scan semantics.
#pragma once
#include "sharded.h"

namespace pattern {

enum class ScanKind { PrefixSum, ScanGlobal, ScanGroup, SegScanImplicit, SegScanFlags };

template <typename T, ScanKind KIND>
class ScanBenchmark {
protected:
  BenchmarkArgs _args;
  PrefetchedBuffer<T> _in, _out;
  PrefetchedBuffer<uint8_t> _flags;
  // --num-gpus=N: global scans through sb_comm_scan (block-cyclic or contiguous shards), restarting scans as
  // independent replicas on shards cut at segment boundaries (sb_comm_foreach)
  int _gpus = 1;
  sb_comm* _comm = nullptr;
  ShardedArray<T> _in_shards, _out_shards;
  int _strategy = SB_SCAN_CYCLIC;
  std::size_t _block = 0;

  static constexpr bool global_scope() { return KIND == ScanKind::PrefixSum || KIND == ScanKind::ScanGlobal; }

  static constexpr bool exclusive() { return KIND == ScanKind::ScanGlobal || KIND == ScanKind::ScanGroup; }
  std::size_t restart() const { return (KIND == ScanKind::PrefixSum || KIND == ScanKind::ScanGlobal) ? 0 : _args.local_size; }

public:
  ScanBenchmark(const BenchmarkArgs& args) : _args{args} {}

  void setup() {
    const std::size_t n = _args.problem_size;
    if(KIND != ScanKind::PrefixSum && KIND != ScanKind::ScanGlobal && _args.local_size == 0)
      throw std::invalid_argument("--local must be positive");
    _gpus = requested_gpus(_args);
    if(_gpus > 1) {
      if(KIND == ScanKind::SegScanFlags) throw std::invalid_argument("--num-gpus: the explicit-flag rows run on one GPU (use --no-flags)");
      _comm = comm_for(_args, _gpus);
      if(global_scope()) {
        const std::string st = _args.cli.getOrDefault<std::string>("--scan-strategy", "cyclic");
        if(st == "cyclic") _strategy = SB_SCAN_CYCLIC;
        else if(st == "reduce_then_scan") _strategy = SB_REDUCE_THEN_SCAN;
        else if(st == "scan_then_carry") _strategy = SB_SCAN_THEN_CARRY;
        else throw std::invalid_argument("--scan-strategy must be cyclic, reduce_then_scan or scan_then_carry");
        if(_strategy == SB_SCAN_THEN_CARRY && exclusive()) throw std::invalid_argument("--scan-strategy=scan_then_carry is inclusive only");
      }
      if(global_scope() && _strategy == SB_SCAN_CYCLIC) {
        _block = sb_cyclic_block_elems(_args.ctx(), DtypeOf<T>::value);
        _in_shards.initialize_cyclic(_comm, n);
        _out_shards.initialize_cyclic(_comm, n);
        _in_shards.fill_iota_cyclic(_block);
      } else {
        // replicas: shards end on a segment boundary (and keep 16-byte alignment)
        const std::size_t align = global_scope() ? 16 : lcm16(_args.local_size);
        if(!global_scope() && n % _args.local_size != 0) throw std::invalid_argument("--num-gpus: --size must be a multiple of --local");
        _in_shards.initialize(_comm, n, align);
        _out_shards.initialize(_comm, n, align);
        _in_shards.fill_iota();
      }
      return;
    }
    _in.initialize(_args.ctx(), n);
    _out.initialize(_args.ctx(), n);
    sb::check(Ops<T>::iota(_args.ctx(), _in.get(), n, 0), "sb_iota");
    if(KIND == ScanKind::SegScanFlags) {
      std::vector<uint8_t> f(n, 0);
      for(std::size_t i = 0; i < n; i += _args.local_size) f[i] = 1;
      _flags.initialize(_args.ctx(), f.data(), n);
    }
  }

  void run(std::vector<sb::event>& events) {
    const std::size_t n = _args.problem_size;
    int rc;
    if(_gpus > 1) {
      if(global_scope()) {
        sb::check(sb_comm_scan(_comm, DtypeOf<T>::value, _in_shards.in_ptrs(), _out_shards.out_ptrs(), _in_shards.counts(), exclusive() ? 1 : 0, _strategy), "sb_comm_scan");
        events.push_back({"sb_comm_scan"});
      } else {
        const int op = KIND == ScanKind::ScanGroup ? SB_OP_SCAN : SB_OP_SEGSCAN;
        sb::check(sb_comm_foreach(_comm, op, DtypeOf<T>::value, _in_shards.in_ptrs(), _out_shards.out_ptrs(), _in_shards.counts(), _args.local_size, exclusive() ? 1 : 0),
                  "sb_comm_foreach");
        events.push_back({"sb_comm_foreach"});
      }
      return;
    }
    if(KIND == ScanKind::PrefixSum) rc = Ops<T>::scan(_args.ctx(), _in.get(), _out.get(), n, 0, 0);
    else if(KIND == ScanKind::ScanGlobal) rc = Ops<T>::scan(_args.ctx(), _in.get(), _out.get(), n, 1, 0);
    else if(KIND == ScanKind::ScanGroup) rc = Ops<T>::scan(_args.ctx(), _in.get(), _out.get(), n, 1, _args.local_size);
    else if(KIND == ScanKind::SegScanImplicit) rc = Ops<T>::segscan(_args.ctx(), _in.get(), _out.get(), n, _args.local_size, nullptr, 0);
    else rc = Ops<T>::segscan(_args.ctx(), _in.get(), _out.get(), n, 0, _flags.get(), 0);
    sb::check(rc, "sb_scan");
    events.push_back({"sb_scan"});
  }

  bool verify(VerificationSetting&) {
    if(_gpus > 1) {
      for(int r = 0; r < _gpus; ++r) {
        const bool ok = (global_scope() && _strategy == SB_SCAN_CYCLIC)
                            ? verify_scan_cyclic<T>(sb_comm_ctx(_comm, r), _out_shards.shard(r), _out_shards.count(r), exclusive(), _block, _gpus, r)
                            : verify_scan<T>(sb_comm_ctx(_comm, r), _out_shards.shard(r), _out_shards.count(r), exclusive(), restart(), _out_shards.begin(r), _args.problem_size);
        if(!ok) return false;
      }
      return true;
    }
    return verify_scan<T>(_args.ctx(), _out, _args.problem_size, exclusive(), restart());
  }

  static ThroughputMetric getThroughputMetric(const BenchmarkArgs& args) {
    const double n = static_cast<double>(args.problem_size);
    return {(2.0 * n * sizeof(T) + (KIND == ScanKind::SegScanFlags ? n : 0.0)) / 1073741824.0, "GiB"};  // 2*N*s (+ N flag bytes)
  }
};

#define SB_SCAN_BENCH(CLASS, KIND_, PREFIX)                                                                             \
  template <class T>                                                                                                    \
  class CLASS : public pattern::ScanBenchmark<T, pattern::ScanKind::KIND_> {                                            \
  public:                                                                                                               \
    CLASS(const BenchmarkArgs& args) : pattern::ScanBenchmark<T, pattern::ScanKind::KIND_>{args} {}                     \
    static std::string getBenchmarkName(BenchmarkArgs&) { return std::string(PREFIX) + ReadableTypename<T>::name; }     \
  };

} // namespace pattern
