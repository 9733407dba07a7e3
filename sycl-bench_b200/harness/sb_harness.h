// sb_harness.h — the sycl-bench plugin surface without SYCL.
//
// Upper drop-in boundary.  A benchmark is still a duck-typed class
//     B(const BenchmarkArgs&), setup(), run() or run(std::vector<sb::event>&), [verify(VerificationSetting&)],
//     [static getThroughputMetric(const BenchmarkArgs&)], getBenchmarkName(BenchmarkArgs&)
// driven by BenchmarkApp::run<B>() / BenchmarkManager<B>::run(), parsed from the same flags and
// reported through the same ResultConsumer rows as the reference:
//     BenchmarkManager / BenchmarkApp   <- /root/reference/include/common.h:30-147, :167-214
//     CommandLine / BenchmarkArgs        <- include/command_line.h:73-120, :129-168
//     TimeMetricsProcessor               <- include/time_metrics.h:31-153
//     ResultConsumer + stdout / CSV      <- include/result_consumer.h:12-106
//     BenchmarkHook                      <- include/benchmark_hook.h:6-16
//     BenchmarkTraits                    <- include/benchmark_traits.h:9-33
//     ReadableTypename                   <- include/type_traits.h:14-23
//     PrefetchedBuffer                   <- include/memory_wrappers.h:39-92
// What changes: the two sycl::queue members of BenchmarkArgs become one `sb_ctx*` (device + stream
// of libsb_cuda), SYCL event profiling becomes CUDA-event timing (kernel-time is always available),
// and the deviations listed in DESIGN.md §Harness (fail-fast flag parsing, working --warmup-run,
// --strict exit code).  Written from scratch; no reference code is included.
#pragma once

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <numeric>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

#include "sb_cuda.h"

// ------------------------------------------------------------------------------------------------------------
// small value types standing in for sycl::id<3> / sycl::range<3> in VerificationSetting
namespace sb {

struct id3 {
  std::size_t v[3] = {0, 0, 0};
  std::size_t& operator[](int i) { return v[i]; }
  const std::size_t& operator[](int i) const { return v[i]; }
};
struct range3 {
  std::size_t v[3] = {1, 1, 1};
  std::size_t& operator[](int i) { return v[i]; }
  const std::size_t& operator[](int i) const { return v[i]; }
  std::size_t size() const { return v[0] * v[1] * v[2]; }
};

// one entry per libsb_cuda call a benchmark submits (the reference collects sycl::event objects;
// here the manager brackets the whole run() with CUDA events, the entries only name the launches)
struct event {
  const char* what = "";
};

inline void check(int status, const char* where) {
  if(status != 0) throw std::runtime_error(std::string(where) + ": " + sb_strerror(status));
}

} // namespace sb

// ------------------------------------------------------------------------------------------------------------
// result sinks
class ResultConsumer {
public:
  virtual ~ResultConsumer() = default;
  virtual void proceedToBenchmark(const std::string& name) = 0;
  virtual void consumeResult(const std::string& key, const std::string& value, const std::string& unit = "") = 0;
  virtual void flush() = 0;
  virtual void discard() {}
};

// human-readable printer: banner + "key: value [unit]" lines
class OstreamResultConsumer : public ResultConsumer {
public:
  explicit OstreamResultConsumer(std::ostream& os) : os_(os) {}
  void proceedToBenchmark(const std::string& name) override { os_ << "********** Results for " << name << "**********" << std::endl; }
  void consumeResult(const std::string& key, const std::string& value, const std::string& unit = "") override {
    os_ << key << ": " << value;
    if(!unit.empty()) os_ << " [" << unit << "]";
    os_ << std::endl;
  }
  void flush() override {}

private:
  std::ostream& os_;
};

// appending CSV writer: every flush emits "# Benchmark name,<sorted columns>" and one row per benchmark;
// units are dropped; commas inside values are replaced so that rows stay parseable
class AppendingCsvResultConsumer : public ResultConsumer {
public:
  explicit AppendingCsvResultConsumer(const std::string& path) : out_(path, std::ios::app) {
    if(!out_) throw std::runtime_error("cannot open output file " + path);
  }
  void proceedToBenchmark(const std::string& name) override { current_ = name; }
  void consumeResult(const std::string& key, const std::string& value, const std::string& = "") override {
    std::string v = value;
    if(v.empty() || v.front() != '"') std::replace(v.begin(), v.end(), ',', ';');
    rows_[current_][key] = v;
  }
  void flush() override {
    std::set<std::string> cols;
    for(const auto& r : rows_)
      for(const auto& kv : r.second) cols.insert(kv.first);
    out_ << "# Benchmark name";
    for(const auto& c : cols) out_ << "," << c;
    out_ << std::endl;
    for(const auto& r : rows_) {
      out_ << r.first;
      for(const auto& c : cols) {
        auto it = r.second.find(c);
        out_ << "," << (it == r.second.end() ? std::string("N/A") : it->second);
      }
      out_ << std::endl;
    }
    rows_.clear();
  }
  void discard() override {
    rows_.erase(current_);
    current_.clear();
  }

private:
  std::ofstream out_;
  std::string current_;
  std::map<std::string, std::map<std::string, std::string>> rows_;
};

// ------------------------------------------------------------------------------------------------------------
// command line: "--key=value" pairs and bare flags; argv[0] is parsed too (harmless, as upstream);
// a key given twice is an error; unknown keys are kept for benchmark-specific use
class CommandLine {
public:
  CommandLine() = default;
  CommandLine(int argc, char** argv) {
    for(int i = 0; i < argc; ++i) {
      const std::string a = argv[i];
      const auto eq = a.find('=');
      if(eq == std::string::npos) {
        flags_.insert(a);
        continue;
      }
      const std::string k = a.substr(0, eq);
      if(!kv_.emplace(k, a.substr(eq + 1)).second) throw std::invalid_argument("Encountered command line argument several times: " + k);
    }
  }
  bool isArgSet(const std::string& k) const { return kv_.count(k) != 0; }
  bool isFlagSet(const std::string& f) const { return flags_.count(f) != 0; }
  template <class T>
  T get(const std::string& k) const {
    auto it = kv_.find(k);
    if(it == kv_.end()) throw std::invalid_argument("Command line argument was requested but missing: " + k);
    return convert<T>(it->second, k);
  }
  template <class T>
  T getOrDefault(const std::string& k, const T& dflt) const {
    return isArgSet(k) ? get<T>(k) : dflt;
  }

private:
  template <class T>
  static T convert(const std::string& s, const std::string& key) {
    if constexpr(std::is_same_v<T, std::string>) {
      return s;
    } else if constexpr(std::is_same_v<T, sb::id3> || std::is_same_v<T, sb::range3>) {
      // "x", "x,y" or "x,y,z" (upstream tests the string length instead of the element count and
      // rejects anything longer than three characters; fixed here)
      T r;
      std::stringstream ss(s);
      std::string cell;
      int n = 0;
      while(std::getline(ss, cell, ',')) {
        if(n >= 3) throw std::invalid_argument("Invalid range/id: " + s);
        r[n++] = static_cast<std::size_t>(std::stoull(cell));
      }
      if(n == 0) throw std::invalid_argument("Invalid range/id: " + s);
      return r;
    } else {
      std::stringstream ss(s);
      T v{};
      ss >> v;
      if(ss.fail()) throw std::invalid_argument("cannot parse value of " + key + ": " + s);
      return v;
    }
  }
  std::unordered_map<std::string, std::string> kv_;
  std::unordered_set<std::string> flags_;
};

struct VerificationSetting {
  bool enabled = true;
  sb::id3 begin;
  sb::range3 range;
};

struct BenchmarkArgs {
  std::size_t problem_size = 3072;
  std::size_t local_size = 256;
  std::size_t num_runs = 5;
  // replaces sycl::queue device_queue / device_queue_in_order: one libsb_cuda context (device + in-order stream)
  std::shared_ptr<sb_ctx> device_ctx;
  int device_index = 0;
  VerificationSetting verification;
  CommandLine cli;  // benchmark-specific flags (--scan-scope=, --reference-verify, --host-input, ...)
  std::shared_ptr<ResultConsumer> result_consumer;
  bool warmup_run = false;
  bool strict = false;  // non-zero exit code on FAIL (upstream always exits 0)

  sb_ctx* ctx() const { return device_ctx.get(); }
};

class BenchmarkCommandLine {
public:
  BenchmarkCommandLine(int argc, char** argv) : cli_(argc, argv) {}

  BenchmarkArgs getBenchmarkArgs() const {
    BenchmarkArgs a;
    a.problem_size = cli_.getOrDefault<std::size_t>("--size", 3072);
    a.local_size = cli_.getOrDefault<std::size_t>("--local", 256);
    a.num_runs = cli_.getOrDefault<std::size_t>("--num-runs", 5);
    a.verification.enabled = !cli_.isFlagSet("--no-verification");
    a.verification.begin = cli_.getOrDefault<sb::id3>("--verification-begin", sb::id3{});
    a.verification.range = cli_.getOrDefault<sb::range3>("--verification-range", sb::range3{});
    a.warmup_run = cli_.isFlagSet("--warmup-run");
    a.strict = cli_.isFlagSet("--strict");
    a.cli = cli_;

    const std::string dev = cli_.getOrDefault<std::string>("--device", "default");
    if(dev == "cpu")
      throw std::invalid_argument("--device=cpu: libsb_cuda has no CPU path (the reference's CPU device is oracle/_ref, see BASELINE.md)");
    if(dev != "gpu" && dev != "default") throw std::invalid_argument("unknown device type: " + dev);
    a.device_index = cli_.getOrDefault<int>("--device-index", 0);
    sb_ctx* raw = nullptr;
    sb::check(sb_ctx_create(a.device_index, SB_STREAM_OWNED, &raw), "sb_ctx_create");
    a.device_ctx = std::shared_ptr<sb_ctx>(raw, [](sb_ctx* c) { sb_ctx_destroy(c); });

    const std::string out = cli_.getOrDefault<std::string>("--output", "stdio");
    if(out == "stdio") a.result_consumer = std::make_shared<OstreamResultConsumer>(std::cout);
    else a.result_consumer = std::make_shared<AppendingCsvResultConsumer>(out);
    return a;
  }

private:
  CommandLine cli_;
};

// ------------------------------------------------------------------------------------------------------------
// hooks
class BenchmarkHook {
public:
  virtual ~BenchmarkHook() = default;
  virtual void atInit() = 0;
  virtual void preSetup() = 0;
  virtual void postSetup() = 0;
  virtual void preKernel() = 0;
  virtual void postKernel() = 0;
  virtual void emitResults(ResultConsumer&) {}
};

// The reference ships no hook implementation (include/nv_energy_meas.h is an empty TODO).  This one
// counts the libsb_cuda kernel launches inside the timed region of every run.
class LaunchCountHook : public BenchmarkHook {
public:
  explicit LaunchCountHook(sb_ctx* ctx) : ctx_(ctx) {}
  void atInit() override {}
  void preSetup() override {}
  void postSetup() override {}
  void preKernel() override { before_ = sb_ctx_launch_count(ctx_); }
  void postKernel() override { last_ = sb_ctx_launch_count(ctx_) - before_; }
  void emitResults(ResultConsumer& rc) override { rc.consumeResult("kernel-launches-per-run", std::to_string(last_)); }

private:
  sb_ctx* ctx_;
  std::uint64_t before_ = 0, last_ = 0;
};

#include "nv_energy_meas.h"  // NvmlEnergyHook (needs BenchmarkHook and ResultConsumer from above)

// ------------------------------------------------------------------------------------------------------------
// traits
namespace detail {

template <typename T, typename = void>
struct SupportsQueueProfiling : std::false_type {};
template <typename T>
struct SupportsQueueProfiling<T, std::void_t<decltype(std::declval<T>().run(std::declval<std::vector<sb::event>&>()))>> : std::true_type {};

template <typename T, typename = void>
struct HasVerify : std::false_type {};
template <typename T>
struct HasVerify<T, std::void_t<decltype(std::declval<T>().verify(std::declval<VerificationSetting&>()))>> : std::true_type {};

template <typename T, typename = void>
struct HasThroughputMetric : std::false_type {};
template <typename T>
struct HasThroughputMetric<T, std::void_t<decltype(T::getThroughputMetric(std::declval<const BenchmarkArgs&>()))>> : std::true_type {};

template <typename T>
struct BenchmarkTraits {
  static constexpr bool supportsQueueProfiling = SupportsQueueProfiling<T>::value;
  static constexpr bool hasVerify = HasVerify<T>::value;
  static constexpr bool hasGetThroughputMetric = HasThroughputMetric<T>::value;
};

} // namespace detail

template <class T>
struct ReadableTypename;
#define SB_READABLE_TYPENAME(T, str)                                                                                   \
  template <>                                                                                                          \
  struct ReadableTypename<T> {                                                                                         \
    static constexpr const char* name = str;                                                                           \
  };
SB_READABLE_TYPENAME(char, "int8")
SB_READABLE_TYPENAME(unsigned char, "uint8")
SB_READABLE_TYPENAME(short, "int16")
SB_READABLE_TYPENAME(unsigned short, "uint16")
SB_READABLE_TYPENAME(int, "int32")
SB_READABLE_TYPENAME(unsigned int, "uint32")
SB_READABLE_TYPENAME(long long, "int64")
SB_READABLE_TYPENAME(unsigned long long, "uint64")
SB_READABLE_TYPENAME(float, "fp32")
SB_READABLE_TYPENAME(double, "fp64")
#undef SB_READABLE_TYPENAME

// ------------------------------------------------------------------------------------------------------------
// timing statistics
struct ThroughputMetric {
  double metric = 0.0;  // the quantity (bytes, elements, ...), not the rate
  std::string unit;
};

template <typename Benchmark>
class TimeMetricsProcessor {
public:
  explicit TimeMetricsProcessor(const BenchmarkArgs& args) : args_(args) {}

  void addTimingResult(const std::string& name, std::chrono::nanoseconds t) {
    if(unavailable_.count(name)) throw std::invalid_argument("Cannot add result for unavailable timing " + name);
    samples_[name].push_back(t);
  }
  void markAsUnavailable(const std::string& name) {
    if(samples_.count(name)) throw std::invalid_argument("Cannot mark timing " + name + " with existing results as unavailable");
    unavailable_.insert(name);
  }

  // columns per timing name: -mean -stddev -median -min -samples -throughput (seconds, std::to_string
  // formatting like upstream), plus -min-ns because six decimals resolve only 1 us
  void emitResults(ResultConsumer& rc) const {
    ThroughputMetric tpm;
    if constexpr(detail::BenchmarkTraits<Benchmark>::hasGetThroughputMetric) {
      tpm = Benchmark::getThroughputMetric(args_);
      rc.consumeResult("throughput-metric", std::to_string(tpm.metric), tpm.unit);
    } else {
      rc.consumeResult("throughput-metric", "N/A");
    }
    std::set<std::string> names(unavailable_.begin(), unavailable_.end());
    for(const auto& kv : samples_) names.insert(kv.first);
    for(const auto& name : names) {
      static const char* suffix[] = {"-mean", "-stddev", "-median", "-min", "-samples", "-throughput"};
      if(unavailable_.count(name)) {
        for(const char* s : suffix) rc.consumeResult(name + s, "N/A");
        continue;
      }
      const auto& raw = samples_.at(name);
      // --warmup-run: the first run is executed but not counted (documented behaviour, README.md:14;
      // upstream never assigns the flag)
      std::size_t first = (args_.warmup_run && raw.size() > 1) ? 1 : 0;
      std::vector<double> sec;
      for(std::size_t i = first; i < raw.size(); ++i) sec.push_back(raw[i].count() / 1.0e9);
      std::sort(sec.begin(), sec.end());
      const double n = static_cast<double>(sec.size());
      const double mean = std::accumulate(sec.begin(), sec.end(), 0.0) / n;
      double var = 0.0;
      for(double x : sec) var += (x - mean) * (x - mean);
      const double stddev = sec.size() > 1 ? std::sqrt(var / (n - 1.0)) : 0.0;
      rc.consumeResult(name + "-mean", std::to_string(mean), "s");
      rc.consumeResult(name + "-stddev", std::to_string(stddev), "s");
      rc.consumeResult(name + "-median", std::to_string(sec[sec.size() / 2]), "s");
      rc.consumeResult(name + "-min", std::to_string(sec.front()), "s");
      std::string all = "\"";
      for(std::size_t i = 0; i < sec.size(); ++i) all += (i ? " " : "") + std::to_string(sec[i]);
      rc.consumeResult(name + "-samples", all + "\"");
      if(tpm.metric > 0.0) rc.consumeResult(name + "-throughput", std::to_string(tpm.metric / sec.front()), tpm.unit + "/s");
      else rc.consumeResult(name + "-throughput", "N/A");
      rc.consumeResult(name + "-min-ns", std::to_string(static_cast<long long>(std::llround(sec.front() * 1e9))), "ns");
    }
  }

private:
  const BenchmarkArgs args_;
  std::unordered_map<std::string, std::vector<std::chrono::nanoseconds>> samples_;
  std::unordered_set<std::string> unavailable_;
};

// ------------------------------------------------------------------------------------------------------------
// device buffer: cudaMalloc + (optional) upload before the timed region
template <class T, int Dimensions = 1>
class PrefetchedBuffer {
public:
  PrefetchedBuffer() = default;
  PrefetchedBuffer(const PrefetchedBuffer&) = delete;
  PrefetchedBuffer& operator=(const PrefetchedBuffer&) = delete;
  ~PrefetchedBuffer() { reset(); }

  // allocate only (forceDataAllocation upstream)
  void initialize(sb_ctx* ctx, std::size_t count) {
    reset();
    ctx_ = ctx;
    count_ = count;
    void* p = nullptr;
    sb::check(sb_malloc(ctx, &p, count * sizeof(T)), "sb_malloc");
    ptr_ = static_cast<T*>(p);
  }
  // allocate + host -> device copy, complete on return (forceDataTransfer upstream)
  void initialize(sb_ctx* ctx, const T* host, std::size_t count) {
    initialize(ctx, count);
    sb::check(sb_memcpy_h2d(ctx, ptr_, host, count * sizeof(T)), "sb_memcpy_h2d");
    sb::check(sb_ctx_sync(ctx), "sb_ctx_sync");
  }
  T* get() const { return ptr_; }
  std::size_t size() const { return count_; }
  // get_host_access(): device -> host copy of [begin, begin + count)
  std::vector<T> get_host_access(std::size_t begin = 0, std::size_t count = static_cast<std::size_t>(-1)) const {
    if(count == static_cast<std::size_t>(-1)) count = count_ - begin;
    std::vector<T> h(count);
    if(count) {
      sb::check(sb_memcpy_d2h(ctx_, h.data(), ptr_ + begin, count * sizeof(T)), "sb_memcpy_d2h");
      sb::check(sb_ctx_sync(ctx_), "sb_ctx_sync");
    }
    return h;
  }
  void reset() {
    if(ptr_) sb_free(ctx_, ptr_);
    ptr_ = nullptr;
    count_ = 0;
  }

private:
  sb_ctx* ctx_ = nullptr;
  T* ptr_ = nullptr;
  std::size_t count_ = 0;
};

// ------------------------------------------------------------------------------------------------------------
// run loop
template <class Benchmark>
class BenchmarkManager {
public:
  explicit BenchmarkManager(const BenchmarkArgs& a) : args(a) {}
  void addHook(BenchmarkHook& h) { hooks.push_back(&h); }

  // returns true unless verification failed
  template <typename... Extra>
  bool run(Extra&&... extra) {
    auto& rc = *args.result_consumer;
    rc.proceedToBenchmark(Benchmark{args, extra...}.getBenchmarkName(args));
    rc.consumeResult("problem-size", std::to_string(args.problem_size));
    rc.consumeResult("local-size", std::to_string(args.local_size));
    char dev[256] = "unknown";
    sb_ctx_device_name(args.ctx(), dev, sizeof(dev));
    rc.consumeResult("device-name", dev);
    rc.consumeResult("sycl-implementation", "libsb_cuda " + std::to_string(sb_version()) + " (CUDA; sm_100a; no SYCL)");

    TimeMetricsProcessor<Benchmark> metrics(args);
    for(auto* h : hooks) h->atInit();

    bool all_pass = true;
    const std::size_t total_runs = args.num_runs + (args.warmup_run ? 1 : 0);
    try {
      for(std::size_t r = 0; r < total_runs && all_pass; ++r) {
        Benchmark b(args, extra...);
        for(auto* h : hooks) h->preSetup();
        b.setup();
        sb::check(sb_ctx_sync(args.ctx()), "setup");
        for(auto* h : hooks) h->postSetup();

        std::vector<sb::event> events;
        events.reserve(1024);
        for(auto* h : hooks) h->preKernel();
        // ---- timed region: host clock around run() + wait (run-time), CUDA events on the stream (kernel-time)
        const auto t0 = std::chrono::steady_clock::now();
        sb::check(sb_timer_begin(args.ctx()), "sb_timer_begin");
        if constexpr(detail::BenchmarkTraits<Benchmark>::supportsQueueProfiling) b.run(events);
        else b.run();
        const auto t_submitted = std::chrono::steady_clock::now();
        float kernel_ms = 0.f;
        sb::check(sb_timer_end(args.ctx(), &kernel_ms), "sb_timer_end");  // waits for the stream
        const auto t1 = std::chrono::steady_clock::now();
        // ---- end of timed region
        for(auto* h : hooks) h->postKernel();

        const auto run_time = std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0);
        const auto kernel_time = std::chrono::nanoseconds(static_cast<long long>(std::llround(static_cast<double>(kernel_ms) * 1e6)));
        metrics.addTimingResult("run-time", run_time);
        metrics.addTimingResult("kernel-time", kernel_time);
        metrics.addTimingResult("submit-time", std::chrono::duration_cast<std::chrono::nanoseconds>(t_submitted - t0));
        metrics.addTimingResult("system-time", run_time > kernel_time ? run_time - kernel_time : std::chrono::nanoseconds(0));

        if constexpr(detail::BenchmarkTraits<Benchmark>::hasVerify) {
          if(args.verification.enabled && args.verification.range.size() > 0) {
            if(!b.verify(args.verification)) all_pass = false;
          }
        }
      }
    } catch(...) {
      rc.discard();
      throw;
    }

    metrics.emitResults(rc);
    for(auto* h : hooks) h->emitResults(rc);
    const bool verified = detail::BenchmarkTraits<Benchmark>::hasVerify && args.verification.enabled && args.verification.range.size() > 0;
    rc.consumeResult("Verification", !verified ? "N/A" : (all_pass ? "PASS" : "FAIL"));
    rc.flush();
    return all_pass;
  }

private:
  BenchmarkArgs args;
  std::vector<BenchmarkHook*> hooks;
};

class BenchmarkApp {
public:
  BenchmarkApp(int argc, char** argv) {
    try {
      args = BenchmarkCommandLine{argc, argv}.getBenchmarkArgs();
    } catch(std::exception& e) {
      // upstream prints and carries on with uninitialised arguments (include/common.h:174-178); fail fast instead
      std::cerr << "Error while parsing command lines: " << e.what() << std::endl;
      std::exit(2);
    }
  }

  const BenchmarkArgs& getArgs() const { return args; }
  bool shouldRunNDRangeKernels() const { return !args.cli.isFlagSet("--no-ndrange-kernels"); }
  bool deviceSupportsFP64() const { return true; }
  // 0, or 1 when --strict was given and a benchmark failed verification or threw
  int exitCode() const { return (args.strict && failures > 0) ? 1 : 0; }

  template <class Benchmark, typename... Extra>
  void run(Extra&&... extra) {
    try {
      const auto name = Benchmark{args, extra...}.getBenchmarkName(args);
      if(!names.insert(name).second) {
        std::cerr << "Benchmark with name '" << name << "' has already been run\n";
        throw std::runtime_error("Duplicate benchmark name");
      }
      BenchmarkManager<Benchmark> mgr(args);
      LaunchCountHook launches(args.ctx());
      mgr.addHook(launches);
      std::unique_ptr<NvmlEnergyHook> nvml;  // energy / clocks / throttle reasons; --no-nvml switches it off
      if(!args.cli.isFlagSet("--no-nvml")) {
        nvml = std::make_unique<NvmlEnergyHook>(args.ctx());
        mgr.addHook(*nvml);
      }
      if(!mgr.run(extra...)) ++failures;
    } catch(std::exception& e) {
      ++failures;
      std::cerr << "Error: " << e.what() << std::endl;
    }
  }

private:
  BenchmarkArgs args;
  std::unordered_set<std::string> names;
  int failures = 0;
};
