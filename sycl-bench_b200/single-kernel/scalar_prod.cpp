// single-kernel/scalar_prod — drop-in for /root/reference/single-kernel/scalar_prod.cpp.
// Rows: ScalarProduct_NDRange_{int32,int64,fp32,fp64} (:205-212,:217-225).  input1 = 1, input2 = 2 (:43-47).
//
// Upstream run() (:54-180) = one product kernel that WRITES N products, then per pass a work-group tree
// kernel and a gather kernel over that buffer, in place: 1 + 2*ceil(log_{2L} N) launches and ~4*N*s bytes.
// Here it is ONE sb_dot_<T> launch: both inputs are read once (2*N*s bytes), nothing but the result is
// written.  Deviations, as for pattern/reduction:
//  * only the NDRange rows (the Hierarchical variant computes the same value, :137-168);
//  * a non-power-of-two --local makes upstream's halving tree skip elements (size 1000, local 100 -> 1024,
//    and upstream's own verify FAILs; tests/golden/single_kernel_golden.json) — not emulated: the result is
//    the scalar product;
//  * verify(): integers exact (2N mod 2^w), floating point within the north-star tolerance of 2N.  Upstream
//    compares with a sequential host sum in T and an absolute tolerance of 1e-5 (:183-202).
#include "../pattern/pattern_common.h"
#include "../sycl2020/ones_reduction.h"  // sycl2020::Fill

template <class T> struct Dot;
template <> struct Dot<int> { static int run(sb_ctx* c, const int* a, const int* b, size_t n, int* o) { return sb_dot_i32(c, (const int32_t*)a, (const int32_t*)b, n, (int32_t*)o); } };
template <> struct Dot<long long> { static int run(sb_ctx* c, const long long* a, const long long* b, size_t n, long long* o) { return sb_dot_i64(c, (const int64_t*)a, (const int64_t*)b, n, (int64_t*)o); } };
template <> struct Dot<float> { static int run(sb_ctx* c, const float* a, const float* b, size_t n, float* o) { return sb_dot_f32(c, a, b, n, o); } };
template <> struct Dot<double> { static int run(sb_ctx* c, const double* a, const double* b, size_t n, double* o) { return sb_dot_f64(c, a, b, n, o); } };

template <class T> struct FillValue;
template <> struct FillValue<int> { static int run(sb_ctx* c, int* p, size_t n, int v) { return sb_fill_i32(c, (int32_t*)p, n, v); } };
template <> struct FillValue<long long> { static int run(sb_ctx* c, long long* p, size_t n, long long v) { return sb_fill_i64(c, (int64_t*)p, n, v); } };
template <> struct FillValue<float> { static int run(sb_ctx* c, float* p, size_t n, float v) { return sb_fill_f32(c, p, n, v); } };
template <> struct FillValue<double> { static int run(sb_ctx* c, double* p, size_t n, double v) { return sb_fill_f64(c, p, n, v); } };

template <class T>
class ScalarProdBench {
protected:
  BenchmarkArgs _args;
  PrefetchedBuffer<T> _in1, _in2, _out;

public:
  ScalarProdBench(const BenchmarkArgs& args) : _args{args} {}

  void setup() {
    const std::size_t n = _args.problem_size;
    if(_args.local_size == 0 || n % _args.local_size != 0) throw std::invalid_argument("nd_range: --size must be a multiple of --local");
    _in1.initialize(_args.ctx(), std::max<std::size_t>(n, 1));
    _in2.initialize(_args.ctx(), std::max<std::size_t>(n, 1));
    sb::check(FillValue<T>::run(_args.ctx(), _in1.get(), n, T(1)), "sb_fill");
    sb::check(FillValue<T>::run(_args.ctx(), _in2.get(), n, T(2)), "sb_fill");
    _out.initialize(_args.ctx(), 1);  // upstream: an N-element buffer of products, result in element 0
  }

  void run(std::vector<sb::event>& events) {
    sb::check(Dot<T>::run(_args.ctx(), _in1.get(), _in2.get(), _args.problem_size, _out.get()), "sb_dot");
    events.push_back({"sb_dot"});
  }

  bool verify(VerificationSetting&) {
    const T result = _out.get_host_access()[0];
    const std::size_t n = _args.problem_size;
    return pattern::value_matches<T>(result, 2 * (uint64_t)n, 2.0L * (long double)n, n);
  }

  static ThroughputMetric getThroughputMetric(const BenchmarkArgs& args) {
    return {(2.0 * args.problem_size * sizeof(T) + sizeof(T)) / 1073741824.0, "GiB"};  // algorithmic bytes: 2*N*s + s
  }
  static std::string getBenchmarkName(BenchmarkArgs&) { return std::string("ScalarProduct_NDRange_") + ReadableTypename<T>::name; }
};

int main(int argc, char** argv) {
  BenchmarkApp app(argc, argv);
  if(app.shouldRunNDRangeKernels()) {
    app.run<ScalarProdBench<int>>();
    app.run<ScalarProdBench<long long>>();
    app.run<ScalarProdBench<float>>();
    app.run<ScalarProdBench<double>>();
  } else {
    std::cerr << "scalar_prod: --no-ndrange-kernels leaves nothing to run (one CUDA kernel backs the NDRange rows; see DESIGN.md)\n";
  }
  return app.exitCode();
}
