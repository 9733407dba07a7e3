// single-kernel/lin_reg_coeff — drop-in for /root/reference/single-kernel/lin_reg_coeff.cpp.
// Rows: LinearRegressionCoeff_{fp32,fp64} (:187-192,:197-204).  x = 1, y = 2 (:41-44).
//
// Upstream run() (:126-146): vec_product(x, y) + reduce, vec_product(x, x) + reduce, reduce(x), reduce(y) —
// two product launches that write N elements each and four in-place multi-pass tree reductions (which
// destroy x and y), a host read-back after every reduction, then the coefficients on the host.
// Here: ONE sb_linreg_<T> launch reads x and y once (2*N*s bytes), keeps the four sums in registers and
// writes {sum xy, sum xx, sum x, sum y, b1, b0}; the coefficient arithmetic is upstream's, operation for
// operation in T (:134-142).  verify(): |b1 - 2| and |b0 - 0| <= 1e-5, upstream's tolerance (:176-178) against
// the closed-form coefficients of the constant inputs, plus the four sums against N, 2N within the
// north-star tolerance.
#include "../pattern/pattern_common.h"

template <class T> struct Linreg;
template <> struct Linreg<float> {
  static int run(sb_ctx* c, const float* x, const float* y, size_t n, float* o) { return sb_linreg_f32(c, x, y, n, o); }
  static int fill(sb_ctx* c, float* p, size_t n, float v) { return sb_fill_f32(c, p, n, v); }
};
template <> struct Linreg<double> {
  static int run(sb_ctx* c, const double* x, const double* y, size_t n, double* o) { return sb_linreg_f64(c, x, y, n, o); }
  static int fill(sb_ctx* c, double* p, size_t n, double v) { return sb_fill_f64(c, p, n, v); }
};

template <class T>
class LinearRegressionCoeffBench {
protected:
  BenchmarkArgs _args;
  PrefetchedBuffer<T> _x, _y, _out;

public:
  LinearRegressionCoeffBench(const BenchmarkArgs& args) : _args{args} {}

  void setup() {
    const std::size_t n = _args.problem_size;
    if(_args.local_size == 0 || n % _args.local_size != 0) throw std::invalid_argument("nd_range: --size must be a multiple of --local");
    _x.initialize(_args.ctx(), std::max<std::size_t>(n, 1));
    _y.initialize(_args.ctx(), std::max<std::size_t>(n, 1));
    sb::check(Linreg<T>::fill(_args.ctx(), _x.get(), n, T(1)), "sb_fill");
    sb::check(Linreg<T>::fill(_args.ctx(), _y.get(), n, T(2)), "sb_fill");
    _out.initialize(_args.ctx(), 6);
  }

  void run(std::vector<sb::event>& events) {
    sb::check(Linreg<T>::run(_args.ctx(), _x.get(), _y.get(), _args.problem_size, _out.get()), "sb_linreg");
    events.push_back({"sb_linreg"});
  }

  bool verify(VerificationSetting&) {
    const auto r = _out.get_host_access();
    const std::size_t n = _args.problem_size;
    const long double ln = (long double)n;
    const bool sums = pattern::value_matches<T>(r[0], 0, 2.0L * ln, n) && pattern::value_matches<T>(r[1], 0, ln, n) &&
                      pattern::value_matches<T>(r[2], 0, ln, n) && pattern::value_matches<T>(r[3], 0, 2.0L * ln, n);
    const T tolerance = T(0.00001);
    return sums && std::fabs(r[4] - T(2)) <= tolerance && std::fabs(r[5] - T(0)) <= tolerance;
  }

  static ThroughputMetric getThroughputMetric(const BenchmarkArgs& args) {
    return {(2.0 * args.problem_size * sizeof(T) + 6 * sizeof(T)) / 1073741824.0, "GiB"};  // algorithmic bytes: 2*N*s + 6*s
  }
  static std::string getBenchmarkName(BenchmarkArgs&) { return std::string("LinearRegressionCoeff_") + ReadableTypename<T>::name; }
};

int main(int argc, char** argv) {
  BenchmarkApp app(argc, argv);
  if(app.shouldRunNDRangeKernels()) {
    app.run<LinearRegressionCoeffBench<float>>();
    app.run<LinearRegressionCoeffBench<double>>();
  }
  return app.exitCode();
}
