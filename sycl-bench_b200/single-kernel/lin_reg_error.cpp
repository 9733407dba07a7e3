// single-kernel/lin_reg_error — drop-in for /root/reference/single-kernel/lin_reg_error.cpp.
// Rows: LinearRegression_{fp32,fp64} (:120-125,:128-135).  Inputs: in1, in2, alpha, beta = rand() / RAND_MAX, four draws
// per element in that order, default seed, one stream over both types (:30-43) — generated the same way here.
//
// Upstream run() (:54-79): one work-item per (alpha, beta) candidate, each looping over all N points: N^2 multiply-adds.
// Here: sb_linreg_error_<T> — the five sums of the quadratic form in one streaming pass over in1 / in2, then one map over
// alpha / beta (include/sb_cuda.h): 5*N*s bytes, two launches.
// verify(): upstream recomputes the N^2 loop on the host in T and accepts a relative L2 error below 1e-6 (:80-118).
// Here (a) the same criterion over ALL candidates against the quadratic form evaluated on the host in long double, and (b) the
// literal O(N) loop in long double for up to 256 evenly spaced candidates, element by element within the north-star tolerance.
#include "../pattern/pattern_common.h"

template <class T> struct LinregError;
template <> struct LinregError<float> {
  static int run(sb_ctx* c, const float* x, const float* y, const float* a, const float* b, size_t n, float* o) { return sb_linreg_error_f32(c, x, y, a, b, n, o); }
};
template <> struct LinregError<double> {
  static int run(sb_ctx* c, const double* x, const double* y, const double* a, const double* b, size_t n, double* o) { return sb_linreg_error_f64(c, x, y, a, b, n, o); }
};

template <class T>
class LinearRegressionBench {
protected:
  BenchmarkArgs _args;
  std::vector<T> _h1, _h2, _ha, _hb;
  PrefetchedBuffer<T> _in1, _in2, _alpha, _beta, _out;

public:
  LinearRegressionBench(const BenchmarkArgs& args) : _args{args} {}

  void setup() {
    const std::size_t n = _args.problem_size;
    _h1.resize(n), _h2.resize(n), _ha.resize(n), _hb.resize(n);
    for(std::size_t i = 0; i < n; ++i) {  // lin_reg_error.cpp:38-43
      _h1[i] = static_cast<T>(rand()) / static_cast<T>(RAND_MAX);
      _h2[i] = static_cast<T>(rand()) / static_cast<T>(RAND_MAX);
      _ha[i] = static_cast<T>(rand()) / static_cast<T>(RAND_MAX);
      _hb[i] = static_cast<T>(rand()) / static_cast<T>(RAND_MAX);
    }
    const std::size_t m = std::max<std::size_t>(n, 1);
    _in1.initialize(_args.ctx(), m), _in2.initialize(_args.ctx(), m), _alpha.initialize(_args.ctx(), m), _beta.initialize(_args.ctx(), m);
    _out.initialize(_args.ctx(), m);
    if(n) {
      sb::check(sb_memcpy_h2d(_args.ctx(), _in1.get(), _h1.data(), n * sizeof(T)), "sb_memcpy_h2d");
      sb::check(sb_memcpy_h2d(_args.ctx(), _in2.get(), _h2.data(), n * sizeof(T)), "sb_memcpy_h2d");
      sb::check(sb_memcpy_h2d(_args.ctx(), _alpha.get(), _ha.data(), n * sizeof(T)), "sb_memcpy_h2d");
      sb::check(sb_memcpy_h2d(_args.ctx(), _beta.get(), _hb.data(), n * sizeof(T)), "sb_memcpy_h2d");
      sb::check(sb_ctx_sync(_args.ctx()), "sb_ctx_sync");
    }
  }

  void run(std::vector<sb::event>& events) {
    sb::check(LinregError<T>::run(_args.ctx(), _in1.get(), _in2.get(), _alpha.get(), _beta.get(), _args.problem_size, _out.get()), "sb_linreg_error");
    events.push_back({"sb_linreg_error: sums"});
    events.push_back({"sb_linreg_error: map"});
  }

  bool verify(VerificationSetting&) {
    const std::size_t n = _args.problem_size;
    if(n == 0) return true;
    const std::vector<T> out = _out.get_host_access(0, n);
    long double sx = 0, sy = 0, sxx = 0, sxy = 0, syy = 0;
    for(std::size_t i = 0; i < n; ++i) {
      const long double x = _h1[i], y = _h2[i];
      sx += x, sy += y, sxx += x * x, sxy += x * y, syy += y * y;
    }
    // (a) upstream's compare() (:80-103) over all candidates
    long double err2 = 0, ref2 = 0;
    for(std::size_t g = 0; g < n; ++g) {
      const long double a = _ha[g], b = _hb[g];
      const long double e = a * a * sxx + 2 * a * b * sx - 2 * a * sxy + (long double)n * b * b - 2 * b * sy + syy;
      const long double d = e - (long double)out[g];
      err2 += d * d, ref2 += e * e;
    }
    if(!(std::fabs((double)ref2) >= 1e-7)) return false;  // :94-96
    if(!(std::sqrt((double)err2) / std::sqrt((double)ref2) < 0.000001)) return false;
    // (b) the literal loop (:64-77) for a sample of candidates
    const double tol = (double)pattern::tolerance<T>(n);
    const std::size_t samples = std::max<std::size_t>(1, std::min<std::size_t>(256, (std::size_t(1) << 28) / n));  // bounded host work
    const std::size_t step = std::max<std::size_t>(n / samples, 1);
    for(std::size_t g = 0; g < n; g += step) {
      long double e = 0;
      for(std::size_t i = 0; i < n; ++i) {
        const long double r = ((long double)_ha[g] * _h1[i] + _hb[g]) - _h2[i];
        e += r * r;
      }
      if(!(std::fabs((double)(e - (long double)out[g])) <= tol * std::max<double>(std::fabs((double)e), 1.0))) return false;
    }
    return true;
  }

  static ThroughputMetric getThroughputMetric(const BenchmarkArgs& args) {
    return {5.0 * args.problem_size * sizeof(T) / 1073741824.0, "GiB"};  // algorithmic bytes: in1, in2, alpha, beta read, out written
  }
  static std::string getBenchmarkName(BenchmarkArgs&) { return std::string("LinearRegression_") + ReadableTypename<T>::name; }
};

int main(int argc, char** argv) {
  BenchmarkApp app(argc, argv);
  app.run<LinearRegressionBench<float>>();
  app.run<LinearRegressionBench<double>>();
  return app.exitCode();
}
