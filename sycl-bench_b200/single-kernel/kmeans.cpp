// single-kernel/kmeans — drop-in for /root/reference/single-kernel/kmeans.cpp.
// Rows: Kmeans_{fp32,fp64} (:118-123,:126-133).  nfeatures = 2, nclusters = 3, features = 2, clusters = 1 (:34-46).
//
// Upstream run() (:49-82) is one map kernel; so is sb_kmeans_<T> (include/sb_cuda.h), including upstream's `index = gid`
// (:74): membership[g] = g when a distance is below FLT_MAX.  The cluster buffer upstream allocates is nclusters * N long
// (:39), of which the kernel reads nclusters * nfeatures elements; the same allocation is kept so that the bytes a run
// touches on the device are comparable.  verify(): upstream's host loop (:84-116), element by element.
#include "../pattern/pattern_common.h"
#include "../sycl2020/ones_reduction.h"  // sycl2020::Fill

template <class T> struct Kmeans;
template <> struct Kmeans<float> {
  static int run(sb_ctx* c, const float* f, const float* cl, size_t n, int nf, int nc, int* m) { return sb_kmeans_f32(c, f, cl, n, nf, nc, (int32_t*)m); }
  static int fill(sb_ctx* c, float* p, size_t n, float v) { return sb_fill_f32(c, p, n, v); }
};
template <> struct Kmeans<double> {
  static int run(sb_ctx* c, const double* f, const double* cl, size_t n, int nf, int nc, int* m) { return sb_kmeans_f64(c, f, cl, n, nf, nc, (int32_t*)m); }
  static int fill(sb_ctx* c, double* p, size_t n, double v) { return sb_fill_f64(c, p, n, v); }
};

template <class T>
class KmeansBench {
protected:
  BenchmarkArgs _args;
  static constexpr int nfeatures = 2, nclusters = 3;
  PrefetchedBuffer<T> _features, _clusters;
  PrefetchedBuffer<int> _membership;

public:
  KmeansBench(const BenchmarkArgs& args) : _args{args} {}

  void setup() {
    const std::size_t n = _args.problem_size;
    _features.initialize(_args.ctx(), std::max<std::size_t>(nfeatures * n, 1));
    _clusters.initialize(_args.ctx(), std::max<std::size_t>(nclusters * n, nclusters * nfeatures));
    _membership.initialize(_args.ctx(), std::max<std::size_t>(n, 1));
    sb::check(Kmeans<T>::fill(_args.ctx(), _features.get(), nfeatures * n, T(2)), "sb_fill");
    sb::check(Kmeans<T>::fill(_args.ctx(), _clusters.get(), std::max<std::size_t>(nclusters * n, nclusters * nfeatures), T(1)), "sb_fill");
    sb::check(sb_memset(_args.ctx(), _membership.get(), 0, std::max<std::size_t>(n, 1) * sizeof(int)), "sb_memset");
  }

  void run(std::vector<sb::event>& events) {
    sb::check(Kmeans<T>::run(_args.ctx(), _features.get(), _clusters.get(), _args.problem_size, nfeatures, nclusters, _membership.get()), "sb_kmeans");
    events.push_back({"sb_kmeans"});
  }

  bool verify(VerificationSetting&) {
    const std::size_t n = _args.problem_size;
    if(n == 0) return true;
    const std::vector<int> got = _membership.get_host_access(0, n);
    const std::vector<T> features = _features.get_host_access(0, nfeatures * n);
    const std::vector<T> clusters = _clusters.get_host_access(0, nclusters * nfeatures);
    for(std::size_t x = 0; x < n; ++x) {  // kmeans.cpp:90-106
      int index = 0;
      T min_dist = T(3.402823466e+38f);
      for(int i = 0; i < nclusters; ++i) {
        T dist = 0;
        for(int l = 0; l < nfeatures; ++l) {
          const T d = features[l * n + x] - clusters[i * nfeatures + l];
          dist += d * d;
        }
        if(dist < min_dist) {
          min_dist = dist;
          index = (int)x;
        }
      }
      if(got[x] != index) {
        std::cout << "Fail at = " << x << " Expected = " << index << " Actual = " << got[x] << std::endl;
        return false;
      }
    }
    return true;
  }

  static ThroughputMetric getThroughputMetric(const BenchmarkArgs& args) {
    return {(double)args.problem_size * (nfeatures * sizeof(T) + sizeof(int)) / 1073741824.0, "GiB"};  // features read, membership written
  }
  static std::string getBenchmarkName(BenchmarkArgs&) { return std::string("Kmeans_") + ReadableTypename<T>::name; }
};

int main(int argc, char** argv) {
  BenchmarkApp app(argc, argv);
  app.run<KmeansBench<float>>();
  app.run<KmeansBench<double>>();
  return app.exitCode();
}
