// Shared pieces of the five pattern/ benchmarks: the C-ABI dispatch per element type and the
// closed-form verification of results computed from the reference's iota input
// (in[i] = static_cast<T>(i), /root/reference/pattern/reduction.cpp:31-34).
//
// verify() here is product code: it uses closed forms and libsb_cuda's on-device
// checker, never the oracle.
#pragma once

#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <type_traits>
#include <vector>

#include "common.h"

namespace pattern {

template <class T> struct Ops;

#define SB_PATTERN_OPS_COMMON(T, ABI, SUF)                                                                              \
  using abi_t = ABI;                                                                                                    \
  static abi_t* p(T* x) { return reinterpret_cast<abi_t*>(x); }                                                         \
  static const abi_t* p(const T* x) { return reinterpret_cast<const abi_t*>(x); }                                       \
  static int iota(sb_ctx* c, T* out, size_t n, uint64_t off) { return sb_iota_##SUF(c, p(out), n, off); }               \
  static int segreduce(sb_ctx* c, T* buf, size_t n, size_t L) { return sb_segreduce_##SUF(c, p(buf), n, L); }

// sb_check_iota_<T>: the integer entry points take no tolerance
template <class ABI> struct CheckCall;
template <> struct CheckCall<int32_t> {
  static int call(sb_ctx* c, const int32_t* d, size_t n, uint64_t off, int kind, size_t L, double, uint64_t* bad) { return sb_check_iota_i32(c, d, n, off, kind, L, bad); }
};
template <> struct CheckCall<int64_t> {
  static int call(sb_ctx* c, const int64_t* d, size_t n, uint64_t off, int kind, size_t L, double, uint64_t* bad) { return sb_check_iota_i64(c, d, n, off, kind, L, bad); }
};
template <> struct CheckCall<float> {
  static int call(sb_ctx* c, const float* d, size_t n, uint64_t off, int kind, size_t L, double tol, uint64_t* bad) { return sb_check_iota_f32(c, d, n, off, kind, L, tol, bad); }
};
template <> struct CheckCall<double> {
  static int call(sb_ctx* c, const double* d, size_t n, uint64_t off, int kind, size_t L, double tol, uint64_t* bad) { return sb_check_iota_f64(c, d, n, off, kind, L, tol, bad); }
};

#define SB_PATTERN_OPS_FULL(T, ABI, SUF)                                                                                \
  template <> struct Ops<T> {                                                                                           \
    SB_PATTERN_OPS_COMMON(T, ABI, SUF)                                                                                  \
    static int check(sb_ctx* c, const T* d, size_t n, uint64_t off, int kind, size_t L, double tol, uint64_t* bad) {    \
      return CheckCall<ABI>::call(c, p(d), n, off, kind, L, tol, bad);                                                  \
    }                                                                                                                   \
    static int iota_cyclic(sb_ctx* c, T* out, size_t n, size_t block, int world, int rank) {                            \
      return sb_iota_cyclic_##SUF(c, p(out), n, block, world, rank);                                                    \
    }                                                                                                                   \
    static int check_cyclic(sb_ctx* c, const T* d, size_t n, size_t block, int world, int rank, int kind, double tol, uint64_t* bad) { \
      return sb_check_iota_cyclic_##SUF(c, p(d), n, block, world, rank, kind, tol, bad);                                \
    }                                                                                                                   \
    static int reduce(sb_ctx* c, const T* in, size_t n, T* out) { return sb_reduce_##SUF(c, p(in), n, p(out)); }        \
    static int scan(sb_ctx* c, const T* in, T* out, size_t n, int excl, size_t group) {                                 \
      return sb_scan_##SUF(c, p(in), p(out), n, excl, group, nullptr);                                                  \
    }                                                                                                                   \
    static int segscan(sb_ctx* c, const T* in, T* out, size_t n, size_t L, const uint8_t* flags, int excl) {            \
      return sb_segscan_##SUF(c, p(in), p(out), n, L, flags, excl);                                                     \
    }                                                                                                                   \
  };

template <> struct Ops<short> {
  SB_PATTERN_OPS_COMMON(short, int16_t, i16)
};
SB_PATTERN_OPS_FULL(int, int32_t, i32)
SB_PATTERN_OPS_FULL(long long, int64_t, i64)
SB_PATTERN_OPS_FULL(float, float, f32)
SB_PATTERN_OPS_FULL(double, double, f64)
#undef SB_PATTERN_OPS_COMMON
#undef SB_PATTERN_OPS_FULL

// ---- closed forms, mod 2^64 ---------------------------------------------------------------------------------
inline uint64_t tri(uint64_t g) { return (g & 1) ? g * ((g + 1) >> 1) : (g >> 1) * (g + 1); }  // g(g+1)/2
inline long double tri_ld(long double g) { return g * (g + 1.0L) / 2.0L; }

// inclusive / exclusive scan value at index g of the iota input; L == 0: global, else restart every L
inline uint64_t scan_value(uint64_t g, bool exclusive, uint64_t L) {
  if(L == 0) return exclusive ? (g == 0 ? 0 : tri(g - 1)) : tri(g);
  const uint64_t q = g / L, j = g % L;
  return exclusive ? q * L * j + (j == 0 ? 0 : tri(j - 1)) : q * L * (j + 1) + tri(j);
}
inline long double scan_value_ld(uint64_t g, bool exclusive, uint64_t L) {
  if(L == 0) return exclusive ? (g == 0 ? 0.0L : tri_ld((long double)(g - 1))) : tri_ld((long double)g);
  const long double q = (long double)(g / L), j = (long double)(g % L), l = (long double)L;
  return exclusive ? q * l * j + (j == 0 ? 0.0L : tri_ld(j - 1)) : q * l * (j + 1) + tri_ld(j);
}

// BASELINE.json north_star tolerance: relative, fp32 <= 1e-5*log2 N, fp64 <= 1e-12*log2 N
template <class T>
inline long double tolerance(size_t n) {
  const long double lg = std::max<long double>(1.0L, std::log2((long double)std::max<size_t>(n, 2)));
  return (std::is_same_v<T, float> ? 1e-5L : 1e-12L) * lg;
}

template <class T>
inline bool value_matches(T got, uint64_t exact_mod64, long double exact_real, size_t n) {
  if constexpr(std::is_integral_v<T>) {
    using U = std::make_unsigned_t<T>;
    return static_cast<U>(got) == static_cast<U>(exact_mod64);
  } else {
    const long double err = std::fabs((long double)got - exact_real);
    return err <= tolerance<T>(n) * std::max<long double>(std::fabs(exact_real), 1.0L);
  }
}

// ---- scan-family verification -------------------------------------------------------------------------------
// EVERY element on the device against the closed form (sb_check_iota_<T>): integers bit-exact, floating point
// within the north-star tolerance.  `out` holds elements [offset, offset + n) of a scan over `total` elements
// (offset = 0, total = n unless sharded).  kind: 0 inclusive scan, 1 exclusive scan, 2 segmented-reduction layout.
template <class T>
bool verify_closed_form(sb_ctx* ctx, const T* data, size_t n, int kind, size_t L, size_t offset = 0, size_t total = 0) {
  if(n == 0) return true;
  if(total == 0) total = n;
  PrefetchedBuffer<unsigned long long> bad;
  bad.initialize(ctx, 1);
  sb::check(Ops<T>::check(ctx, data, n, offset, kind, L, (double)tolerance<T>(total), reinterpret_cast<uint64_t*>(bad.get())), "sb_check_iota");
  return bad.get_host_access()[0] == 0;
}
template <class T>
bool verify_scan(sb_ctx* ctx, const PrefetchedBuffer<T>& out, size_t n, bool exclusive, size_t L, size_t offset = 0, size_t total = 0) {
  return verify_closed_form<T>(ctx, out.get(), n, exclusive ? 1 : 0, L, offset, total);
}
// the same for a shard of the block-cyclic layout of sb_scan_cyclic (global scans only)
template <class T>
bool verify_scan_cyclic(sb_ctx* ctx, const PrefetchedBuffer<T>& out, size_t n, bool exclusive, size_t block, int world, int rank) {
  if(n == 0) return true;
  PrefetchedBuffer<unsigned long long> bad;
  bad.initialize(ctx, 1);
  sb::check(Ops<T>::check_cyclic(ctx, out.get(), n, block, world, rank, exclusive ? 1 : 0, (double)tolerance<T>(n * (size_t)world), reinterpret_cast<uint64_t*>(bad.get())),
            "sb_check_iota_cyclic");
  return bad.get_host_access()[0] == 0;
}

} // namespace pattern
