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

#define SB_PATTERN_OPS_FULL(T, ABI, SUF)                                                                                \
  template <> struct Ops<T> {                                                                                           \
    SB_PATTERN_OPS_COMMON(T, ABI, SUF)                                                                                  \
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
// integers: every element on the device against the closed form (sb_check_iota); floating point: the
// first and last 2^16 elements plus 2^16 strided samples against the closed form within the tolerance.
// `out` holds elements [offset, offset + n) of a scan over `total` elements (offset = 0, total = n unless sharded).
template <class T>
bool verify_scan(sb_ctx* ctx, const PrefetchedBuffer<T>& out, size_t n, bool exclusive, size_t L, size_t offset = 0, size_t total = 0) {
  if(n == 0) return true;
  if(total == 0) total = n;
  if constexpr(std::is_same_v<T, int> || std::is_same_v<T, long long>) {
    PrefetchedBuffer<unsigned long long> bad;
    bad.initialize(ctx, 1);
    int rc;
    if constexpr(std::is_same_v<T, int>) rc = sb_check_iota_i32(ctx, reinterpret_cast<const int32_t*>(out.get()), n, offset, exclusive ? 1 : 0, L, reinterpret_cast<uint64_t*>(bad.get()));
    else rc = sb_check_iota_i64(ctx, reinterpret_cast<const int64_t*>(out.get()), n, offset, exclusive ? 1 : 0, L, reinterpret_cast<uint64_t*>(bad.get()));
    sb::check(rc, "sb_check_iota");
    return bad.get_host_access()[0] == 0;
  } else {
    const size_t win = std::min<size_t>(n, size_t(1) << 16);
    auto check_window = [&](size_t begin, size_t count) {
      const auto h = out.get_host_access(begin, count);
      for(size_t i = 0; i < count; ++i)
        if(!value_matches<T>(h[i], 0, scan_value_ld(offset + begin + i, exclusive, L), total)) return false;
      return true;
    };
    if(!check_window(0, win) || !check_window(n - win, win)) return false;
    const size_t stride = std::max<size_t>(1, n / win);
    for(size_t g = stride / 2; g < n; g += stride * 64) {
      const size_t count = std::min<size_t>(64, n - g);
      if(!check_window(g, count)) return false;
    }
    return true;
  }
}

} // namespace pattern
