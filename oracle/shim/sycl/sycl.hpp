// Host-only stand-in for <sycl/sycl.hpp>, just large enough to compile the UNMODIFIED reference
// translation units /root/reference/pattern/{reduction,segmentedreduction}.cpp, the reduction trio
// /root/reference/sycl2020/{atomics/atomic_reduction,group_algorithms/reduce_over_group,
// kernel_reduction/kernel_reduction}.cpp (and the harness headers they include) with plain g++.  No SYCL implementation exists in this image
// (SURVEY.md F5), so this is how the reference's own arithmetic is executed for the oracle and for
// the CPU baseline ("reference source on an in-repo host shim").
//
// TEST INFRASTRUCTURE ONLY — never included by the product (libsb_cuda / harness / pattern).
//
// Execution model (what a CPU SYCL backend such as AdaptiveCpp's OpenMP backend does):
//   * queue::submit runs the command group immediately and synchronously;
//   * parallel_for_work_group: `#pragma omp parallel for` over work-groups,
//     parallel_for_work_item = a plain loop over the local range;
//   * parallel_for(nd_range) with group_barrier: one ucontext fiber per work-item, round-robin
//     between barriers (slow; for small parity checks only);
//   * local_accessor memory is a per-thread arena;
//   * buffers have reference semantics (shared implementation), "device" memory is host memory.
//
// Extra, not SYCL: when the environment variable SYCL_SHIM_DUMP names a directory, every host
// access to a buffer writes the buffer's bytes to <dir>/<sequence>.bin so that tests can compare
// the reference's RESULT VALUES (which it never prints) with the oracle and with libsb_cuda.
#pragma once

#include <ucontext.h>

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace sycl {

// ---- exception ------------------------------------------------------------------------------------
class exception : public std::runtime_error {
public:
  explicit exception(const std::string& what) : std::runtime_error(what) {}
};

// ---- range / id -------------------------------------------------------------------------------------
namespace detail {
template <int D>
struct array_base {
  std::size_t v[D];
  std::size_t& operator[](int i) { return v[i]; }
  const std::size_t& operator[](int i) const { return v[i]; }
  std::size_t get(int i) const { return v[i]; }
};
} // namespace detail

template <int D = 1>
class range : public detail::array_base<D> {
public:
  template <int DD = D, typename = std::enable_if_t<DD == 1>>
  range(std::size_t a) { this->v[0] = a; }
  template <int DD = D, typename = std::enable_if_t<DD == 2>>
  range(std::size_t a, std::size_t b) { this->v[0] = a; this->v[1] = b; }
  template <int DD = D, typename = std::enable_if_t<DD == 3>>
  range(std::size_t a, std::size_t b, std::size_t c) { this->v[0] = a; this->v[1] = b; this->v[2] = c; }
  std::size_t size() const {
    std::size_t s = 1;
    for(int i = 0; i < D; ++i) s *= this->v[i];
    return s;
  }
};

template <int D = 1>
class id : public detail::array_base<D> {
public:
  id() { for(int i = 0; i < D; ++i) this->v[i] = 0; }
  template <int DD = D, typename = std::enable_if_t<DD == 1>>
  id(std::size_t a) { this->v[0] = a; }
  template <int DD = D, typename = std::enable_if_t<DD == 2>>
  id(std::size_t a, std::size_t b) { this->v[0] = a; this->v[1] = b; }
  template <int DD = D, typename = std::enable_if_t<DD == 3>>
  id(std::size_t a, std::size_t b, std::size_t c) { this->v[0] = a; this->v[1] = b; this->v[2] = c; }
  template <int DD = D, typename = std::enable_if_t<DD == 1>>
  operator std::size_t() const { return this->v[0]; }
};

// ---- device / selectors / properties ------------------------------------------------------------------
enum class aspect { cpu, gpu, fp16, fp64, usm_device_allocations, usm_host_allocations, usm_shared_allocations };

namespace info {
namespace device {
struct name { using return_type = std::string; };
struct vendor { using return_type = std::string; };
} // namespace device
namespace event_profiling {
struct command_submit { using return_type = std::uint64_t; };
struct command_start { using return_type = std::uint64_t; };
struct command_end { using return_type = std::uint64_t; };
} // namespace event_profiling
} // namespace info

class device {
public:
  template <typename Param>
  typename Param::return_type get_info() const {
    if constexpr(std::is_same_v<Param, info::device::name>) {
      int threads = 1;
#ifdef _OPENMP
      threads = omp_get_max_threads();
#endif
      return "sycl-bench host shim (OpenMP; " + std::to_string(threads) + " threads)";
    } else {
      return typename Param::return_type{};
    }
  }
  bool has(aspect a) const { return a == aspect::cpu || a == aspect::fp64; }
  bool is_cpu() const { return true; }
  bool is_gpu() const { return false; }
};

struct cpu_selector_t {};
struct gpu_selector_t {};
struct default_selector_t {};
inline constexpr cpu_selector_t cpu_selector_v{};
inline constexpr gpu_selector_t gpu_selector_v{};
inline constexpr default_selector_t default_selector_v{};

namespace property {
namespace queue {
struct in_order {};
struct enable_profiling {};
} // namespace queue
} // namespace property

class property_list {
public:
  template <typename... P>
  property_list(P...) {}
};

// ---- event ---------------------------------------------------------------------------------------------
class event {
public:
  std::uint64_t t_submit = 0, t_start = 0, t_end = 0;
  void wait() {}
  void wait_and_throw() {}
  template <typename Param>
  typename Param::return_type get_profiling_info() const {
    if constexpr(std::is_same_v<Param, info::event_profiling::command_submit>) return t_submit;
    else if constexpr(std::is_same_v<Param, info::event_profiling::command_start>) return t_start;
    else return t_end;
  }
};

// ---- access enums ----------------------------------------------------------------------------------------
namespace access {
enum class mode { read, write, read_write, discard_write, discard_read_write, atomic };
enum class target { device, host_task, constant_buffer, local, host_buffer, global_buffer = device };
} // namespace access
using access_mode = access::mode;
using target = access::target;

class handler;

namespace detail {

inline std::uint64_t now_ns() {
  return (std::uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(
      std::chrono::steady_clock::now().time_since_epoch()).count();
}

// per-thread local-memory arena for local_accessor
inline thread_local char* tls_local_base = nullptr;
inline thread_local std::vector<char> tls_local_arena;

inline void dump_host_access(const void* data, std::size_t bytes) {
  const char* dir = std::getenv("SYCL_SHIM_DUMP");
  if(!dir || !*dir) return;
  static std::atomic<int> seq{0};
  std::size_t limit = bytes;
  if(const char* mx = std::getenv("SYCL_SHIM_DUMP_MAX_BYTES")) limit = std::min<std::size_t>(bytes, std::strtoull(mx, nullptr, 10));
  char path[4096];
  std::snprintf(path, sizeof(path), "%s/%04d.bin", dir, seq.fetch_add(1));
  if(FILE* f = std::fopen(path, "wb")) {
    std::fwrite(data, 1, limit, f);
    std::fclose(f);
  }
}

template <typename T>
struct buffer_impl {
  std::vector<T> storage; // the "device" allocation
  std::size_t count = 0;
};

} // namespace detail

// ---- accessors ---------------------------------------------------------------------------------------------
template <typename T, int D = 1, access::mode M = access::mode::read_write>
class accessor {
public:
  accessor() = default;
  accessor(T* p, std::size_t n, std::size_t off = 0) : ptr(p), count(n), offset(off) {}
  T& operator[](std::size_t i) const { return ptr[offset + i]; }
  std::size_t size() const { return count; }
  T* get_pointer() const { return ptr + offset; }

private:
  T* ptr = nullptr;
  std::size_t count = 0;
  std::size_t offset = 0;
};

template <typename T, int D = 1>
class host_accessor {
public:
  host_accessor(T* p, std::size_t n) : ptr(p), count(n) {}
  T& operator[](std::size_t i) const { return ptr[i]; }
  std::size_t size() const { return count; }

private:
  T* ptr;
  std::size_t count;
};

template <typename T, int D = 1>
class local_accessor {
public:
  local_accessor() = default;
  inline local_accessor(range<D> r, handler& cgh);
  T& operator[](std::size_t i) const { return reinterpret_cast<T*>(detail::tls_local_base + offset)[i]; }
  std::size_t size() const { return count; }

private:
  std::size_t offset = 0;
  std::size_t count = 0;
};

// ---- buffer (reference semantics) ------------------------------------------------------------------------------
template <typename T, int D = 1>
class buffer {
public:
  explicit buffer(range<D> r) : impl(std::make_shared<detail::buffer_impl<T>>()), rng(r) {
    impl->count = r.size();
    impl->storage.resize(impl->count);
  }
  buffer(T* host, range<D> r) : buffer(static_cast<const T*>(host), r) {}
  buffer(const T* host, range<D> r) : impl(std::make_shared<detail::buffer_impl<T>>()), rng(r) {
    impl->count = r.size();
    impl->storage.assign(host, host + impl->count); // the host -> device copy
  }
  void set_write_back(bool) {}
  range<D> get_range() const { return rng; }
  std::size_t size() const { return impl->count; }

  template <access::mode M, access::target Tg = access::target::device>
  accessor<T, D, M> get_access(handler&) { return accessor<T, D, M>(impl->storage.data(), impl->count); }
  template <access::mode M, access::target Tg = access::target::device>
  accessor<T, D, M> get_access(handler&, range<D> r, id<D> off = {}) {
    return accessor<T, D, M>(impl->storage.data(), r.size(), off[0]);
  }
  template <access::mode M>
  host_accessor<T, D> get_access() {
    detail::dump_host_access(impl->storage.data(), impl->count * sizeof(T));
    return host_accessor<T, D>(impl->storage.data(), impl->count);
  }
  template <access::mode M>
  host_accessor<T, D> get_access(range<D> r, id<D> off = {}) {
    detail::dump_host_access(impl->storage.data(), impl->count * sizeof(T));
    return host_accessor<T, D>(impl->storage.data() + off[0], r.size());
  }
  host_accessor<T, D> get_host_access() {
    detail::dump_host_access(impl->storage.data(), impl->count * sizeof(T));
    return host_accessor<T, D>(impl->storage.data(), impl->count);
  }
  T* device_data() const { return impl->storage.data(); } // shim-internal (sycl::reduction)

private:
  std::shared_ptr<detail::buffer_impl<T>> impl;
  range<D> rng;
};

// ---- work-item index classes ---------------------------------------------------------------------------------------
template <int D = 1>
class nd_range {
public:
  nd_range(range<D> g, range<D> l) : global(g), local(l) {}
  range<D> get_global_range() const { return global; }
  range<D> get_local_range() const { return local; }

private:
  range<D> global, local;
};

template <int D = 1>
class h_item {
public:
  h_item(std::size_t l, std::size_t g) : lid(l), gid(g) {}
  std::size_t get_local_id(int) const { return lid; }
  id<D> get_local_id() const { return id<D>(lid); }
  id<D> get_global_id() const { return id<D>(gid); }
  std::size_t get_global_id(int) const { return gid; }

private:
  std::size_t lid, gid;
};

template <int D = 1>
class group {
public:
  group(std::size_t g, std::size_t l, std::size_t n) : gid(g), lsize(l), ngroups(n) {}
  std::size_t get_group_id(int) const { return gid; }
  id<D> get_group_id() const { return id<D>(gid); }
  std::size_t get_local_range(int) const { return lsize; }
  std::size_t get_group_range(int) const { return ngroups; }
  template <typename F>
  void parallel_for_work_item(F&& f) const {
    for(std::size_t l = 0; l < lsize; ++l) f(h_item<D>(l, gid * lsize + l));
  }

private:
  std::size_t gid, lsize, ngroups;
};

template <int D = 1>
class nd_item {
public:
  nd_item(std::size_t l, std::size_t g, std::size_t ls, std::size_t ng) : lid(l), grp(g), lsize(ls), ngroups(ng) {}
  std::size_t get_local_id(int) const { return lid; }
  id<D> get_local_id() const { return id<D>(lid); }
  id<D> get_global_id() const { return id<D>(grp * lsize + lid); }
  std::size_t get_global_id(int) const { return grp * lsize + lid; }
  std::size_t get_group(int) const { return grp; }
  std::size_t get_global_linear_id() const { return grp * lsize + lid; }
  std::size_t get_local_linear_id() const { return lid; }
  std::size_t get_group_linear_id() const { return grp; }
  group<D> get_group() const { return group<D>(grp, lsize, ngroups); }
  std::size_t get_local_range(int) const { return lsize; }

private:
  std::size_t lid, grp, lsize, ngroups;
};

// ---- fibers for nd_range kernels with barriers ---------------------------------------------------------------------------
namespace detail {
struct fiber_group {
  ucontext_t scheduler;
  std::vector<ucontext_t> ctx;
  std::vector<char> stacks;
  std::vector<char> done;
  std::vector<char> collective; // one 16-byte slot per work-item + one for the result (group algorithms)
  std::size_t current = 0;
  std::function<void(std::size_t)> body;
};
inline thread_local fiber_group* tls_fibers = nullptr;

inline void fiber_entry() {
  fiber_group* fg = tls_fibers;
  const std::size_t me = fg->current;
  fg->body(me);
  fg->done[me] = 1;
  swapcontext(&fg->ctx[me], &fg->scheduler);
}

inline void fiber_barrier() {
  fiber_group* fg = tls_fibers;
  if(!fg) return; // hierarchical / single work-item: nothing to wait for
  const std::size_t me = fg->current;
  swapcontext(&fg->ctx[me], &fg->scheduler);
}

inline void run_group_with_fibers(fiber_group& fg, std::size_t lsize) {
  constexpr std::size_t STACK = 32 * 1024;
  fg.ctx.resize(lsize);
  fg.stacks.resize(lsize * STACK);
  fg.done.assign(lsize, 0);
  if(fg.collective.size() < (lsize + 1) * 16) fg.collective.resize((lsize + 1) * 16);
  for(std::size_t l = 0; l < lsize; ++l) {
    getcontext(&fg.ctx[l]);
    fg.ctx[l].uc_stack.ss_sp = fg.stacks.data() + l * STACK;
    fg.ctx[l].uc_stack.ss_size = STACK;
    fg.ctx[l].uc_link = &fg.scheduler;
    makecontext(&fg.ctx[l], fiber_entry, 0);
  }
  tls_fibers = &fg;
  bool any = true;
  while(any) {
    any = false;
    for(std::size_t l = 0; l < lsize; ++l) {
      if(fg.done[l]) continue;
      fg.current = l;
      swapcontext(&fg.scheduler, &fg.ctx[l]);
      if(!fg.done[l]) any = true;
    }
  }
  tls_fibers = nullptr;
}
} // namespace detail

template <int D>
inline void group_barrier(const group<D>&) { detail::fiber_barrier(); }

// ---- SYCL 2020 pieces used by sycl2020/{atomics,group_algorithms,kernel_reduction} ------------------------------
template <typename T = void>
struct plus {
  T operator()(const T& a, const T& b) const { return a + b; }
};

enum class memory_order { relaxed, acquire, release, acq_rel, seq_cst };
enum class memory_scope { work_item, sub_group, work_group, device, system };
namespace access {
enum class address_space { global_space, local_space, constant_space, private_space, generic_space };
}

// fetch_add only: integers through the GCC atomic builtins, floating point through a CAS loop on the bits
template <typename T, memory_order O, memory_scope S, access::address_space A = access::address_space::generic_space>
class atomic_ref {
public:
  explicit atomic_ref(T& r) : p(&r) {}
  T fetch_add(T v) const {
    if constexpr(std::is_integral_v<T>) {
      return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
    } else {
      using U = std::conditional_t<sizeof(T) == 4, std::uint32_t, std::uint64_t>;
      U* up = reinterpret_cast<U*>(p);
      U old = __atomic_load_n(up, __ATOMIC_RELAXED);
      for(;;) {
        T cur;
        std::memcpy(&cur, &old, sizeof(T));
        const T next = cur + v;
        U bits;
        std::memcpy(&bits, &next, sizeof(T));
        if(__atomic_compare_exchange_n(up, &old, bits, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return cur;
      }
    }
  }

private:
  T* p;
};

// reduce_over_group inside an nd_range kernel: every work-item parks its value, work-item 0 folds them in
// local-id order, everybody reads the total (two fiber barriers)
template <int D, typename T, typename Op>
inline T reduce_over_group(const group<D>&, T x, Op op) {
  static_assert(sizeof(T) <= 16, "collective slot too small");
  detail::fiber_group* fg = detail::tls_fibers;
  if(!fg) return x;
  const std::size_t L = fg->ctx.size(), me = fg->current;
  auto slot = [&](std::size_t i) { return reinterpret_cast<T*>(fg->collective.data() + i * 16); };
  *slot(me) = x;
  detail::fiber_barrier();
  if(me == 0) {
    T acc = *slot(0);
    for(std::size_t l = 1; l < L; ++l) acc = op(acc, *slot(l));
    *slot(L) = acc;
  }
  detail::fiber_barrier();
  return *slot(L);
}

namespace detail {
template <typename T, typename Op>
struct reduction_desc {
  T* target; // combined with the value already stored there (no initialize_to_identity property)
  Op op;
};
template <typename T, typename Op>
struct reducer {
  T value;
  Op op;
  void combine(const T& v) { value = op(value, v); }
  reducer& operator+=(const T& v) {
    combine(v);
    return *this;
  }
};
} // namespace detail

// ---- handler ----------------------------------------------------------------------------------------------------------------
class handler {
public:
  std::size_t local_bytes = 0;

  std::size_t reserve_local(std::size_t bytes) {
    const std::size_t off = (local_bytes + 63) & ~std::size_t(63);
    local_bytes = off + bytes;
    return off;
  }

  template <typename F>
  void single_task(F&& f) { f(); }
  template <typename Name, typename F>
  void single_task(F&& f) { f(); }

  template <typename Name = void, typename F>
  void parallel_for_work_group(range<1> num_groups, range<1> local, F f) {
    const std::size_t G = num_groups[0], L = local[0];
    const std::size_t lb = local_bytes + 64;
#pragma omp parallel
    {
      if(detail::tls_local_arena.size() < lb) detail::tls_local_arena.resize(lb);
      detail::tls_local_base = detail::tls_local_arena.data();
#pragma omp for schedule(static)
      for(std::size_t g = 0; g < G; ++g) f(group<1>(g, L, G));
    }
  }

  // basic parallel_for(range, f(id)).  Executed in ascending index order on one thread: the only user on the
  // path, the gather kernel of single-kernel/scalar_prod.cpp:172-178 (`g[idx] = g[idx * L]` for all idx in
  // parallel), reads elements that other work-items write; ascending order is the schedule under which
  // every read sees the value of the previous pass.
  template <typename Name = void, typename F>
  void parallel_for(range<1> r, F f) {
    for(std::size_t i = 0; i < r[0]; ++i) f(id<1>(i));
  }

  // parallel_for(range, reduction, f(id, reducer&)): one private reducer per OpenMP thread over a static
  // schedule, partials folded in thread order (the combination order is implementation-defined in SYCL)
  template <typename Name = void, typename T, typename Op, typename F>
  void parallel_for(range<1> r, detail::reduction_desc<T, Op> red, F f) {
    const std::size_t n = r[0];
    std::vector<T> partial;
#pragma omp parallel
    {
#ifdef _OPENMP
      const int tid = omp_get_thread_num(), nth = omp_get_num_threads();
#else
      const int tid = 0, nth = 1;
#endif
#pragma omp single
      partial.assign(nth, T(0));
      detail::reducer<T, Op> rd{T(0), red.op};
#pragma omp for schedule(static)
      for(std::size_t i = 0; i < n; ++i) f(id<1>(i), rd);
      partial[tid] = rd.value;
    }
    T acc = *red.target;
    for(const T& p : partial) acc = red.op(acc, p);
    *red.target = acc;
  }

  template <typename Name = void, typename F>
  void parallel_for(nd_range<1> ndr, F f) {
    const std::size_t L = ndr.get_local_range()[0];
    if(L == 0 || ndr.get_global_range()[0] % L != 0) throw exception("nd_range: global size is not a multiple of the local size");
    const std::size_t G = ndr.get_global_range()[0] / L;
    const std::size_t lb = local_bytes + 64;
#pragma omp parallel
    {
      if(detail::tls_local_arena.size() < lb) detail::tls_local_arena.resize(lb);
      detail::tls_local_base = detail::tls_local_arena.data();
      detail::fiber_group fg;
#pragma omp for schedule(static)
      for(std::size_t g = 0; g < G; ++g) {
        fg.body = [&, g](std::size_t l) { f(nd_item<1>(l, g, L, G)); };
        detail::run_group_with_fibers(fg, L);
      }
    }
  }
};

template <typename T, int D, typename Op>
inline detail::reduction_desc<T, Op> reduction(buffer<T, D>& b, handler&, Op op) {
  return detail::reduction_desc<T, Op>{b.device_data(), op};
}

template <typename T, int D>
inline local_accessor<T, D>::local_accessor(range<D> r, handler& cgh) : offset(cgh.reserve_local(r.size() * sizeof(T))), count(r.size()) {}

// ---- queue ---------------------------------------------------------------------------------------------------------------------
class queue {
public:
  queue() = default;
  queue(const property_list&) {}
  template <typename Selector>
  queue(const Selector&, const property_list& = {}) {
    if(std::is_same_v<Selector, gpu_selector_t>) throw exception("host shim: no GPU device (this build of the reference is CPU-only)");
  }

  template <typename F>
  event submit(F&& cgf) {
    event e;
    e.t_submit = detail::now_ns();
    handler h;
    e.t_start = detail::now_ns();
    cgf(h);
    e.t_end = detail::now_ns();
    return e;
  }
  void wait() {}
  void wait_and_throw() {}
  device get_device() const { return device{}; }

  template <typename T>
  event copy(const T* src, T* dst, std::size_t count) {
    std::memcpy(dst, src, count * sizeof(T));
    return event{};
  }
  template <typename T>
  event copy(const T* src, T* dst, std::size_t count, event) { return copy(src, dst, count); }
};

// ---- USM (only so that USMBuffer in include/memory_wrappers.h parses) ------------------------------------------------------------
namespace usm {
enum class alloc { host, device, shared, unknown };
}
inline void* malloc(std::size_t bytes, const queue&, usm::alloc) { return std::malloc(bytes); }
inline void* malloc_host(std::size_t bytes, const queue&) { return std::malloc(bytes); }
inline void* malloc_device(std::size_t bytes, const queue&) { return std::malloc(bytes); }
inline void free(void* p, const queue&) { std::free(p); }

} // namespace sycl
