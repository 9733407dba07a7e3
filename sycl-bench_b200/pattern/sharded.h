// --num-gpus=N: the problem is sharded over N devices, shard r lives on device (--device-index + r).
//   reduction, prefixsum, scan --scan-scope=global : sb_comm_reduce / sb_comm_scan (include/sb_cuda.h).
//       --scan-strategy=cyclic (default): block-cyclic shards, one pass over the input, the round totals cross NVLink
//       inside the kernel; reduce_then_scan / scan_then_carry: contiguous shards + ONE NCCL collective of one element
//       per rank (BASELINE.json north_star).
//   segmentedreduction, segmentedscan, scan --scan-scope=group : shards cut on segment boundaries, independent
//       replicas of the 1-GPU kernel (sb_comm_foreach), nothing is exchanged.
// The reference has no multi-device mode; with --num-gpus=1 (the default) nothing here is used.
#pragma once

#include <map>
#include <memory>
#include <utility>

#include "pattern_common.h"

namespace pattern {

inline int requested_gpus(const BenchmarkArgs& args) {
  const int n = args.cli.getOrDefault<int>("--num-gpus", 1);
  if(n < 1) throw std::invalid_argument("--num-gpus must be >= 1");
  return n;
}

// one communicator per (context, size) for the life of the process: ncclCommInitAll costs ~1 s.
// main() calls release_comms() before it returns; the table itself is never destroyed, so nothing touches
// CUDA or NCCL during static destruction.
struct CommEntry {
  std::shared_ptr<sb_ctx> keep_alive;  // rank 0's context must outlive the communicator
  sb_comm* comm = nullptr;
};
inline std::map<std::pair<sb_ctx*, int>, CommEntry>& comm_table() {
  static auto* table = new std::map<std::pair<sb_ctx*, int>, CommEntry>();
  return *table;
}
inline sb_comm* comm_for(const BenchmarkArgs& args, int ngpus) {
  auto& slot = comm_table()[{args.ctx(), ngpus}];
  if(!slot.comm) {
    sb::check(sb_comm_create(args.ctx(), ngpus, &slot.comm), "sb_comm_create");
    slot.keep_alive = args.device_ctx;
  }
  return slot.comm;
}
inline void release_comms() {
  for(auto& kv : comm_table()) sb_comm_destroy(kv.second.comm);
  comm_table().clear();
}

// contiguous shard [begin, end) of `rank`; every shard but the last is a multiple of `align` elements: 16 keeps the
// 16-byte alignment the vector paths of sb_reduce / sb_scan want, the segment length cuts on segment boundaries
inline std::pair<std::size_t, std::size_t> shard_bounds(std::size_t n, int world, int rank, std::size_t align = 16) {
  std::size_t per = (n + world - 1) / world;
  per = (per + align - 1) / align * align;
  const std::size_t begin = std::min(n, per * static_cast<std::size_t>(rank));
  return {begin, std::min(n, begin + per)};
}
inline std::size_t lcm16(std::size_t L) {
  std::size_t a = L, b = 16;
  while(b) {
    const std::size_t t = a % b;
    a = b;
    b = t;
  }
  return L / a * 16;
}

template <class T>
class ShardedArray {
public:
  void initialize(sb_comm* comm, std::size_t n, std::size_t align = 16) {
    comm_ = comm;
    const int world = sb_comm_size(comm);
    shards_.clear();
    begin_.assign(world, 0);
    count_.assign(world, 0);
    ptrs_.assign(world, nullptr);
    for(int r = 0; r < world; ++r) {
      const auto [b, e] = shard_bounds(n, world, r, align);
      begin_[r] = b;
      count_[r] = e - b;
      shards_.push_back(std::make_unique<PrefetchedBuffer<T>>());
      shards_.back()->initialize(sb_comm_ctx(comm, r), std::max<std::size_t>(count_[r], 1));
      ptrs_[r] = shards_.back()->get();
    }
  }
  // one element per rank (result slots)
  void initialize_scalars(sb_comm* comm) {
    comm_ = comm;
    const int world = sb_comm_size(comm);
    shards_.clear();
    begin_.assign(world, 0);
    count_.assign(world, 1);
    ptrs_.assign(world, nullptr);
    for(int r = 0; r < world; ++r) {
      shards_.push_back(std::make_unique<PrefetchedBuffer<T>>());
      shards_.back()->initialize(sb_comm_ctx(comm, r), 1);
      ptrs_[r] = shards_.back()->get();
    }
  }
  // block-cyclic layout of sb_scan_cyclic: every rank holds n / world elements (n must be a multiple of world)
  void initialize_cyclic(sb_comm* comm, std::size_t n) {
    comm_ = comm;
    const int world = sb_comm_size(comm);
    if(n % world != 0) throw std::invalid_argument("--scan-strategy=cyclic needs --size to be a multiple of --num-gpus");
    shards_.clear();
    begin_.assign(world, 0);
    count_.assign(world, n / world);
    ptrs_.assign(world, nullptr);
    for(int r = 0; r < world; ++r) {
      shards_.push_back(std::make_unique<PrefetchedBuffer<T>>());
      shards_.back()->initialize(sb_comm_ctx(comm, r), std::max<std::size_t>(count_[r], 1));
      ptrs_[r] = shards_.back()->get();
    }
  }
  void fill_iota_cyclic(std::size_t block) {
    for(int r = 0; r < world(); ++r) {
      sb::check(Ops<T>::iota_cyclic(sb_comm_ctx(comm_, r), shards_[r]->get(), count_[r], block, world(), r), "sb_iota_cyclic");
      sb::check(sb_ctx_sync(sb_comm_ctx(comm_, r)), "sb_ctx_sync");
    }
  }
  void fill_iota() {
    for(int r = 0; r < world(); ++r) {
      sb::check(Ops<T>::iota(sb_comm_ctx(comm_, r), shards_[r]->get(), count_[r], begin_[r]), "sb_iota");
      sb::check(sb_ctx_sync(sb_comm_ctx(comm_, r)), "sb_ctx_sync");
    }
  }
  int world() const { return static_cast<int>(shards_.size()); }
  const PrefetchedBuffer<T>& shard(int r) const { return *shards_[r]; }
  std::size_t begin(int r) const { return begin_[r]; }
  const std::size_t* counts() const { return count_.data(); }
  std::size_t count(int r) const { return count_[r]; }
  const void* const* in_ptrs() const { return reinterpret_cast<const void* const*>(ptrs_.data()); }
  void* const* out_ptrs() const { return reinterpret_cast<void* const*>(ptrs_.data()); }

private:
  sb_comm* comm_ = nullptr;
  std::vector<std::unique_ptr<PrefetchedBuffer<T>>> shards_;
  std::vector<std::size_t> begin_, count_;
  std::vector<T*> ptrs_;
};

template <class T> struct DtypeOf;
template <> struct DtypeOf<int> { static constexpr int value = SB_I32; };
template <> struct DtypeOf<long long> { static constexpr int value = SB_I64; };
template <> struct DtypeOf<float> { static constexpr int value = SB_F32; };
template <> struct DtypeOf<double> { static constexpr int value = SB_F64; };

} // namespace pattern
