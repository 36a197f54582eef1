// arena.h -- sub-allocation of device memory.  cudaMalloc / cudaFree cost milliseconds each on a B200 box (they
// synchronise the device and map memory: 2-9 ms per call measured, tools/time_to_solution.py), and an AMR hierarchy of
// the reference's default run needs about 450 arrays: the library takes device memory from the driver in a few large
// chunks and hands out ranges of them.  This header is the bookkeeping only -- offsets in one chunk, no CUDA calls -- so
// that it can be exercised on a machine without a GPU (mgic_arena_selftest).
#ifndef MGIC_ARENA_H
#define MGIC_ARENA_H

#include <cstddef>
#include <map>

// first-fit free list over [0, size) with coalescing; every range is a multiple of `align` bytes
class RangeAllocator {
 public:
  RangeAllocator(size_t size, size_t align) : size_(size), align_(align) { free_[0] = size; }
  // offset of a free range of at least `bytes`, or (size_t)-1
  size_t take(size_t bytes) {
    const size_t need = round_up(bytes ? bytes : 1);
    for (auto it = free_.begin(); it != free_.end(); ++it) {
      if (it->second < need) continue;
      const size_t off = it->first, left = it->second - need;
      free_.erase(it);
      if (left) free_[off + need] = left;
      used_[off] = need;
      inUse_ += need;
      return off;
    }
    return (size_t)-1;
  }
  // false: `off` is not the start of a range handed out by take()
  bool give(size_t off) {
    auto u = used_.find(off);
    if (u == used_.end()) return false;
    size_t len = u->second;
    used_.erase(u);
    inUse_ -= len;
    auto nx = free_.lower_bound(off);
    if (nx != free_.end() && off + len == nx->first) { len += nx->second; nx = free_.erase(nx); }   // merge with the next free range
    if (nx != free_.begin()) {
      auto pv = std::prev(nx);
      if (pv->first + pv->second == off) { pv->second += len; return true; }                         // ... and with the previous
    }
    free_[off] = len;
    return true;
  }
  size_t in_use() const { return inUse_; }
  size_t size() const { return size_; }
  size_t ranges_in_use() const { return used_.size(); }
  size_t free_ranges() const { return free_.size(); }
  const std::map<size_t, size_t> &used() const { return used_; }

 private:
  size_t round_up(size_t b) const { return (b + align_ - 1) / align_ * align_; }
  size_t size_, align_, inUse_ = 0;
  std::map<size_t, size_t> free_, used_;   // offset -> length
};

#endif
