// alloc.cu -- where the library's device arrays come from: a sub-allocator over a few large cudaMalloc chunks (bookkeeping in
// arena.h), optional guard bands around every array, and the counters behind mgic_alloc_stats.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

#include "mgic_internal.h"
#include "arena.h"

// ---- device allocations of the library.  cudaMalloc / cudaFree synchronise the device and map memory: 2-9 ms per call on
// a B200 box, and the reference's default hierarchy (7 levels, 11 arrays-of-boxes) needs about 450 arrays -- 3 of the 4.3 s
// from set_grids to the converged nonlinear loop were spent there (tools/time_to_solution.py).  Requests below
// ARENA_DIRECT are ranges of a few large chunks per device (arena.h; chunks are kept until the process ends), larger ones
// and arrays that CUDA IPC has to name (multi-rank fields: `plain`) are blocks of their own.  MGIC_ARENA=0 turns it off.
// mgic_alloc_stats counts the driver calls and the time inside them.
static std::atomic<long long> g_allocCalls{0}, g_allocNs{0}, g_freeCalls{0}, g_freeNs{0};
static cudaError_t timed_cuda_malloc(void **p, size_t bytes) {
  const auto t0 = std::chrono::steady_clock::now();
  const cudaError_t e = cudaMalloc(p, bytes);
  g_allocNs += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
  g_allocCalls++;
  return e;
}
static cudaError_t timed_cuda_free(void *p) {
  const auto t0 = std::chrono::steady_clock::now();
  const cudaError_t e = cudaFree(p);
  g_freeNs += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
  g_freeCalls++;
  return e;
}
namespace {
constexpr size_t ARENA_ALIGN = 512, ARENA_DIRECT = (size_t)256 << 20, ARENA_FIRST = (size_t)256 << 20, ARENA_MAX_CHUNK = (size_t)4 << 30;
struct DevChunk {
  char *base;
  RangeAllocator ra;
  DevChunk(char *b, size_t n) : base(b), ra(n, ARENA_ALIGN) {}
};
struct DevArena {
  std::vector<DevChunk *> chunks;
  size_t next = ARENA_FIRST;
};
std::mutex g_arenaMu;
std::map<int, DevArena> g_arenas;   // by device
bool arena_on() {
  static const bool on = [] { const char *e = getenv("MGIC_ARENA"); return !(e && e[0] == '0'); }();
  return on;
}
}  // namespace
static cudaError_t raw_dev_malloc(void **p, size_t bytes, bool plain) {
  if (plain || !arena_on() || bytes >= ARENA_DIRECT) return timed_cuda_malloc(p, bytes);
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(g_arenaMu);
  DevArena &A = g_arenas[dev];
  for (DevChunk *c : A.chunks) {
    const size_t off = c->ra.take(bytes);
    if (off != (size_t)-1) { *p = c->base + off; return cudaSuccess; }
  }
  size_t want = std::max(A.next, (bytes + ARENA_ALIGN) * 2);
  want = (want + (((size_t)2 << 20) - 1)) / ((size_t)2 << 20) * ((size_t)2 << 20);
  void *base = nullptr;
  e = timed_cuda_malloc(&base, want);
  if (e != cudaSuccess) {   // no room for a chunk: the request alone
    cudaGetLastError();
    return timed_cuda_malloc(p, bytes);
  }
  A.next = std::min(A.next * 2, ARENA_MAX_CHUNK);
  DevChunk *c = new DevChunk((char *)base, want);
  A.chunks.push_back(c);
  c->ra.take(1);   // no array starts at the chunk's own address: CUDA IPC would take it for a block of its own (comm.cu owns_its_block)
  *p = c->base + c->ra.take(bytes);
  return cudaSuccess;
}
static cudaError_t raw_dev_free(void *p) {
  if (!p) return cudaSuccess;
  {
    std::lock_guard<std::mutex> lk(g_arenaMu);
    for (auto &da : g_arenas)
      for (DevChunk *c : da.second.chunks)
        if ((char *)p >= c->base && (char *)p < c->base + c->ra.size()) {
          // like cudaFree: nothing in flight on the chunk's device may still use the range when it is handed out again
          int cur = da.first;
          cudaGetDevice(&cur);
          if (cur != da.first) cudaSetDevice(da.first);
          const cudaError_t e = cudaDeviceSynchronize();
          if (cur != da.first) cudaSetDevice(cur);
          if (!c->ra.give((size_t)((char *)p - c->base))) return cudaErrorInvalidDevicePointer;
          return e;
        }
  }
  return timed_cuda_free(p);
}
// ---- guard bands (MGIC_ARENA_GUARD=<bytes> in the environment; the GPU test suite sets it): every array gets that many bytes
// of 0xA5 in front of it and behind it, checked when the array is freed and by mgic_arena_guard_check -- a write past either
// end of an array shows up as a violation instead of as a wrong number in some other array.  (compute-sanitizer is not
// available on the GPU pool.)  Arrays exported by CUDA IPC keep their own block and carry no bands.
namespace {
size_t guard_bytes() {
  static const size_t g = [] {
    const char *e = getenv("MGIC_ARENA_GUARD");
    const long long v = e ? atoll(e) : 0;
    return v > 0 ? (size_t)((v + 511) / 512 * 512) : (size_t)0;
  }();
  return g;
}
std::mutex g_guardMu;
std::map<void *, size_t> g_guarded;   // user pointer -> bytes
std::atomic<long long> g_guardViolations{0};
bool guard_intact(char *user, size_t bytes, const char *when) {
  const size_t G = guard_bytes();
  std::vector<unsigned char> h(2 * G);
  cudaDeviceSynchronize();
  if (cudaMemcpy(h.data(), user - G, G, cudaMemcpyDeviceToHost) != cudaSuccess ||
      cudaMemcpy(h.data() + G, user + bytes, G, cudaMemcpyDeviceToHost) != cudaSuccess) {
    cudaGetLastError();
    return true;   // cannot tell (the device is in an error state: the caller sees that elsewhere)
  }
  for (size_t i = 0; i < 2 * G; i++)
    if (h[i] != 0xA5) {
      g_guardViolations++;
      fprintf(stderr, "mgic: guard band violated (%s): array of %zu bytes, byte %lld %s it\n", when, bytes,
              i < G ? (long long)(G - i) : (long long)(i - G), i < G ? "before" : "after the end of");
      return false;
    }
  return true;
}
}  // namespace
cudaError_t mgic_dev_malloc_(void **p, size_t bytes, bool plain) {
  const size_t G = guard_bytes();
  if (!G || plain) return raw_dev_malloc(p, bytes, plain);
  void *raw = nullptr;
  cudaError_t e = raw_dev_malloc(&raw, bytes + 2 * G, false);
  if (e != cudaSuccess) return e;
  char *user = (char *)raw + G;
  if ((e = cudaMemset(raw, 0xA5, G)) != cudaSuccess || (e = cudaMemset(user + bytes, 0xA5, G)) != cudaSuccess) return e;
  cudaDeviceSynchronize();
  {
    std::lock_guard<std::mutex> lk(g_guardMu);
    g_guarded[user] = bytes;
  }
  *p = user;
  return cudaSuccess;
}
cudaError_t mgic_dev_free(void *p) {
  if (!p) return cudaSuccess;
  size_t bytes = 0;
  bool guarded = false;
  {
    std::lock_guard<std::mutex> lk(g_guardMu);
    auto it = g_guarded.find(p);
    if (it != g_guarded.end()) { guarded = true; bytes = it->second; g_guarded.erase(it); }
  }
  if (!guarded) return raw_dev_free(p);
  guard_intact((char *)p, bytes, "at free");
  return raw_dev_free((char *)p - guard_bytes());
}
// checks the bands of every live array now; *violations = all violations seen so far in this process (0 bands: always 0)
extern "C" int mgic_arena_guard_check(long long *violations, long long *arrays_checked) {
  std::vector<std::pair<void *, size_t>> live;
  {
    std::lock_guard<std::mutex> lk(g_guardMu);
    live.assign(g_guarded.begin(), g_guarded.end());
  }
  for (auto &l : live) guard_intact((char *)l.first, l.second, "mgic_arena_guard_check");
  if (violations) *violations = g_guardViolations.load();
  if (arrays_checked) *arrays_checked = (long long)live.size();
  return MGIC_OK;
}
// the bands do their job: an array is overrun by one byte on purpose (and underrun), both must be seen; the two violations are
// taken off the process's count again.  0 = detected; 1 = no bands configured; 2 = missed
extern "C" int mgic_arena_guard_selftest(int device) {
  if (!guard_bytes()) return 1;
  if (cudaSetDevice(device) != cudaSuccess) { mgic_set_error("no CUDA device"); return MGIC_ERR_NO_DEVICE; }
  int seen = 0;
  for (int side = 0; side < 2; side++) {
    char *a = nullptr;
    const size_t bytes = 1000;
    if (mgic_dev_malloc(&a, bytes) != cudaSuccess) return MGIC_ERR_CUDA;
    cudaMemset(side ? a - 1 : a + bytes, 0, 1);
    const long long before = g_guardViolations.load();
    fprintf(stderr, "mgic: (self-test: the next guard-band report is provoked)\n");
    mgic_dev_free(a);
    if (g_guardViolations.load() == before + 1) { seen++; g_guardViolations--; }
  }
  return seen == 2 ? 0 : 2;
}
// bookkeeping self-test without a device: `ops` random takes / gives on a 1 MiB range, checked against a byte map
extern "C" int mgic_arena_selftest(unsigned seed, int ops) {
  const size_t N = (size_t)1 << 20, AL = 512;
  RangeAllocator ra(N, AL);
  std::vector<unsigned char> owner(N / AL, 0);
  std::vector<std::pair<size_t, size_t>> live;   // offset, bytes
  unsigned long long st = seed * 2654435761ull + 12345;
  auto rnd = [&st]() { st = st * 6364136223846793005ull + 1442695040888963407ull; return (unsigned)(st >> 33); };
  for (int it = 0; it < ops; it++) {
    if (live.empty() || rnd() % 3) {
      const size_t bytes = 1 + rnd() % (64 * 1024);
      const size_t off = ra.take(bytes);
      if (off == (size_t)-1) continue;
      if (off % AL || off + bytes > N) return 1;
      for (size_t b = off / AL; b < (off + bytes + AL - 1) / AL; b++) {
        if (owner[b]) return 2;   // overlap with a live range
        owner[b] = 1;
      }
      live.push_back({off, bytes});
    } else {
      const size_t k = rnd() % live.size();
      const size_t off = live[k].first, bytes = live[k].second;
      if (!ra.give(off)) return 3;
      if (ra.give(off)) return 4;   // double free must be refused
      for (size_t b = off / AL; b < (off + bytes + AL - 1) / AL; b++) owner[b] = 0;
      live[k] = live.back(); live.pop_back();
    }
  }
  for (auto &l : live) if (!ra.give(l.first)) return 5;
  if (ra.in_use() != 0 || ra.free_ranges() != 1) return 6;   // everything coalesced back into one range
  if (ra.take(N) != 0) return 7;
  return 0;
}
extern "C" int mgic_alloc_stats(long long *alloc_calls, double *alloc_seconds, long long *free_calls, double *free_seconds) {
  if (alloc_calls) *alloc_calls = g_allocCalls.load();
  if (alloc_seconds) *alloc_seconds = 1e-9 * (double)g_allocNs.load();
  if (free_calls) *free_calls = g_freeCalls.load();
  if (free_seconds) *free_seconds = 1e-9 * (double)g_freeNs.load();
  return MGIC_OK;
}

