// Runtime singleton, logging and the error path.
#include <chrono>
#include "common.h"
#include <map>
#include <unordered_map>
#include <vector>

namespace qb {

Runtime &rt() {
  static Runtime r;
  return r;
}

void fatal(const char *file, int line, const char *func, const char *fmt, ...) {
  Runtime &r = rt();
  FILE *o = r.out ? r.out : stderr;
  fprintf(o, "%sERROR: ", r.prefix);
  va_list ap;
  va_start(ap, fmt);
  vfprintf(o, fmt, ap);
  va_end(ap);
  fprintf(o, " (rank %d, %s:%d in %s())\n", r.rank, file, line, func);
  fflush(o);
  // same behaviour as the reference's errorQuda -> comm_abort: the process ends
  exit(1);
}

void log_msg(int level, const char *fmt, ...) {
  Runtime &r = rt();
  if (r.verbosity < level || r.rank != 0) return;
  FILE *o = r.out ? r.out : stdout;
  fprintf(o, "%s", r.prefix);
  va_list ap;
  va_start(ap, fmt);
  vfprintf(o, fmt, ap);
  va_end(ap);
  fflush(o);
}

static std::multimap<size_t, void *> pool_cache;          // free blocks by size
static std::unordered_map<void *, size_t> pool_live;       // blocks handed out
static double pool_driver_seconds = 0.0;                   // wall time spent inside cudaMalloc / cudaFree on behalf of the pool
static long pool_driver_calls = 0;
static double wall_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
double pool_driver_time(long *calls) { if (calls) *calls = pool_driver_calls; return pool_driver_seconds; }

void *pool_malloc(size_t bytes) {
  if (bytes == 0) bytes = 16;
  auto it = pool_cache.find(bytes);
  void *p = nullptr;
  if (it != pool_cache.end()) {
    p = it->second;
    pool_cache.erase(it);
  } else {
    const double t0 = wall_s();
    cudaError_t e = cudaMalloc(&p, bytes);
    pool_driver_seconds += wall_s() - t0; pool_driver_calls++;
    if (e != cudaSuccess) {
      // out of memory: drop the cache and retry once
      cudaGetLastError();
      for (auto &kv : pool_cache) cudaFree(kv.second);
      pool_cache.clear();
      e = cudaMalloc(&p, bytes);
      if (e != cudaSuccess) QB_ERROR("cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
  }
  pool_live[p] = bytes;
  return p;
}

void pool_free(void *ptr) {
  if (!ptr) return;
  auto it = pool_live.find(ptr);
  if (it == pool_live.end()) { cudaFree(ptr); return; }
  // the block may still be in use by kernels queued on the library's streams; a later user only ever touches it
  // through the same in-order compute stream (or after an event wait on it), so reuse is stream-ordered
  pool_cache.emplace(it->second, ptr);
  pool_live.erase(it);
}

size_t pool_cached_bytes() {
  size_t n = 0;
  for (auto &kv : pool_cache) n += kv.first;
  return n;
}

void pool_release_all() {
  const double t0 = wall_s();
  for (auto &kv : pool_cache) { cudaFree(kv.second); pool_driver_calls++; }
  pool_driver_seconds += wall_s() - t0;
  pool_cache.clear();
}

}  // namespace qb
