// Common definitions for the B200 twisted-mass / multigrid engine (host + device).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace qb {

// Error model of the reference: formatted message, then the process exits
// (/root/reference/include/util_quda.h:50-60 errorQuda -> comm_abort).
[[noreturn]] void fatal(const char *file, int line, const char *func, const char *fmt, ...);
void log_msg(int level, const char *fmt, ...);  // level: 1 summarize, 2 verbose, 3 debug

#define QB_ERROR(...) ::qb::fatal(__FILE__, __LINE__, __func__, __VA_ARGS__)
#define QB_CUDA(call)                                                         \
  do {                                                                        \
    cudaError_t qb_err_ = (call);                                             \
    if (qb_err_ != cudaSuccess) QB_ERROR("%s failed: %s", #call, cudaGetErrorString(qb_err_)); \
  } while (0)
#define QB_CHECK_LAUNCH()                                                      \
  do {                                                                         \
    cudaError_t qb_err_ = cudaGetLastError();                                  \
    if (qb_err_ != cudaSuccess) QB_ERROR("kernel launch failed: %s", cudaGetErrorString(qb_err_)); \
    ::qb::rt().launches++;                                                     \
  } while (0)

enum Prec { PREC_HALF = 2, PREC_SINGLE = 4, PREC_DOUBLE = 8 };

// process-global runtime state (the reference keeps the same things in file-static globals,
// /root/reference/lib/interface_quda.cpp:80-160, :462-499)
struct Runtime {
  bool device_ready = false, memory_ready = false;
  int device = 0;
  int num_sms = 148;
  cudaStream_t compute = nullptr;  // all operator / BLAS kernels
  cudaStream_t halo = nullptr;     // pack + exchange, overlapped with interior compute
  cudaEvent_t ev_pack_ready = nullptr, ev_halo_done = nullptr, ev_in_ready = nullptr;
  long long launches = 0;
  int verbosity = 1;
  FILE *out = nullptr;
  char prefix[64] = "";
  // comms
  int rank = 0, size = 1;
  int grid[4] = {1, 1, 1, 1};
  int coord[4] = {0, 0, 0, 0};
  int part_mask = 0;  // bit d: dimension d is partitioned (real ranks or forced self-exchange)
  bool grid_set = false;
};
Runtime &rt();

// Size-bucketed caching allocator for field storage: cudaMalloc / cudaFree synchronise the device and cost
// milliseconds at field sizes, and solvers create and drop work vectors constantly (the reference has the same
// design: lib/malloc.cpp + the pinned/device pools of later QUDA).  Freed blocks are kept and reused by exact size.
void *pool_malloc(size_t bytes);
void pool_free(void *ptr);
size_t pool_cached_bytes();  // bytes sitting in the cache (reusable without a driver call)
void pool_release_all();   // return everything cached to the driver (freeGaugeQuda / endQuda)
double pool_driver_time(long *calls);  // seconds (and calls) spent in cudaMalloc / cudaFree on behalf of the pool so far

inline int div_up(long a, long b) { return (int)((a + b - 1) / b); }

}  // namespace qb
