// Device side of the all-reduce over the ranks that runs INSIDE a reduction kernel (comm.h: PeerReduce).
#pragma once
#include "comm.h"

namespace qb {

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Called by ALL threads of ONE CTA per rank (the last CTA of a reduction), with block-uniform arguments.  vals: n <= PEER_MAX_RED doubles in
// shared memory holding this rank's sums, written before the call (the function synchronises first).  Every rank stores its sums into
// every rank's mailbox over NVLink, raises its sequence flag there, waits for all flags in its own mailbox and adds the contributions in
// rank order: on return (after the trailing barrier) vals holds the global sums, bit-identical on all ranks.
__device__ inline void peer_allreduce_cta(const PeerReduce &pr, double *vals, int n) {
  __syncthreads();
  const int slot = (int)(pr.seq & 1);
  for (int i = threadIdx.x; i < pr.size * n; i += blockDim.x) {
    const int p = i / n, k = i - p * n;
    pr.box[p][((size_t)slot * pr.size + pr.rank) * PEER_MAX_RED + k] = vals[k];
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < pr.size) {
    st_release_sys(pr.flag[threadIdx.x] + slot * pr.size + pr.rank, pr.seq);
    const unsigned long long *mine = pr.flag[pr.rank] + slot * pr.size + threadIdx.x;
    while (ld_acquire_sys(mine) < pr.seq) {}
  }
  __syncthreads();
  const volatile double *box = pr.box[pr.rank] + (size_t)slot * pr.size * PEER_MAX_RED;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    double v = 0.0;
    for (int p = 0; p < pr.size; p++) v += box[(size_t)p * PEER_MAX_RED + k];
    vals[k] = v;
  }
  __syncthreads();
}

}  // namespace qb
