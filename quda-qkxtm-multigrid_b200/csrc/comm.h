// Inter-GPU communication: one process per GPU, NCCL over NVLink/NVSwitch.
// Replaces the MPI/QMP layer of the reference (/root/reference/include/comm_quda.h:56-194,
// lib/comm_mpi.cpp, lib/comm_common.cpp) for the hot path: halo send/recv per partitioned
// (dim, dir) and small all-reduces of doubles.
#pragma once
#include "common.h"

namespace qb {

struct Lattice;

// bootstrap (called from the C API)
void comm_unique_id(void *out128);
void comm_bootstrap(int rank, int size, const void *unique_id128);
void comm_set_grid(const int *dims, int (*rank_from_coords)(const int *, void *), void *fdata);
void comm_finalize();

bool comm_self_exchange();          // true when partitioned dims are emulated on a single rank
int comm_neighbor_rank(int dim, int dir);  // dir 0: backward, 1: forward

void *comm_alloc_halo(size_t bytes);
void comm_free_halo(void *p);

// exchange all partitioned faces of one precision arena (send_arena -> neighbours' recv_arena)
void comm_exchange_halo(Lattice &lat, int prec_idx, cudaStream_t s);
// generic point-to-point pair on a stream (gauge ghost links, coarse halos)
void comm_sendrecv(const void *sendbuf, int to_rank, void *recvbuf, int from_rank, size_t bytes, cudaStream_t s);
void comm_sendrecv_group(int n, const void *const *sendbuf, const int *to_rank, void *const *recvbuf, const int *from_rank, const size_t *bytes, cudaStream_t s);

// ---- all-reduce fused into the reduction kernels (blas.cu) ----------------------------------------------------------------------
// Every rank exposes a small mailbox in its HBM to all peers (CUDA IPC over NVLink / NVSwitch, opened once at bootstrap).  The last CTA
// of a reduction kernel stores its rank's partial sums straight into every peer's mailbox, raises a sequence flag there, waits for the
// flags of all peers in its own mailbox and adds the contributions in rank order (deterministic, identical on all ranks): the global sum
// is complete when the kernel ends -- no NCCL launch, no host staging, one stream synchronisation where the host needs the scalar.
// (Reference: kernel -> mapped host memory spin -> MPI_Allreduce, lib/reduce_core.cuh:72-93, lib/face_buffer.cpp:407-428.)
constexpr int PEER_MAX_RANKS = 16;
constexpr int PEER_MAX_RED = 192;      // doubles per reduction (block dot products of 64 right-hand sides: 3 x 64)
struct PeerReduce {
  int rank = 0, size = 1;
  unsigned long long seq = 0;          // sequence number of this reduction (identical on all ranks: reductions are collective)
  double *box[PEER_MAX_RANKS];         // box[p]: mailbox of rank p as seen from this device, [2 slots][size][PEER_MAX_RED]
  unsigned long long *flag[PEER_MAX_RANKS];  // flag[p]: [2 slots][size] sequence flags of rank p's mailbox
};
// false: single rank, or the mailboxes could not be mapped (then reductions fall back to ncclAllReduce on the compute stream)
bool comm_peer_reduce_ready();
void comm_peer_reduce_enable(bool on);   // measurement switch (all ranks alike)
// arguments for the next collective reduction (advances the sequence number)
PeerReduce comm_peer_reduce_next();
// fallback: in-place sum of n doubles in device memory, enqueued on stream s
void comm_allreduce_sum_device(double *d_data, int n, cudaStream_t s);

// Maps one device allocation of EVERY rank into this process (CUDA IPC over NVLink / NVSwitch).  Collective: all ranks call it with their
// own allocation.  mapped[r] (r < size) = pointer valid on this device for rank r's allocation (own rank: `local` itself).  Returns false,
// on all ranks alike, if any rank could not export or open a handle (then nothing stays mapped).
bool comm_ipc_map(void *local, void **mapped);
void comm_ipc_unmap(void **mapped);

// Ghost-zone arena the neighbours store their faces into directly (coarse-level halos; the fine Dslash keeps the same scheme in
// Lattice, dslash_host.cu).  Two buffers selected by the parity of the exchange sequence number (a rank may start packing exchange s + 1
// while a neighbour still reads exchange s) followed by 8 arrival flags, slot = dim * 2 + dir.  create / destroy are collective.
struct PeerArena {
  char *local = nullptr;
  size_t bytes = 0;                    // one buffer
  void *mapped[PEER_MAX_RANKS] = {};
  bool peer = false;                   // neighbours' arenas mapped here (else: plain local receive buffer for NCCL send / recv)
  unsigned long long seq = 0;          // exchanges so far (identical on all ranks: exchanges are collective)
  void create(size_t buffer_bytes);
  void destroy();
  size_t buf() const { return peer ? (size_t)(seq & 1) * bytes : 0; }
  char *recv_base() const { return local + buf(); }
  char *send_base(int rank) const { return (char *)mapped[rank] + buf(); }
  unsigned long long *flag_mine(int slot) const { return (unsigned long long *)(local + 2 * bytes) + slot; }
  unsigned long long *flag_of(int rank, int slot) const { return (unsigned long long *)((char *)mapped[rank] + 2 * bytes) + slot; }
};
// arrival flags: signal = release-store seq into up to 8 remote flags (after the stores of the pack kernel before it in stream order);
// wait = spin until up to 8 local flags have reached seq
struct HaloFlags { unsigned long long *p[8]; unsigned long long seq; int n; };
void comm_halo_signal(const HaloFlags &f, cudaStream_t s);
void comm_halo_wait(const HaloFlags &f, cudaStream_t s);
bool comm_peer_halo_wanted();          // more than one rank, <= PEER_MAX_RANKS, not switched off (QB_PEER_HALO=0)

// sum / max all-reduce of n doubles held on the host (blocking; used by reductions in solvers)
void comm_allreduce_sum(double *data, int n);
void comm_allreduce_max(double *data, int n);
void comm_barrier();

}  // namespace qb
