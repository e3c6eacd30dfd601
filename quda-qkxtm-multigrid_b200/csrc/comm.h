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

// sum / max all-reduce of n doubles held on the host (blocking; used by reductions in solvers)
void comm_allreduce_sum(double *data, int n);
void comm_allreduce_max(double *data, int n);
void comm_barrier();

}  // namespace qb
