// NCCL-backed communication layer (dlopen'ed: no link-time dependency, so the single-GPU path
// works on a box without NCCL, and inside a PyTorch process the already-loaded libnccl is reused).
#include <dlfcn.h>
#include <vector>
#include "comm.h"
#include "dslash_api.h"

namespace qb {

// minimal NCCL ABI (stable across 2.x)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclMax = 2 };

struct NcclApi {
  void *handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi nccl;
static ncclComm_t comm = nullptr;
static double *d_red = nullptr;   // device scratch for all-reduces
static double *h_red = nullptr;   // pinned host mirror
static const int RED_MAX = 256;

#define QB_NCCL(call)                                                                      \
  do {                                                                                     \
    int e_ = (call);                                                                       \
    if (e_ != ncclSuccess) QB_ERROR("%s failed: %s", #call, nccl.GetErrorString ? nccl.GetErrorString(e_) : "?"); \
  } while (0)

static void load_nccl() {
  if (nccl.handle) return;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (nccl.handle) break;
  }
  if (!nccl.handle) QB_ERROR("multi-GPU run requested but libnccl.so.2 could not be loaded: %s", dlerror());
#define SYM(field, name)                                              \
  *(void **)(&nccl.field) = dlsym(nccl.handle, name);                 \
  if (!nccl.field) QB_ERROR("symbol %s missing from libnccl", name)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(AllReduce, "ncclAllReduce");
  SYM(AllGather, "ncclAllGather");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
}

static void peer_reduce_setup();

void comm_unique_id(void *out128) {
  load_nccl();
  ncclUniqueId id;
  QB_NCCL(nccl.GetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
}

void comm_bootstrap(int rank, int size, const void *unique_id128) {
  Runtime &r = rt();
  if (size < 1 || rank < 0 || rank >= size) QB_ERROR("invalid rank %d / size %d", rank, size);
  r.rank = rank;
  r.size = size;
  if (size == 1) return;
  if (!unique_id128) return;  // bookkeeping only (CPU tests of the rank grid): any exchange will fail loudly
  if (!r.device_ready) QB_ERROR("call initQudaDevice before commsBootstrapQudaB200 (the NCCL communicator binds to the current device)");
  load_nccl();
  ncclUniqueId id;
  memcpy(&id, unique_id128, sizeof(id));
  QB_NCCL(nccl.CommInitRank(&comm, size, id, rank));
  peer_reduce_setup();
}

// ---- peer mailboxes for the fused all-reduce (comm.h) ----------------------------------------------------------------------------
static PeerReduce peer;
static bool peer_ready = false;
static void *peer_local = nullptr;           // my mailbox: boxes then flags
static void *peer_mapped[PEER_MAX_RANKS];    // IPC mappings of the peers' mailboxes
static size_t peer_box_bytes(int size) { return sizeof(double) * 2 * size * PEER_MAX_RED; }

bool comm_ipc_map(void *local, void **mapped) {
  Runtime &r = rt();
  for (int p = 0; p < r.size; p++) mapped[p] = nullptr;
  if (r.size == 1) { mapped[0] = local; return true; }
  cudaIpcMemHandle_t mine;
  bool ok = cudaIpcGetMemHandle(&mine, local) == cudaSuccess;
  // exchange the handles (and whether every rank got one) through NCCL
  struct Msg { cudaIpcMemHandle_t h; int ok; int pad[3]; };
  Msg m{}; m.h = mine; m.ok = ok ? 1 : 0;
  Msg *d_all = nullptr;
  QB_CUDA(cudaMalloc((void **)&d_all, sizeof(Msg) * r.size));
  QB_CUDA(cudaMemcpy(d_all + r.rank, &m, sizeof(Msg), cudaMemcpyHostToDevice));
  QB_NCCL(nccl.AllGather(d_all + r.rank, d_all, sizeof(Msg), ncclInt8, comm, r.compute));
  QB_CUDA(cudaStreamSynchronize(r.compute));
  std::vector<Msg> all(r.size);
  QB_CUDA(cudaMemcpy(all.data(), d_all, sizeof(Msg) * r.size, cudaMemcpyDeviceToHost));
  QB_CUDA(cudaFree(d_all));
  for (int p = 0; p < r.size; p++) ok = ok && all[p].ok;
  int mapped_ok = ok ? 1 : 0;
  if (ok) {
    for (int p = 0; p < r.size; p++) {
      if (p == r.rank) { mapped[p] = local; continue; }
      if (cudaIpcOpenMemHandle(&mapped[p], all[p].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { mapped_ok = 0; mapped[p] = nullptr; }
    }
  }
  cudaGetLastError();
  // all ranks must take the same path
  double v = mapped_ok ? 0.0 : 1.0;
  comm_allreduce_sum(&v, 1);
  if (v != 0.0) {
    for (int p = 0; p < r.size; p++)
      if (p != r.rank && mapped[p]) { cudaIpcCloseMemHandle(mapped[p]); mapped[p] = nullptr; }
    return false;
  }
  return true;
}

void comm_ipc_unmap(void **mapped) {
  Runtime &r = rt();
  for (int p = 0; p < r.size; p++) {
    if (p != r.rank && mapped[p]) cudaIpcCloseMemHandle(mapped[p]);
    mapped[p] = nullptr;
  }
}

// ---- direct halo delivery (comm.h) -------------------------------------------------------------------------------------------------
bool comm_peer_halo_wanted() {
  const char *env = getenv("QB_PEER_HALO");
  return rt().size > 1 && rt().size <= PEER_MAX_RANKS && !(env && atoi(env) == 0);
}

__global__ void halo_signal_kernel(const HaloFlags f) {
  if ((int)threadIdx.x < f.n) {
    __threadfence_system();   // the faces were stored by the pack kernel before this one in stream order
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f.p[threadIdx.x]), "l"(f.seq) : "memory");
  }
}
__global__ void halo_wait_kernel(const HaloFlags f) {
  if ((int)threadIdx.x < f.n) {
    unsigned long long v;
    do { asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f.p[threadIdx.x]) : "memory"); } while (v < f.seq);
  }
}
void comm_halo_signal(const HaloFlags &f, cudaStream_t s) {
  halo_signal_kernel<<<1, 32, 0, s>>>(f);
  QB_CHECK_LAUNCH();
}
void comm_halo_wait(const HaloFlags &f, cudaStream_t s) {
  halo_wait_kernel<<<1, 32, 0, s>>>(f);
  QB_CHECK_LAUNCH();
}

void PeerArena::create(size_t buffer_bytes) {
  destroy();
  bytes = (buffer_bytes + 255) & ~(size_t)255;
  const size_t total = 2 * bytes + 256;
  local = (char *)comm_alloc_halo(total);
  QB_CUDA(cudaMemset(local, 0, total));
  seq = 0;
  peer = comm_peer_halo_wanted() && comm_ipc_map(local, mapped);
}
void PeerArena::destroy() {
  if (!local) return;
  if (peer) {
    // nobody may still be storing into (or have mapped) an arena that goes away
    cudaDeviceSynchronize();
    comm_barrier();
    comm_ipc_unmap(mapped);
    comm_barrier();
  }
  cudaFree(local);
  local = nullptr; bytes = 0; peer = false; seq = 0;
}

static void peer_reduce_setup() {
  Runtime &r = rt();
  peer_ready = false;
  const char *env = getenv("QB_PEER_REDUCE");
  if (env && atoi(env) == 0) return;
  if (r.size > PEER_MAX_RANKS) return;
  const size_t bytes = peer_box_bytes(r.size) + sizeof(unsigned long long) * 2 * r.size;
  peer_local = comm_alloc_halo(bytes);
  QB_CUDA(cudaMemset(peer_local, 0, bytes));
  if (!comm_ipc_map(peer_local, peer_mapped)) {
    log_msg(1, "peer mailboxes for the fused all-reduce could not be mapped on every rank: reductions use ncclAllReduce\n");
    return;
  }
  peer.rank = r.rank; peer.size = r.size; peer.seq = 0;
  for (int p = 0; p < r.size; p++) {
    peer.box[p] = (double *)peer_mapped[p];
    peer.flag[p] = (unsigned long long *)((char *)peer_mapped[p] + peer_box_bytes(r.size));
  }
  peer_ready = true;
  log_msg(2, "fused all-reduce: mailboxes of %d ranks mapped over CUDA IPC\n", r.size);
}

static bool peer_enabled = true;
bool comm_peer_reduce_ready() { return peer_ready && peer_enabled; }
void comm_peer_reduce_enable(bool on) { peer_enabled = on; }
PeerReduce comm_peer_reduce_next() {
  peer.seq++;
  return peer;
}

void comm_allreduce_sum_device(double *d_data, int n, cudaStream_t s) {
  if (rt().size == 1) return;
  QB_NCCL(nccl.AllReduce(d_data, d_data, n, ncclFloat64, ncclSum, comm, s));
}

static int default_rank_from_coords(const int *c, void *fdata) {
  const int *dims = (const int *)fdata;
  int rank = c[0];
  for (int i = 1; i < 4; i++) rank = dims[i] * rank + c[i];  // t fastest (interface_quda.cpp:261-274)
  return rank;
}

static int (*rank_map)(const int *, void *) = nullptr;
static void *rank_map_data = nullptr;
static int grid_dims[4] = {1, 1, 1, 1};

void comm_set_grid(const int *dims, int (*func)(const int *, void *), void *fdata) {
  Runtime &r = rt();
  int n = 1;
  for (int d = 0; d < 4; d++) {
    if (dims[d] < 1) QB_ERROR("invalid comm grid dimension %d", dims[d]);
    n *= dims[d];
    grid_dims[d] = dims[d];
    r.grid[d] = dims[d];
  }
  if (n != r.size) QB_ERROR("communication grid %dx%dx%dx%d does not match the number of ranks %d", dims[0], dims[1], dims[2], dims[3], r.size);
  rank_map = func ? func : default_rank_from_coords;
  rank_map_data = func ? fdata : (void *)grid_dims;
  // find my coordinates by scanning the map
  int c[4];
  bool found = false;
  for (c[0] = 0; c[0] < dims[0]; c[0]++)
    for (c[1] = 0; c[1] < dims[1]; c[1]++)
      for (c[2] = 0; c[2] < dims[2]; c[2]++)
        for (c[3] = 0; c[3] < dims[3]; c[3]++)
          if (rank_map(c, rank_map_data) == r.rank) {
            for (int d = 0; d < 4; d++) r.coord[d] = c[d];
            found = true;
          }
  if (!found) QB_ERROR("rank %d not found in the rank map", r.rank);
  // the grid may be set again (another decomposition of the same ranks, before the next loadGaugeQuda): with real ranks the
  // partitioned dimensions are exactly those of the grid
  if (r.size > 1) r.part_mask = 0;
  for (int d = 0; d < 4; d++)
    if (dims[d] > 1) r.part_mask |= 1 << d;
  r.grid_set = true;
}

bool comm_self_exchange() { return rt().size == 1; }

int comm_neighbor_rank(int dim, int dir) {
  Runtime &r = rt();
  int c[4] = {r.coord[0], r.coord[1], r.coord[2], r.coord[3]};
  c[dim] = (c[dim] + (dir ? 1 : r.grid[dim] - 1)) % r.grid[dim];
  return rank_map ? rank_map(c, rank_map_data) : 0;
}

// Buffers that may be exported to the peers with CUDA IPC: whole multiples of 2 MiB, so that an exported handle covers this buffer alone
// (the driver packs smaller allocations into shared 2 MiB blocks, and an IPC handle always exports the whole block)
static size_t ipc_round(size_t bytes) { const size_t g = (size_t)2 << 20; return ((bytes ? bytes : 1) + g - 1) / g * g; }
void *comm_alloc_halo(size_t bytes) {
  void *p = nullptr;
  QB_CUDA(cudaMalloc(&p, ipc_round(bytes)));
  return p;
}
void comm_free_halo(void *p) {
  if (p) cudaFree(p);
}

void comm_sendrecv(const void *sendbuf, int to_rank, void *recvbuf, int from_rank, size_t bytes, cudaStream_t s) {
  Runtime &r = rt();
  if (r.size == 1 || (to_rank == r.rank && from_rank == r.rank)) {
    QB_CUDA(cudaMemcpyAsync(recvbuf, sendbuf, bytes, cudaMemcpyDeviceToDevice, s));
    return;
  }
  QB_NCCL(nccl.GroupStart());
  QB_NCCL(nccl.Send(sendbuf, bytes, ncclInt8, to_rank, comm, s));
  QB_NCCL(nccl.Recv(recvbuf, bytes, ncclInt8, from_rank, comm, s));
  QB_NCCL(nccl.GroupEnd());
}

// n send / receive pairs in ONE NCCL group (one launch on the stream instead of n)
void comm_sendrecv_group(int n, const void *const *sendbuf, const int *to_rank, void *const *recvbuf, const int *from_rank, const size_t *bytes, cudaStream_t s) {
  Runtime &r = rt();
  if (r.size == 1) {
    for (int i = 0; i < n; i++) QB_CUDA(cudaMemcpyAsync(recvbuf[i], sendbuf[i], bytes[i], cudaMemcpyDeviceToDevice, s));
    return;
  }
  if (!comm) QB_ERROR("exchange requested but the NCCL communicator was not created");
  QB_NCCL(nccl.GroupStart());
  for (int i = 0; i < n; i++) {
    QB_NCCL(nccl.Send(sendbuf[i], bytes[i], ncclInt8, to_rank[i], comm, s));
    QB_NCCL(nccl.Recv(recvbuf[i], bytes[i], ncclInt8, from_rank[i], comm, s));
  }
  QB_NCCL(nccl.GroupEnd());
}

void comm_exchange_halo(Lattice &lat, int pi, cudaStream_t s) {
  Runtime &r = rt();
  if (r.size == 1) return;
  if (!comm) QB_ERROR("halo exchange requested but the NCCL communicator was not created (commsBootstrapQudaB200 without a unique id)");
  const Geom &g = lat.geom;
  char *send = (char *)lat.send_arena[pi], *recv = (char *)lat.recv_arena[pi];
  QB_NCCL(nccl.GroupStart());
  for (int d = 0; d < 4; d++) {
    if (!g.part[d]) continue;
    const int back = comm_neighbor_rank(d, 0), fwd = comm_neighbor_rank(d, 1);
    // block size: from this (d,dir) offset to the next block start
    const size_t o0 = lat.face_off[pi][d][0], o1 = lat.face_off[pi][d][1];
    const size_t len = o1 - o0;  // both directions have identical extents
    // my back face -> backward neighbour's "from forward" slot; my forward face -> forward neighbour's "from back" slot
    QB_NCCL(nccl.Send(send + o0, len, ncclInt8, back, comm, s));
    QB_NCCL(nccl.Recv(recv + o1, len, ncclInt8, fwd, comm, s));
    QB_NCCL(nccl.Send(send + o1, len, ncclInt8, fwd, comm, s));
    QB_NCCL(nccl.Recv(recv + o0, len, ncclInt8, back, comm, s));
  }
  QB_NCCL(nccl.GroupEnd());
}

static void ensure_red() {
  if (!d_red) {
    QB_CUDA(cudaMalloc((void **)&d_red, sizeof(double) * RED_MAX));
    QB_CUDA(cudaMallocHost((void **)&h_red, sizeof(double) * RED_MAX));
  }
}

static void allreduce(double *data, int n, int op) {
  Runtime &r = rt();
  if (r.size == 1) return;
  if (n > RED_MAX) QB_ERROR("all-reduce of %d doubles exceeds the scratch size", n);
  ensure_red();
  memcpy(h_red, data, sizeof(double) * n);
  QB_CUDA(cudaMemcpyAsync(d_red, h_red, sizeof(double) * n, cudaMemcpyHostToDevice, r.compute));
  QB_NCCL(nccl.AllReduce(d_red, d_red, n, ncclFloat64, op, comm, r.compute));
  QB_CUDA(cudaMemcpyAsync(h_red, d_red, sizeof(double) * n, cudaMemcpyDeviceToHost, r.compute));
  QB_CUDA(cudaStreamSynchronize(r.compute));
  memcpy(data, h_red, sizeof(double) * n);
}

void comm_allreduce_sum(double *data, int n) { allreduce(data, n, ncclSum); }
void comm_allreduce_max(double *data, int n) { allreduce(data, n, ncclMax); }

void comm_barrier() {
  double x = 0;
  allreduce(&x, 1, ncclSum);
}

void comm_finalize() {
  if (peer_local) {
    cudaDeviceSynchronize();
    if (rt().size <= PEER_MAX_RANKS) comm_ipc_unmap(peer_mapped);
    cudaFree(peer_local);
    peer_local = nullptr; peer_ready = false;
  }
  if (comm) nccl.CommDestroy(comm);
  comm = nullptr;
  if (d_red) cudaFree(d_red);
  if (h_red) cudaFreeHost(h_red);
  d_red = h_red = nullptr;
  Runtime &r = rt();
  r.rank = 0; r.size = 1; r.part_mask = 0; r.grid_set = false;
  for (int d = 0; d < 4; d++) { r.grid[d] = 1; r.coord[d] = 0; }
}

}  // namespace qb
