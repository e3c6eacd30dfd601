// Multigrid coarse-level operator (replaces /root/reference/lib/dirac_coarse.cpp, lib/dslash_coarse.cu
// and the coarse-link construction of lib/coarse_op.cu, lib/coarsecoarse_op.cu, lib/coarse_op.cuh).
//
// Stencil form used at every level:   (M psi)(x) = L_8(x) psi(x) + sum_{d=0..7} L_d(x) psi(x + e_d)
// with e_d = +mu for d = 2 mu and -mu for d = 2 mu + 1.  On a coarse level L_d(x) are dense N x N complex
// matrices (N = 2 * n_vec: chirality x null-vector index) kept in HBM as
//     Y[site][d][col][row pair] float4 = (Y[2rp][col], Y[2rp+1][col])
// i.e. one contiguous 9*N*N*8-byte record per site, column-major in row pairs, so that the matvec of the
// coarse Dslash streams each matrix once with 128-bit loads and needs no cross-thread reduction inside a
// direction.  The links are the Galerkin product  L^c = R L P  of the next-finer level, so
//     M_c = R M P   holds exactly (up to rounding) -- checked by tests (the identity of MG::verify,
// lib/multigrid.cpp:372-486).  The -kappa of the reference's  X - kappa * sum Y  is folded into the links.
#pragma once
#include "comm.h"
#include <complex>
#include <memory>
#include "dirac.h"
#include "transfer.h"

namespace qb {

struct CoarseOperator {
  LevelGeom geom;
  int nvec = 0;   // coarse colours
  int N = 0;      // 2 * nvec
  float *Y = nullptr;     // [V][9][N][N/2] float4
  float *Xinv = nullptr;  // [V][N][N/2] float4, inverse of the site-diagonal block L_8 (for even-odd preconditioning)
  void *Y16 = nullptr, *Xinv16 = nullptr;   // optional fp16 copies used by the single-RHS kernel (enable_half_links)
  void enable_half_links();
  // Preconditioned links Yhat_d(x) = Xinv(x) L_d(x), d < 8, same record layout as Y with the identity in slot 8
  // (createYpreconditioned, lib/coarse_op.cuh:1217-1283; here every link sits on its OUTPUT site, so the backward links need no
  // Xinv of a neighbour and no ghost exchange).  Two uses: the even-odd operator 1 - Yhat_pq Yhat_qp in two launches instead of four,
  // and the preconditioned coarsening of this level: the Galerkin product of Xinv M = 1 + sum_d Yhat_d (DiracCoarsePC::createCoarseOp,
  // lib/dirac_coarse.cpp:377-380).
  float *Yhat = nullptr;
  void *Yhat16 = nullptr;
  void compute_yhat();
  // halo buffers of coarse spinors for partitioned dimensions: [d][dir] -> [parity][plane N/2][faceVh] float4
  // send[d][0] = my slice x_d = 0 (goes backward), send[d][1] = my slice x_d = X_d - 1 (goes forward);
  // recv[d][0] = from the backward neighbour, recv[d][1] = from the forward neighbour (alias of send in self-exchange mode)
  float *send[4][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
  float *recv[4][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
  // with real ranks the receive blocks live in one arena that the neighbours' pack kernels store into directly over NVLink
  // (PeerArena, comm.h); recv_off[d][k] = byte offset of block (d, k) inside one buffer of it.  NCCL send / recv when it is not mapped.
  mutable PeerArena ghost_arena;
  size_t recv_off[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
  size_t link_bytes() const { return (size_t)geom.V() * 9 * N * N * 8; }
  void allocate(const LevelGeom &g, int nvec_);
  // pack the boundary slices of the parities in `parity_mask` of a coarse field and exchange them with the neighbours
  void exchange_ghost(const float *field, const long *poff, int parity_mask) const;
  void compute_xinv();    // batched in-kernel Gauss-Jordan with partial pivoting (the reference calls MAGMA, coarse_op.cuh:1466-1474)
  // ---- tensor-core (multi-RHS) view of the same links, coarse_mrhs.cu ----
  float *Ymma = nullptr;      // [V][9][N/2][N] float4: K-major UMMA operand image of every link matrix
  float *Xinv_mma = nullptr;  // [V][N/2][N] float4
  int *nbr = nullptr;         // [V][8] full-site index of x + e_d
  bool mrhs_ready = false;
  void prepare_mrhs();        // (re)build the three arrays above from Y / Xinv
  // Partitioned lattices: ghost zone of BLOCK fields.  The neighbour table points boundary hops at "ghost sites" numbered from 2 Vh:
  // ghost site = mrhs_goff[d][dir] + parity * faceVh[d] + face index, dir 0 = slice X_d - 1 of the backward neighbour, dir 1 = slice 0
  // of the forward neighbour; the send / receive arenas hold one contiguous [kc][r] block of R vectors per ghost site, exactly what the
  // kernel's producer bulk-copies for a local neighbour (reference: ghost zones of composite fields, lib/color_spinor_pack.cu,
  // lib/dslash_coarse.cu:707)
  long mrhs_goff[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
  long mrhs_ghost_sites = 0;
  mutable float *mrhs_send = nullptr, *mrhs_recv = nullptr;   // mrhs_recv: current buffer of mrhs_arena
  mutable size_t mrhs_arena_bytes = 0;
  mutable PeerArena mrhs_arena;
  void exchange_block_ghost(const float *field, const long *poff, int parity_mask, int R) const;
  ~CoarseOperator();
};

// Builders of the Galerkin coarse links (coarse_op.cu)
// clover_site: site-major fp32 packed clover term [V][72] (nullptr: the site-local term is 1 + i a gamma5)
// VL: left vectors of the product in the layout of T.V (nullptr: V itself), see DiracTM::create_coarse_op
void build_coarse_from_fine(CoarseOperator &out, const Transfer &T, const GaugeField &gauge, const Geom &fine_geom, double kappa, double twist_a,
                            const float *clover_site = nullptr, const float *VL = nullptr);
// preconditioned = true: coarsen Xinv M (the links Yhat of `fine`) instead of M
void build_coarse_from_coarse(CoarseOperator &out, const Transfer &T, const CoarseOperator &fine, bool preconditioned = false);
// rows of chirality 0 / 1 of every link matrix (slots 0..8) times c0 / c1: turns R M P into R A^-1 M P when A is a constant per chirality
void scale_coarse_rows(CoarseOperator &op, std::complex<double> c0, std::complex<double> c1);
// tensor-core variant of build_coarse_from_fine (coarse_op_mma.cu); out.Y allocated and zeroed
bool galerkin_mma_supported(const Transfer &T);
void build_coarse_from_fine_mma(CoarseOperator &out, const Transfer &T, const GaugeField &gauge, const Geom &fine_geom, double kappa, double twist_a,
                                const float *clover_site);

// Coarse Dirac operator.  Full operator on [even | odd] fields, or (pc = true) the symmetric even-odd
// Schur complement  1 - Xinv_pp Y_pq Xinv_qq Y_qp  on single-parity fields (DiracCoarsePC, dirac_coarse.cpp:226-372).
class DiracCoarse : public Dirac {
 public:
  std::shared_ptr<CoarseOperator> op;
  bool pc;
  int matpc_type;
  mutable std::unique_ptr<SpinorField> tmp1, tmp2;
  DiracCoarse(std::shared_ptr<CoarseOperator> op_, bool pc_, int matpc_) : op(op_), pc(pc_), matpc_type(matpc_) {}
  DiracType type() const override { return pc ? DIRAC_COARSE_PC : DIRAC_COARSE; }
  bool is_pc() const override { return pc; }
  int matpc() const override { return matpc_type; }
  Prec precision() const override { return PREC_SINGLE; }
  SpinorField *new_field(Prec) const override { return new SpinorField(op->geom.Vh, pc ? 1 : 2, PREC_SINGLE, 2, op->nvec); }
  SpinorField *new_parity_field(Prec) const override { return new SpinorField(op->geom.Vh, 1, PREC_SINGLE, 2, op->nvec); }

  // out(parity) = sum_d Y_d in(other parity)   [hopping part only]
  void Dslash(SpinorField &out, const SpinorField &in, int parity) const override;
  void DslashXpay(SpinorField &out, const SpinorField &in, int parity, const SpinorField &x, double k) const override;
  void Clover(SpinorField &out, const SpinorField &in, int parity) const;      // X in
  void CloverInv(SpinorField &out, const SpinorField &in, int parity) const;   // X^-1 in
  void M(SpinorField &out, const SpinorField &in) const override;
  void prepare(SpinorField &src, SpinorField &sol, SpinorField &x, SpinorField &b, SolutionType sol_type) const override;
  void reconstruct(SpinorField &x, const SpinorField &b, SolutionType sol_type) const override;
  void create_coarse_op(CoarseOperator &coarse, const Transfer &T, bool preconditioned = false) const override;
  void DiagInv(SpinorField &out, const SpinorField &in) const override;
  void Diag(SpinorField &out, const SpinorField &in, int parity) const override { Clover(out, in, parity); }
  int p_parity() const { return (matpc_type == MATPC_EVEN_EVEN || matpc_type == MATPC_EVEN_EVEN_ASYM) ? 0 : 1; }
};

// low-level launcher: out(sites of `parity`, or all sites if parity < 0) =
//    [use_x]  (Xinv or L_8) in_diag(x)   +   [use_y]  sum_d L_d(x) in_hop(x + e_d)
// optionally followed by out = a * out + b * xpay
struct CoarseApplyArgs {
  const CoarseOperator *op;
  float *out;                // full-field base pointers ([parity][plane][cb]); parity fields are addressed through poff
  const float *in_hop;       // field supplying the neighbours
  const float *in_diag;      // field supplying the site-diagonal term
  const float *xpay;
  long out_poff[2], hop_poff[2], diag_poff[2], xpay_poff[2];  // float4 offset of each parity block inside the buffers
  int parity;                // -1: all sites
  bool use_y, use_x, use_xinv;
  float a, b;
  bool force_fp32 = false;   // ignore the fp16 link copies (residual operator of the K-cycle, verify)
  bool use_yhat = false;     // hopping term from Yhat = Xinv Y instead of Y
};
void coarse_apply(const CoarseApplyArgs &args);

// R coarse vectors side by side: [parity][cb][kc = N/2][r] float4 = components 2kc, 2kc+1 of vector r.  All vectors of a
// site are one contiguous block, which is what the multi-RHS kernel stages as its MMA operand.
struct CoarseBlockField {
  long Vh;
  int nparity, N, R;
  float *v = nullptr;
  void **ptrs = nullptr;  // device scratch: member field pointers for pack / unpack
  bool owner = true;
  CoarseBlockField(long Vh, int nparity, int N, int R);
  CoarseBlockField(long Vh, int nparity, int N, int R, float *storage);   // non-owning view of an existing block buffer
  ~CoarseBlockField();
  CoarseBlockField(const CoarseBlockField &) = delete;
  size_t parity_float4() const { return (size_t)Vh * (N / 2) * R; }
  size_t bytes() const { return parity_float4() * 16 * nparity; }
  void pack(SpinorField *const *fields);          // R single fields -> block
  void unpack(SpinorField *const *fields) const;  // block -> R single fields
  // the same with raw device pointers to [plane][cb] float4 blocks of matching parity count (e.g. one parity of full fields)
  void pack_ptrs(const void *const *ptrs);
  void unpack_ptrs(const void *const *ptrs) const;
};

// Same contract as CoarseApplyArgs on block fields; mode = 1: one tf32 pass (11-bit operands), 3: split tf32 (fp32-accurate)
struct CoarseMrhsArgs {
  const CoarseOperator *op;
  float *out;
  const float *in_hop, *in_diag, *xpay;
  long out_poff[2], hop_poff[2], diag_poff[2], xpay_poff[2];  // float4 offsets of the parity blocks
  int parity;
  bool use_y, use_x, use_xinv;
  float a, b;
  int R, mode;
};
void coarse_apply_mrhs(const CoarseMrhsArgs &args);
int coarse_mrhs_max_rhs(int N, int mode);

}  // namespace qb
