/* Extensions of the QUDA C interface for device-resident operation on B200.
 *
 * The reference API (include/quda.h) moves every operand host -> device -> host on each call
 * (/root/reference/lib/interface_quda.cpp:1509-1557).  These entry points expose the same
 * operators on fields that stay resident in HBM, which is what the reference does internally
 * between `cudaColorSpinorField in(*in_h, cudaParam)` (:1513) and `*out_h = out` (:1554), and what
 * QKXTM does through the internal C++ classes (interface_quda.cpp:6289-6449).  They exist so that
 * benchmarks can time the kernels without PCIe traffic and so that multi-GPU runs can be
 * bootstrapped without MPI (NCCL unique id handed in by the launcher).
 *
 * Plain C ABI: opaque handles are void*, no C++/torch types.
 */
#ifndef QUDA_B200_EXT_H
#define QUDA_B200_EXT_H
#ifdef __cplusplus
extern "C" {
#endif

/* --- resident spinor fields ------------------------------------------------------------------
 * site_subset: QUDA_PARITY_SITE_SUBSET (1) or QUDA_FULL_SITE_SUBSET (2); precision: device precision.
 * The lattice is the one of the gauge field loaded with loadGaugeQuda. */
void *newSpinorQudaB200(QudaSiteSubset site_subset, QudaPrecision precision);
void freeSpinorQudaB200(void *field);
/* QudaInvertParam::make_resident_solution = 1 (reference include/quda.h:293-297, lib/interface_quda.cpp:2493-2508): invertQuda leaves the
 * solution on the device and does not write h_x.  This returns that field (owned by the library; replaced by the next such solve, released by
 * freeGaugeQuda / endQuda), NULL if there is none; use it with saveSpinorQudaB200 or the *Resident* operators below. */
void *residentSolutionQudaB200(void);
/* host (param->cpu_prec, dirac_order, gamma_basis) <-> resident field; mirrors the
 * cudaColorSpinorField <- cpuColorSpinorField assignment (lib/cuda_color_spinor_field.cu:513-552) */
void loadSpinorQudaB200(void *field, const void *h_in, QudaInvertParam *param);
void saveSpinorQudaB200(void *h_out, const void *field, QudaInvertParam *param);

/* same semantics as dslashQuda / MatQuda / MatDagMatQuda, operands resident */
void dslashResidentQudaB200(void *out, void *in, QudaInvertParam *param, QudaParity parity);
void matResidentQudaB200(void *out, void *in, QudaInvertParam *param);
void matDagMatResidentQudaB200(void *out, void *in, QudaInvertParam *param);

/* Repeat dslashResident `niter` times on the library's compute stream and return the mean device
 * time per application in milliseconds, measured with CUDA events on that stream (what
 * tests/dslash_test.cpp:455-616 does).  per_iter_ms (may be NULL) receives each iteration's time. */
double timeDslashQudaB200(void *out, void *in, QudaInvertParam *param, QudaParity parity, int niter,
                          float *per_iter_ms);

/* mean device time in ms of the halo part of one hop alone on a partitioned lattice (face pack kernel + exchange of every
 * partitioned face; no interior / boundary kernel); bytes_sent (may be NULL) receives the bytes this rank sends per hop */
double timeHaloQudaB200(void *out, void *in, QudaInvertParam *param, QudaParity parity, int niter, double *bytes_sent);

/* Global reductions: 1 when the all-reduce over the ranks runs INSIDE the reduction kernel (the last CTA stores the rank's
 * sums into every peer's HBM mailbox over NVLink, waits for the peers' flags and adds in rank order), 0 when it is an ncclAllReduce on the
 * compute stream (mailboxes not mappable, QB_PEER_REDUCE=0) or a single rank.  Reference: lib/reduce_core.cuh:72-93 + MPI_Allreduce. */
int commPeerReduceActiveQudaB200(void);
/* mean host-visible latency (microseconds) of one global norm2 of an fp32 field of n_reals reals: kernel + all-reduce + the stream
 * synchronisation that hands the sum to the host; use_peer = 0 forces the ncclAllReduce path for comparison (collective call) */
double timeReduceQudaB200(long n_reals, int niter, int use_peer);

/* launch geometry of the fine Dslash kernels (the reference autotunes this, lib/tune.cpp:480-655;
 * here a fixed default is used and this knob exists for tuning runs).  Call after loadGaugeQuda. */
/* mean ms of one batched hop over `nbatch` fp32 parity fields (multi-RHS fine Dslash: links fetched once for all members);
   max_dev_vs_single: largest relative L2 deviation of a member from the single-field kernel on the same input (expected: 0) */
double timeDslashBatchQudaB200(QudaInvertParam *param, QudaParity parity, int nbatch, int niter, double *max_dev_vs_single /* may be NULL */);
void setDslashBlockSizeQudaB200(int threads_per_block);
/* number of kernels this library has launched since initQuda (monotonic counter) */
long long kernelLaunchCountQudaB200(void);
/* the CUDA stream (cudaStream_t) compute kernels are launched on */
void *computeStreamQudaB200(void);
/* blocks until all work queued by the library has finished */
void syncQudaB200(void);

/* --- multi-GPU bootstrap (replaces MPI_Init/QMP of tests/test_util.cpp:44-67) -------------------
 * Call on every rank before initCommsGridQuda.  `unique_id` is the 128-byte ncclUniqueId produced
 * by ncclUniqueIdQudaB200 on rank 0 and broadcast by the launcher (torch.distributed, files, ...). */
void ncclUniqueIdQudaB200(void *unique_id_out_128B);
void commsBootstrapQudaB200(int rank, int size, const void *unique_id_128B);
/* info10 = {rank, size, coords[4], grid[4]} after initCommsGridQuda (rank <-> coordinate map is
 * lexicographic with t fastest, lib/interface_quda.cpp:261-274) */
void commRankInfoQudaB200(int *info10);
/* checkerboard index of every site of face `face_num` (0: x_dim = 0, 1: x_dim = X_dim - 1) of the given
 * parity in face-index order, as used by the pack kernel (cf. indexFromFaceIndex, lib/dslash_index.cuh:13-96) */
void faceIndexMapQudaB200(int dim, int face_num, int parity, int *h_cb_out);
/* single-process emulation of a partitioned lattice: halos are packed, "exchanged" with the rank
 * itself and consumed by the boundary kernels (the reference's --partition test trick,
 * tests/test_util.cpp:2047-2065, lib/comm_common.cpp:420-433).  mask bit d = dimension d. */
void commDimPartitionedSetQudaB200(int mask);

/* --- BLAS-1 / reduction test hook (the reference drives these through tests/blas_test.cu) -------------
 * name = a function of quda::blas (include/blas_quda.h:33-144): "axpy", "caxpy", "cDotProductNormA", ...;
 * x, y, z, w: host arrays of n complex numbers of `prec` bytes per real (updated in place), coef = {a_re, a_im,
 * b_re, b_im}.  Returns how many doubles were written to result. */
int blasQudaB200(const char *name, long n, int prec, const double *coef, void *x, void *y, void *z, void *w, double *result);

/* --- multigrid introspection (what MG::verify, lib/multigrid.cpp:372-486, checks inside the library) ----
 * `mg` is the handle returned by newMultigridQuda; level 0 is the fine grid.  Generic host field order:
 * [parity][checkerboard site][component k = spin * nColor + colour][re, im], float32. */
/* relative deviations {|R P eta - eta|, max_k |P R v_k - v_k|, |R M P eta - M_c eta|} of level `level` */
void mgVerifyQudaB200(void *mg, int level, double *dev3);
/* info8 = {coarse X[0..3], n_vec, fine components per site, sites per aggregate, coarse components per site} */
void mgLevelInfoQudaB200(void *mg, int level, int *info8);
void mgProlongQudaB200(void *mg, int level, float *h_fine_out, const float *h_coarse_in);
void mgRestrictQudaB200(void *mg, int level, float *h_coarse_out, const float *h_fine_in);
/* operator of a level: pc = 0 the full operator (level 0: fine M, level >= 1: coarse M_c), pc = 1 the smoother's operator */
void mgMatQudaB200(void *mg, int level, int pc, float *h_out, const float *h_in);
void mgNullVectorQudaB200(void *mg, int level, int k, float *h_out);
/* wall-clock profile of the multigrid cycle: enable the stream-synchronising section timers, then read (and reset) a level's accumulated
 * seconds t6 = {pre-smooth or coarsest solve, residual, restrict, coarse solve (all levels below), prolong, post-smooth} and its cycle count */
void mgProfileEnableQudaB200(int on);
void mgProfileGetQudaB200(void *mg, int level, double *t6, long *ncycle);
/* link matrices of the coarse operator on level `level` >= 1, row-major h_out[site][d][row][col][re,im] (site = parity * Vh + x_cb;
 * d = 0..7: hop to x + e_d with e_d = +mu (d = 2 mu) / -mu (d = 2 mu + 1), d = 8: site-diagonal block).  which = 0: the links L (the -kappa
 * of the reference's X - kappa sum Y folded in: Y_{mu+4}(x) = -L_{2mu}(x)/kappa, Y_mu(x) = -L_{2mu+1}(x+mu)^dag/kappa, X = L_8,
 * lib/dslash_coarse.cu:49-203), 1: Xinv ([site][row][col]), 2: Yhat = Xinv L.  Test hook for the element-wise comparison with the reference's calculateY. */
void mgCoarseLinksQudaB200(void *mg, int level, int which, float *h_out);
/* mean device time in ms (CUDA events on the compute stream) of `niter` applications on level `level` of
 * what = 0: full operator, 1: smoother operator, 2: prolongator, 3: restrictor */
double mgTimeQudaB200(void *mg, int level, int what, int niter);
/* Multi-right-hand-side coarse operator on the tensor cores (tcgen05, tf32) of coarse level `level` >= 1, applied to `nrhs`
 * vectors at once (h_in / h_out = [rhs][generic host field]); the link matrices of a site are read once for all of them.
 * what = 0: full operator M_c; 1: hopping term into the even sites; 2: Xinv on the odd sites.
 * mode = 1: one tf32 pass (11 significant bits per operand); 3: split tf32, fp32-accurate (hi/lo operands, all four products).
 * Reference: one right-hand side per call, lib/dslash_coarse.cu:293-333; the multi-RHS coarse grid is BASELINE config 5. */
void mgMatMrhsQudaB200(void *mg, int level, int what, int nrhs, int mode, float *h_out, const float *h_in);
/* largest nrhs one call accepts on that level (bounded by the 128-row MMA tile and by shared memory) */
int mgMrhsMaxRhsQudaB200(void *mg, int level, int mode);
/* mean device time in ms of `niter` applications of the operator above on resident block fields */
double mgTimeMrhsQudaB200(void *mg, int level, int what, int nrhs, int mode, int niter);
/* one multigrid cycle of level `level`:  x = MG(b) */
void mgCycleQudaB200(void *mg, int level, float *h_x, const float *h_b);

#ifdef __cplusplus
}
#endif
#endif
