/* Drop-in C interface of the B200 twisted-mass / multigrid engine.
 *
 * This is the subset of the reference's public C API (/root/reference/include/quda.h) that the
 * hot path of SURVEY.md section 8 needs, with *byte-identical* parameter-struct layouts so that a
 * caller built against the reference header (tests/dslash_test.cpp, tests/multigrid_invert_test.cpp,
 * qkxtm/CalcMG_*.cpp) links against libquda_b200.so unchanged.  tests/test_abi_layout.py compiles
 * this header and the reference's side by side and compares sizeof/offsetof of every field.
 *
 * All functions are extern "C", take plain pointers, and follow the reference's error model:
 * no return codes, a formatted message and process exit on error (include/util_quda.h:50-60).
 * Entry points that are outside the hot path are exported too but abort with a clear message.
 */
#ifndef QUDA_B200_API_H
#define QUDA_B200_API_H

#include <stdio.h>
#include "quda_b200_enums.h"

#define QUDA_VERSION_MAJOR 0
#define QUDA_VERSION_MINOR 9
#define QUDA_VERSION_SUBMINOR 0
#define QUDA_VERSION ((QUDA_VERSION_MAJOR << 16) | (QUDA_VERSION_MINOR << 8) | QUDA_VERSION_SUBMINOR)
#define QUDA_MAX_DIM 6
#define QUDA_MAX_GEOMETRY 8
#define QUDA_MAX_MULTI_SHIFT 32
#define QUDA_MAX_DWF_LS 128
#define QUDA_MAX_MG_LEVEL 4
#define QUDA_MAX_MULTI_REDUCE 16

#ifdef __cplusplus
extern "C" {
#endif

/* replaces /root/reference/include/quda.h:25-80 */
typedef struct QudaGaugeParam_s {
  QudaFieldLocation location;
  int X[4];
  double anisotropy, tadpole_coeff, scale;
  QudaLinkType type;
  QudaGaugeFieldOrder gauge_order;
  QudaTboundary t_boundary;
  QudaPrecision cpu_prec;
  QudaPrecision cuda_prec;
  QudaReconstructType reconstruct;
  QudaPrecision cuda_prec_sloppy;
  QudaReconstructType reconstruct_sloppy;
  QudaPrecision cuda_prec_precondition;
  QudaReconstructType reconstruct_precondition;
  QudaGaugeFixed gauge_fix;
  int ga_pad, site_ga_pad, staple_pad, llfat_ga_pad, mom_ga_pad;
  double gaugeGiB; /* written back by loadGaugeQuda */
  int preserve_gauge;
  QudaStaggeredPhase staggered_phase_type;
  int staggered_phase_applied;
  double i_mu;
  int overlap, overwrite_mom;
  int use_resident_gauge, use_resident_mom, make_resident_gauge, make_resident_mom;
  int return_result_gauge, return_result_mom;
} QudaGaugeParam;

/* replaces /root/reference/include/quda.h:86-299 (incl. the fork's preconditionerUP/DN :225-228) */
typedef struct QudaInvertParam_s {
  QudaFieldLocation input_location, output_location;
  QudaDslashType dslash_type;
  QudaInverterType inv_type;
  double mass, kappa, m5;
  int Ls;
  double b_5[QUDA_MAX_DWF_LS];
  double c_5[QUDA_MAX_DWF_LS];
  double mu, epsilon;
  QudaTwistFlavorType twist_flavor;
  double tol, tol_restart, tol_hq;
  double true_res, true_res_hq; /* written back */
  int maxiter;
  double reliable_delta;
  int use_sloppy_partial_accumulator;
  int max_res_increase, max_res_increase_total;
  int heavy_quark_check;
  int pipeline;
  int num_offset;
  int num_src;
  int overlap;
  double offset[QUDA_MAX_MULTI_SHIFT];
  double tol_offset[QUDA_MAX_MULTI_SHIFT];
  double tol_hq_offset[QUDA_MAX_MULTI_SHIFT];
  double true_res_offset[QUDA_MAX_MULTI_SHIFT];
  double iter_res_offset[QUDA_MAX_MULTI_SHIFT];
  double true_res_hq_offset[QUDA_MAX_MULTI_SHIFT];
  QudaSolutionType solution_type;
  QudaSolveType solve_type;
  QudaMatPCType matpc_type;
  QudaDagType dagger;
  QudaMassNormalization mass_normalization;
  QudaSolverNormalization solver_normalization;
  QudaPreserveSource preserve_source;
  QudaPrecision cpu_prec, cuda_prec, cuda_prec_sloppy, cuda_prec_precondition;
  QudaDiracFieldOrder dirac_order;
  QudaGammaBasis gamma_basis;
  QudaFieldLocation clover_location;
  QudaPrecision clover_cpu_prec, clover_cuda_prec, clover_cuda_prec_sloppy, clover_cuda_prec_precondition;
  QudaCloverFieldOrder clover_order;
  QudaUseInitGuess use_init_guess;
  double clover_coeff;
  int compute_clover_trlog;
  double trlogA[2];
  int compute_clover, compute_clover_inverse, return_clover, return_clover_inverse;
  QudaVerbosity verbosity;
  int sp_pad, cl_pad;
  int iter;                                  /* written back (accumulated) */
  double spinorGiB, cloverGiB, gflops, secs; /* written back */
  QudaTune tune;
  int Nsteps;
  int gcrNkrylov;
  QudaInverterType inv_type_precondition;
  void *preconditioner;
  void *preconditionerUP;
  void *preconditionerDN;
  QudaDslashType dslash_type_precondition;
  QudaVerbosity verbosity_precondition;
  double tol_precondition;
  int maxiter_precondition;
  double omega;
  int precondition_cycle;
  QudaSchwarzType schwarz_type;
  QudaResidualType residual_type;
  QudaPrecision cuda_prec_ritz;
  int nev, max_search_dim, rhs_idx, deflation_grid, use_reduced_vector_set;
  double eigenval_tol;
  int use_cg_updates;
  double cg_iterref_tol;
  int eigcg_max_restarts, max_restart_num;
  double inc_tol;
  int make_resident_solution, use_resident_solution;
} QudaInvertParam;

/* replaces /root/reference/include/quda.h:302-325 (declared for layout completeness; unused) */
typedef struct QudaEigParam_s {
  QudaInvertParam *invert_param;
  QudaSolutionType RitzMat_lanczos, RitzMat_Convcheck;
  QudaEigType eig_type;
  double *MatPoly_param;
  int NPoly;
  double Stp_residual;
  int nk, np, f_size;
  double eigen_shift;
} QudaEigParam;

/* replaces /root/reference/include/quda.h:327-409 (incl. the fork's setup_maxiter/tol, delta_*) */
typedef struct QudaMultigridParam_s {
  QudaInvertParam *invert_param;
  int n_level;
  int geo_block_size[QUDA_MAX_MG_LEVEL][QUDA_MAX_DIM];
  int spin_block_size[QUDA_MAX_MG_LEVEL];
  int n_vec[QUDA_MAX_MG_LEVEL];
  QudaInverterType smoother[QUDA_MAX_MG_LEVEL];
  QudaSolutionType coarse_grid_solution_type[QUDA_MAX_MG_LEVEL];
  QudaSolveType smoother_solve_type[QUDA_MAX_MG_LEVEL];
  QudaMultigridCycleType cycle_type[QUDA_MAX_MG_LEVEL];
  int nu_pre[QUDA_MAX_MG_LEVEL];
  int nu_post[QUDA_MAX_MG_LEVEL];
  double smoother_tol[QUDA_MAX_MG_LEVEL];
  int setup_maxiter;
  double setup_tol;
  double omega[QUDA_MAX_MG_LEVEL];
  QudaBoolean global_reduction[QUDA_MAX_MG_LEVEL];
  QudaFieldLocation location[QUDA_MAX_MG_LEVEL];
  QudaComputeNullVector compute_null_vector;
  QudaBoolean generate_all_levels;
  QudaBoolean run_verify;
  char vec_infile[256];
  char vec_outfile[256];
  double gflops, secs; /* written back */
  double delta_muPR, delta_kappaPR, delta_cswPR, delta_muCG, delta_kappaCG, delta_cswCG;
} QudaMultigridParam;

typedef int (*QudaCommsMap)(const int *coords, void *fdata);

/* ---- hot-path entry points (reference file:line each one replaces) --------------------- */
void setVerbosityQuda(QudaVerbosity verbosity, const char prefix[], FILE *outfile);  /* quda.h:442 */
void initCommsGridQuda(int nDim, const int *dims, QudaCommsMap func, void *fdata);     /* quda.h:483, interface_quda.cpp:261-330 */
void initQudaDevice(int device);                                                       /* quda.h:495, interface_quda.cpp:359-460 */
void initQudaMemory(void);                                                             /* quda.h:503, interface_quda.cpp:462-499 */
void initQuda(int device);                                                             /* quda.h:514, interface_quda.cpp:501-519 */
void endQuda(void);                                                                    /* quda.h:519 */
QudaGaugeParam newQudaGaugeParam(void);                                                /* quda.h:528, check_params.h */
QudaInvertParam newQudaInvertParam(void);                                              /* quda.h:537 */
QudaMultigridParam newQudaMultigridParam(void);                                        /* quda.h:546 */
QudaEigParam newQudaEigParam(void);
void printQudaGaugeParam(QudaGaugeParam *param);                                       /* quda.h:561 */
void printQudaInvertParam(QudaInvertParam *param);                                     /* quda.h:567 */
void printQudaMultigridParam(QudaMultigridParam *param);                               /* quda.h:573 */
void loadGaugeQuda(void *h_gauge, QudaGaugeParam *param);                              /* quda.h:586, interface_quda.cpp:521-692 */
void freeGaugeQuda(void);                                                              /* quda.h:591 */
void saveGaugeQuda(void *h_gauge, QudaGaugeParam *param);                              /* quda.h:598 */
void invertQuda(void *h_x, void *h_b, QudaInvertParam *param);                         /* quda.h:636, interface_quda.cpp:2276-2543 */
void *newMultigridQuda(QudaMultigridParam *param);                                     /* quda.h:666, interface_quda.cpp:2257-2269 */
void destroyMultigridQuda(void *mg_instance);                                          /* quda.h:671 */
void dslashQuda(void *h_out, void *h_in, QudaInvertParam *inv_param, QudaParity parity); /* quda.h:692, interface_quda.cpp:1496-1569 */
void MatQuda(void *h_out, void *h_in, QudaInvertParam *inv_param);                     /* quda.h:738, interface_quda.cpp:1716-1784 */
void MatDagMatQuda(void *h_out, void *h_in, QudaInvertParam *inv_param);               /* quda.h:747, interface_quda.cpp:1786-1863 */

/* ---- exported for link compatibility; outside SURVEY section 8 => abort with a message -------- */
void loadCloverQuda(void *h_clover, void *h_clovinv, QudaInvertParam *inv_param);
void freeCloverQuda(void);
void invertMultiSrcQuda(void **_hp_x, void **_hp_b, QudaInvertParam *param);
void invertMultiShiftQuda(void **_hp_x, void *_hp_b, QudaInvertParam *param);
void cloverQuda(void *h_out, void *h_in, QudaInvertParam *inv_param, QudaParity *parity, int inverse);

#ifdef __cplusplus
}
#endif

#include "quda_b200_ext.h"
#endif /* QUDA_B200_API_H */
