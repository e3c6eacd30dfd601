"""ctypes binding of libquda_b200.so: the reference's public C interface for the hot path.

Mirrors /root/reference/include/quda.h (structs :25-80, :86-299, :327-409; functions :442-747) and
/root/reference/include/enum_quda.h name for name, so the parity tests read like the reference's own
tests (tests/dslash_test.cpp, tests/invert_test.cpp, tests/multigrid_invert_test.cpp).

There is NO fallback: if the CUDA library is missing or cannot be loaded, importing `lib()` raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QUDA_B200_LIB", os.path.join(_HERE, "libquda_b200.so"))  # override: tuning builds only

QUDA_MAX_DIM = 6
QUDA_MAX_MULTI_SHIFT = 32
QUDA_MAX_DWF_LS = 128
QUDA_MAX_MG_LEVEL = 4
QUDA_INVALID_ENUM = -(2 ** 31)

# --- enums (values of enum_quda.h) ---------------------------------------------------------------
QUDA_CPU_FIELD_LOCATION, QUDA_CUDA_FIELD_LOCATION = 1, 2
QUDA_SU3_LINKS = QUDA_WILSON_LINKS = 0
QUDA_QDP_GAUGE_ORDER = 5
QUDA_QDPJIT_GAUGE_ORDER, QUDA_CPS_WILSON_GAUGE_ORDER, QUDA_MILC_GAUGE_ORDER = 6, 7, 8
QUDA_ANTI_PERIODIC_T, QUDA_PERIODIC_T = -1, 1
QUDA_HALF_PRECISION, QUDA_SINGLE_PRECISION, QUDA_DOUBLE_PRECISION = 2, 4, 8
QUDA_RECONSTRUCT_NO, QUDA_RECONSTRUCT_12, QUDA_RECONSTRUCT_8 = 18, 12, 8
QUDA_GAUGE_FIXED_NO, QUDA_GAUGE_FIXED_YES = 0, 1
QUDA_WILSON_DSLASH, QUDA_CLOVER_WILSON_DSLASH = 0, 1
QUDA_TWISTED_MASS_DSLASH, QUDA_TWISTED_CLOVER_DSLASH = 7, 8
QUDA_FLOAT_CLOVER_ORDER, QUDA_FLOAT2_CLOVER_ORDER, QUDA_FLOAT4_CLOVER_ORDER, QUDA_PACKED_CLOVER_ORDER = 1, 2, 4, 5
QUDA_CG_INVERTER, QUDA_BICGSTAB_INVERTER, QUDA_GCR_INVERTER, QUDA_MR_INVERTER = 0, 1, 2, 3
QUDA_MG_INVERTER = 15
QUDA_INVALID_INVERTER = QUDA_INVALID_ENUM
(QUDA_MAT_SOLUTION, QUDA_MATDAG_MAT_SOLUTION, QUDA_MATPC_SOLUTION, QUDA_MATPC_DAG_SOLUTION,
 QUDA_MATPCDAG_MATPC_SOLUTION) = range(5)
(QUDA_DIRECT_SOLVE, QUDA_NORMOP_SOLVE, QUDA_DIRECT_PC_SOLVE, QUDA_NORMOP_PC_SOLVE) = range(4)
(QUDA_MG_CYCLE_VCYCLE, QUDA_MG_CYCLE_FCYCLE, QUDA_MG_CYCLE_WCYCLE, QUDA_MG_CYCLE_RECURSIVE) = range(4)
QUDA_ADDITIVE_SCHWARZ, QUDA_MULTIPLICATIVE_SCHWARZ = 0, 1
QUDA_L2_RELATIVE_RESIDUAL, QUDA_L2_ABSOLUTE_RESIDUAL, QUDA_HEAVY_QUARK_RESIDUAL = 1, 2, 4
(QUDA_MATPC_EVEN_EVEN, QUDA_MATPC_ODD_ODD, QUDA_MATPC_EVEN_EVEN_ASYMMETRIC,
 QUDA_MATPC_ODD_ODD_ASYMMETRIC) = range(4)
QUDA_DAG_NO, QUDA_DAG_YES = 0, 1
QUDA_KAPPA_NORMALIZATION, QUDA_MASS_NORMALIZATION, QUDA_ASYMMETRIC_MASS_NORMALIZATION = 0, 1, 2
QUDA_DEFAULT_NORMALIZATION, QUDA_SOURCE_NORMALIZATION = 0, 1
QUDA_PRESERVE_SOURCE_NO, QUDA_PRESERVE_SOURCE_YES = 0, 1
QUDA_INTERNAL_DIRAC_ORDER, QUDA_DIRAC_ORDER, QUDA_QDP_DIRAC_ORDER = 0, 1, 2
QUDA_SILENT, QUDA_SUMMARIZE, QUDA_VERBOSE, QUDA_DEBUG_VERBOSE = 0, 1, 2, 3
QUDA_TUNE_NO, QUDA_TUNE_YES = 0, 1
QUDA_EVEN_PARITY, QUDA_ODD_PARITY = 0, 1
QUDA_PARITY_SITE_SUBSET, QUDA_FULL_SITE_SUBSET = 1, 2
QUDA_DEGRAND_ROSSI_GAMMA_BASIS, QUDA_UKQCD_GAMMA_BASIS, QUDA_CHIRAL_GAMMA_BASIS = 0, 1, 2
QUDA_TWIST_MINUS, QUDA_TWIST_PLUS, QUDA_TWIST_NO = -1, 1, 0
QUDA_TWIST_NONDEG_DOUBLET, QUDA_TWIST_DEG_DOUBLET = 2, -2
QUDA_USE_INIT_GUESS_NO, QUDA_USE_INIT_GUESS_YES = 0, 1
QUDA_COMPUTE_NULL_VECTOR_NO, QUDA_COMPUTE_NULL_VECTOR_YES = 0, 1
QUDA_BOOLEAN_NO, QUDA_BOOLEAN_YES = 0, 1

_i, _d, _p = C.c_int, C.c_double, C.c_void_p


class QudaGaugeParam(C.Structure):
    _fields_ = [
        ("location", _i), ("X", _i * 4), ("anisotropy", _d), ("tadpole_coeff", _d), ("scale", _d),
        ("type", _i), ("gauge_order", _i), ("t_boundary", _i), ("cpu_prec", _i), ("cuda_prec", _i),
        ("reconstruct", _i), ("cuda_prec_sloppy", _i), ("reconstruct_sloppy", _i),
        ("cuda_prec_precondition", _i), ("reconstruct_precondition", _i), ("gauge_fix", _i),
        ("ga_pad", _i), ("site_ga_pad", _i), ("staple_pad", _i), ("llfat_ga_pad", _i), ("mom_ga_pad", _i),
        ("gaugeGiB", _d), ("preserve_gauge", _i), ("staggered_phase_type", _i),
        ("staggered_phase_applied", _i), ("i_mu", _d), ("overlap", _i), ("overwrite_mom", _i),
        ("use_resident_gauge", _i), ("use_resident_mom", _i), ("make_resident_gauge", _i),
        ("make_resident_mom", _i), ("return_result_gauge", _i), ("return_result_mom", _i),
    ]


class QudaInvertParam(C.Structure):
    _fields_ = [
        ("input_location", _i), ("output_location", _i), ("dslash_type", _i), ("inv_type", _i),
        ("mass", _d), ("kappa", _d), ("m5", _d), ("Ls", _i),
        ("b_5", _d * QUDA_MAX_DWF_LS), ("c_5", _d * QUDA_MAX_DWF_LS),
        ("mu", _d), ("epsilon", _d), ("twist_flavor", _i),
        ("tol", _d), ("tol_restart", _d), ("tol_hq", _d), ("true_res", _d), ("true_res_hq", _d),
        ("maxiter", _i), ("reliable_delta", _d), ("use_sloppy_partial_accumulator", _i),
        ("max_res_increase", _i), ("max_res_increase_total", _i), ("heavy_quark_check", _i),
        ("pipeline", _i), ("num_offset", _i), ("num_src", _i), ("overlap", _i),
        ("offset", _d * QUDA_MAX_MULTI_SHIFT), ("tol_offset", _d * QUDA_MAX_MULTI_SHIFT),
        ("tol_hq_offset", _d * QUDA_MAX_MULTI_SHIFT), ("true_res_offset", _d * QUDA_MAX_MULTI_SHIFT),
        ("iter_res_offset", _d * QUDA_MAX_MULTI_SHIFT), ("true_res_hq_offset", _d * QUDA_MAX_MULTI_SHIFT),
        ("solution_type", _i), ("solve_type", _i), ("matpc_type", _i), ("dagger", _i),
        ("mass_normalization", _i), ("solver_normalization", _i), ("preserve_source", _i),
        ("cpu_prec", _i), ("cuda_prec", _i), ("cuda_prec_sloppy", _i), ("cuda_prec_precondition", _i),
        ("dirac_order", _i), ("gamma_basis", _i), ("clover_location", _i), ("clover_cpu_prec", _i),
        ("clover_cuda_prec", _i), ("clover_cuda_prec_sloppy", _i), ("clover_cuda_prec_precondition", _i),
        ("clover_order", _i), ("use_init_guess", _i), ("clover_coeff", _d), ("compute_clover_trlog", _i),
        ("trlogA", _d * 2), ("compute_clover", _i), ("compute_clover_inverse", _i), ("return_clover", _i),
        ("return_clover_inverse", _i), ("verbosity", _i), ("sp_pad", _i), ("cl_pad", _i), ("iter", _i),
        ("spinorGiB", _d), ("cloverGiB", _d), ("gflops", _d), ("secs", _d), ("tune", _i), ("Nsteps", _i),
        ("gcrNkrylov", _i), ("inv_type_precondition", _i), ("preconditioner", _p),
        ("preconditionerUP", _p), ("preconditionerDN", _p), ("dslash_type_precondition", _i),
        ("verbosity_precondition", _i), ("tol_precondition", _d), ("maxiter_precondition", _i),
        ("omega", _d), ("precondition_cycle", _i), ("schwarz_type", _i), ("residual_type", _i),
        ("cuda_prec_ritz", _i), ("nev", _i), ("max_search_dim", _i), ("rhs_idx", _i),
        ("deflation_grid", _i), ("use_reduced_vector_set", _i), ("eigenval_tol", _d),
        ("use_cg_updates", _i), ("cg_iterref_tol", _d), ("eigcg_max_restarts", _i),
        ("max_restart_num", _i), ("inc_tol", _d), ("make_resident_solution", _i),
        ("use_resident_solution", _i),
    ]


class QudaMultigridParam(C.Structure):
    _L = QUDA_MAX_MG_LEVEL
    _fields_ = [
        ("invert_param", C.POINTER(QudaInvertParam)), ("n_level", _i),
        ("geo_block_size", (_i * QUDA_MAX_DIM) * QUDA_MAX_MG_LEVEL),
        ("spin_block_size", _i * QUDA_MAX_MG_LEVEL), ("n_vec", _i * QUDA_MAX_MG_LEVEL),
        ("smoother", _i * QUDA_MAX_MG_LEVEL), ("coarse_grid_solution_type", _i * QUDA_MAX_MG_LEVEL),
        ("smoother_solve_type", _i * QUDA_MAX_MG_LEVEL), ("cycle_type", _i * QUDA_MAX_MG_LEVEL),
        ("nu_pre", _i * QUDA_MAX_MG_LEVEL), ("nu_post", _i * QUDA_MAX_MG_LEVEL),
        ("smoother_tol", _d * QUDA_MAX_MG_LEVEL), ("setup_maxiter", _i), ("setup_tol", _d),
        ("omega", _d * QUDA_MAX_MG_LEVEL), ("global_reduction", _i * QUDA_MAX_MG_LEVEL),
        ("location", _i * QUDA_MAX_MG_LEVEL), ("compute_null_vector", _i), ("generate_all_levels", _i),
        ("run_verify", _i), ("vec_infile", C.c_char * 256), ("vec_outfile", C.c_char * 256),
        ("gflops", _d), ("secs", _d), ("delta_muPR", _d), ("delta_kappaPR", _d), ("delta_cswPR", _d),
        ("delta_muCG", _d), ("delta_kappaCG", _d), ("delta_cswCG", _d),
    ]


# functions include/quda.h + include/quda_b200_ext.h declare; tests check every one is exported
EXPORTS = [
    "setVerbosityQuda", "initCommsGridQuda", "initQudaDevice", "initQudaMemory", "initQuda", "endQuda",
    "newQudaGaugeParam", "newQudaInvertParam", "newQudaMultigridParam", "newQudaEigParam",
    "printQudaGaugeParam", "printQudaInvertParam", "printQudaMultigridParam",
    "loadGaugeQuda", "freeGaugeQuda", "saveGaugeQuda", "invertQuda", "newMultigridQuda",
    "destroyMultigridQuda", "dslashQuda", "MatQuda", "MatDagMatQuda",
    "loadCloverQuda", "freeCloverQuda", "invertMultiSrcQuda", "invertMultiShiftQuda", "cloverQuda",
    "newSpinorQudaB200", "freeSpinorQudaB200", "loadSpinorQudaB200", "saveSpinorQudaB200",
    "dslashResidentQudaB200", "matResidentQudaB200", "matDagMatResidentQudaB200", "timeDslashQudaB200", "timeDslashBatchQudaB200", "timeHaloQudaB200",
    "setDslashBlockSizeQudaB200", "kernelLaunchCountQudaB200", "computeStreamQudaB200", "syncQudaB200",
    "ncclUniqueIdQudaB200", "commsBootstrapQudaB200", "commDimPartitionedSetQudaB200",
    "commRankInfoQudaB200", "faceIndexMapQudaB200",
    "blasQudaB200", "mgVerifyQudaB200", "mgLevelInfoQudaB200", "mgProlongQudaB200", "mgRestrictQudaB200", "mgMatQudaB200",
    "mgNullVectorQudaB200", "mgCycleQudaB200", "mgTimeQudaB200", "mgMatMrhsQudaB200", "mgTimeMrhsQudaB200", "mgMrhsMaxRhsQudaB200",
    "mgCoarseLinksQudaB200", "residentSolutionQudaB200", "commPeerReduceActiveQudaB200", "timeReduceQudaB200",
    "mgProfileEnableQudaB200", "mgProfileGetQudaB200",
]

_lib = None


def lib():
    """Load libquda_b200.so (once) and declare prototypes.  Raises if the extension is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    GP, IP, MP = C.POINTER(QudaGaugeParam), C.POINTER(QudaInvertParam), C.POINTER(QudaMultigridParam)
    L.setVerbosityQuda.argtypes = [_i, C.c_char_p, _p]
    L.initCommsGridQuda.argtypes = [_i, C.POINTER(_i), _p, _p]
    L.initQudaDevice.argtypes = [_i]
    L.initQuda.argtypes = [_i]
    L.newQudaGaugeParam.restype = QudaGaugeParam
    L.newQudaInvertParam.restype = QudaInvertParam
    L.newQudaMultigridParam.restype = QudaMultigridParam
    L.printQudaGaugeParam.argtypes = [GP]
    L.printQudaInvertParam.argtypes = [IP]
    L.printQudaMultigridParam.argtypes = [MP]
    L.loadGaugeQuda.argtypes = [_p, GP]
    L.saveGaugeQuda.argtypes = [_p, GP]
    L.invertQuda.argtypes = [_p, _p, IP]
    L.newMultigridQuda.argtypes = [MP]
    L.newMultigridQuda.restype = _p
    L.destroyMultigridQuda.argtypes = [_p]
    L.dslashQuda.argtypes = [_p, _p, IP, _i]
    L.MatQuda.argtypes = [_p, _p, IP]
    L.MatDagMatQuda.argtypes = [_p, _p, IP]
    L.loadCloverQuda.argtypes = [_p, _p, IP]
    L.invertMultiSrcQuda.argtypes = [_p, _p, IP]
    L.newSpinorQudaB200.argtypes = [_i, _i]
    L.newSpinorQudaB200.restype = _p
    L.residentSolutionQudaB200.restype = _p
    L.residentSolutionQudaB200.argtypes = []
    L.freeSpinorQudaB200.argtypes = [_p]
    L.loadSpinorQudaB200.argtypes = [_p, _p, IP]
    L.saveSpinorQudaB200.argtypes = [_p, _p, IP]
    L.dslashResidentQudaB200.argtypes = [_p, _p, IP, _i]
    L.matResidentQudaB200.argtypes = [_p, _p, IP]
    L.matDagMatResidentQudaB200.argtypes = [_p, _p, IP]
    L.timeDslashQudaB200.argtypes = [_p, _p, IP, _i, _i, C.POINTER(C.c_float)]
    L.timeDslashQudaB200.restype = _d
    L.timeDslashBatchQudaB200.argtypes = [IP, _i, _i, _i, C.POINTER(_d)]
    L.timeDslashBatchQudaB200.restype = _d
    L.timeHaloQudaB200.argtypes = [_p, _p, IP, _i, _i, C.POINTER(_d)]
    L.timeHaloQudaB200.restype = _d
    L.timeReduceQudaB200.restype = _d
    L.timeReduceQudaB200.argtypes = [C.c_long, _i, _i]
    L.setDslashBlockSizeQudaB200.argtypes = [_i]
    L.kernelLaunchCountQudaB200.restype = C.c_longlong
    L.computeStreamQudaB200.restype = _p
    L.ncclUniqueIdQudaB200.argtypes = [_p]
    L.commsBootstrapQudaB200.argtypes = [_i, _i, _p]
    L.commDimPartitionedSetQudaB200.argtypes = [_i]
    L.commRankInfoQudaB200.argtypes = [C.POINTER(_i)]
    L.faceIndexMapQudaB200.argtypes = [_i, _i, _i, C.POINTER(_i)]
    L.blasQudaB200.argtypes = [C.c_char_p, C.c_long, _i, C.POINTER(_d), _p, _p, _p, _p, C.POINTER(_d)]
    L.mgVerifyQudaB200.argtypes = [_p, _i, C.POINTER(_d)]
    L.mgLevelInfoQudaB200.argtypes = [_p, _i, C.POINTER(_i)]
    L.mgProlongQudaB200.argtypes = [_p, _i, _p, _p]
    L.mgRestrictQudaB200.argtypes = [_p, _i, _p, _p]
    L.mgMatQudaB200.argtypes = [_p, _i, _i, _p, _p]
    L.mgNullVectorQudaB200.argtypes = [_p, _i, _i, _p]
    L.mgCoarseLinksQudaB200.argtypes = [_p, _i, _i, _p]
    L.mgProfileEnableQudaB200.argtypes = [_i]
    L.mgProfileGetQudaB200.argtypes = [_p, _i, C.POINTER(C.c_double), C.POINTER(C.c_long)]
    L.mgCycleQudaB200.argtypes = [_p, _i, _p, _p]
    L.mgTimeQudaB200.argtypes = [_p, _i, _i, _i]
    L.mgTimeQudaB200.restype = _d
    L.mgMatMrhsQudaB200.argtypes = [_p, _i, _i, _i, _i, _p, _p]
    L.mgTimeMrhsQudaB200.argtypes = [_p, _i, _i, _i, _i, _i]
    L.mgTimeMrhsQudaB200.restype = _d
    L.mgMrhsMaxRhsQudaB200.argtypes = [_p, _i, _i]
    _lib = L
    return L


# --- helpers shaped like the reference tests' setup code -------------------------------------------
def gauge_param(X, cpu_prec=QUDA_DOUBLE_PRECISION, cuda_prec=QUDA_DOUBLE_PRECISION,
                reconstruct=QUDA_RECONSTRUCT_NO, t_boundary=QUDA_ANTI_PERIODIC_T, anisotropy=1.0,
                cuda_prec_sloppy=None, reconstruct_sloppy=None, cuda_prec_precondition=None,
                reconstruct_precondition=None):
    """What tests/dslash_test.cpp:111-174 fills in."""
    g = lib().newQudaGaugeParam()
    for d in range(4):
        g.X[d] = X[d]
    g.anisotropy = anisotropy
    g.type = QUDA_WILSON_LINKS
    g.gauge_order = QUDA_QDP_GAUGE_ORDER
    g.t_boundary = t_boundary
    g.cpu_prec = cpu_prec
    g.cuda_prec = cuda_prec
    g.reconstruct = reconstruct
    g.cuda_prec_sloppy = cuda_prec_sloppy if cuda_prec_sloppy is not None else cuda_prec
    g.reconstruct_sloppy = reconstruct_sloppy if reconstruct_sloppy is not None else reconstruct
    g.cuda_prec_precondition = cuda_prec_precondition if cuda_prec_precondition is not None else g.cuda_prec_sloppy
    g.reconstruct_precondition = reconstruct_precondition if reconstruct_precondition is not None else g.reconstruct_sloppy
    g.gauge_fix = QUDA_GAUGE_FIXED_NO
    g.ga_pad = 0
    return g


def invert_param(kappa=0.1, mu=0.01, flavor=QUDA_TWIST_PLUS, dslash_type=QUDA_TWISTED_MASS_DSLASH,
                 matpc=QUDA_MATPC_EVEN_EVEN, dagger=QUDA_DAG_NO, cpu_prec=QUDA_DOUBLE_PRECISION,
                 cuda_prec=QUDA_DOUBLE_PRECISION, solution_type=QUDA_MATPC_SOLUTION,
                 mass_normalization=QUDA_KAPPA_NORMALIZATION, gamma_basis=QUDA_DEGRAND_ROSSI_GAMMA_BASIS,
                 dirac_order=QUDA_DIRAC_ORDER):
    """What tests/dslash_test.cpp:111-220 fills in."""
    p = lib().newQudaInvertParam()
    p.kappa = kappa
    p.mu = mu
    p.epsilon = 0.0
    p.twist_flavor = flavor
    p.dslash_type = dslash_type
    p.matpc_type = matpc
    p.dagger = dagger
    p.cpu_prec = cpu_prec
    p.cuda_prec = cuda_prec
    p.cuda_prec_sloppy = cuda_prec
    p.cuda_prec_precondition = cuda_prec
    p.solution_type = solution_type
    p.solve_type = QUDA_DIRECT_PC_SOLVE
    p.mass_normalization = mass_normalization
    p.gamma_basis = gamma_basis
    p.dirac_order = dirac_order
    p.input_location = QUDA_CPU_FIELD_LOCATION
    p.output_location = QUDA_CPU_FIELD_LOCATION
    p.tune = QUDA_TUNE_NO
    p.sp_pad = 0
    p.cl_pad = 0
    p.verbosity = QUDA_SILENT
    p.Ls = 1
    p.mass = 0.0
    p.m5 = 0.0
    p.inv_type = QUDA_GCR_INVERTER
    p.preserve_source = QUDA_PRESERVE_SOURCE_YES
    return p


def multigrid_param(inv_param, n_level=2, geo_block=((4, 4, 4, 4),), n_vec=(24,), nu_pre=2, nu_post=2,
                    smoother_tol=0.25, omega=0.85, setup_maxiter=500, setup_tol=5e-6,
                    cycle=QUDA_MG_CYCLE_RECURSIVE, generate_all_levels=True, run_verify=True, solve_type=None):
    """What tests/multigrid_invert_test.cpp:161-286 (setMultigridParam) fills in.
    `inv_param` must stay alive as long as the returned struct is used.
    solve_type = the OUTER solve type the hierarchy is meant for (the test's global `solve_type`, tests/test_util.cpp:1600 defaults to
    QUDA_DIRECT_PC_SOLVE): an even-odd outer solve injects single-parity fields into the coarse grids (multigrid_invert_test.cpp:252),
    i.e. the preconditioned coarsening.  None = QUDA_DIRECT_SOLVE (full-field injection)."""
    pc_inject = solve_type == QUDA_DIRECT_PC_SOLVE
    m = lib().newQudaMultigridParam()
    m.invert_param = C.pointer(inv_param)
    m.n_level = n_level
    for l in range(n_level):
        gb = geo_block[min(l, len(geo_block) - 1)]
        for d in range(QUDA_MAX_DIM):
            m.geo_block_size[l][d] = gb[d] if d < 4 else 1
        m.spin_block_size[l] = 2 if l == 0 else 1
        m.n_vec[l] = n_vec[min(l, len(n_vec) - 1)]
        m.nu_pre[l] = nu_pre
        m.nu_post[l] = nu_post
        m.cycle_type[l] = cycle
        m.smoother[l] = QUDA_MR_INVERTER
        m.smoother_tol[l] = smoother_tol
        m.global_reduction[l] = QUDA_BOOLEAN_YES
        m.smoother_solve_type[l] = QUDA_DIRECT_PC_SOLVE
        m.coarse_grid_solution_type[l] = QUDA_MATPC_SOLUTION if pc_inject else QUDA_MAT_SOLUTION
        m.omega[l] = omega
        m.location[l] = QUDA_CUDA_FIELD_LOCATION
    m.smoother[n_level - 1] = QUDA_GCR_INVERTER
    m.compute_null_vector = QUDA_COMPUTE_NULL_VECTOR_YES
    m.generate_all_levels = QUDA_BOOLEAN_YES if generate_all_levels else QUDA_BOOLEAN_NO
    m.run_verify = QUDA_BOOLEAN_YES if run_verify else QUDA_BOOLEAN_NO
    m.setup_maxiter = setup_maxiter
    m.setup_tol = setup_tol
    m.delta_muPR = 1.0
    m.delta_kappaPR = 1.0
    return m
