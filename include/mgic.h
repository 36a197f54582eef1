/*
 * mgic.h -- C ABI of the B200-native multigrid hot path of MG_IC_code.
 *
 * Two nested boundaries, both exported by libmgic_b200.so:
 *
 *  (A) mgic_*  : device-resident operator API.  One mgic_op is one
 *      VariableCoeffPoissonOperator (Source/VariableCoeffPoissonOperator.H:25)
 *      on one multigrid/AMR level; fields live in HBM; no per-call PCIe
 *      traffic.  The C++ host mirror (mg_ic_code_b200/host) wraps these in the
 *      reference's AMRLevelOp<LevelData<FArrayBox>> interface.
 *
 *  (B) the Fortran symbols of the reference's .ChF kernels with their exact
 *      argument lists (host pointers; H2D, CUDA kernel, D2H) -- the link-time
 *      drop-in for Source/VariableCoeffPoissonOperatorF_F.H and
 *      Source/SetLevelDataF_F.H.  Declared in mgic_chf.h.
 *
 * Plain C: pointers and sizes only.  Every mgic_* function returns 0 on
 * success, nonzero on error (message via mgic_last_error()); the reference's
 * error behaviour (MayDay::Error/Abort == process abort) is re-created by the
 * C++ host mirror on a nonzero return.  There is NO CPU fallback: without a
 * CUDA device every compute entry point fails with MGIC_ERR_NO_DEVICE.
 *
 * All arithmetic is FP64; kernels are compiled without FMA contraction and
 * follow the operation order of the .ChF sources.
 */
#ifndef MGIC_H
#define MGIC_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGIC_OK 0
#define MGIC_ERR_NO_DEVICE 1
#define MGIC_ERR_CUDA 2
#define MGIC_ERR_ARG 3
#define MGIC_ERR_STATE 4

/* BC flags: params.txt bc_lo/bc_hi (Source/SetBCs.cpp:70-94) */
#define MGIC_BC_DIRICHLET 0
#define MGIC_BC_NEUMANN 1
#define MGIC_BC_PERIODIC 2

/* CoarseAverage::averageType [Chombo]; params.txt coefficient_average_type */
#define MGIC_AVG_ARITHMETIC 0
#define MGIC_AVG_HARMONIC 1

typedef struct mgic_ctx mgic_ctx;     /* device + streams (+ peers)                         */
typedef struct mgic_op mgic_op;       /* VariableCoeffPoissonOperator on one level          */
typedef struct mgic_field mgic_field; /* one-component FP64 level array resident in HBM     */
typedef struct mgic_mg mgic_mg;       /* MultiGrid hierarchy built by the factory           */
typedef struct mgic_vars mgic_vars;   /* multigrid_vars (8 comps, MultigridUserVariables.hpp) */

/* Mirrors PoissonParameters (Source/PoissonParameters.H) + the solver knobs read at
 * Main_PoissonSolver.cpp:106-126 and the BC keys of Source/SetBCs.cpp:45-56. */
typedef struct mgic_params {
  double alpha, beta;
  double G_Newton, phi_amplitude, phi_wavelength;
  double bh1_bare_mass, bh1_spin, bh1_momentum, bh1_offset;
  double bh2_bare_mass, bh2_spin, bh2_momentum, bh2_offset;
  double L;
  double bc_value;
  double tolerance;
  int N[3];
  int max_level;
  int block_factor, max_grid_size;
  int coefficient_average_type;
  int is_periodic;
  int bc_lo[3], bc_hi[3];
  int numMGsmooth, numMGIterations, preCondSolverDepth;
  int max_iterations, max_NL_iterations;
  int verbosity;
} mgic_params;

const char *mgic_last_error(void);
const char *mgic_version(void);

/* ------------------------------------------------------------------ context */
int mgic_ctx_create(int device, mgic_ctx **out);
int mgic_ctx_destroy(mgic_ctx *);
int mgic_ctx_sync(mgic_ctx *);
/* use an existing CUDA stream (e.g. torch's current stream) for all launches; 0 = own stream */
int mgic_ctx_set_stream(mgic_ctx *, void *cuda_stream);
void *mgic_ctx_stream(mgic_ctx *);
/* number of kernels this library has launched on the context since creation (bench.py gpu_launches) */
long long mgic_ctx_launch_count(mgic_ctx *);
/* tuning knobs: "fused_cfg" (tile shape of the fused GSRB sweep), "fused_min_cells" (smaller levels use the
 * per-colour kernel), "bottom_kernel" (bottom BiCGStab as 1: all vectors in the distributed shared memory of one
 * cluster when the level fits, else bricks held by 16-CTA clusters, else 4; 5: cluster-held bricks first; 4: per-CTA
 * bricks with four grid barriers per iteration; 2: one kernel in a thread-block cluster; 3: the same as a cooperative
 * grid; 0: host-driven launches), "use_graph" (1: V-cycles replayed as CUDA graphs), "fuse_transfers" (1: setToZero /
 * prolongIncrement folded into the following fused sweep), "agglo_cells" (multi-rank: MG depths whose slab has at most
 * this many cells are gathered onto every rank), "overlap_halo" (multi-rank: exchange on a second stream while the
 * interior planes are swept), "p2p_halo" (multi-rank: 1 halo planes by NVLink peer stores, 0 ncclSend/ncclRecv; same
 * value on every rank), "fold_halo" (multi-rank: 1 the fused sweep stores its boundary planes into the neighbours' ghost planes
 * itself instead of an exchange kernel before every sweep; measured slower, default 0), "restrict_tma" (1: restrictResidual of rectangular, non-periodic levels of at least 8 x fused_min_cells cells (256^3) by the
 * plane-streaming kernel; 0: one thread per coarse cell everywhere; same bits), "fused_patch" (1: AMR levels that are
 * one box are swept by the fused kernel too, homogeneousCFInterp evaluated in the sweep; 0: per-colour kernel).  Fields do not depend on fused_* / use_graph / *_halo; bottom_kernel changes only the
 * summation order of the bottom solver's dot products. */
int mgic_ctx_set_option(mgic_ctx *, const char *name, long long value);
/* reads an option back; also "last_bottom_kernel": which bottom solver ran last (0 host-driven, 1 one cluster's shared
 * memory, 2 cluster kernel, 3 cooperative grid, 4 brick kernel, 5 cluster-sized bricks); -1 = unknown name */
long long mgic_ctx_get_option(mgic_ctx *, const char *name);
/* per-launch CUDA-event timing of the dominant kernel (the finest level's GSRB launches): arm with enable = 1,
 * run, then read the number of timed launches and their summed device time (bench.py roofline) */
int mgic_ctx_profile(mgic_ctx *, int enable);
int mgic_ctx_profile_read(mgic_ctx *, long long *launches, double *total_ms);
/* same per category: 0 finest GSRB, 1 halo exchange, 2 all-gather, 3 bottom solve, 4 restrict, 5 coarser GSRB, 6 prolong */
int mgic_ctx_profile_read_tag(mgic_ctx *, int tag, long long *launches, double *total_ms);
/* driver allocations the library has made so far in this process (calls, seconds spent inside cudaMalloc / cudaFree).
 * Arrays below 256 MiB are ranges of a few large chunks per device (kept until the process ends; MGIC_ARENA=0 in the
 * environment: one cudaMalloc per array), because a B200 box pays milliseconds per cudaMalloc and a hierarchy of many small
 * levels needs hundreds of arrays (tools/time_to_solution.py) */
int mgic_alloc_stats(long long *alloc_calls, double *alloc_seconds, long long *free_calls, double *free_seconds);
/* MGIC_ARENA_GUARD=<bytes> in the environment puts that many bytes of 0xA5 in front of and behind every array (except the
 * IPC-exported fields of multi-rank z-slab levels); they are checked when an array is freed and here, for every live array.
 * *violations = arrays found with a damaged band so far in this process (each is also reported on stderr). */
int mgic_arena_guard_check(long long *violations, long long *arrays_checked);
/* overruns and underruns a scratch array by one byte on purpose: 0 = both were detected (and taken off the count again),
 * 1 = no bands configured, 2 = missed */
int mgic_arena_guard_selftest(int device);
/* the sub-allocator's bookkeeping checked on the host alone (no device): `ops` random allocations / frees; 0 = consistent */
int mgic_arena_selftest(unsigned seed, int ops);
/* multi-GPU z-slab decomposition: this context is rank `rank` of `nranks` (one process per GPU);
 * peers are wired with mgic_ctx_set_peer_halo(); see mgic_comm.h */
int mgic_ctx_set_rank(mgic_ctx *, int rank, int nranks);

/* ----------------------------------------------------------------- operator
 * replaces: VariableCoeffPoissonOperator + AMRPoissonOp::define
 * (Source/VariableCoeffPoissonOperatorFactory.cpp:187-192).  n = cells of the level's
 * (rectangular, single-box-per-rank) domain; the rank owns global planes
 * [k0, k0+nz_local).  Single GPU: k0 = 0, nz_local = n[2]. */
int mgic_op_create(mgic_ctx *, const int n[3], int k0, int nz_local, double dx, double alpha, double beta,
                   const int bc_lo[3], const int bc_hi[3], double bc_value, mgic_op **out);
/* One AMR level > 0 (first half): the operator on a box [lo, hi] (inclusive, level index space,
 * coarsenable by 2) of the refined domain n_domain.  Faces of the box that are not domain faces are coarse-fine
 * interfaces; there the operator does what the reference's class does itself on such a level --
 * [Chombo] AMRPoissonOp::homogeneousCFInterp before each colour pass of levelGSRB
 * (Source/VariableCoeffPoissonOperator.cpp:296) and before restrictResidual (:156): the ghost value is the parabola
 * through the two interior cells and a ZERO coarse value.  mgic_op_relax / _level_gsrb / _precond /
 * _restrict_residual and the vector operations work on it; mgic_op_residual / _apply fail with MGIC_ERR_ARG: on a
 * patch the coarse-fine ghost values come from the coarser level (next two functions).  Fields are patch-shaped. */
int mgic_op_create_patch(mgic_ctx *, const int n_domain[3], const int lo[3], const int hi[3], double dx, double dx_coarse,
                         double alpha, double beta, const int bc_lo[3], const int bc_hi[3], double bc_value, mgic_op **out);
/* The same for an AMR level (or one connected part of it) made of SEVERAL boxes -- what BRMeshRefine hands to
 * VariableCoeffPoissonOperatorFactory::define (Source/SetGrids.cpp:113-114, ...Factory.cpp:59-106): boxes of at most
 * max_grid_size cells that touch and whose union need not be a rectangle.  boxes = nboxes x {lo0,lo1,lo2,hi0,hi1,hi2}
 * (inclusive, the level's index space, disjoint, coarsenable by 2).  The level lives in ONE array over the union's
 * bounding box with a cell mask: [Chombo] LevelData::exchange between the level's boxes (VariableCoeffPoissonOperator.cpp:
 * 48,131,163,301) is a neighbour read, coarse-fine ghosts (homogeneousCFInterp / QuadCFInterp) are evaluated per cell.
 * Fields have the bounding box's shape; cells outside the boxes are kept at zero.  Results are independent of how the
 * union is cut into boxes.  A union that fills its bounding box is the rectangular patch above. */
int mgic_op_create_patch_boxes(mgic_ctx *, const int n_domain[3], int nboxes, const int *boxes, double dx, double dx_coarse,
                               double alpha, double beta, const int bc_lo[3], const int bc_hi[3], double bc_value, mgic_op **out);
long long mgic_op_valid_cells(const mgic_op *);            /* cells of the level's boxes (the bounding box for a rectangle) */
int mgic_op_get_mask(const mgic_op *, unsigned char *host);  /* 1 = cell of the level, over the bounding box, x fastest */
/* [Chombo] AMRPoissonOp::AMROperatorNF / AMRResidualNF (inherited by the reference's operator, VariableCoeffPoissonOperator.H:25)
 * on a patch that has a coarser but no finer level: QuadCFInterp::coarseFineInterp(phi, phi_coarse) -- per fine ghost cell
 * a second-order tangential Taylor interpolation of the coarse field, then the parabola through it and the two interior
 * fine cells -- followed by applyOpI / residualI.  phi_coarse is a field of the coarser level;
 * coarse_lo = the index, in that level's index space, of its cell (0,0,0): {0,0,0} for a level that covers the domain,
 * the coarser patch's lo otherwise.  It must cover the coarsened patch grown by two cells (proper nesting). */
int mgic_op_amr_operator_nf(mgic_op *patch, mgic_field *lhs, mgic_field *phi, const mgic_field *phi_coarse, const int coarse_lo[3],
                            int homogeneous);
int mgic_op_amr_residual_nf(mgic_op *patch, mgic_field *lhs, mgic_field *phi, const mgic_field *phi_coarse, const int coarse_lo[3],
                            const mgic_field *rhs, int homogeneous);
/* ----------------------------------------------------------------- AMR hierarchy
 * replaces: [Chombo] AMRMultiGrid::AMRVCycle as MultilevelLinearOp::preCond drives it, MultilevelLinearOp itself and the
 * outer BiCGStabSolver<Vector<LevelData<FArrayBox>*>> (Main_PoissonSolver.cpp:103-117,169-184) on a
 * hierarchy of levels: level 0 = the MG hierarchy `base` (one array), finer level l = a list of patch operators
 * (mgic_op_create_patch, coefficients set), each nested with refinement ratio 2 in ONE array of the level below
 * (proper nesting: its coarse cells and their face neighbours are cells of that array, except at domain faces) and not
 * touching its siblings -- boxes that touch form ONE node (mgic_op_create_patch_boxes), so a level is given as its
 * connected components (config C4: one box on level 1, two disjoint boxes on level 2).
 * reflux is the reference's no-op (VariableCoeffPoissonOperator.cpp:264-271).
 * A LEVEL VECTOR is an array of mgic_amr_nodes() fields: [0] on the base level, then one per patch in creation order
 * (level 1's patches first).  npatches[l-1] = number of patches of level l; `patches` is the flattened list.
 * mgic_amr_create = a chain, one patch per level. */
typedef struct mgic_amr mgic_amr;
int mgic_amr_create(mgic_mg *base, int nfiner, mgic_op *const *patches, mgic_amr **out);
int mgic_amr_create_levels(mgic_mg *base, int nfiner, const int *npatches, mgic_op *const *patches, mgic_amr **out);
int mgic_amr_destroy(mgic_amr *);
int mgic_amr_levels(const mgic_amr *);
int mgic_amr_nodes(const mgic_amr *);
int mgic_amr_node_info(const mgic_amr *, int node, int *level, int *parent_node);
/* AMRVCycle: corr (out) = the correction of one cycle for the residuals res, pre = post = numMGsmooth, MultiGrid::oneCycle
 * on the base level; res of a coarser array under a finer patch is replaced by the averaged fine residual */
int mgic_amr_vcycle(mgic_amr *, mgic_field *const *corr, mgic_field *const *res);
/* MultilevelLinearOp::applyOp / residual: per level applyOpI / residualI with the coarse-fine ghost cells interpolated
 * from the level below (AMROperatorNF / AMRResidualNF) */
int mgic_amr_apply(mgic_amr *, mgic_field *const *lhs, mgic_field *const *phi, int homogeneous);
int mgic_amr_residual(mgic_amr *, mgic_field *const *res, mgic_field *const *phi, mgic_field *const *rhs, int homogeneous);
/* AMRPoissonOp::zeroCovered level by level; CoarseAverage::averageToCoarse from the finest level down */
int mgic_amr_zero_covered(mgic_amr *, mgic_field *const *x);
int mgic_amr_average_down(mgic_amr *, mgic_field *const *x);
/* over the valid cells NOT covered by a finer level: ord 0 max |x| (the outer solver's norm, Main_PoissonSolver.cpp:176);
 * ord 1, 2 computeNorm's (sum |x|^p dx_l^3)^(1/p) (:208); dot = sum x*y*dx_l^3 (MultilevelLinearOp::dotProduct) */
int mgic_amr_norm(mgic_amr *, mgic_field *const *x, int ord, double *out);
int mgic_amr_dot(mgic_amr *, mgic_field *const *x, mgic_field *const *y, double *out);
/* MultilevelLinearOp::preCond: cor = 0, then numMGIterations AMR V-cycles, each on the residual of the correction so far */
int mgic_amr_precond(mgic_amr *, mgic_field *const *cor, mgic_field *const *res);
/* solver.solve(dpsi, rhs) (Main_PoissonSolver.cpp:184) on the hierarchy: BiCGStab over the level vectors, AMR V-cycle
 * preconditioner, max-norm, eps = tolerance, imax = max_iterations; exit_status / norms as mgic_mg_outer_solve */
int mgic_amr_outer_solve(mgic_amr *, mgic_field *const *dpsi, mgic_field *const *rhs, int *iterations, int *exit_status,
                         double *norms, int max_norms);
/* the coarse-fine ghost values of one face (0 x-lo, 1 x-hi, 2 y-lo, 3 y-hi, 4 z-lo, 5 z-hi) left by the last of the two
 * calls above: x faces [j + ny*k], y faces [i + nx*k], z faces [i + nx*j] */
int mgic_op_cf_ghosts(mgic_op *patch, int face, double *host);
int mgic_op_destroy(mgic_op *);
/* setCoefs (VariableCoeffPoissonOperator.cpp:208-218): coefficients are SHARED (caller keeps them alive);
 * bCoef may be NULL == the constant 1 (set_b_coef, Source/SetLevelData.cpp:330-340) */
int mgic_op_set_coefs(mgic_op *, mgic_field *aCoef, mgic_field *bCoef, double alpha, double beta);
int mgic_op_set_alpha_beta(mgic_op *, double alpha, double beta);       /* :196-206 */
int mgic_op_reset_lambda(mgic_op *);                                    /* :220-249 */
int mgic_op_compute_lambda(mgic_op *);                                  /* :252-260 */
int mgic_op_get_lambda(mgic_op *, mgic_field **lambda);
int mgic_op_dims(const mgic_op *, int n[3], int *k0, int *nz_local, double *dx);

/* fields: AMRLevelOp::create / createCoarser */
int mgic_field_create(mgic_op *like, mgic_field **out);
int mgic_field_destroy(mgic_field *);
/* global ghost-free host array, Fortran order idx = i + nx*(j + ny*k), k global; the rank moves its slab */
int mgic_field_upload(mgic_field *, const double *host);
int mgic_field_download(const mgic_field *, double *host);
/* same, ordered on the context stream without a host sync (pinned host memory; pair with mgic_ctx_sync) */
int mgic_field_upload_async(mgic_field *, const double *host);
int mgic_field_download_async(const mgic_field *, double *host);
/* the same copies on the context's transfer streams (one per PCIe direction) instead of the compute stream, so that
 * the upload of the next right-hand side, a V-cycle and the download of the previous correction overlap: each copy is
 * ordered after everything issued on the compute stream before the call; mgic_field_wait() makes the compute stream
 * wait for the field's last prefetch / writeback (call it before the field is used or overwritten by a kernel);
 * mgic_ctx_sync() / mgic_field_sync() also drain the transfer streams.  Pinned host memory. */
int mgic_field_prefetch(mgic_field *, const double *host);
int mgic_field_writeback(mgic_field *, double *host);
int mgic_field_wait(mgic_field *);
/* one FArrayBox (host, inclusive bounds incl. ghosts, Fortran order): copies fab ∩ region ∩ (this rank's slab) */
int mgic_field_upload_fab(mgic_field *, const double *fab, const int fab_lo[3], const int fab_hi[3],
                          const int region_lo[3], const int region_hi[3]);
int mgic_field_download_fab(const mgic_field *, double *fab, const int fab_lo[3], const int fab_hi[3],
                            const int region_lo[3], const int region_hi[3]);
/* wait for the asynchronous fab copies issued on the field's context */
int mgic_field_sync(const mgic_field *);
/* device pointer of local cell (0,0,0) and strides in doubles (plumbing for torch / tests) */
int mgic_field_devptr(const mgic_field *, void **ptr, long long *stride_y, long long *stride_z);

/* operator methods; names follow Source/VariableCoeffPoissonOperator.H:40-168 and AMRPoissonOp [Chombo] */
int mgic_op_relax(mgic_op *, mgic_field *e, const mgic_field *residual, int iterations);   /* relax -> levelGSRB :273 */
int mgic_op_gsrb_color(mgic_op *, mgic_field *e, const mgic_field *residual, int whichPass); /* one pass of :290-331 */
int mgic_op_level_jacobi(mgic_op *, mgic_field *e, const mgic_field *residual);             /* :360-385 */
int mgic_op_residual(mgic_op *, mgic_field *lhs, mgic_field *phi, const mgic_field *rhs, int homogeneous); /* :30-67 */
int mgic_op_apply(mgic_op *, mgic_field *lhs, mgic_field *phi, int homogeneous);            /* applyOpI :106-121 */
int mgic_op_apply_no_boundary(mgic_op *, mgic_field *lhs, mgic_field *phi);                 /* :123-149 */
int mgic_op_restrict_residual(mgic_op *fine, mgic_field *resCoarse, mgic_field *phiFine, const mgic_field *rhsFine); /* :151-194 */
int mgic_op_prolong_increment(mgic_op *fine, mgic_field *phiFine, const mgic_field *correctCoarse); /* [Chombo] FORT_PROLONG */
int mgic_op_precond(mgic_op *, mgic_field *phi, const mgic_field *rhs);                     /* :72-104 */
/* BLAS-1 over valid cells [Chombo AMRPoissonOp] */
int mgic_op_norm(mgic_op *, const mgic_field *x, int ord, double *out);  /* ord 0: max|x|, 1: sum|x|, 2: sqrt(sum x^2) */
int mgic_op_dot(mgic_op *, const mgic_field *x, const mgic_field *y, double *out);
int mgic_op_incr(mgic_op *, mgic_field *y, const mgic_field *x, double scale);             /* y += scale*x */
int mgic_op_axby(mgic_op *, mgic_field *y, const mgic_field *x1, const mgic_field *x2, double a, double b);
int mgic_op_scale(mgic_op *, mgic_field *y, double s);
int mgic_op_assign(mgic_op *, mgic_field *y, const mgic_field *x);
int mgic_op_set_to_zero(mgic_op *, mgic_field *y);
int mgic_op_set_val(mgic_op *, mgic_field *y, double v);
/* smoother implementation used by relax: 0 = one launch per colour pass, 1 = fused red+black plane-streaming sweep (default) */
int mgic_op_set_smoother(mgic_op *, int kind);

/* ---------------------------------------------------------------- factory / MG
 * replaces VariableCoeffPoissonOperatorFactory::define + MGnewOp (Factory.cpp:59-106,139-234) driven by
 * [Chombo] MultiGrid::define: ops for depth 0,1,.. are built until the depth limit
 * "boxes coarsenable by 2^depth * s_maxCoarse" (Factory.cpp:168-172) fails; boxes are the
 * domainSplit lattice of max_grid_size (Source/SetGrids.cpp:54-58) unless given explicitly.
 * Coefficients of depth>0 are CoarseAverage'd directly from depth 0 (Factory.cpp:199-227). */
int mgic_mg_create(mgic_ctx *, const mgic_params *, mgic_field *aCoef0, mgic_field *bCoef0, mgic_mg **out);
/* flags: MGIC_MG_KEEP_B = always stream bCoef (by default a bCoef that is identically 1 -- set_b_coef,
 * Source/SetLevelData.cpp:330-340 -- is dropped from the kernels; x*1 == x, so results are bit-identical) */
#define MGIC_MG_KEEP_B 1
int mgic_mg_create_ex(mgic_ctx *, const mgic_params *, mgic_field *aCoef0, mgic_field *bCoef0, int flags, mgic_mg **out);
int mgic_mg_b_is_one(const mgic_mg *);
int mgic_mg_destroy(mgic_mg *);
int mgic_mg_depths(const mgic_mg *);
int mgic_mg_op(mgic_mg *, int depth, mgic_op **op);               /* MGnewOp(domain, depth) result; NULL past the end */
int mgic_mg_scratch(mgic_mg *, int depth, mgic_field **e, mgic_field **r);  /* MultiGrid m_correction/m_residual */
/* (re)coarsen coefficients after aCoef0 changed + recompute lambda (what re-defining the factory does per NL iteration) */
int mgic_mg_refresh_coefs(mgic_mg *);
/* [Chombo] MultiGrid::oneCycle (homogeneous) == cycle(0, e, r): relax/restrict/recurse/prolong/relax + bottom solve */
int mgic_mg_vcycle(mgic_mg *, mgic_field *e, const mgic_field *r);
/* setToZero(e) + oneCycle(e, r) in one call: what [Chombo] MultilevelLinearOp::preCond issues first; the zero fill and the
 * first read of e are skipped */
int mgic_mg_vcycle_from_zero(mgic_mg *, mgic_field *e, const mgic_field *r);
int mgic_mg_bottom_solve(mgic_mg *, mgic_field *e, const mgic_field *r, int *iterations);
int mgic_mg_last_bottom_iterations(mgic_mg *);
/* select the smoother implementation on every depth (see mgic_op_set_smoother) */
int mgic_mg_set_smoother(mgic_mg *, int kind);

/* f1: BiCGStabSolver<Vector<LevelData*>>::solve preconditioned by numMGIterations V-cycles
 * (Main_PoissonSolver.cpp:103-117,173-184); norms[0..iterations] = residual history (normType 0) */
int mgic_mg_outer_solve(mgic_mg *, mgic_field *dpsi, const mgic_field *rhs, int *iterations, int *exit_status,
                        double *norms, int max_norms);

/* ---------------------------------------------------------------- source terms
 * multigrid_vars: 8 components (MultigridUserVariables.hpp:10-23) with one ghost layer on every side */
int mgic_vars_create(mgic_ctx *, const mgic_params *, int k0, int nz_local, mgic_vars **out);
int mgic_vars_destroy(mgic_vars *);
int mgic_vars_download(const mgic_vars *, int comp, double *host /* ghost-free global array */);
int mgic_vars_download_ghosted(const mgic_vars *, int comp, double *host /* (n+2)^3 */);
/* set_initial_conditions (Source/SetLevelData.cpp:32-71): psi=1, dpsi=0, phi, Aij over the ghosted box */
int mgic_set_initial_conditions(mgic_vars *, mgic_field *dpsi);
/* set_a_coef / set_b_coef / set_rhs (Source/SetLevelData.cpp:73-127,281-340) */
int mgic_set_a_coef(mgic_vars *, mgic_field *aCoef, double constant_K);
int mgic_set_b_coef(mgic_vars *, mgic_field *bCoef);
int mgic_set_rhs(mgic_vars *, mgic_field *rhs, double constant_K);
int mgic_set_rhs_and_a_coef(mgic_vars *, mgic_field *rhs, mgic_field *aCoef, double constant_K);  /* both in one pass */
/* set_update_psi0 (Source/SetLevelData.cpp:243-263) + computeNorm (Main_PoissonSolver.cpp:208) */
int mgic_update_psi0(mgic_vars *, mgic_op *op0, mgic_field *dpsi, double *dpsi_norm);
/* ----------------------------------------------------------------- the problem on an AMR hierarchy (max_level > 0)
 * replaces: poissonSolve, Main_PoissonSolver.cpp:45-256, for a hierarchy: per level multigrid_vars / dpsi / rhs / aCoef /
 * bCoef (:79-88), set_initial_conditions (:93, Source/SetLevelData.cpp:32-71 with the LEVEL's dx) and per nonlinear
 * iteration (:131-216): set_rhs / set_a_coef / set_b_coef on every level (Source/SetLevelData.cpp:73-127,281-340),
 * defineOperatorFactory + MultilevelLinearOp + BiCGStabSolver rebuilt (:163-178), solver.solve (:184), then per level
 * QuadCFInterp::coarseFineInterp(dpsi, dpsi_coarser) + exchange + set_update_psi0 (:189-205, Source/SetLevelData.cpp:243-263)
 * and computeNorm(dpsi, p = 2) over the cells no finer level covers (:208).
 * A level > 0 is given as its connected components ("nodes"); a node is a list of boxes that may touch, held in ONE
 * masked array (mgic_op_create_patch_boxes).  nnodes[l-1] = nodes of level l, nboxes[q] = boxes of finer node q (levels
 * flattened, level 1 first), boxes = 6 ints each {lo0,lo1,lo2,hi0,hi1,hi2} in the level's index space (refinement ratio 2
 * on every level, Source/PoissonParameters.cpp:75-79).  Node 0 is the base level (P->N).  One GPU. */
typedef struct mgic_hier mgic_hier;
int mgic_hier_create(mgic_ctx *, const mgic_params *P, int nfiner, const int *nnodes, const int *nboxes, const int *boxes, mgic_hier **out);
int mgic_hier_destroy(mgic_hier *);
int mgic_hier_nodes(const mgic_hier *);
int mgic_hier_node_info(const mgic_hier *, int node, int *level, int lo[3], int n[3], long long *valid_cells);
int mgic_hier_get_mask(const mgic_hier *, int node, unsigned char *host);            /* 1 = cell of the level, bounding-box shaped */
int mgic_hier_set_initial_conditions(mgic_hier *);                                    /* Main_PoissonSolver.cpp:90-96 */
int mgic_hier_nl_iteration(mgic_hier *, double *dpsi_norm, int *solver_iterations, int *solver_status);   /* :131-212 body */
/* the same body step by step, as the reference's driver calls it (host/dropin maps Main_PoissonSolver.cpp's calls onto these) */
int mgic_hier_set_solver_params(mgic_hier *, int numMGsmooth, int numMGIterations, int preCondSolverDepth, double tolerance,
                                int max_iterations);                         /* :108-123 */
int mgic_hier_set_sources(mgic_hier *, double constant_K);                   /* set_a_coef / set_b_coef / set_rhs, :154-160 */
int mgic_hier_define_solver(mgic_hier *);                                    /* defineOperatorFactory + mlOp.define, :163-170 */
int mgic_hier_solve(mgic_hier *, int *iterations, int *exit_status);         /* solver.solve(dpsi, rhs), :173-184 */
int mgic_hier_update_psi(mgic_hier *);                                       /* QuadCFInterp + set_update_psi0 per level, :189-205 */
int mgic_hier_dpsi_norm(mgic_hier *, double *out);                           /* computeNorm(dpsi), :208 */
int mgic_hier_release_solver(mgic_hier *);                                   /* no-op: the solver objects are reused, coefficients refreshed in place */
int mgic_hier_nl_solve(mgic_hier *, double *dpsi_norms, int max_out, int *nl_iterations);                 /* :93 + the loop */
/* replaces: output_final_data + set_output_data (Source/WriteOutput.H:127-227, Source/SetLevelData.cpp:343-396): the GRChombo
 * checkpoint (32 variables, three ghost layers per box, header / per-level attributes as the reference sets them).  No HDF5
 * in this build: a self-describing container ("MGICCHK1" + JSON header + the doubles in Chombo's dataset order) that
 * tools/mgic2hdf5.py converts to vcPoissonFinal.3d.hdf5.  Multi-rank context: collective -- psi of the z-slab-distributed
 * base level is gathered (the boxes' ghost layers reach into the neighbours' planes), rank 0 writes `path`, the others write
 * nothing. */
int mgic_hier_write_checkpoint(mgic_hier *, const char *path, double constant_K);
/* what: 0..7 multigrid_vars component (0 = psi, MultigridUserVariables.hpp), 8 dpsi, 9 rhs, 10 aCoef; bounding-box shaped */
int mgic_hier_download(const mgic_hier *, int node, int what, double *host);
/* ----------------------------------------------------------------- grid generation
 * replaces: set_grids, set_tag_cells (Source/SetGrids.cpp:31-148,172-207) with set_regrid_condition
 * (Source/SetLevelData.cpp:188-240) and [Chombo] BRMeshRefine::regrid (nesting radius 2, :64-68,113-114): base level =
 * domainSplit lattice of max_grid_size boxes; while a new level appears, on every level tag the cells whose |regrid
 * condition| >= refine_threshold * max over the level, grow the tags by two cells, and re-cluster (Berger-Rigoutsos
 * signatures, fill_ratio, boxes multiples of block_factor and at most max_grid_size, proper nesting).  P->max_level is the
 * deepest level allowed; refine_threshold / fill_ratio are params.txt's keys of that name (buffer_size is read by the
 * reference but never used: Source/PoissonParameters.cpp:55).  BRMeshRefine is Chombo's (not vendored): restated, see
 * csrc/grids.cu.  mgic_grids_regrid = the clustering alone on caller-supplied tags (host only, no device needed). */
typedef struct mgic_grids mgic_grids;
int mgic_grids_generate(mgic_ctx *, const mgic_params *P, double refine_threshold, double fill_ratio, mgic_grids **out);
int mgic_grids_regrid(const mgic_params *P, double fill_ratio, int top_level, const int *nboxes, const int *boxes, const int *ntags,
                      const int *tags, mgic_grids **out);
int mgic_grids_destroy(mgic_grids *);
int mgic_grids_levels(const mgic_grids *);
int mgic_grids_num_boxes(const mgic_grids *, int level);
/* boxes: 6 ints each {lo0,lo1,lo2,hi0,hi1,hi2}; part_of_box: the connected part (0 .. *nparts-1) each box belongs to */
int mgic_grids_get_boxes(const mgic_grids *, int level, int *boxes, int *part_of_box, int *nparts);
int mgic_grids_level_stats(const mgic_grids *, int level, double *max_condition, long long *tagged_cells, long long *cells);
/* the problem object on these grids: every level > 0 as its connected parts, each in one masked array */
int mgic_hier_create_from_grids(mgic_ctx *, const mgic_params *P, const mgic_grids *, mgic_hier **out);
/* multigrid_vars of one AMR level > 0 and its psi update (the pieces mgic_hier composes) */
int mgic_vars_create_patch(mgic_ctx *, const mgic_params *P, const mgic_op *patch, mgic_vars **out);
int mgic_update_psi0_patch(mgic_vars *, mgic_op *patch, mgic_field *dpsi, const mgic_field *dpsi_coarse, const int coarse_lo[3]);

/* the NL loop of Main_PoissonSolver.cpp:131-216 for a single level, fully device resident */
int mgic_nl_solve(mgic_ctx *, const mgic_params *, double *dpsi_norms, int max_out, int *nl_iterations,
                  double *psi_out /* optional ghost-free host array */);

#ifdef __cplusplus
}
#endif
#endif /* MGIC_H */
