/*
 * mgic_oracle.h -- C interface of the CPU ORACLE.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library, and only as the checker / the timed
 * CPU arm.  The product (mg_ic_code_b200/) never links, imports or executes it.
 *
 * PARITY.  The reference's own C++ compiles unmodified against a stand-in for
 * the Chombo API it touches (oracle/_ref, Makefile target `ref`:
 * SetLevelData.cpp + SetBinaryBH.H + MyPhiFunction.H, PoissonParameters.cpp,
 * SetBCs.cpp, VariableCoeffPoissonOperator.cpp,
 * VariableCoeffPoissonOperatorFactory.cpp), and this oracle reproduces it BIT FOR
 * BIT (tests/test_reference_pins.py, tests/golden/reference_sources_16.npz):
 *   PINNED to the reference's code: the source terms, the
 *   parameter block, ParseBC's dispatch (a14), the operator class's orchestration
 *   -- levelGSRB, residualI, applyOpI, restrictResidual, preCond, levelJacobi,
 *   lambda (a6 - a12) -- and the factory's MG depth limit and coefficient
 *   coarsening (a17).
 *   PINNED through a mechanical translation: the six .ChF kernels.  There is no
 *   Fortran compiler, so oracle/chf2c.py translates the reference's .ChF files to
 *   C++ at build time (syntax only, every expression copied through); the
 *   reference's classes run on that translation, and this file's restatement of
 *   the kernels is bit-identical to it.
 *   UNPINNED: everything that is Chombo 3.2's, which is not vendored under
 *   /root/reference and is restated from its published algorithm: DiriBC/NeumBC,
 *   exchange, CoarseAverage, FORT_PROLONG, MultiGrid::cycle,
 *   BiCGStabSolver::solve, AMRMultiGrid, MultilevelLinearOp, QuadCFInterp.  The
 *   reference ships no tests and no golden vectors.  Pins there: the known-answer
 *   tests in tests/ (trivial KAT, trace-free KAT, manufactured solution,
 *   decomposition invariance) and an independent numpy twin (tests/np_twin.py).
 */
#ifndef MGIC_ORACLE_H
#define MGIC_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors PoissonParameters (Source/PoissonParameters.H:20-63) plus the solver
 * knobs Main_PoissonSolver.cpp:106-126 and Source/SetBCs.cpp:42-58 read. */
typedef struct orc_params {
  double alpha, beta;
  double G_Newton, phi_amplitude, phi_wavelength;
  double bh1_bare_mass, bh1_spin, bh1_momentum, bh1_offset;
  double bh2_bare_mass, bh2_spin, bh2_momentum, bh2_offset;
  double L;           /* domain length of direction 0; dx = L / N[0]       */
  double bc_value;
  double tolerance;
  int N[3];
  int max_level;
  int block_factor, max_grid_size;
  int coefficient_average_type; /* 0 arithmetic, 1 harmonic (CoarseAverage enum) */
  int is_periodic;
  int bc_lo[3], bc_hi[3];       /* 0 Dirichlet, 1 Neumann, 2 periodic      */
  int numMGsmooth, numMGIterations, preCondSolverDepth;
  int max_iterations, max_NL_iterations;
  int verbosity;
} orc_params;

/* ---- kernel level: Fortran calling convention of the .ChF routines ------
 * FRA  = ptr, lo0,lo1,lo2, hi0,hi1,hi2, ncomp      (all by pointer)
 * FRA1 = ptr, lo0,lo1,lo2, hi0,hi1,hi2
 * BOX  = lo0,lo1,lo2, hi0,hi1,hi2                                          */
#define ORC_FRA(a)  double *a, const int *a##lo0, const int *a##lo1, const int *a##lo2, \
                    const int *a##hi0, const int *a##hi1, const int *a##hi2, const int *a##nc
#define ORC_CFRA(a) const double *a, const int *a##lo0, const int *a##lo1, const int *a##lo2, \
                    const int *a##hi0, const int *a##hi1, const int *a##hi2, const int *a##nc
#define ORC_FRA1(a)  double *a, const int *a##lo0, const int *a##lo1, const int *a##lo2, \
                     const int *a##hi0, const int *a##hi1, const int *a##hi2
#define ORC_CFRA1(a) const double *a, const int *a##lo0, const int *a##lo1, const int *a##lo2, \
                     const int *a##hi0, const int *a##hi1, const int *a##hi2
#define ORC_BOX(b)  const int *b##lo0, const int *b##lo1, const int *b##lo2, \
                    const int *b##hi0, const int *b##hi1, const int *b##hi2

void orc_gsrbhelmholtzvc3d(ORC_FRA(dpsi), ORC_CFRA(rhs), ORC_BOX(region), const double *dx,
                           const double *alpha, ORC_CFRA(aCoef), const double *beta,
                           ORC_CFRA(bCoef), ORC_CFRA(lambda), const int *redBlack);
void orc_vccomputeop3d(ORC_FRA(lofdpsi), ORC_CFRA(dpsi), const double *alpha, ORC_CFRA(aCoef),
                       const double *beta, ORC_CFRA(bCoef), ORC_BOX(region), const double *dx);
void orc_vccomputeres3d(ORC_FRA(res), ORC_CFRA(dpsi), ORC_CFRA(rhs), const double *alpha,
                        ORC_CFRA(aCoef), const double *beta, ORC_CFRA(bCoef), ORC_BOX(region),
                        const double *dx);
void orc_restrictresvc3d(ORC_FRA(res), ORC_CFRA(dpsi), ORC_CFRA(rhs), const double *alpha,
                         ORC_CFRA(aCoef), const double *beta, ORC_CFRA(bCoef), ORC_BOX(region),
                         const double *dx);
void orc_getlaplacianpsif(ORC_FRA1(lap), ORC_CFRA1(psi), const double *dx, ORC_BOX(box));
void orc_getrhogradphif(ORC_FRA1(rho), ORC_CFRA1(phi), const double *dx, ORC_BOX(box));
void orc_prolong(ORC_FRA(phi), ORC_CFRA(coarse), ORC_BOX(region), const int *m);

/* ---- problem level -------------------------------------------------------
 * Field ids for orc_get_field / orc_set_field (global, ghost-free arrays in
 * Fortran order: idx = i + nx*(j + ny*k)).                                   */
enum {
  ORC_F_E = 0,      /* MG correction e[depth]   (depth 0: the vector being solved for) */
  ORC_F_R = 1,      /* MG residual / rhs r[depth]                                      */
  ORC_F_A = 2,      /* aCoef[depth]                                                    */
  ORC_F_B = 3,      /* bCoef[depth]                                                    */
  ORC_F_LAMBDA = 4, /* lambda[depth]                                                   */
  ORC_F_TMP = 5,    /* scratch, output of residual/applyOp                             */
  ORC_F_DPSI = 6,   /* level-0 only: dpsi                                              */
  ORC_F_RHS = 7,    /* level-0 only: rhs from set_rhs                                  */
  ORC_F_MGVAR0 = 16 /* + comp (0..7): multigrid_vars component                          */
};

typedef struct orc_problem orc_problem;

orc_problem *orc_create(const orc_params *p);
void orc_destroy(orc_problem *);
int orc_num_threads(void);
void orc_set_num_threads(int n);   /* overrides OMP_NUM_THREADS (bench.py reference arm under torchrun) */

/* Main_PoissonSolver.cpp:79-95 */
void orc_set_initial_conditions(orc_problem *);
/* Main_PoissonSolver.cpp:154-160: set_a_coef, set_b_coef, set_rhs */
void orc_set_coefs_and_rhs(orc_problem *, double constant_K);
/* Main_PoissonSolver.cpp:163-170: factory + MG hierarchy (MGnewOp until NULL);
 * returns the number of MG levels (depths). */
int orc_define_solver(orc_problem *);
int orc_mg_depths(const orc_problem *);
void orc_level_dims(const orc_problem *, int depth, int n[3], double *dx);

void orc_get_field(orc_problem *, int depth, int field, double *out);
void orc_set_field(orc_problem *, int depth, int field, const double *in);
/* ghosted read (ng ghost layers, array extents n+2*ng) of e/dpsi/mgvars */
void orc_get_field_ghosted(orc_problem *, int depth, int field, int ng, double *out);

/* operator methods on MG depth d (reference: VariableCoeffPoissonOperator.cpp) */
void orc_op_relax(orc_problem *, int depth, int iterations);           /* e, r            */
void orc_op_gsrb_color(orc_problem *, int depth, int whichPass);       /* one colour pass */
void orc_op_residual(orc_problem *, int depth, int homogeneous);       /* tmp = r - L e   */
void orc_op_apply(orc_problem *, int depth, int homogeneous);          /* tmp = L e       */
void orc_op_restrict(orc_problem *, int depth);                        /* r[d+1] from e[d], r[d] */
void orc_op_prolong(orc_problem *, int depth);                         /* e[d] += P e[d+1] */
void orc_op_precond(orc_problem *, int depth);                         /* e = lambda*r; relax 2 */
double orc_op_norm(orc_problem *, int depth, int field, int ord);
double orc_op_dot(orc_problem *, int depth, int field1, int field2);

/* MultiGrid::cycle(0, e[0], r[0]) incl. bottom solve; returns bottom BiCGStab iterations */
int orc_vcycle(orc_problem *);
/* bottom solver alone on depth = last: BiCGStab(e, r); returns iterations */
int orc_bottom_solve(orc_problem *);
/* copy level-0 rhs into r[0], zero e[0] */
void orc_load_rhs_zero_e(orc_problem *);

/* f1: solver.solve(dpsi, rhs): outer BiCGStab preconditioned by numMGIterations V-cycles.
 * Returns iterations; fills exit status and final residual norm. */
int orc_outer_solve(orc_problem *, int *exit_status, double *final_norm, double *norms, int max_norms);
/* Main_PoissonSolver.cpp:189-205 + computeNorm :208 ; returns dpsi_norm */
double orc_update_psi0(orc_problem *);
/* full NL loop (Main_PoissonSolver.cpp:131-216); dpsi_norms[NL_iter]; returns #NL iterations */
/* set_regrid_condition (mode 0) / set_constant_K_integrand (mode 1) on fresh initial data over the index box [lo, hi] */
void orc_condition_box(const orc_params *, double dx, const int lo[3], const int hi[3], int mode, double *out);
/* set_output_data: the 32 GRChombo variables over [lo, hi] from the given psi over the same box */
void orc_output_box(const orc_params *, double dx, const int lo[3], const int hi[3], const double *psi, double constant_K, double *out);
void orc_set_dpsi_with_bc(orc_problem *, const double *valid_cells);   /* then orc_update_psi0 */
int orc_nl_solve(orc_problem *, double *dpsi_norms, int max_out);

/* ---- one AMR level > 0 (a box of the refined domain, split into max_grid_size boxes): what the reference's operator
 * class does there itself -- levelGSRB and restrictResidual with [Chombo] homogeneousCFInterp
 * (VariableCoeffPoissonOperator.cpp:156,296).  Fields: ORC_F_E, ORC_F_R, ORC_F_A, ORC_F_B, ORC_F_LAMBDA as patch-shaped
 * arrays (x fastest); ORC_F_TMP = the restricted residual on the patch coarsened by 2. */
typedef struct orc_patch orc_patch;
orc_patch *orc_patch_create(const int n_domain[3], const int lo[3], const int hi[3], int max_grid_size, double dx, double dx_crse,
                            double alpha, double beta, const int bc_lo[3], const int bc_hi[3], double bc_value);
/* the level as a list of boxes (nboxes x {lo0,lo1,lo2,hi0,hi1,hi2}) that may touch; fields travel as arrays over the bounding box */
orc_patch *orc_patch_create_boxes(const int n_domain[3], int nboxes, const int *boxes, int max_grid_size, double dx, double dx_crse,
                                  double alpha, double beta, const int bc_lo[3], const int bc_hi[3], double bc_value);
void orc_patch_destroy(orc_patch *);
void orc_patch_set(orc_patch *, int field, const double *in);
void orc_patch_get(orc_patch *, int field, double *out);
/* the nonlinear loop's per-level steps on an AMR level > 0 (Main_PoissonSolver.cpp:93,154-160,189-205) */
void orc_patch_set_initial_conditions(orc_patch *, const orc_params *);
void orc_patch_set_coefs_and_rhs(orc_patch *, double constant_K);
void orc_patch_update_psi(orc_patch *, const double *dpsi_bounding_box);
void orc_patch_get_var(orc_patch *, int comp /* 0..7 multigrid_vars, 8 rhs */, double *out);
int orc_patch_num_boxes(const orc_patch *);
void orc_patch_relax(orc_patch *, int iterations);
void orc_patch_gsrb_color(orc_patch *, int whichPass);
void orc_patch_restrict(orc_patch *);
void orc_patch_precond(orc_patch *);
/* [Chombo] QuadCFInterp + AMRPoissonOp::AMROperatorNF / AMRResidualNF: the coarser level's field over its whole domain
 * (n_domain / 2); the result is read back as field ORC_F_RHS */
void orc_patch_set_coarse(orc_patch *, const double *in);
void orc_patch_amr_operator_nf(orc_patch *, int homogeneous_phys_bc);
void orc_patch_amr_residual_nf(orc_patch *, int homogeneous_phys_bc);
double orc_interp_homo(double dx, double dx_crse, double far_value, double near_value);

#ifdef __cplusplus
}
#endif
#endif
