// ref_harness.cpp -- TEST INFRASTRUCTURE.  C entry points around the REFERENCE's own code: this file is linked with the
// reference's Source/SetLevelData.cpp (which includes SetBinaryBH.H and MyPhiFunction.H), PoissonParameters.cpp, SetBCs.cpp,
// VariableCoeffPoissonOperator.cpp and VariableCoeffPoissonOperatorFactory.cpp, compiled unmodified where they lie against
// oracle/ref_shim/chombo_standin.H, and with the mechanical C++ translation of its two .ChF kernel files (oracle/chf2c.py).
// Nothing in the resulting library comes from the oracle (mgic_oracle.h is included for its field ids only), so
// tests/test_reference_pins.py compares two independent things: the reference's own code and the oracle's restatement.
// Not the reference's: Chombo (stand-in, restated from the published algorithms) and Fortran's place in the tool chain
// (translated, not compiled).
#include "SetLevelData.H"   // the reference's prototypes (Source/SetLevelData.H:27-71)
#include "SetBCs.H"         // ParseBC, GlobalBCRS (Source/SetBCs.H)
#include "VariableCoeffPoissonOperatorFactory.H"

#include "mgic_oracle.h"

#include <cstdio>

int AMRPoissonOp::s_relaxMode = 1;
int AMRPoissonOp::s_maxCoarse = 2;
const IntVect IntVect::Unit(1, 1, 1), IntVect::Zero(0, 0, 0);
const RealVect RealVect::Unit(1.0, 1.0, 1.0), RealVect::Zero(0.0, 0.0, 0.0);

// defined in Source/SetBinaryBH.H / MyPhiFunction.H (non-inline, compiled in SetLevelData.cpp's translation unit)
Real get_bh_radius(RealVect &loc_bh, const Real bh_x_offset);
Real get_Aij(const int i, const int j, const Real &rbh1, const Real &rbh2, const RealVect &n1, const RealVect &n2,
             const RealVect &J1, const RealVect &J2, const RealVect &P1, const RealVect &P2, const PoissonParameters &a_params);
Real set_binary_bh_psi(const RealVect &loc, const PoissonParameters &a_params);
Real my_phi_function(RealVect loc, Real amplitude, Real wavelength, RealVect L);

// The Fortran symbols the reference's generated prototypes declare (SetLevelDataF_F.H, VariableCoeffPoissonOperatorF_F.H) --
// getlaplacianpsif_, getrhogradphif_, gsrbhelmholtzvc3d_, vccomputeop3d_, vccomputeres3d_, restrictresvc3d_ -- are defined
// in oracle/_ref/gen/*.cpp, which chf2c.py generates from the reference's .ChF files at build time.

// ---- the reference's operator class (Source/VariableCoeffPoissonOperator.{H,cpp}, compiled unmodified) on one level that
// covers its domain, split into boxes of max_grid_size; physical BCs by the reference's ParseBC (Source/SetBCs.cpp, compiled
// unmodified) reading bc_lo / bc_hi / bc_value from the ParmParse table.  AMRPoissonOp::define is Chombo's, so the harness
// sets the members define would set.
class RefOp : public VariableCoeffPoissonOperator {
public:
  DisjointBoxLayout grids, coarseGrids;
  LevelData<FArrayBox> e, r, tmp, resCoarse;
  IntVect n;
  RefOp(const int N[3], int maxGrid, Real dx, Real alpha, Real beta) : n(N[0], N[1], N[2]) {
    std::vector<Box> boxes, cboxes;
    for (int k = 0; k < N[2]; k += maxGrid)
      for (int j = 0; j < N[1]; j += maxGrid)
        for (int i = 0; i < N[0]; i += maxGrid) {
          Box b(IntVect(i, j, k), IntVect(std::min(i + maxGrid, N[0]) - 1, std::min(j + maxGrid, N[1]) - 1, std::min(k + maxGrid, N[2]) - 1));
          boxes.push_back(b);
          cboxes.push_back(coarsen(b, 2));
        }
    grids = DisjointBoxLayout(boxes);
    coarseGrids = DisjointBoxLayout(cboxes);
    m_dx = dx; m_dxCrse = 2 * dx;
    m_domain = ProblemDomain(Box(IntVect::Zero, IntVect(N[0] - 1, N[1] - 1, N[2] - 1)));
    m_bc = BCHolder(ParseBC);
    RefCountedPtr<LevelData<FArrayBox>> a(new LevelData<FArrayBox>(grids, 1, IntVect::Zero));
    RefCountedPtr<LevelData<FArrayBox>> b(new LevelData<FArrayBox>(grids, 1, IntVect::Zero));
    setCoefs(a, b, alpha, beta);
    e.define(grids, 1, IntVect::Unit);
    r.define(grids, 1, IntVect::Zero);
    tmp.define(grids, 1, IntVect::Zero);
    resCoarse.define(coarseGrids, 1, IntVect::Zero);
    computeLambda();
  }
  LevelData<FArrayBox> *field(int which) {
    switch (which) {
      case ORC_F_E: return &e;
      case ORC_F_R: return &r;
      case ORC_F_A: return &*m_aCoef;
      case ORC_F_B: return &*m_bCoef;
      case ORC_F_LAMBDA: return &m_lambda;
      case ORC_F_TMP: return &tmp;
      default: return nullptr;
    }
  }
};

static void level_copy(LevelData<FArrayBox> &ld, double *out, const double *in, const IntVect &n) {
  for (DataIterator dit = ld.dataIterator(); dit.ok(); ++dit) {
    const Box &b = ld.disjointBoxLayout()[dit()];
    for (BoxIterator bit(b); bit.ok(); ++bit) {
      const IntVect &iv = bit();
      const size_t q = iv[0] + (size_t)n[0] * (iv[1] + (size_t)n[1] * iv[2]);
      if (in) ld[dit](iv, 0) = in[q];
      else out[q] = ld[dit](iv, 0);
    }
  }
}

extern "C" {

// bcs: bc_lo[3], bc_hi[3] (0 Dirichlet, 1 Neumann), bc_value -- handed to ParseBC / ParseValue through the ParmParse table
void *ref_op_create(const int N[3], int max_grid_size, double dx, double alpha, double beta, const int bc_lo[3], const int bc_hi[3],
                    double bc_value) {
  char line[128];
  std::snprintf(line, sizeof line, "bc_lo = %d %d %d", bc_lo[0], bc_lo[1], bc_lo[2]); ParmParse::standin_add_line(line);
  std::snprintf(line, sizeof line, "bc_hi = %d %d %d", bc_hi[0], bc_hi[1], bc_hi[2]); ParmParse::standin_add_line(line);
  std::snprintf(line, sizeof line, "bc_value = %.17g", bc_value); ParmParse::standin_add_line(line);
  GlobalBCRS::s_areBCsParsed = false;      // ParseBC caches bc_lo / bc_hi in statics (SetBCs.cpp:53-58)
  return new RefOp(N, max_grid_size, dx, alpha, beta);
}
void ref_op_destroy(void *h) { delete (RefOp *)h; }
int ref_op_num_boxes(void *h) { return ((RefOp *)h)->grids.size(); }
// fields as global ghost-free arrays (first index fastest); ids as in mgic_oracle.h (ORC_F_E, _R, _A, _B, _LAMBDA, _TMP)
int ref_op_set(void *h, int field, const double *in) {
  RefOp *o = (RefOp *)h;
  LevelData<FArrayBox> *ld = o->field(field);
  if (!ld || field == ORC_F_LAMBDA) return 1;
  level_copy(*ld, nullptr, in, o->n);
  if (field == ORC_F_A || field == ORC_F_B) o->setCoefs(o->m_aCoef, o->m_bCoef, o->m_alpha, o->m_beta);   // "if you change this call resetLambda()"
  return 0;
}
int ref_op_get(void *h, int field, double *out) {
  RefOp *o = (RefOp *)h;
  LevelData<FArrayBox> *ld = o->field(field);
  if (!ld) return 1;
  if (field == ORC_F_LAMBDA) o->resetLambda();
  level_copy(*ld, out, nullptr, o->n);
  return 0;
}
void ref_op_get_coarse_residual(void *h, double *out) {
  RefOp *o = (RefOp *)h;
  level_copy(o->resCoarse, out, nullptr, IntVect(o->n[0] / 2, o->n[1] / 2, o->n[2] / 2));
}
void ref_op_relax(void *h, int iterations, int relax_mode) {       // AMRPoissonOp::relax -> levelGSRB (1) / levelJacobi (4)
  RefOp *o = (RefOp *)h;
  AMRPoissonOp::s_relaxMode = relax_mode;
  o->relax(o->e, o->r, iterations);
  AMRPoissonOp::s_relaxMode = 1;
}
// same, reporting a MayDay::Abort / Error instead of unwinding through C: 0 ok, 1 aborted (message in msg)
int ref_op_relax_status_msg(void *h, int iterations, int relax_mode, char *msg, int msglen) {
  try {
    ref_op_relax(h, iterations, relax_mode);
    return 0;
  } catch (const std::exception &e) {
    AMRPoissonOp::s_relaxMode = 1;
    if (msg && msglen > 0) std::snprintf(msg, msglen, "%s", e.what());
    return 1;
  }
}
int ref_op_relax_status(void *h, int iterations, int relax_mode) { return ref_op_relax_status_msg(h, iterations, relax_mode, nullptr, 0); }
void ref_op_residual(void *h, int homogeneous) { RefOp *o = (RefOp *)h; o->residualI(o->tmp, o->e, o->r, homogeneous != 0); }
void ref_op_apply(void *h, int homogeneous) { RefOp *o = (RefOp *)h; o->applyOpI(o->tmp, o->e, homogeneous != 0); }
void ref_op_restrict(void *h) { RefOp *o = (RefOp *)h; o->restrictResidual(o->resCoarse, o->e, o->r); }
void ref_op_precond(void *h) { RefOp *o = (RefOp *)h; o->preCond(o->e, o->r); }

// ---- the reference's factory (Source/VariableCoeffPoissonOperatorFactory.cpp, compiled unmodified) through the function
// Main_PoissonSolver.cpp calls, defineOperatorFactory (:163-166), on one AMR level: how deep MGnewOp goes, and the
// coefficients / lambda / dx of every depth.  coefficient_average_type as PoissonParameters holds it (-1: not in the input).
struct RefFactory {
  AMRLevelOpFactory<LevelData<FArrayBox>> *f = nullptr;
  ProblemDomain domain;
  Vector<RefCountedPtr<LevelData<FArrayBox>>> a, b;
  std::vector<VariableCoeffPoissonOperator *> ops;     // depth 0, 1, ... until MGnewOp returned NULL
  IntVect n;
  ~RefFactory() { for (auto *o : ops) delete o; delete f; }
};
void *ref_factory_create(const int N[3], int max_grid_size, double dx, double alpha, double beta, int coefficient_average_type,
                         const double *aCoef, const double *bCoef) {
  RefFactory *F = new RefFactory;
  F->n = IntVect(N[0], N[1], N[2]);
  std::vector<Box> boxes;
  for (int k = 0; k < N[2]; k += max_grid_size)
    for (int j = 0; j < N[1]; j += max_grid_size)
      for (int i = 0; i < N[0]; i += max_grid_size)
        boxes.push_back(Box(IntVect(i, j, k), IntVect(std::min(i + max_grid_size, N[0]) - 1, std::min(j + max_grid_size, N[1]) - 1,
                                                     std::min(k + max_grid_size, N[2]) - 1)));
  Vector<DisjointBoxLayout> grids(1, DisjointBoxLayout(boxes));
  F->domain = ProblemDomain(Box(IntVect::Zero, IntVect(N[0] - 1, N[1] - 1, N[2] - 1)));
  Vector<ProblemDomain> domains(1, F->domain);
  F->a.push_back(RefCountedPtr<LevelData<FArrayBox>>(new LevelData<FArrayBox>(grids[0], 1, IntVect::Zero)));
  F->b.push_back(RefCountedPtr<LevelData<FArrayBox>>(new LevelData<FArrayBox>(grids[0], 1, IntVect::Zero)));
  level_copy(*F->a[0], nullptr, aCoef, F->n);
  level_copy(*F->b[0], nullptr, bCoef, F->n);
  PoissonParameters p;
  p.coarsestDomain = F->domain; p.coarsestDx = dx; p.alpha = alpha; p.beta = beta;
  p.refRatio.resize(1); p.refRatio.assign(2);
  p.coefficient_average_type = coefficient_average_type;
  F->f = defineOperatorFactory(grids, domains, F->a, F->b, p);
  for (int depth = 0;; depth++) {                         // what MultiGrid::define does
    MGLevelOp<LevelData<FArrayBox>> *op = F->f->MGnewOp(F->domain, depth, true);
    if (!op) break;
    F->ops.push_back(dynamic_cast<VariableCoeffPoissonOperator *>(op));
  }
  return F;
}
void ref_factory_destroy(void *h) { delete (RefFactory *)h; }
int ref_factory_depths(void *h) { return (int)((RefFactory *)h)->ops.size(); }
int ref_factory_average_type(void *h) { return dynamic_cast<VariableCoeffPoissonOperatorFactory *>(((RefFactory *)h)->f)->m_coefficient_average_type; }
// depth's cells per direction, dx, number of boxes
void ref_factory_level(void *h, int depth, int n[3], double *dx, int *boxes) {
  VariableCoeffPoissonOperator *o = ((RefFactory *)h)->ops[depth];
  const Box &d = o->standin_domain().domainBox();
  for (int q = 0; q < 3; q++) n[q] = d.hi[q] - d.lo[q] + 1;
  *dx = o->standin_dx();
  *boxes = o->standin_grids().size();
}
// field: ORC_F_A, ORC_F_B, ORC_F_LAMBDA of a depth, as a global array
int ref_factory_get(void *h, int depth, int field, double *out) {
  VariableCoeffPoissonOperator *o = ((RefFactory *)h)->ops[depth];
  LevelData<FArrayBox> *ld = field == ORC_F_A ? &*o->m_aCoef : field == ORC_F_B ? &*o->m_bCoef : field == ORC_F_LAMBDA ? &o->m_lambda : nullptr;
  if (!ld) return 1;
  const Box &d = o->standin_domain().domainBox();
  level_copy(*ld, out, nullptr, IntVect(d.hi[0] + 1, d.hi[1] + 1, d.hi[2] + 1));
  return 0;
}

typedef struct {
  double G_Newton, phi_amplitude, phi_wavelength;
  double bh1_bare_mass, bh1_spin, bh1_momentum, bh1_offset;
  double bh2_bare_mass, bh2_spin, bh2_momentum, bh2_offset;
  double L[3];
} ref_params;

static PoissonParameters to_params(const ref_params *p) {
  PoissonParameters q;
  q.G_Newton = p->G_Newton; q.phi_amplitude = p->phi_amplitude; q.phi_wavelength = p->phi_wavelength;
  q.bh1_bare_mass = p->bh1_bare_mass; q.bh1_spin = p->bh1_spin; q.bh1_momentum = p->bh1_momentum; q.bh1_offset = p->bh1_offset;
  q.bh2_bare_mass = p->bh2_bare_mass; q.bh2_spin = p->bh2_spin; q.bh2_momentum = p->bh2_momentum; q.bh2_offset = p->bh2_offset;
  q.domainLength = RealVect(p->L[0], p->L[1], p->L[2]);
  q.alpha = 1.0; q.beta = -1.0;
  return q;
}

static void copy_out(const FArrayBox &f, int comp, double *out) {
  if (!out) return;
  const Real *s = f.dataPtr(comp);
  for (long q = 0; q < f.box().numPts(); q++) out[q] = s[q];
}

// One level, one box of N cells, the reference's ghost widths (Main_PoissonSolver.cpp:79-88: multigrid_vars and dpsi 3
// ghosts, rhs / aCoef / bCoef none).  set_initial_conditions; then, if dpsi_ghosted is given ((N+6)^3, first index
// fastest), set_update_psi0 with it; then set_a_coef, set_b_coef, set_rhs (the order of Main_PoissonSolver.cpp:154-160).
// Outputs (any may be NULL): mgvars 8 x (N+6)^3, rhs / acoef / bcoef N^3.
int ref_set_level_data(const ref_params *p, const int N[3], double dx, double constant_K, const double *dpsi_ghosted,
                       double *mgvars, double *rhs, double *acoef, double *bcoef) {
  PoissonParameters params = to_params(p);
  Box dom(IntVect(0, 0, 0), IntVect(N[0] - 1, N[1] - 1, N[2] - 1));
  const IntVect g3(3, 3, 3);
  LevelData<FArrayBox> vars(dom, NUM_MULTIGRID_VARS, g3), dpsi(dom, 1, g3);
  LevelData<FArrayBox> r(dom, 1, IntVect::Zero), a(dom, 1, IntVect::Zero), b(dom, 1, IntVect::Zero);
  RealVect vdx(dx, dx, dx);
  set_initial_conditions(vars, dpsi, vdx, params);
  DataIterator dit = vars.dataIterator();
  dit.begin();
  if (dpsi_ghosted) {
    FArrayBox &d = dpsi[dit()];
    Real *dst = d.dataPtr(0);
    for (long q = 0; q < d.box().numPts(); q++) dst[q] = dpsi_ghosted[q];
    Copier none;
    set_update_psi0(vars, dpsi, none);
  }
  set_a_coef(a, vars, params, vdx, constant_K);
  set_b_coef(b, params, vdx);
  set_rhs(r, vars, vdx, params, constant_K);
  if (mgvars)
    for (int c = 0; c < NUM_MULTIGRID_VARS; c++) copy_out(vars[dit()], c, mgvars + (size_t)c * vars[dit()].box().numPts());
  copy_out(r[dit()], 0, rhs);
  copy_out(a[dit()], 0, acoef);
  copy_out(b[dit()], 0, bcoef);
  return 0;
}

// set_regrid_condition (mode 0, Source/SetLevelData.cpp:188-240) / set_constant_K_integrand (mode 1, :128-186) of the
// reference on freshly initialised data (what set_grids does, Source/SetGrids.cpp:86-95); one box of N cells; out N^3
int ref_condition(const ref_params *p, const int N[3], double dx, int mode, double *out) {
  PoissonParameters params = to_params(p);
  Box dom(IntVect(0, 0, 0), IntVect(N[0] - 1, N[1] - 1, N[2] - 1));
  const IntVect g3(3, 3, 3);
  LevelData<FArrayBox> vars(dom, NUM_MULTIGRID_VARS, g3), dpsi(dom, 1, g3), cond(dom, 1, IntVect::Zero);
  RealVect vdx(dx, dx, dx);
  set_initial_conditions(vars, dpsi, vdx, params);
  if (mode == 0) set_regrid_condition(cond, vars, vdx, params);
  else set_constant_K_integrand(cond, vars, vdx, params);
  DataIterator dit = cond.dataIterator();
  dit.begin();
  copy_out(cond[dit()], 0, out);
  return 0;
}

// set_output_data (Source/SetLevelData.cpp:343-396): the 32 GRChombo variables from multigrid_vars after psi += dpsi_ghosted
// (may be NULL); one box of N cells with the checkpoint's three ghost layers (Source/WriteOutput.H:180); out 32 x (N+6)^3
int ref_output_data(const ref_params *p, const int N[3], double dx, double constant_K, const double *dpsi_ghosted, double *out) {
  PoissonParameters params = to_params(p);
  Box dom(IntVect(0, 0, 0), IntVect(N[0] - 1, N[1] - 1, N[2] - 1));
  const IntVect g3(3, 3, 3);
  LevelData<FArrayBox> vars(dom, NUM_MULTIGRID_VARS, g3), dpsi(dom, 1, g3), gr(dom, NUM_GRCHOMBO_VARS, g3);
  RealVect vdx(dx, dx, dx);
  set_initial_conditions(vars, dpsi, vdx, params);
  DataIterator dit = vars.dataIterator();
  dit.begin();
  if (dpsi_ghosted) {
    FArrayBox &d = dpsi[dit()];
    Real *dst = d.dataPtr(0);
    for (long q = 0; q < d.box().numPts(); q++) dst[q] = dpsi_ghosted[q];
    Copier none;
    set_update_psi0(vars, dpsi, none);
  }
  set_output_data(gr, vars, params, vdx, constant_K);
  for (int c = 0; c < NUM_GRCHOMBO_VARS; c++) copy_out(gr[dit()], c, out + (size_t)c * gr[dit()].box().numPts());
  return 0;
}

// the point functions of Source/SetBinaryBH.H and MyPhiFunction.H at one location (relative to the domain centre)
void ref_point_values(const ref_params *p, const double loc[3], double Aij[6], double *psi_bh, double *phi) {
  PoissonParameters params = to_params(p);
  RealVect x(loc[0], loc[1], loc[2]);
  RealVect l1 = x, l2 = x;
  Real r1 = get_bh_radius(l1, params.bh1_offset), r2 = get_bh_radius(l2, params.bh2_offset);
  RealVect n1(l1[0] / r1, l1[1] / r1, l1[2] / r1), n2(l2[0] / r2, l2[1] / r2, l2[2] / r2);
  RealVect J1(0.0, 0.0, params.bh1_spin), J2(0.0, 0.0, params.bh2_spin);
  RealVect P1(0.0, params.bh1_momentum, 0.0), P2(0.0, params.bh2_momentum, 0.0);
  const int ij[6][2] = {{0, 0}, {0, 1}, {0, 2}, {1, 1}, {1, 2}, {2, 2}};   // c_A11_0 .. c_A33_0 (MultigridUserVariables.hpp)
  for (int q = 0; q < 6; q++) Aij[q] = get_Aij(ij[q][0], ij[q][1], r1, r2, n1, n2, J1, J2, P1, P2, params);
  *psi_bh = set_binary_bh_psi(x, params);
  *phi = my_phi_function(x, params.phi_amplitude, params.phi_wavelength, params.domainLength);
}

// set_m_value (Source/SetLevelData.cpp:266-279)
double ref_m_value(const ref_params *p, double phi_here, double constant_K) {
  PoissonParameters params = to_params(p);
  Real m = 0;
  set_m_value(m, phi_here, params, constant_K);
  return m;
}

// getPoissonParameters (Source/PoissonParameters.cpp:26-131, compiled unmodified) on an input file in params.txt's format
// plus `key = value` override lines (the command line of Main_PoissonSolver.cpp:272).  Returns 0, or 1 with the
// MayDay::Error / ParmParse message in err.
typedef struct {
  int nCells[3], maxGridSize, blockFactor, bufferSize, coefficient_average_type, verbosity, periodic[3], maxLevel, numLevels;
  int refRatio0, refRatioLast, domainLo[3], domainHi[3], domainPeriodic[3];
  double fillRatio, refineThresh, coarsestDx, domainLength[3], probLo[3], probHi[3], alpha, beta, G_Newton;
  double phi_amplitude, phi_wavelength, bh1_bare_mass, bh2_bare_mass, bh1_spin, bh2_spin, bh1_momentum, bh2_momentum;
  double bh1_offset, bh2_offset;
} ref_poisson_parameters;

int ref_get_poisson_parameters(const char *file, int noverrides, const char *const *overrides, ref_poisson_parameters *o,
                               char *err, int errlen) {
  try {
    if (!ParmParse::standin_load(file)) throw std::runtime_error(std::string("cannot open ") + file);
    for (int q = 0; q < noverrides; q++) ParmParse::standin_add_line(overrides[q]);
    PoissonParameters p;
    getPoissonParameters(p);
    for (int d = 0; d < 3; d++) {
      o->nCells[d] = p.nCells[d]; o->periodic[d] = p.periodic[d];
      o->domainLo[d] = p.coarsestDomain.domainBox().lo[d]; o->domainHi[d] = p.coarsestDomain.domainBox().hi[d];
      o->domainPeriodic[d] = p.coarsestDomain.isPeriodic(d);
      o->domainLength[d] = p.domainLength[d]; o->probLo[d] = p.probLo[d]; o->probHi[d] = p.probHi[d];
    }
    o->maxGridSize = p.maxGridSize; o->blockFactor = p.blockFactor; o->bufferSize = p.bufferSize;
    o->coefficient_average_type = p.coefficient_average_type; o->verbosity = p.verbosity;
    o->maxLevel = p.maxLevel; o->numLevels = p.numLevels;
    o->refRatio0 = p.refRatio.empty() ? -1 : p.refRatio.front(); o->refRatioLast = p.refRatio.empty() ? -1 : p.refRatio.back();
    o->fillRatio = p.fillRatio; o->refineThresh = p.refineThresh; o->coarsestDx = p.coarsestDx;
    o->alpha = p.alpha; o->beta = p.beta; o->G_Newton = p.G_Newton;
    o->phi_amplitude = p.phi_amplitude; o->phi_wavelength = p.phi_wavelength;
    o->bh1_bare_mass = p.bh1_bare_mass; o->bh2_bare_mass = p.bh2_bare_mass; o->bh1_spin = p.bh1_spin; o->bh2_spin = p.bh2_spin;
    o->bh1_momentum = p.bh1_momentum; o->bh2_momentum = p.bh2_momentum; o->bh1_offset = p.bh1_offset; o->bh2_offset = p.bh2_offset;
    return 0;
  } catch (const std::exception &e) {
    if (err && errlen > 0) { std::snprintf(err, errlen, "%s", e.what()); }
    return 1;
  }
}

}  // extern "C"
