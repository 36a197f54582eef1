// VariableCoeffPoissonOperatorFactory.cpp -- see the header.  Replaces Source/VariableCoeffPoissonOperatorFactory.cpp.
#include "VariableCoeffPoissonOperatorFactory.H"
#include "HierarchySession.H"

#include <cstring>

AMRLevelOpFactory<LevelData<FArrayBox>> *defineOperatorFactory(const Vector<DisjointBoxLayout> &a_grids,
                                                               const Vector<ProblemDomain> &a_vectDom,
                                                               Vector<RefCountedPtr<LevelData<FArrayBox>>> &a_aCoef,
                                                               Vector<RefCountedPtr<LevelData<FArrayBox>>> &a_bCoef,
                                                               const PoissonParameters &a_params) {
  VariableCoeffPoissonOperatorFactory *opFactory = new VariableCoeffPoissonOperatorFactory;
  opFactory->maxGridSize = a_params.maxGridSize;
  opFactory->define(a_params.coarsestDomain, a_grids, a_params.refRatio, a_params.coarsestDx, BCHolder::fromParmParse(),
                    a_params.alpha, a_aCoef, a_params.beta, a_bCoef);
  if (a_params.coefficient_average_type >= 0) opFactory->m_coefficient_average_type = a_params.coefficient_average_type;
  (void)a_vectDom;
  return (AMRLevelOpFactory<LevelData<FArrayBox>> *)opFactory;
}

void VariableCoeffPoissonOperatorFactory::define(const ProblemDomain &a_coarseDomain, const Vector<DisjointBoxLayout> &a_grids,
                                                 const Vector<int> &a_refRatios, const Real &a_coarsedx, BCHolder a_bc,
                                                 const Real &a_alpha, Vector<RefCountedPtr<LevelData<FArrayBox>>> &a_aCoef,
                                                 const Real &a_beta, Vector<RefCountedPtr<LevelData<FArrayBox>>> &a_bCoef) {
  setDefaultValues();
  m_hierarchyMode = false;
  if (a_grids.size() != 1) {
    // a hierarchy: the operators of every level live in the session's mgic_hier (HierarchySession.H); this factory is the
    // token MultilevelLinearOp::define receives
    if (!hierarchySession().active()) MayDay::Error("VariableCoeffPoissonOperatorFactory (B200): an AMR hierarchy needs the grids set_grids produced");
    m_hierarchyMode = true;
    return;
  }
  m_boxes = a_grids;
  m_refRatios = a_refRatios;
  m_bc = a_bc;
  m_domains.assign(1, a_coarseDomain);
  m_dx.assign(1, a_coarsedx);
  m_alpha = a_alpha; m_beta = a_beta;
  m_aCoef = a_aCoef; m_bCoef = a_bCoef;
  m_hier.reset();
}

std::shared_ptr<DeviceHierarchy> VariableCoeffPoissonOperatorFactory::hierarchy() {
  if (m_hier) return m_hier;
  if (!m_dev) m_dev = defaultDevice();
  auto h = std::make_shared<DeviceHierarchy>();
  h->dev = m_dev;
  h->aCoef = m_aCoef[0]; h->bCoef = m_bCoef[0];
  mgic_params &P = h->P;
  memset(&P, 0, sizeof(P));
  const ProblemDomain &dom = m_domains[0];
  for (int d = 0; d < 3; d++) {
    P.N[d] = dom.domainBox().size(d);
    P.bc_lo[d] = m_bc.bc_lo[d]; P.bc_hi[d] = m_bc.bc_hi[d];
  }
  P.alpha = m_alpha; P.beta = m_beta; P.bc_value = m_bc.bc_value;
  P.L = m_dx[0] * P.N[0];
  P.is_periodic = dom.isPeriodic() ? 1 : 0;
  P.max_level = 0;
  // the depth limit is a property of the box lattice (Factory.cpp:168-172): largest size all boxes are coarsenable by
  int mgs = maxGridSize > 0 ? maxGridSize : m_boxes[0].boxArray()[0].size(0);
  for (const Box &b : m_boxes[0].boxArray())
    for (int d = 0; d < 3; d++)
      if (b.size(d) != mgs || b.smallEnd(d) % mgs != 0) MayDay::Error("B200 operator: boxes must form the uniform domainSplit lattice of max_grid_size");
  P.max_grid_size = mgs;
  P.block_factor = mgs;
  P.coefficient_average_type = m_coefficient_average_type;
  P.numMGsmooth = m_numSmooth; P.numMGIterations = 1; P.preCondSolverDepth = m_maxDepth;
  P.tolerance = 1e-7; P.max_iterations = 100; P.max_NL_iterations = 1;
  // level-0 coefficients: upload through a geometry-only operator, then build the hierarchy on them
  mgic_op *lvl = nullptr;
  int lo[3], hi[3];
  for (int d = 0; d < 3; d++) { lo[d] = P.is_periodic ? MGIC_BC_PERIODIC : P.bc_lo[d]; hi[d] = P.is_periodic ? MGIC_BC_PERIODIC : P.bc_hi[d]; }
  MGIC_CALL(mgic_op_create(m_dev->ctx, P.N, 0, P.N[2], m_dx[0], m_alpha, m_beta, lo, hi, P.bc_value, &lvl));
  mgic_field *a = AMRPoissonOp::twinOn(*m_aCoef[0], lvl), *b = AMRPoissonOp::twinOn(*m_bCoef[0], lvl);
  for (auto *ld : {m_aCoef[0].get(), m_bCoef[0].get()})
    if (ld->twin()->fresh == DeviceTwin::HOST) { ld->upload(); ld->twin()->fresh = DeviceTwin::BOTH; }
  mgic_op_destroy(lvl);
  MGIC_CALL(mgic_mg_create(m_dev->ctx, &P, a, b, &h->mg));
  m_hier = h;
  return m_hier;
}

// replaces Factory.cpp:139-234: the depth loop, coefficient coarsening and lambda already ran in mgic_mg_create
MGLevelOp<LevelData<FArrayBox>> *VariableCoeffPoissonOperatorFactory::MGnewOp(const ProblemDomain &a_indexSpace, int a_depth,
                                                                               bool /*a_homoOnly*/) {
  if (!(a_indexSpace == m_domains[0])) MayDay::Abort("No corresponding AMRLevel to starting point of MGnewOp");  // Factory.cpp:153
  std::shared_ptr<DeviceHierarchy> h = hierarchy();
  mgic_op *op = nullptr, *coarser = nullptr;
  MGIC_CALL(mgic_mg_op(h->mg, a_depth, &op));
  if (!op) return NULL;  // cannot coarsen further (:168-172)
  MGIC_CALL(mgic_mg_op(h->mg, a_depth + 1, &coarser));
  VariableCoeffPoissonOperator *newOp = new VariableCoeffPoissonOperator;
  newOp->m_op = op; newOp->m_coarserOp = coarser; newOp->m_hier = h;
  const int coarsening = 1 << a_depth;
  newOp->m_domain = coarsen(m_domains[0], coarsening);
  newOp->m_dx = m_dx[0] * coarsening; newOp->m_dxCrse = 2 * newOp->m_dx;
  newOp->m_alpha = m_alpha; newOp->m_beta = m_beta; newOp->m_bc = m_bc;
  if (a_depth == 0) { newOp->m_aCoef = m_aCoef[0]; newOp->m_bCoef = m_bCoef[0]; }  // :194-197
  return (MGLevelOp<LevelData<FArrayBox>> *)newOp;
}

AMRLevelOp<LevelData<FArrayBox>> *VariableCoeffPoissonOperatorFactory::AMRnewOp(const ProblemDomain &a_indexSpace) {
  return (AMRLevelOp<LevelData<FArrayBox>> *)MGnewOp(a_indexSpace, 0, true);
}

int VariableCoeffPoissonOperatorFactory::refToFiner(const ProblemDomain &a_domain) const {
  for (size_t i = 0; i < m_domains.size(); i++)
    if (m_domains[i] == a_domain) return i < m_refRatios.size() ? m_refRatios[i] : 1;
  MayDay::Abort("Domain not found in AMR hierarchy");  // Factory.cpp:311-313
}
