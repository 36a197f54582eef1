// VariableCoeffPoissonOperator.cpp -- every method is a thin forward to the C ABI (include/mgic.h); see the header.
#include "VariableCoeffPoissonOperator.H"

#include "VariableCoeffPoissonOperatorFactory.H"

int AMRPoissonOp::s_relaxMode = 1;
int AMRPoissonOp::s_maxCoarse = 2;

// ---- coherence between a LevelData's host FABs and its device twin -------------------------------------------------
mgic_field *AMRPoissonOp::twinOn(const LevelData<FArrayBox> &ld, mgic_op *op) {
  CH_assert(ld.isDefined());
  std::shared_ptr<DeviceTwin> &tw = ld.twin();
  if (!tw) {
    tw = std::make_shared<DeviceTwin>();
    MGIC_CALL(mgic_field_create(op, &tw->field));
    // a LevelData that has never been touched on the host starts life on the device (zero filled)
    tw->fresh = ld.hostAllocated() ? DeviceTwin::HOST : DeviceTwin::DEVICE;
  }
  return tw->field;
}
mgic_field *AMRPoissonOp::devIn(const LevelData<FArrayBox> &ld) const {
  mgic_field *f = twinOn(ld, m_op);
  if (ld.twin()->fresh == DeviceTwin::HOST) { ld.upload(); ld.twin()->fresh = DeviceTwin::BOTH; }
  return f;
}
mgic_field *AMRPoissonOp::devOut(LevelData<FArrayBox> &ld) const {
  mgic_field *f = twinOn(ld, m_op);
  ld.twin()->fresh = DeviceTwin::DEVICE;
  return f;
}
mgic_field *AMRPoissonOp::devInOut(LevelData<FArrayBox> &ld) const {
  mgic_field *f = devIn(ld);
  ld.twin()->fresh = DeviceTwin::DEVICE;
  return f;
}

// ---- [Chombo] AMRPoissonOp members the reference inherits ---------------------------------------------------------------
void AMRPoissonOp::create(LevelData<FArrayBox> &a_lhs, const LevelData<FArrayBox> &a_rhs) {
  a_lhs.define(a_rhs.disjointBoxLayout(), a_rhs.nComp(), a_rhs.ghostVect());
  devOut(a_lhs);
}
void AMRPoissonOp::createCoarser(LevelData<FArrayBox> &a_coarse, const LevelData<FArrayBox> &a_fine, bool a_ghosted) {
  CH_assert(m_coarserOp);
  DisjointBoxLayout dbl;
  coarsen(dbl, a_fine.disjointBoxLayout(), 2);  // multigrid, so coarsen by 2
  a_coarse.define(dbl, a_fine.nComp(), a_ghosted ? a_fine.ghostVect() : IntVect::Zero);
  twinOn(a_coarse, m_coarserOp);
  a_coarse.twin()->fresh = DeviceTwin::DEVICE;
}
void AMRPoissonOp::assign(LevelData<FArrayBox> &a_lhs, const LevelData<FArrayBox> &a_rhs) {
  mgic_field *x = devIn(a_rhs);
  MGIC_CALL(mgic_op_assign(m_op, devOut(a_lhs), x));
}
Real AMRPoissonOp::dotProduct(const LevelData<FArrayBox> &a_1, const LevelData<FArrayBox> &a_2) {
  Real v = 0;
  MGIC_CALL(mgic_op_dot(m_op, devIn(a_1), devIn(a_2), &v));
  return v;
}
void AMRPoissonOp::incr(LevelData<FArrayBox> &a_lhs, const LevelData<FArrayBox> &a_x, Real a_scale) {
  mgic_field *x = devIn(a_x);
  MGIC_CALL(mgic_op_incr(m_op, devInOut(a_lhs), x, a_scale));
}
void AMRPoissonOp::axby(LevelData<FArrayBox> &a_lhs, const LevelData<FArrayBox> &a_x, const LevelData<FArrayBox> &a_y, Real a, Real b) {
  mgic_field *x = devIn(a_x), *y = devIn(a_y);
  MGIC_CALL(mgic_op_axby(m_op, devOut(a_lhs), x, y, a, b));
}
void AMRPoissonOp::scale(LevelData<FArrayBox> &a_lhs, const Real &a_scale) { MGIC_CALL(mgic_op_scale(m_op, devInOut(a_lhs), a_scale)); }
Real AMRPoissonOp::norm(const LevelData<FArrayBox> &a_x, int a_ord) {
  Real v = 0;
  MGIC_CALL(mgic_op_norm(m_op, devIn(a_x), a_ord, &v));
  return v;
}
void AMRPoissonOp::setToZero(LevelData<FArrayBox> &a_x) { MGIC_CALL(mgic_op_set_to_zero(m_op, devOut(a_x))); }
void AMRPoissonOp::relax(LevelData<FArrayBox> &a_e, const LevelData<FArrayBox> &a_residual, int a_iterations) {
  for (int i = 0; i < a_iterations; i++) {
    switch (s_relaxMode) {
      case 1: levelGSRB(a_e, a_residual); break;
      case 4: levelJacobi(a_e, a_residual); break;
      default: MayDay::Abort("unrecognized relaxation mode");
    }
  }
}
void AMRPoissonOp::prolongIncrement(LevelData<FArrayBox> &a_phiThisLevel, const LevelData<FArrayBox> &a_correctCoarse) {
  CH_assert(m_coarserOp);
  mgic_field *c = twinOn(a_correctCoarse, m_coarserOp);
  if (a_correctCoarse.twin()->fresh == DeviceTwin::HOST) { a_correctCoarse.upload(); a_correctCoarse.twin()->fresh = DeviceTwin::BOTH; }
  MGIC_CALL(mgic_op_prolong_increment(m_op, devInOut(a_phiThisLevel), c));
}

// ---- VariableCoeffPoissonOperator (replaces Source/VariableCoeffPoissonOperator.cpp) --------------------------------------
void VariableCoeffPoissonOperator::residualI(LevelData<FArrayBox> &a_lhs, const LevelData<FArrayBox> &a_dpsi,
                                             const LevelData<FArrayBox> &a_rhs, bool a_homogeneous) {
  mgic_field *phi = devIn(a_dpsi), *rhs = devIn(a_rhs);
  MGIC_CALL(mgic_op_residual(m_op, devOut(a_lhs), phi, rhs, a_homogeneous ? 1 : 0));  // replaces :30-67
}
void VariableCoeffPoissonOperator::preCond(LevelData<FArrayBox> &a_correction, const LevelData<FArrayBox> &a_residual) {
  CH_assert(a_residual.nComp() == a_correction.nComp());
  resetLambda();
  mgic_field *r = devIn(a_residual);
  MGIC_CALL(mgic_op_precond(m_op, devOut(a_correction), r));  // replaces :72-104
}
void VariableCoeffPoissonOperator::applyOpI(LevelData<FArrayBox> &a_lhs, const LevelData<FArrayBox> &a_dpsi, bool a_homogeneous) {
  mgic_field *phi = devIn(a_dpsi);
  MGIC_CALL(mgic_op_apply(m_op, devOut(a_lhs), phi, a_homogeneous ? 1 : 0));  // replaces :106-121
}
void VariableCoeffPoissonOperator::applyOpNoBoundary(LevelData<FArrayBox> &a_lhs, const LevelData<FArrayBox> &a_dpsi) {
  mgic_field *phi = devIn(a_dpsi);
  MGIC_CALL(mgic_op_apply_no_boundary(m_op, devOut(a_lhs), phi));  // replaces :123-149
}
void VariableCoeffPoissonOperator::restrictResidual(LevelData<FArrayBox> &a_resCoarse, LevelData<FArrayBox> &a_dpsiFine,
                                                    const LevelData<FArrayBox> &a_rhsFine) {
  CH_assert(m_coarserOp);
  mgic_field *phi = devIn(a_dpsiFine), *rhs = devIn(a_rhsFine);
  mgic_field *rc = twinOn(a_resCoarse, m_coarserOp);
  a_resCoarse.twin()->fresh = DeviceTwin::DEVICE;
  MGIC_CALL(mgic_op_restrict_residual(m_op, rc, phi, rhs));  // replaces :151-194
}
void VariableCoeffPoissonOperator::setAlphaAndBeta(const Real &a_alpha, const Real &a_beta) {
  m_alpha = a_alpha; m_beta = a_beta;
  MGIC_CALL(mgic_op_set_alpha_beta(m_op, a_alpha, a_beta));
  m_lambdaNeedsResetting = true;  // :204-205
}
void VariableCoeffPoissonOperator::setCoefs(const RefCountedPtr<LevelData<FArrayBox>> &a_aCoef,
                                            const RefCountedPtr<LevelData<FArrayBox>> &a_bCoef, const Real &a_alpha,
                                            const Real &a_beta) {
  m_alpha = a_alpha; m_beta = a_beta; m_aCoef = a_aCoef; m_bCoef = a_bCoef;
  MGIC_CALL(mgic_op_set_coefs(m_op, devIn(*a_aCoef), devIn(*a_bCoef), a_alpha, a_beta));
  m_lambdaNeedsResetting = true;  // :216-217
}
void VariableCoeffPoissonOperator::resetLambda() {
  if (m_lambdaNeedsResetting) {
    MGIC_CALL(mgic_op_compute_lambda(m_op));  // replaces :220-249
    m_lambdaNeedsResetting = false;
  }
}
void VariableCoeffPoissonOperator::computeLambda() {
  m_lambdaNeedsResetting = true;
  resetLambda();  // replaces :252-260
}
void VariableCoeffPoissonOperator::levelGSRB(LevelData<FArrayBox> &a_dpsi, const LevelData<FArrayBox> &a_rhs) {
  CH_assert(a_dpsi.isDefined() && a_rhs.isDefined());
  CH_assert(a_dpsi.nComp() == a_rhs.nComp());
  resetLambda();
  mgic_field *r = devIn(a_rhs);
  MGIC_CALL(mgic_op_relax(m_op, devInOut(a_dpsi), r, 1));  // one red+black sweep, replaces :273-332
}
void VariableCoeffPoissonOperator::levelJacobi(LevelData<FArrayBox> &a_dpsi, const LevelData<FArrayBox> &a_rhs) {
  resetLambda();
  mgic_field *r = devIn(a_rhs);
  MGIC_CALL(mgic_op_level_jacobi(m_op, devInOut(a_dpsi), r));  // replaces :360-385
}
