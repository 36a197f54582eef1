// mgic_device.cuh -- device helpers shared by the stencil kernels: the 7-point bracket of the reference's
// Fortran and the GSRB point update, written once so every kernel variant produces the same bits.
#ifndef MGIC_DEVICE_CUH
#define MGIC_DEVICE_CUH

// The CHF_DTERM block common to the four operator kernels (VariableCoeffPoissonOperatorF.ChF:111-120,
// 219-228, 322-330, 415-424): each bracket left to right, brackets added in x, y, z order.
__device__ __forceinline__ double lap7(double c, double xm, double xp, double ym, double yp, double zm, double zp) {
  const double t = 2.0 * c;
  return ((xp + xm) - t) + ((yp + ym) - t) + ((zp + zm) - t);
}

// GSRBHELMHOLTZVC3D point update (VariableCoeffPoissonOperatorF.ChF:107-128)
template <bool HAS_B>
__device__ __forceinline__ double gsrb_point(double c, double xm, double xp, double ym, double yp, double zm, double zp, double av,
                                             double bv, double lv, double rv, double alpha, double beta, double dxinv) {
  double lof = alpha * av * c;              // :107-108
  double l = lap7(c, xm, xp, ym, yp, zm, zp);  // :111-120
  l = l * dxinv;                            // :122  (ldpsi*dxinv)*bCoef
  if (HAS_B) l = l * bv;
  lof = lof - beta * l;                     // :124
  return c - lv * (lof - rv);               // :127-128
}
#endif
