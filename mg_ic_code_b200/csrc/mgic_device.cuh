// mgic_device.cuh -- device helpers shared by the stencil kernels: the 7-point bracket of the reference's
// Fortran and the GSRB point update, written once so every kernel variant produces the same bits.
#ifndef MGIC_DEVICE_CUH
#define MGIC_DEVICE_CUH

#include "mgic_internal.h"

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

// Neighbour values of cell (i,j,k) (local indices) with the physical BC folded in.
struct Nb { double xm, xp, ym, yp, zm, zp; };

// [Chombo] AMRPoissonOpF.ChF INTERPHOMO: the parabola through pa (second interior cell), pb (first interior cell) and a
// zero coarse value, at the ghost cell; operation order of the Fortran (constants precomputed on the host the same way)
__device__ __forceinline__ double cf_homog(const double *cf, double pa, double pb) {
  const double a = ((pb - pa) * cf[0] - (pb)*cf[1]) * cf[2];
  const double b = (pb)*cf[3] - a * cf[4];
  return a * cf[6] + b * cf[5] + pa;
}

__device__ __forceinline__ double ghost(const BCk &bc, int f, double c, const double *p, long long wrapIdx, long long farIdx,
                                        int faceIdx) {
  // Dirichlet / Neumann: a*c + b  (DiriBC order 1: 2v - near;  NeumBC: near + sign*dx*v  [Chombo BCFunc])
  if (bc.type[f] == MGIC_BC_PERIODIC) return p[wrapIdx];
  if (bc.type[f] == MGIC_FACE_CF) return cf_homog(bc.cf, p[farIdx], c);
  if (bc.type[f] == MGIC_FACE_GHOST) return bc.face[f][faceIdx];
  return bc.a[f] * c + bc.b[f];
}

// masked AMR level: neighbour beyond face f of cell idx (see BCk::mask)
__device__ __forceinline__ double nb_masked(const BCk &bc, int f, bool inside, long long nidx, bool domainFace, double c, const double *p,
                                            long long farIdx, long long idx) {
  if (inside && bc.mask[nidx]) return p[nidx];               // a cell of the level: the fine-fine exchange
  if (domainFace) return bc.a[f] * c + bc.b[f];              // ParseBC
  if (bc.face[f]) return bc.face[f][idx];                    // QuadCFInterp's stored ghost
  return cf_homog(bc.cf, p[farIdx], c);                      // homogeneousCFInterp
}

// MASKED is a compile-time switch: the kernels of plain levels carry none of the masked branch (with a run-time test on
// bc.mask instead, k_restrict at 512^3 went from 0.62 to 0.84 ms: profiles/r1b_ vs r2m_launches_bench_512.csv)
template <bool MASKED = false>
__device__ __forceinline__ Nb neighbours(const double *p, long long idx, int i, int j, int k, const Geom &g,
                                         const BCk &bc, double c) {
  Nb n;
  if (MASKED) {
    n.xm = nb_masked(bc, 0, i > 0, idx - 1, bc.plo[0] + i == 0, c, p, idx + 1, idx);
    n.xp = nb_masked(bc, 1, i < g.nx - 1, idx + 1, bc.plo[0] + i == bc.ndom[0] - 1, c, p, idx - 1, idx);
    n.ym = nb_masked(bc, 2, j > 0, idx - g.sy, bc.plo[1] + j == 0, c, p, idx + g.sy, idx);
    n.yp = nb_masked(bc, 3, j < g.ny - 1, idx + g.sy, bc.plo[1] + j == bc.ndom[1] - 1, c, p, idx - g.sy, idx);
    n.zm = nb_masked(bc, 4, k > 0, idx - g.sz, bc.plo[2] + k == 0, c, p, idx + g.sz, idx);
    n.zp = nb_masked(bc, 5, k < g.nz - 1, idx + g.sz, bc.plo[2] + k == bc.ndom[2] - 1, c, p, idx - g.sz, idx);
    return n;
  }
  n.xm = (i > 0) ? p[idx - 1] : ghost(bc, 0, c, p, idx + (g.nx - 1), idx + 1, j + g.ny * k);
  n.xp = (i < g.nx - 1) ? p[idx + 1] : ghost(bc, 1, c, p, idx - (g.nx - 1), idx - 1, j + g.ny * k);
  n.ym = (j > 0) ? p[idx - g.sy] : ghost(bc, 2, c, p, idx + (long long)(g.ny - 1) * g.sy, idx + g.sy, i + g.nx * k);
  n.yp = (j < g.ny - 1) ? p[idx + g.sy] : ghost(bc, 3, c, p, idx - (long long)(g.ny - 1) * g.sy, idx - g.sy, i + g.nx * k);
  // z: ghost planes exist in memory; MGIC_FACE_INTERIOR means they hold the neighbour slab's planes
  n.zm = (k > 0 || bc.type[4] == MGIC_FACE_INTERIOR) ? p[idx - g.sz]
                                                    : ghost(bc, 4, c, p, idx + (long long)(g.nz - 1) * g.sz, idx + g.sz, i + g.nx * j);
  n.zp = (k < g.nz - 1 || bc.type[5] == MGIC_FACE_INTERIOR) ? p[idx + g.sz]
                                                           : ghost(bc, 5, c, p, idx - (long long)(g.nz - 1) * g.sz, idx - g.sz, i + g.nx * j);
  return n;
}

#endif
