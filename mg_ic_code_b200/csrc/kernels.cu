// kernels.cu -- sm_100a CUDA kernels of the multigrid hot path (level-wide, device-resident fields).
//
// Every kernel keeps the operation order of the reference's Fortran and the file is
// compiled with -fmad=false, so results are bit-identical to a non-contracting CPU evaluation.
//
// Data layout: one FP64 array per level field, x fastest, no ghost cells in x/y (rows stay 16/32-byte
// aligned for vector loads), MGIC_GZ ghost planes below and above the rank's z-slab.  Physical boundary
// ghosts are never stored: the value the reference's ParseBC (Source/SetBCs.cpp:49-131) would have written
// into the ghost cell, ghost = a*near + b, is recomputed on the fly from the cell's own current value.
#include "mgic_internal.h"
#include "mgic_device.cuh"

namespace {

// ---- launch bookkeeping ---------------------------------------------------------------------------------
inline int post_launch(mgic_ctx *c, const char *what) {
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    mgic_set_error("kernel %s: %s", what, cudaGetErrorString(e));
    return MGIC_ERR_CUDA;
  }
  return MGIC_OK;
}

// ---- GSRB colour pass: GSRBHELMHOLTZVC3D (VariableCoeffPoissonOperatorF.ChF:56-139) ------------------------
// One thread per cell of the colour: i = 2t + parity so that (i + j + k_global + color) is even (:98-106).
template <bool HAS_B, bool MASKED>
__global__ void __launch_bounds__(256) k_gsrb_color(Geom g, BCk bc, double *__restrict__ phi, const double *__restrict__ rhs,
                                                    const double *__restrict__ a, const double *__restrict__ b,
                                                    const double *__restrict__ lam, double alpha, double beta,
                                                    double dxinv, int color) {
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= g.ny) return;
  const int i = 2 * t + ((j + k + g.k0 + color) & 1);
  if (i >= g.nx) return;
  const long long idx = i + j * g.sy + k * g.sz;
  if (MASKED && !bc.mask[idx]) return;   // masked AMR level: not a cell of the level's boxes
  const double c = phi[idx];
  const Nb n = neighbours<MASKED>(phi, idx, i, j, k, g, bc, c);
  phi[idx] = gsrb_point<HAS_B>(c, n.xm, n.xp, n.ym, n.yp, n.zm, n.zp, a[idx], HAS_B ? b[idx] : 1.0, lam[idx], rhs[idx], alpha, beta,
                               dxinv);
}

// ---- applyOp / residual: VCCOMPUTEOP3D (:181-237), VCCOMPUTERES3D (:283-339) --------------------------------
// MODE 0: lhs = alpha*a*phi - S*dxinv*beta*b ; MODE 1: lhs = (rhs - alpha*a*phi) + S*dxinv*beta*b
template <int MODE, bool HAS_B, bool MASKED>
__global__ void __launch_bounds__(256) k_op(Geom g, BCk bc, double *__restrict__ lhs, const double *__restrict__ phi,
                                            const double *__restrict__ rhs, const double *__restrict__ a,
                                            const double *__restrict__ b, double alpha, double beta, double dxinv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i >= g.nx || j >= g.ny) return;
  const long long idx = i + j * g.sy + k * g.sz;
  if (MASKED && !bc.mask[idx]) return;
  const double c = phi[idx];
  const Nb n = neighbours<MASKED>(phi, idx, i, j, k, g, bc, c);
  double l = lap7(c, n.xm, n.xp, n.ym, n.yp, n.zm, n.zp);
  l = l * dxinv * beta;                                  // :227 / :331  ((ldpsi*dxinv)*beta)*bCoef
  if (HAS_B) l = l * b[idx];
  if (MODE == 0) {
    lhs[idx] = alpha * a[idx] * c - l;                   // :211-212, :229
  } else {
    lhs[idx] = (rhs[idx] - alpha * a[idx] * c) + l;      // :314-316, :333
  }
}

// ---- restrictResidual: RESTRICTRESVC3D (:379-437) ------------------------------------------------------------
// One thread per coarse cell; the eight fine contributions are accumulated in the order the Fortran loop
// nest delivers them to that coarse cell: (0,0,0),(1,0,0),(0,1,0),(1,1,0),(0,0,1),... starting from the
// zero the caller stored (VariableCoeffPoissonOperator.cpp:177).
// (Measured alternative: 16-byte loads of the x-pairs with the outer x-neighbours by warp shuffle, 20 vector loads per
// coarse cell instead of 72 scalar ones -- bit-identical, but slower: 0.80 vs 0.75 ms per V-cycle at 512^3.)
template <bool HAS_B, bool MASKED>
__global__ void __launch_bounds__(128) k_restrict(Geom g, BCk bc, double *__restrict__ resC, long long csy, long long csz,
                                                  const double *__restrict__ phi, const double *__restrict__ rhs,
                                                  const double *__restrict__ a, const double *__restrict__ b, double alpha,
                                                  double beta, double dxinv) {
  const int I = blockIdx.x * blockDim.x + threadIdx.x;
  const int J = blockIdx.y * blockDim.y + threadIdx.y;
  const int K = blockIdx.z;
  if (2 * I >= g.nx || 2 * J >= g.ny) return;
  if (MASKED && !bc.mask[2 * I + 2 * J * g.sy + 2 * K * g.sz]) return;   // boxes are coarsenable: all eight cells or none
  double acc = 0.0;
#pragma unroll
  for (int dk = 0; dk < 2; dk++)
#pragma unroll
    for (int dj = 0; dj < 2; dj++)
#pragma unroll
      for (int di = 0; di < 2; di++) {
        const int i = 2 * I + di, j = 2 * J + dj, k = 2 * K + dk;
        const long long idx = i + j * g.sy + k * g.sz;
        const double c = phi[idx];
        const Nb n = neighbours<MASKED>(phi, idx, i, j, k, g, bc, c);
        double lof = alpha * a[idx] * c;                       // :411-412
        double l = lap7(c, n.xm, n.xp, n.ym, n.yp, n.zm, n.zp);
        l = l * dxinv * beta;                                  // :427
        if (HAS_B) l = l * b[idx];
        lof = lof - l;                                         // :429
        acc = acc + (rhs[idx] - lof) / 8.0;                    // :431-432
      }
  resC[I + J * csy + K * csz] = acc;
}

// ---- prolongIncrement: [Chombo] AMRPoissonOpF.ChF PROLONG, m = 2 ---------------------------------------------
__global__ void __launch_bounds__(128) k_prolong(Geom g, double *__restrict__ phi, const double *__restrict__ coarse,
                                                 long long csy, long long csz, const unsigned char *__restrict__ fineMask) {
  const int I = blockIdx.x * blockDim.x + threadIdx.x;
  const int J = blockIdx.y * blockDim.y + threadIdx.y;
  const int K = blockIdx.z;
  if (2 * I >= g.nx || 2 * J >= g.ny) return;
  if (fineMask && !fineMask[2 * I + 2 * J * g.sy + 2 * K * g.sz]) return;
  const double c = coarse[I + J * csy + K * csz];
#pragma unroll
  for (int dk = 0; dk < 2; dk++)
#pragma unroll
    for (int dj = 0; dj < 2; dj++) {
      double2 *q = reinterpret_cast<double2 *>(phi + 2 * I + (2 * J + dj) * g.sy + (2 * K + dk) * g.sz);
      double2 v = *q;
      v.x = v.x + c;
      v.y = v.y + c;
      *q = v;
    }
}

// ---- [Chombo] QuadCFInterp::coarseFineInterp on one coarse-fine face of an AMR patch, ratio 2 -------------------------
// One thread per ghost cell of the face.  phiStar = the coarse field taken to the ghost cell's tangential position by a
// second-order Taylor expansion (centred differences, one-sided next to a domain face, mixed term dropped where a
// diagonal coarse cell is outside the domain), then QuadCFInterpF.ChF QUADINTERP: the parabola through the two interior
// fine cells and phiStar.  Operation order of the oracle (Op::quadCFInterp).
struct QcfArgs {
  Geom g;                 // the patch
  int plo[3], cdom[3];    // patch lower corner (fine level index space), coarse domain size
  int clo[3];             // coarse-level index of coarse[0]
  long long csy, csz;
  int dir, side, ta, tb;  // tangential directions ta < tb
  double h;
};
__device__ __forceinline__ void quad_cf_face_cell(const QcfArgs &A, int qa, int qb, const double *__restrict__ phi,
                                                  const double *__restrict__ coarse, double *__restrict__ face) {
  // geometry per direction without dynamically indexed local arrays: (value for x, y, z) selected by compile-time-simple tests
  const int dir = A.dir, ta = A.ta, tb = A.tb;
  const int nA = ta == 0 ? A.g.nx : A.g.ny;                  // ta is 0 or 1
  const int nB = tb == 1 ? A.g.ny : A.g.nz;                  // tb is 1 or 2
  const int nD = dir == 0 ? A.g.nx : (dir == 1 ? A.g.ny : A.g.nz);
  if (qa >= nA || qb >= nB) return;
  const long long fsA = ta == 0 ? 1 : A.g.sy, fsB = tb == 1 ? A.g.sy : A.g.sz, fsD = dir == 0 ? 1 : (dir == 1 ? A.g.sy : A.g.sz);
  const long long csA = ta == 0 ? 1 : A.csy, csB = tb == 1 ? A.csy : A.csz, csD = dir == 0 ? 1 : (dir == 1 ? A.csy : A.csz);
  const int ploA = ta == 0 ? A.plo[0] : A.plo[1], ploB = tb == 1 ? A.plo[1] : A.plo[2];
  const int ploD = dir == 0 ? A.plo[0] : (dir == 1 ? A.plo[1] : A.plo[2]);
  const int cloA = ta == 0 ? A.clo[0] : A.clo[1], cloB = tb == 1 ? A.clo[1] : A.clo[2];
  const int cloD = dir == 0 ? A.clo[0] : (dir == 1 ? A.clo[1] : A.clo[2]);
  const int cdA = ta == 0 ? A.cdom[0] : A.cdom[1], cdB = tb == 1 ? A.cdom[1] : A.cdom[2];
  const double h = A.h, H = 2.0 * A.h;
  // the ghost cell in the level's index space and the coarse cell that contains it
  const int fA = ploA + qa, fB = ploB + qb, fD = ploD + (A.side < 0 ? -1 : nD);
  const int cA = fA >> 1, cB = fB >> 1, cD = fD >> 1;
  const double *cc = coarse + ((cA - cloA) * csA + (cB - cloB) * csB + (cD - cloD) * csD);   // C(oa, ob) = cc[oa*csA + ob*csB]
  const double c0 = cc[0];
  double phistar = c0;
  // tangential direction ta
  const double xa = (fA + 0.5) * h - (cA + 0.5) * H;
  {
    const bool hasLo = cA - 1 >= 0, hasHi = cA + 1 <= cdA - 1;
    double d1, d2;
    if (hasLo && hasHi) {
      d1 = (cc[csA] - cc[-csA]) / (2.0 * H);
      d2 = ((cc[csA] - 2.0 * c0) + cc[-csA]) / (H * H);
    } else if (hasHi) {
      d1 = ((4.0 * cc[csA] - 3.0 * c0) - cc[2 * csA]) / (2.0 * H);
      d2 = ((c0 - 2.0 * cc[csA]) + cc[2 * csA]) / (H * H);
    } else {
      d1 = ((3.0 * c0 - 4.0 * cc[-csA]) + cc[-2 * csA]) / (2.0 * H);
      d2 = ((c0 - 2.0 * cc[-csA]) + cc[-2 * csA]) / (H * H);
    }
    phistar = phistar + (d1 * xa + 0.5 * d2 * xa * xa);
  }
  // tangential direction tb
  const double xb = (fB + 0.5) * h - (cB + 0.5) * H;
  {
    const bool hasLo = cB - 1 >= 0, hasHi = cB + 1 <= cdB - 1;
    double d1, d2;
    if (hasLo && hasHi) {
      d1 = (cc[csB] - cc[-csB]) / (2.0 * H);
      d2 = ((cc[csB] - 2.0 * c0) + cc[-csB]) / (H * H);
    } else if (hasHi) {
      d1 = ((4.0 * cc[csB] - 3.0 * c0) - cc[2 * csB]) / (2.0 * H);
      d2 = ((c0 - 2.0 * cc[csB]) + cc[2 * csB]) / (H * H);
    } else {
      d1 = ((3.0 * c0 - 4.0 * cc[-csB]) + cc[-2 * csB]) / (2.0 * H);
      d2 = ((c0 - 2.0 * cc[-csB]) + cc[-2 * csB]) / (H * H);
    }
    phistar = phistar + (d1 * xb + 0.5 * d2 * xb * xb);
  }
  const bool corners = cA - 1 >= 0 && cA + 1 <= cdA - 1 && cB - 1 >= 0 && cB + 1 <= cdB - 1;
  if (corners) {
    const double mixed = (((cc[csA + csB] - cc[csA - csB]) - cc[-csA + csB]) + cc[-csA - csB]) / (4.0 * H * H);
    phistar = phistar + mixed * xa * xb;
  }
  // QUADINTERP along the normal: pb = first interior cell, pa = second
  const long long nearIdx = qa * fsA + qb * fsB + (A.side < 0 ? 0 : (long long)(nD - 1) * fsD);
  const long long inward = A.side < 0 ? fsD : -fsD;
  const double pb = phi[nearIdx], pa = phi[nearIdx + inward];
  const double x = 2.0 * h;
  const double nref = 2.0;
  const double a = (2.0 / h / h) * ((2.0 * phistar + pa * (nref + 1.0)) - pb * (nref + 3.0)) / (nref * nref + 4.0 * nref + 3.0);
  const double b = (pb - pa) / h - a * h;
  face[qa + (long long)nA * qb] = (pa + b * x) + a * x * x;
}
__global__ void __launch_bounds__(256) k_quad_cf_face(QcfArgs A, const double *__restrict__ phi, const double *__restrict__ coarse,
                                                      double *__restrict__ face) {
  quad_cf_face_cell(A, blockIdx.x * blockDim.x + threadIdx.x, blockIdx.y * blockDim.y + threadIdx.y, phi, coarse, face);   // local tangential indices
}
// every coarse-fine face of a rectangular patch in ONE launch: blockIdx.z = face (0 x-lo ... 5 z-hi), faces without a
// coarse-fine boundary have a null array.  (Config C4 spent 2.5 of 22 ms per nonlinear iteration in 450 one-face launches.)
struct QcfFaces { double *face[6]; };
__global__ void __launch_bounds__(256) k_quad_cf_faces(QcfArgs A, QcfFaces F, const double *__restrict__ phi, const double *__restrict__ coarse) {
  const int f = blockIdx.z;
  if (!F.face[f]) return;
  A.dir = f >> 1; A.side = (f & 1) ? +1 : -1;
  A.ta = A.dir == 0 ? 1 : 0; A.tb = A.dir == 2 ? 1 : 2;
  quad_cf_face_cell(A, blockIdx.x * blockDim.x + threadIdx.x, blockIdx.y * blockDim.y + threadIdx.y, phi, coarse, F.face[f]);
}

// ---- the same interpolation for a masked AMR level (a union of boxes in one bounding-box array): one thread per cell of
// the level, one value per face whose neighbour is a coarse-fine ghost (neither a cell of the level nor outside the
// domain), stored cell-indexed: face[f][idx].  Arithmetic and its order are k_quad_cf_face's (the oracle's Op::quadCFInterp).
struct QcfmArgs {
  Geom g;
  int plo[3], ndom[3], clo[3];
  long long csy, csz;
  double h;
  double *face[6];
};
__device__ double quad_cf_eval(const QcfmArgs &A, int dir, int side, const int iv[3], long long idx, const double *__restrict__ phi,
                               const double *__restrict__ coarse) {
  const int ta = dir == 0 ? 1 : 0, tb = dir == 2 ? 1 : 2;
  const long long fs[3] = {1, A.g.sy, A.g.sz}, cs[3] = {1, A.csy, A.csz};
  const double h = A.h, H = 2.0 * A.h;
  int fg[3] = {A.plo[0] + iv[0], A.plo[1] + iv[1], A.plo[2] + iv[2]};   // the ghost cell in the level's index space
  fg[dir] += side;
  const int cA = fg[ta] >> 1, cB = fg[tb] >> 1, cD = fg[dir] >> 1;
  const int cdA = A.ndom[ta] / 2, cdB = A.ndom[tb] / 2;
  const long long csA = cs[ta], csB = cs[tb];
  const double *cc = coarse + ((cA - A.clo[ta]) * csA + (cB - A.clo[tb]) * csB + (cD - A.clo[dir]) * cs[dir]);
  const double c0 = cc[0];
  double phistar = c0;
  const double xa = (fg[ta] + 0.5) * h - (cA + 0.5) * H;
  {
    const bool hasLo = cA - 1 >= 0, hasHi = cA + 1 <= cdA - 1;
    double d1, d2;
    if (hasLo && hasHi) {
      d1 = (cc[csA] - cc[-csA]) / (2.0 * H);
      d2 = ((cc[csA] - 2.0 * c0) + cc[-csA]) / (H * H);
    } else if (hasHi) {
      d1 = ((4.0 * cc[csA] - 3.0 * c0) - cc[2 * csA]) / (2.0 * H);
      d2 = ((c0 - 2.0 * cc[csA]) + cc[2 * csA]) / (H * H);
    } else {
      d1 = ((3.0 * c0 - 4.0 * cc[-csA]) + cc[-2 * csA]) / (2.0 * H);
      d2 = ((c0 - 2.0 * cc[-csA]) + cc[-2 * csA]) / (H * H);
    }
    phistar = phistar + (d1 * xa + 0.5 * d2 * xa * xa);
  }
  const double xb = (fg[tb] + 0.5) * h - (cB + 0.5) * H;
  {
    const bool hasLo = cB - 1 >= 0, hasHi = cB + 1 <= cdB - 1;
    double d1, d2;
    if (hasLo && hasHi) {
      d1 = (cc[csB] - cc[-csB]) / (2.0 * H);
      d2 = ((cc[csB] - 2.0 * c0) + cc[-csB]) / (H * H);
    } else if (hasHi) {
      d1 = ((4.0 * cc[csB] - 3.0 * c0) - cc[2 * csB]) / (2.0 * H);
      d2 = ((c0 - 2.0 * cc[csB]) + cc[2 * csB]) / (H * H);
    } else {
      d1 = ((3.0 * c0 - 4.0 * cc[-csB]) + cc[-2 * csB]) / (2.0 * H);
      d2 = ((c0 - 2.0 * cc[-csB]) + cc[-2 * csB]) / (H * H);
    }
    phistar = phistar + (d1 * xb + 0.5 * d2 * xb * xb);
  }
  const bool corners = cA - 1 >= 0 && cA + 1 <= cdA - 1 && cB - 1 >= 0 && cB + 1 <= cdB - 1;
  if (corners) {
    const double mixed = (((cc[csA + csB] - cc[csA - csB]) - cc[-csA + csB]) + cc[-csA - csB]) / (4.0 * H * H);
    phistar = phistar + mixed * xa * xb;
  }
  const double pb = phi[idx], pa = phi[idx - side * fs[dir]];   // first / second interior cell along the normal
  const double x = 2.0 * h;
  const double nref = 2.0;
  const double a = (2.0 / h / h) * ((2.0 * phistar + pa * (nref + 1.0)) - pb * (nref + 3.0)) / (nref * nref + 4.0 * nref + 3.0);
  const double b = (pb - pa) / h - a * h;
  return (pa + b * x) + a * x * x;
}
__global__ void __launch_bounds__(256) k_quad_cf_masked(QcfmArgs A, const unsigned char *__restrict__ mask, const double *__restrict__ phi,
                                                        const double *__restrict__ coarse) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i >= A.g.nx || j >= A.g.ny) return;
  const long long idx = i + j * A.g.sy + k * A.g.sz;
  if (mask && !mask[idx]) return;   // mask == null: a rectangular level, every cell of the array
  const int iv[3] = {i, j, k}, n[3] = {A.g.nx, A.g.ny, A.g.nz};
  const long long fs[3] = {1, A.g.sy, A.g.sz};
  for (int f = 0; f < 6; f++) {
    const int dir = f >> 1, side = (f & 1) ? 1 : -1;
    const int q = iv[dir] + side;
    if (q >= 0 && q < n[dir] && (!mask || mask[idx + side * fs[dir]])) continue;      // a cell of the level
    const int gq = A.plo[dir] + q;
    if (gq < 0 || gq >= A.ndom[dir]) continue;                             // physical boundary
    A.face[f][idx] = quad_cf_eval(A, dir, side, iv, idx, phi, coarse);
  }
}

// ---- lambda: resetLambda (VariableCoeffPoissonOperator.cpp:220-249) ------------------------------------------
__global__ void k_lambda(long long n, double *__restrict__ lam, const double *__restrict__ a, double alpha, double plus) {
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
    double v = a[q];   // copy   :234
    v = v * alpha;     // mult   :235
    v = v + plus;      // plus   :241
    lam[q] = 1.0 / v;  // invert :244
  }
}

// ---- BLAS-1 over the contiguous valid cells of the slab ---------------------------------------------------
enum { EW_MULT, EW_INCR, EW_AXBY, EW_SCALE, EW_ASSIGN, EW_SETVAL, EW_JACOBI };
template <int OP>
__global__ void k_ew(long long n, double *__restrict__ y, const double *__restrict__ x1, const double *__restrict__ x2,
                     double s1, double s2) {
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
    if (OP == EW_MULT) y[q] = x1[q] * x2[q];
    else if (OP == EW_INCR) y[q] = y[q] + s1 * x1[q];
    else if (OP == EW_AXBY) y[q] = s1 * x1[q] + s2 * x2[q];
    else if (OP == EW_SCALE) y[q] = y[q] * s1;
    else if (OP == EW_ASSIGN) y[q] = x1[q];
    else if (OP == EW_SETVAL) y[q] = s1;
    else if (OP == EW_JACOBI) y[q] = y[q] + s1 * (x2[q] * x1[q]);  // phi += 0.5 * (lambda * resid)  (levelJacobi :375-381)
  }
}

// ---- reductions: warp shuffle -> block -> fixed-order final pass by the last block (deterministic) ----------
// kind 0 max|x|, 1 sum|x|, 2 sum x^2, 3 sum x*y, 4 count(x != s)
template <int KIND>
__device__ __forceinline__ double red_elem(double x, double y, double s) {
  if (KIND == 0 || KIND == 1) return fabs(x);
  if (KIND == 2) return x * x;
  if (KIND == 3) return x * y;
  return (x != s) ? 1.0 : 0.0;
}
template <int KIND>
__device__ __forceinline__ double red_comb(double a, double b) { return KIND == 0 ? fmax(a, b) : a + b; }

template <int KIND>
__device__ __forceinline__ double block_reduce(double v, double *sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = red_comb<KIND>(v, __shfl_down_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    v = (l < (blockDim.x >> 5)) ? sh[l] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = red_comb<KIND>(v, __shfl_down_sync(0xffffffffu, v, o));
  }
  return v;  // valid in thread 0
}

template <int KIND>
__global__ void __launch_bounds__(256) k_reduce(long long n, const double *__restrict__ x, const double *__restrict__ y, double s,
                                                double *__restrict__ part, unsigned int *__restrict__ count,
                                                double *__restrict__ out, const unsigned char *__restrict__ skip) {
  __shared__ double sh[32];
  __shared__ bool last;
  double v = 0.0;
  // skip[q] != 0: the cell counts as x = 0 (a cell a finer AMR level covers: the composite norms / dot products) -- the same
  // bits as reducing a copy with those cells zeroed
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x)
    v = red_comb<KIND>(v, red_elem<KIND>((skip && skip[q]) ? 0.0 : x[q], KIND == 3 ? y[q] : 0.0, s));
  v = block_reduce<KIND>(v, sh);
  if (threadIdx.x == 0) {
    part[blockIdx.x] = v;
    __threadfence();
    last = (atomicAdd(count, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    double w = 0.0;
    for (int q = threadIdx.x; q < (int)gridDim.x; q += blockDim.x) w = red_comb<KIND>(w, part[q]);
    __syncthreads();
    w = block_reduce<KIND>(w, sh);
    if (threadIdx.x == 0) {
      *out = w;
      *count = 0;
    }
  }
}

// ---- coefficient coarsening: [Chombo] CoarseAverage (AverageF.ChF AVERAGE / AVERAGEHARMONIC) -----------------
// refScale = 1/nRef^3; fine cells summed ii fastest; arithmetic: sum*refScale; harmonic: 1/(sum(1/f)*refScale)
__global__ void __launch_bounds__(128) k_coarse_average(Geom gc, double *__restrict__ c, const double *__restrict__ f,
                                                        long long fsy, long long fsz, int nref, int harmonic,
                                                        const unsigned char *__restrict__ fineMask) {
  const int I = blockIdx.x * blockDim.x + threadIdx.x;
  const int J = blockIdx.y * blockDim.y + threadIdx.y;
  const int K = blockIdx.z;
  if (I >= gc.nx || J >= gc.ny) return;
  if (fineMask && !fineMask[I * nref + (J * nref) * fsy + (long long)(K * nref) * fsz]) return;   // not under the fine level
  const double refScale = 1.0 / (double)(nref * nref * nref);
  double sum = 0.0;
  for (int kk = 0; kk < nref; kk++)
    for (int jj = 0; jj < nref; jj++)
      for (int ii = 0; ii < nref; ii++) {
        const double fv = f[(I * nref + ii) + (J * nref + jj) * fsy + (long long)(K * nref + kk) * fsz];
        sum = sum + (harmonic ? 1.0 / fv : fv);
      }
  c[I + J * gc.sy + K * gc.sz] = harmonic ? 1.0 / (sum * refScale) : sum * refScale;
}

// y = v on a sub-box (g.nx x g.ny x g.nz cells, strides g.sy / g.sz of the array that contains it): [Chombo]
// AMRPoissonOp::zeroCovered -- the cells of a level that a finer level covers -- for the composite norms and dot products
__global__ void __launch_bounds__(128) k_box_set(Geom g, double *__restrict__ y, double v, const unsigned char *__restrict__ fineMask,
                                                 long long msy, long long msz) {
  const int I = blockIdx.x * blockDim.x + threadIdx.x;
  const int J = blockIdx.y * blockDim.y + threadIdx.y;
  const int K = blockIdx.z;
  if (I >= g.nx || J >= g.ny) return;
  if (fineMask && !fineMask[2 * I + 2 * J * msy + 2 * K * msz]) return;   // ratio 2: the coarse cell is not under the fine level
  y[I + J * g.sy + K * g.sz] = v;
}

// the same on a byte array: the "covered by a finer level" marks the composite reductions skip
__global__ void __launch_bounds__(128) k_box_set_u8(Geom g, unsigned char *__restrict__ y, unsigned char v,
                                                    const unsigned char *__restrict__ fineMask, long long msy, long long msz) {
  const int I = blockIdx.x * blockDim.x + threadIdx.x;
  const int J = blockIdx.y * blockDim.y + threadIdx.y;
  const int K = blockIdx.z;
  if (I >= g.nx || J >= g.ny) return;
  if (fineMask && !fineMask[2 * I + 2 * J * msy + 2 * K * msz]) return;
  y[I + J * g.sy + K * g.sz] = v;
}

// dst box := src box (nx x ny x nz cells, each array with its own strides): the cells of a z-slab-distributed coarser level
// that lie under / around a replicated finer patch, on their way into the patch's staging array (capi.cu amr_coarse_source)
__global__ void __launch_bounds__(128) k_copy_box(int nx, int ny, const double *__restrict__ src, long long ssy, long long ssz,
                                                  double *__restrict__ dst, long long dsy, long long dsz) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i >= nx || j >= ny) return;
  dst[i + j * dsy + k * dsz] = src[i + j * ssy + k * ssz];
}

// y = 0 outside the mask (uploads and constant fills of a masked AMR level: cells outside its boxes stay zero)
__global__ void __launch_bounds__(256) k_apply_mask(long long n, double *__restrict__ y, const unsigned char *__restrict__ mask) {
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x)
    if (!mask[q]) y[q] = 0.0;
}

inline dim3 grid3(int nx, int ny, int nz, dim3 b) { return dim3((nx + b.x - 1) / b.x, (ny + b.y - 1) / b.y, nz); }
inline int ew_grid(mgic_ctx *c, long long n) {
  long long b = (n + 255) / 256;
  long long cap = (long long)c->numSMs * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

// ================================================================================================================
namespace mgk {

int gsrb_color(mgic_ctx *c, const Geom &g, const BCk &bc, double *phi, const double *rhs, const double *a, const double *b,
               const double *lam, double alpha, double beta, double dx, int color) {
  const double dxinv = 1.0 / (dx * dx);  // :89
  dim3 blk(64, 4, 1);
  dim3 grd = grid3((g.nx + 1) / 2, g.ny, g.nz, blk);
// K<PRE..., HAS_B, MASKED><<<...>>>ARGS: the four builds of a stencil kernel (bCoef stream or not, masked AMR level or not)
#define MGIC_UNPAREN(...) __VA_ARGS__
#define MGIC_STENCIL_LAUNCH(K, PRE, ARGS)                                                    \
  do {                                                                                       \
    if (bc.mask) {                                                                           \
      if (b) K<MGIC_UNPAREN PRE true, true><<<grd, blk, 0, c->stream>>> ARGS;                 \
      else K<MGIC_UNPAREN PRE false, true><<<grd, blk, 0, c->stream>>> ARGS;                  \
    } else {                                                                                 \
      if (b) K<MGIC_UNPAREN PRE true, false><<<grd, blk, 0, c->stream>>> ARGS;                \
      else K<MGIC_UNPAREN PRE false, false><<<grd, blk, 0, c->stream>>> ARGS;                 \
    }                                                                                        \
  } while (0)
  MGIC_STENCIL_LAUNCH(k_gsrb_color, (), (g, bc, phi, rhs, a, b, lam, alpha, beta, dxinv, color));
  return post_launch(c, "gsrb_color");
}

int apply_op(mgic_ctx *c, const Geom &g, const BCk &bc, double *lhs, const double *phi, const double *a, const double *b,
             double alpha, double beta, double dx) {
  const double dxinv = 1.0 / (dx * dx);
  dim3 blk(64, 4, 1);
  dim3 grd = grid3(g.nx, g.ny, g.nz, blk);
  MGIC_STENCIL_LAUNCH(k_op, (0,), (g, bc, lhs, phi, nullptr, a, b, alpha, beta, dxinv));
  return post_launch(c, "apply_op");
}

int residual(mgic_ctx *c, const Geom &g, const BCk &bc, double *res, const double *phi, const double *rhs, const double *a,
             const double *b, double alpha, double beta, double dx) {
  const double dxinv = 1.0 / (dx * dx);
  dim3 blk(64, 4, 1);
  dim3 grd = grid3(g.nx, g.ny, g.nz, blk);
  MGIC_STENCIL_LAUNCH(k_op, (1,), (g, bc, res, phi, rhs, a, b, alpha, beta, dxinv));
  return post_launch(c, "residual");
}

int restrict_res(mgic_ctx *c, const Geom &g, const BCk &bc, double *resC, long long csy, long long csz, const double *phi,
                 const double *rhs, const double *a, const double *b, double alpha, double beta, double dx) {
  const double dxinv = 1.0 / (dx * dx);
  dim3 blk(32, 4, 1);
  dim3 grd = grid3(g.nx / 2, g.ny / 2, g.nz / 2, blk);
  MGIC_STENCIL_LAUNCH(k_restrict, (), (g, bc, resC, csy, csz, phi, rhs, a, b, alpha, beta, dxinv));
  return post_launch(c, "restrict");
}

int quad_cf_face(mgic_ctx *c, const Geom &g, const int plo[3], const int ndom[3], double h, int dir, int side, const double *phi,
                 const double *coarse, long long csy, long long csz, const int clo[3], double *face) {
  QcfArgs A;
  A.g = g;
  for (int d = 0; d < 3; d++) { A.plo[d] = plo[d]; A.cdom[d] = ndom[d] / 2; A.clo[d] = clo[d]; }
  A.csy = csy; A.csz = csz;
  A.dir = dir; A.side = side;
  A.ta = dir == 0 ? 1 : 0; A.tb = dir == 2 ? 1 : 2;
  A.h = h;
  const int n[3] = {g.nx, g.ny, g.nz};
  dim3 blk(32, 8, 1);
  dim3 grd((n[A.ta] + blk.x - 1) / blk.x, (n[A.tb] + blk.y - 1) / blk.y, 1);
  k_quad_cf_face<<<grd, blk, 0, c->stream>>>(A, phi, coarse, face);
  return post_launch(c, "quad_cf_face");
}

int quad_cf_faces(mgic_ctx *c, const Geom &g, const int plo[3], const int ndom[3], double h, const double *phi, const double *coarse,
                  long long csy, long long csz, const int clo[3], double *const face[6]) {
  QcfArgs A;
  A.g = g;
  for (int d = 0; d < 3; d++) { A.plo[d] = plo[d]; A.cdom[d] = ndom[d] / 2; A.clo[d] = clo[d]; }
  A.csy = csy; A.csz = csz;
  A.dir = 0; A.side = -1; A.ta = 1; A.tb = 2;
  A.h = h;
  QcfFaces F;
  bool any = false;
  for (int f = 0; f < 6; f++) { F.face[f] = face[f]; any = any || face[f]; }
  if (!any) return MGIC_OK;
  const int na = g.nx > g.ny ? g.nx : g.ny, nb = g.ny > g.nz ? g.ny : g.nz;   // ta in {x, y}, tb in {y, z}
  dim3 blk(32, 8, 1);
  dim3 grd((na + blk.x - 1) / blk.x, (nb + blk.y - 1) / blk.y, 6);
  k_quad_cf_faces<<<grd, blk, 0, c->stream>>>(A, F, phi, coarse);
  return post_launch(c, "quad_cf_faces");
}

int quad_cf_masked(mgic_ctx *c, const Geom &g, const unsigned char *mask, const int plo[3], const int ndom[3], double h, const double *phi,
                   const double *coarse, long long csy, long long csz, const int clo[3], double *const face[6]) {
  QcfmArgs A;
  A.g = g;
  for (int d = 0; d < 3; d++) { A.plo[d] = plo[d]; A.ndom[d] = ndom[d]; A.clo[d] = clo[d]; }
  A.csy = csy; A.csz = csz; A.h = h;
  for (int f = 0; f < 6; f++) A.face[f] = face[f];
  dim3 blk(64, 4, 1);
  dim3 grd = grid3(g.nx, g.ny, g.nz, blk);
  k_quad_cf_masked<<<grd, blk, 0, c->stream>>>(A, mask, phi, coarse);
  return post_launch(c, "quad_cf_masked");
}

int prolong(mgic_ctx *c, const Geom &g, double *phi, const double *coarse, long long csy, long long csz, const unsigned char *fineMask) {
  dim3 blk(32, 4, 1);
  dim3 grd = grid3(g.nx / 2, g.ny / 2, g.nz / 2, blk);
  k_prolong<<<grd, blk, 0, c->stream>>>(g, phi, coarse, csy, csz, fineMask);
  return post_launch(c, "prolong");
}

static inline long long ncells(const Geom &g) { return (long long)g.nx * g.ny * g.nz; }

int compute_lambda(mgic_ctx *c, const Geom &g, double *lam, const double *a, double alpha, double beta, double dx) {
  const double plus = 2.0 * 3 * beta / (dx * dx);  // :241
  const long long n = ncells(g);
  k_lambda<<<ew_grid(c, n), 256, 0, c->stream>>>(n, lam, a, alpha, plus);
  return post_launch(c, "lambda");
}

#define EW_LAUNCH(OP, y, x1, x2, s1, s2)                                              \
  const long long n = ncells(g);                                                      \
  k_ew<OP><<<ew_grid(c, n), 256, 0, c->stream>>>(n, y, x1, x2, s1, s2);                \
  return post_launch(c, #OP)

int mult(mgic_ctx *c, const Geom &g, double *y, const double *x, const double *l) { EW_LAUNCH(EW_MULT, y, x, l, 0.0, 0.0); }
int incr(mgic_ctx *c, const Geom &g, double *y, const double *x, double s) { EW_LAUNCH(EW_INCR, y, x, nullptr, s, 0.0); }
int axby(mgic_ctx *c, const Geom &g, double *y, const double *x1, const double *x2, double a, double b) {
  EW_LAUNCH(EW_AXBY, y, x1, x2, a, b);
}
int scale(mgic_ctx *c, const Geom &g, double *y, double s) { EW_LAUNCH(EW_SCALE, y, nullptr, nullptr, s, 0.0); }
int assign(mgic_ctx *c, const Geom &g, double *y, const double *x) { EW_LAUNCH(EW_ASSIGN, y, x, nullptr, 0.0, 0.0); }
int set_val(mgic_ctx *c, const Geom &g, double *y, double v) { EW_LAUNCH(EW_SETVAL, y, nullptr, nullptr, v, 0.0); }
int jacobi_update(mgic_ctx *c, const Geom &g, double *phi, const double *res, const double *lam, double w) {
  EW_LAUNCH(EW_JACOBI, phi, res, lam, w, 0.0);
}

int reduce(mgic_ctx *c, const Geom &g, const double *x, const double *y, int kind, int slot, const unsigned char *skip) {
  const long long n = ncells(g);
  // grid depends only on n -> summation tree (and result bits) are a function of the level size alone
  long long nb = (n + 256 * 8 - 1) / (256 * 8);
  if (nb < 1) nb = 1;
  if (nb > (long long)c->partCap) nb = (long long)c->partCap;
  const int grd = (int)nb;
  double *out = c->d_scal + slot;
  switch (kind) {
    case 0: k_reduce<0><<<grd, 256, 0, c->stream>>>(n, x, y, 0.0, c->d_part, c->d_count, out, skip); break;
    case 1: k_reduce<1><<<grd, 256, 0, c->stream>>>(n, x, y, 0.0, c->d_part, c->d_count, out, skip); break;
    case 2: k_reduce<2><<<grd, 256, 0, c->stream>>>(n, x, y, 0.0, c->d_part, c->d_count, out, skip); break;
    case 3: k_reduce<3><<<grd, 256, 0, c->stream>>>(n, x, y, 0.0, c->d_part, c->d_count, out, skip); break;
    default: mgic_set_error("reduce: bad kind %d", kind); return MGIC_ERR_ARG;
  }
  return post_launch(c, "reduce");
}

int is_constant(mgic_ctx *c, const Geom &g, const double *x, double value, int slot) {
  const long long n = ncells(g);
  long long nb = (n + 256 * 8 - 1) / (256 * 8);
  if (nb < 1) nb = 1;
  if (nb > (long long)c->partCap) nb = (long long)c->partCap;
  k_reduce<4><<<(int)nb, 256, 0, c->stream>>>(n, x, nullptr, value, c->d_part, c->d_count, c->d_scal + slot, nullptr);
  return post_launch(c, "is_constant");
}

int box_set_val(mgic_ctx *c, const Geom &g, double *y, double v, const unsigned char *fineMask, long long msy, long long msz) {
  dim3 blk(32, 4, 1);
  dim3 grd = grid3(g.nx, g.ny, g.nz, blk);
  k_box_set<<<grd, blk, 0, c->stream>>>(g, y, v, fineMask, msy, msz);
  return post_launch(c, "box_set_val");
}

int box_set_u8(mgic_ctx *c, const Geom &g, unsigned char *y, unsigned char v, const unsigned char *fineMask, long long msy, long long msz) {
  dim3 blk(32, 4, 1);
  dim3 grd = grid3(g.nx, g.ny, g.nz, blk);
  k_box_set_u8<<<grd, blk, 0, c->stream>>>(g, y, v, fineMask, msy, msz);
  return post_launch(c, "box_set_u8");
}

int copy_box(mgic_ctx *c, int nx, int ny, int nz, const double *src, long long ssy, long long ssz, double *dst, long long dsy, long long dsz) {
  if (nx <= 0 || ny <= 0 || nz <= 0) return MGIC_OK;
  dim3 blk(32, 4, 1);
  dim3 grd = grid3(nx, ny, nz, blk);
  k_copy_box<<<grd, blk, 0, c->stream>>>(nx, ny, src, ssy, ssz, dst, dsy, dsz);
  return post_launch(c, "copy_box");
}

int apply_mask(mgic_ctx *c, const Geom &g, double *y, const unsigned char *mask) {
  const long long n = ncells(g);
  k_apply_mask<<<ew_grid(c, n), 256, 0, c->stream>>>(n, y, mask);
  return post_launch(c, "apply_mask");
}

int coarse_average(mgic_ctx *c, const Geom &gc, double *cp, const double *fine, long long fsy, long long fsz, int nref,
                   int harmonic, const unsigned char *fineMask) {
  dim3 blk(32, 4, 1);
  dim3 grd = grid3(gc.nx, gc.ny, gc.nz, blk);
  k_coarse_average<<<grd, blk, 0, c->stream>>>(gc, cp, fine, fsy, fsz, nref, harmonic, fineMask);
  return post_launch(c, "coarse_average");
}

}  // namespace mgk
