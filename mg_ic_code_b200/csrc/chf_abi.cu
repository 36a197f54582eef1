// chf_abi.cu -- link-time drop-ins for the reference's Chombo-Fortran kernels (include/mgic_chf.h).
//
// Each symbol has exactly the argument list the ChF preprocessor generates (prototypes:
// Source/VariableCoeffPoissonOperatorF_F.H:107-117,233-241,359-368,488-497; Source/SetLevelDataF_F.H:15-19,
// 43-47).  Semantics: host pointers in, host pointers out; the FABs are staged to HBM, one sm_100a kernel runs
// over `region`, the output FAB is copied back.  Errors abort() like the Fortran MAYDAYERROR().  No CPU path.
#include <cstdlib>

#include "mgic_chf.h"
#include "mgic_internal.h"

namespace {

using mgk::FabView;

__device__ __forceinline__ double &at(const FabView &v, int i, int j, int k, int n) {
  return v.p[(i - v.lo[0]) + v.s1 * (j - v.lo[1]) + v.s2 * (long long)(k - v.lo[2]) + v.sc * n];
}
__device__ __forceinline__ double lap7f(const FabView &u, int i, int j, int k, int n) {
  const double t = 2.0 * at(u, i, j, k, n);
  return ((at(u, i + 1, j, k, n) + at(u, i - 1, j, k, n)) - t) + ((at(u, i, j + 1, k, n) + at(u, i, j - 1, k, n)) - t) +
         ((at(u, i, j, k + 1, n) + at(u, i, j, k - 1, n)) - t);
}

struct Region { int lo[3], hi[3]; };

// GSRBHELMHOLTZVC3D (VariableCoeffPoissonOperatorF.ChF:56-139)
__global__ void __launch_bounds__(128) kf_gsrb(FabView dpsi, FabView rhs, Region r, double dxinv, double alpha, FabView a,
                                               double beta, FabView b, FabView lam, int redBlack, int ncomp) {
  const int j = r.lo[1] + blockIdx.y * blockDim.y + threadIdx.y;
  const int k = r.lo[2] + blockIdx.z;
  if (j > r.hi[1]) return;
  const int imin = r.lo[0] + abs((r.lo[0] + j + k + redBlack) % 2);  // :98-104
  const int i = imin + 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (i > r.hi[0]) return;
  for (int n = 0; n < ncomp; n++) {
    double lof = alpha * at(a, i, j, k, n) * at(dpsi, i, j, k, n);
    double l = lap7f(dpsi, i, j, k, n);
    l = l * dxinv * at(b, i, j, k, n);
    lof = lof - beta * l;
    at(dpsi, i, j, k, n) = at(dpsi, i, j, k, n) - at(lam, i, j, k, n) * (lof - at(rhs, i, j, k, n));
  }
}

// VCCOMPUTEOP3D (:181-237) mode 0, VCCOMPUTERES3D (:283-339) mode 1
__global__ void __launch_bounds__(128) kf_op(FabView out, FabView dpsi, FabView rhs, Region r, double dxinv, double alpha,
                                             FabView a, double beta, FabView b, int ncomp, int mode) {
  const int i = r.lo[0] + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = r.lo[1] + blockIdx.y * blockDim.y + threadIdx.y;
  const int k = r.lo[2] + blockIdx.z;
  if (i > r.hi[0] || j > r.hi[1]) return;
  for (int n = 0; n < ncomp; n++) {
    double l = lap7f(dpsi, i, j, k, n);
    l = l * dxinv * beta * at(b, i, j, k, n);
    if (mode == 0) at(out, i, j, k, n) = alpha * at(a, i, j, k, n) * at(dpsi, i, j, k, n) - l;
    else at(out, i, j, k, n) = (at(rhs, i, j, k, n) - alpha * at(a, i, j, k, n) * at(dpsi, i, j, k, n)) + l;
  }
}

// RESTRICTRESVC3D (:379-437): res(i/2,j/2,k/2) += (rhs - L dpsi)/8 -- one thread per coarse cell touched by
// `region`, fine cells visited in the Fortran loop order; accumulates ONTO the incoming res like the Fortran.
__global__ void __launch_bounds__(128) kf_restrict(FabView res, FabView dpsi, FabView rhs, Region r, double dxinv, double alpha,
                                                   FabView a, double beta, FabView b, int ncomp) {
  const int I = r.lo[0] / 2 + blockIdx.x * blockDim.x + threadIdx.x;
  const int J = r.lo[1] / 2 + blockIdx.y * blockDim.y + threadIdx.y;
  const int K = r.lo[2] / 2 + blockIdx.z;
  if (I > r.hi[0] / 2 || J > r.hi[1] / 2) return;
  for (int n = 0; n < ncomp; n++) {
    double acc = at(res, I, J, K, n);
    for (int k = max(2 * K, r.lo[2]); k <= min(2 * K + 1, r.hi[2]); k++)
      for (int j = max(2 * J, r.lo[1]); j <= min(2 * J + 1, r.hi[1]); j++)
        for (int i = max(2 * I, r.lo[0]); i <= min(2 * I + 1, r.hi[0]); i++) {
          double lof = alpha * at(a, i, j, k, n) * at(dpsi, i, j, k, n);
          double l = lap7f(dpsi, i, j, k, n);
          l = l * dxinv * beta * at(b, i, j, k, n);
          lof = lof - l;
          acc = acc + (at(rhs, i, j, k, n) - lof) / 8.0;
        }
    at(res, I, J, K, n) = acc;
  }
}

// GETLAPLACIANPSIF (SetLevelDataF.ChF:15-58) mode 0, GETRHOGRADPHIF (:65-103) mode 1
__global__ void __launch_bounds__(128) kf_src(FabView out, FabView in, Region r, double dx, int mode) {
  const int i = r.lo[0] + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = r.lo[1] + blockIdx.y * blockDim.y + threadIdx.y;
  const int k = r.lo[2] + blockIdx.z;
  if (i > r.hi[0] || j > r.hi[1]) return;
  double acc = 0.0;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const int e0 = (d == 0), e1 = (d == 1), e2 = (d == 2);
    const double m = at(in, i - e0, j - e1, k - e2, 0), p = at(in, i + e0, j + e1, k + e2, 0);
    if (mode == 0) acc = acc + 1.0 / dx / dx * (+1.0 * m - 2.0 * at(in, i, j, k, 0) + 1.0 * p);
    else {
      const double g = 0.5 / dx * (p - m);
      acc = acc + 0.5 * g * g;
    }
  }
  at(out, i, j, k, 0) = acc;
}

// [Chombo] AMRPoissonOpF.ChF PROLONG: phi(i,j,k) += coarse(i/m, j/m, k/m)
__global__ void __launch_bounds__(128) kf_prolong(FabView phi, FabView coarse, Region r, int m, int ncomp) {
  const int i = r.lo[0] + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = r.lo[1] + blockIdx.y * blockDim.y + threadIdx.y;
  const int k = r.lo[2] + blockIdx.z;
  if (i > r.hi[0] || j > r.hi[1]) return;
  for (int n = 0; n < ncomp; n++) at(phi, i, j, k, n) = at(phi, i, j, k, n) + at(coarse, i / m, j / m, k / m, n);
}

// ---- host staging -----------------------------------------------------------------------------------------
[[noreturn]] void mayday(const char *what) {
  fprintf(stderr, "mgic_b200: %s: %s\n", what, mgic_last_error());
  abort();
}
#define CK(call)                                                              \
  do {                                                                        \
    cudaError_t e__ = (call);                                                 \
    if (e__ != cudaSuccess) {                                                 \
      mgic_set_error("%s -> %s", #call, cudaGetErrorString(e__));             \
      mayday("CUDA error (there is no CPU fallback)");                        \
    }                                                                         \
  } while (0)

struct Staged {
  FabView v;
  size_t bytes = 0;
  double *host = nullptr;
  Staged(const double *p, int l0, int l1, int l2, int h0, int h1, int h2, int nc, cudaStream_t s) {
    v.lo[0] = l0; v.lo[1] = l1; v.lo[2] = l2;
    v.s1 = h0 - l0 + 1; v.s2 = v.s1 * (h1 - l1 + 1); v.sc = v.s2 * (h2 - l2 + 1);
    bytes = (size_t)v.sc * nc * sizeof(double);
    host = const_cast<double *>(p);
    CK(cudaMalloc(&v.p, bytes));
    CK(cudaMemcpyAsync(v.p, p, bytes, cudaMemcpyHostToDevice, s));
  }
  void back(cudaStream_t s) { CK(cudaMemcpyAsync(host, v.p, bytes, cudaMemcpyDeviceToHost, s)); }
  ~Staged() { cudaFree(v.p); }
};
#define STAGE(a, nc) Staged s_##a(a, *i##a##lo0, *i##a##lo1, *i##a##lo2, *i##a##hi0, *i##a##hi1, *i##a##hi2, nc, st)
#define REGION(b) Region rg = {{*i##b##lo0, *i##b##lo1, *i##b##lo2}, {*i##b##hi0, *i##b##hi1, *i##b##hi2}}

inline dim3 rgrid(int nx, int ny, int nz, dim3 blk) { return dim3((nx + blk.x - 1) / blk.x, (ny + blk.y - 1) / blk.y, nz); }
inline bool empty(const Region &r) { return r.hi[0] < r.lo[0] || r.hi[1] < r.lo[1] || r.hi[2] < r.lo[2]; }
void finish(cudaStream_t st) {
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
}

}  // namespace

extern "C" {

void gsrbhelmholtzvc3d_(MGIC_FRA(dpsi), MGIC_CFRA(rhs), MGIC_BOX(region), const double *dx, const double *alpha,
                        MGIC_CFRA(aCoef), const double *beta, MGIC_CFRA(bCoef), MGIC_CFRA(lambda), const int *redBlack) {
  const int ncomp = *ndpsicomp;
  if (ncomp != *nrhscomp || ncomp != *nbCoefcomp) { mgic_set_error("ncomp mismatch"); mayday("GSRBHELMHOLTZVC3D MAYDAYERROR"); }  // :77-87
  cudaStream_t st = 0;
  REGION(region);
  if (empty(rg)) return;
  STAGE(dpsi, ncomp); STAGE(rhs, *nrhscomp); STAGE(aCoef, *naCoefcomp); STAGE(bCoef, *nbCoefcomp); STAGE(lambda, *nlambdacomp);
  dim3 blk(32, 4, 1);
  dim3 grd = rgrid((rg.hi[0] - rg.lo[0] + 2) / 2, rg.hi[1] - rg.lo[1] + 1, rg.hi[2] - rg.lo[2] + 1, blk);
  kf_gsrb<<<grd, blk, 0, st>>>(s_dpsi.v, s_rhs.v, rg, 1.0 / (*dx * *dx), *alpha, s_aCoef.v, *beta, s_bCoef.v, s_lambda.v, *redBlack,
                               ncomp);
  s_dpsi.back(st);
  finish(st);
}

void vccomputeop3d_(MGIC_FRA(lofdpsi), MGIC_CFRA(dpsi), const double *alpha, MGIC_CFRA(aCoef), const double *beta,
                    MGIC_CFRA(bCoef), MGIC_BOX(region), const double *dx) {
  const int ncomp = *ndpsicomp;
  if (ncomp != *nlofdpsicomp || ncomp != *nbCoefcomp) { mgic_set_error("ncomp mismatch"); mayday("VCCOMPUTEOP3D MAYDAYERROR"); }  // :200-206
  cudaStream_t st = 0;
  REGION(region);
  if (empty(rg)) return;
  STAGE(lofdpsi, ncomp); STAGE(dpsi, ncomp); STAGE(aCoef, *naCoefcomp); STAGE(bCoef, *nbCoefcomp);
  dim3 blk(32, 4, 1);
  dim3 grd = rgrid(rg.hi[0] - rg.lo[0] + 1, rg.hi[1] - rg.lo[1] + 1, rg.hi[2] - rg.lo[2] + 1, blk);
  kf_op<<<grd, blk, 0, st>>>(s_lofdpsi.v, s_dpsi.v, s_dpsi.v, rg, 1.0 / (*dx * *dx), *alpha, s_aCoef.v, *beta, s_bCoef.v, ncomp, 0);
  s_lofdpsi.back(st);
  finish(st);
}

void vccomputeres3d_(MGIC_FRA(res), MGIC_CFRA(dpsi), MGIC_CFRA(rhs), const double *alpha, MGIC_CFRA(aCoef), const double *beta,
                     MGIC_CFRA(bCoef), MGIC_BOX(region), const double *dx) {
  const int ncomp = *ndpsicomp;
  if (ncomp != *nrescomp || ncomp != *nbCoefcomp) { mgic_set_error("ncomp mismatch"); mayday("VCCOMPUTERES3D MAYDAYERROR"); }  // :303-309
  cudaStream_t st = 0;
  REGION(region);
  if (empty(rg)) return;
  STAGE(res, ncomp); STAGE(dpsi, ncomp); STAGE(rhs, *nrhscomp); STAGE(aCoef, *naCoefcomp); STAGE(bCoef, *nbCoefcomp);
  dim3 blk(32, 4, 1);
  dim3 grd = rgrid(rg.hi[0] - rg.lo[0] + 1, rg.hi[1] - rg.lo[1] + 1, rg.hi[2] - rg.lo[2] + 1, blk);
  kf_op<<<grd, blk, 0, st>>>(s_res.v, s_dpsi.v, s_rhs.v, rg, 1.0 / (*dx * *dx), *alpha, s_aCoef.v, *beta, s_bCoef.v, ncomp, 1);
  s_res.back(st);
  finish(st);
}

void restrictresvc3d_(MGIC_FRA(res), MGIC_CFRA(dpsi), MGIC_CFRA(rhs), const double *alpha, MGIC_CFRA(aCoef), const double *beta,
                      MGIC_CFRA(bCoef), MGIC_BOX(region), const double *dx) {
  const int ncomp = *ndpsicomp;
  cudaStream_t st = 0;
  REGION(region);
  if (empty(rg)) return;
  if (rg.lo[0] < 0 || rg.lo[1] < 0 || rg.lo[2] < 0) {  // :406-409 integer division needs shifted, non-negative indices
    mgic_set_error("region must be shifted to non-negative indices (CHF_FRA_SHIFT, VariableCoeffPoissonOperator.cpp:188-192)");
    mayday("RESTRICTRESVC3D");
  }
  STAGE(res, *nrescomp); STAGE(dpsi, ncomp); STAGE(rhs, *nrhscomp); STAGE(aCoef, *naCoefcomp); STAGE(bCoef, *nbCoefcomp);
  dim3 blk(32, 4, 1);
  dim3 grd = rgrid(rg.hi[0] / 2 - rg.lo[0] / 2 + 1, rg.hi[1] / 2 - rg.lo[1] / 2 + 1, rg.hi[2] / 2 - rg.lo[2] / 2 + 1, blk);
  kf_restrict<<<grd, blk, 0, st>>>(s_res.v, s_dpsi.v, s_rhs.v, rg, 1.0 / (*dx * *dx), *alpha, s_aCoef.v, *beta, s_bCoef.v, ncomp);
  s_res.back(st);
  finish(st);
}

void getlaplacianpsif_(MGIC_FRA1(l_of_psi), MGIC_CFRA1(psi), const double *dx, MGIC_BOX(box)) {
  cudaStream_t st = 0;
  REGION(box);
  if (empty(rg)) return;
  STAGE(l_of_psi, 1); STAGE(psi, 1);
  dim3 blk(32, 4, 1);
  dim3 grd = rgrid(rg.hi[0] - rg.lo[0] + 1, rg.hi[1] - rg.lo[1] + 1, rg.hi[2] - rg.lo[2] + 1, blk);
  kf_src<<<grd, blk, 0, st>>>(s_l_of_psi.v, s_psi.v, rg, *dx, 0);
  s_l_of_psi.back(st);
  finish(st);
}

void getrhogradphif_(MGIC_FRA1(rho_grad_phi), MGIC_CFRA1(phi), const double *dx, MGIC_BOX(box)) {
  cudaStream_t st = 0;
  REGION(box);
  if (empty(rg)) return;
  STAGE(rho_grad_phi, 1); STAGE(phi, 1);
  dim3 blk(32, 4, 1);
  dim3 grd = rgrid(rg.hi[0] - rg.lo[0] + 1, rg.hi[1] - rg.lo[1] + 1, rg.hi[2] - rg.lo[2] + 1, blk);
  kf_src<<<grd, blk, 0, st>>>(s_rho_grad_phi.v, s_phi.v, rg, *dx, 1);
  s_rho_grad_phi.back(st);
  finish(st);
}

void prolong_(MGIC_FRA(phi), MGIC_CFRA(coarse), MGIC_BOX(region), const int *m) {
  cudaStream_t st = 0;
  REGION(region);
  if (empty(rg)) return;
  if (rg.lo[0] < 0 || rg.lo[1] < 0 || rg.lo[2] < 0 || *m < 1) { mgic_set_error("region must be shifted to non-negative indices"); mayday("PROLONG"); }
  STAGE(phi, *nphicomp); STAGE(coarse, *ncoarsecomp);
  dim3 blk(32, 4, 1);
  dim3 grd = rgrid(rg.hi[0] - rg.lo[0] + 1, rg.hi[1] - rg.lo[1] + 1, rg.hi[2] - rg.lo[2] + 1, blk);
  kf_prolong<<<grd, blk, 0, st>>>(s_phi.v, s_coarse.v, rg, *m, *nphicomp);
  s_phi.back(st);
  finish(st);
}

}  // extern "C"
