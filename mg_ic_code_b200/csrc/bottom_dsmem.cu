// bottom_dsmem.cu -- the bottom BiCGStab with every vector resident in the shared memory of ONE thread-block cluster.
//
// For bottom levels of up to 32^3 cells (the single-GPU and strong-scaling case) the eleven vectors of the solve
// (phi, rhs, aCoef, lambda, r, r~, e, p, v, p~/s~, t) fit in the distributed shared memory of a 16-CTA cluster: each CTA
// owns nz/16 planes of every vector, x/y neighbours are local, z neighbours across a plane-slab boundary are read from
// the neighbouring CTA's shared memory (ld.shared::cluster), dot products are reduced through DSMEM, and the 15 barriers
// per BiCGStab iteration are hardware cluster barriers with nothing to flush to L2.  Global memory is touched twice: to
// load the level and to store the corrected phi.  Same arithmetic and control flow as bottom.cu / the oracle.
#include <cooperative_groups.h>

#include "mgic_internal.h"
#include "mgic_device.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int NT = 1024;
constexpr int MAXC = 4;  // cells per thread

struct DsArgs {
  Geom g;
  BCk bc;  // homogeneous, physical on all six faces
  double alpha, beta, dxinv;
  double *phi;
  const double *rhs, *a, *b, *lam;
  int pl;  // planes per CTA
  int imax;
  double eps, reps, hang, small;
  int numRestarts;
  int *out;
};

template <bool HAS_B>
struct Ds {
  const DsArgs &A;
  cg::cluster_group cl;
  int rank, csize, nloc, nxy, nmine;
  // shared-memory vectors of this CTA (nloc doubles each)
  double *phi, *rhs, *a, *lam, *b, *r, *rt, *e, *p, *v, *x, *t;
  double *red;  // 2 buffers x 2 values
  int nred = 0;
  int li[MAXC];      // local index of the thread's cells (-1: none)
  int ijk[MAXC];     // packed global (i, j, k)

  __device__ Ds(const DsArgs &a_, double *smem) : A(a_), cl(cg::this_cluster()) {
    rank = (int)cl.block_rank(); csize = (int)cl.num_blocks();
    nxy = A.g.nx * A.g.ny; nloc = nxy * A.pl;
    double *q = smem;
    phi = q; q += nloc; rhs = q; q += nloc; a = q; q += nloc; lam = q; q += nloc;
    if (HAS_B) { b = q; q += nloc; } else b = nullptr;
    r = q; q += nloc; rt = q; q += nloc; e = q; q += nloc; p = q; q += nloc; v = q; q += nloc; x = q; q += nloc; t = q; q += nloc;
    red = q;
    nmine = 0;
#pragma unroll
    for (int m = 0; m < MAXC; m++) {
      const int l = threadIdx.x + m * NT;
      li[m] = -1; ijk[m] = 0;
      if (l < nloc) {
        const int kl = l / nxy, rem = l - kl * nxy, j = rem / A.g.nx, i = rem - j * A.g.nx;
        li[m] = l; ijk[m] = i | (j << 10) | ((rank * A.pl + kl) << 20);
        nmine = m + 1;
      }
    }
  }
#define FOR_CELLS(m) _Pragma("unroll") for (int m = 0; m < MAXC; m++) if (li[m] >= 0)

  // neighbours of the thread's cell m in vector X (the same offset in every CTA's shared memory)
  __device__ __forceinline__ Nb nbr(double *X, int m, double c) const {
    const int l = li[m], i = ijk[m] & 1023, j = (ijk[m] >> 10) & 1023, k = ijk[m] >> 20;
    const int kl = k - rank * A.pl;
    const BCk &bc = A.bc;
    Nb n;
    n.xm = (i > 0) ? X[l - 1] : bc.a[0] * c + bc.b[0];
    n.xp = (i < A.g.nx - 1) ? X[l + 1] : bc.a[1] * c + bc.b[1];
    n.ym = (j > 0) ? X[l - A.g.nx] : bc.a[2] * c + bc.b[2];
    n.yp = (j < A.g.ny - 1) ? X[l + A.g.nx] : bc.a[3] * c + bc.b[3];
    if (kl > 0) n.zm = X[l - nxy];
    else if (k > 0) n.zm = cl.map_shared_rank(X, rank - 1)[l + (A.pl - 1) * nxy];
    else n.zm = bc.a[4] * c + bc.b[4];
    if (kl < A.pl - 1) n.zp = X[l + nxy];
    else if (k < A.g.nz - 1) n.zp = cl.map_shared_rank(X, rank + 1)[l - (A.pl - 1) * nxy];
    else n.zp = bc.a[5] * c + bc.b[5];
    return n;
  }
  __device__ __forceinline__ double op_point(double *X, int m) const {  // VCCOMPUTEOP3D (:209-234)
    const int l = li[m];
    const double c = X[l];
    const Nb nb = nbr(X, m, c);
    double lp = lap7(c, nb.xm, nb.xp, nb.ym, nb.yp, nb.zm, nb.zp);
    lp = lp * A.dxinv * A.beta;
    if (HAS_B) lp = lp * b[l];
    return A.alpha * a[l] * c - lp;
  }
  __device__ __forceinline__ double res_point(double *X, int m) const {  // VCCOMPUTERES3D (:312-336)
    const int l = li[m];
    const double c = X[l];
    const Nb nb = nbr(X, m, c);
    double lp = lap7(c, nb.xm, nb.xp, nb.ym, nb.yp, nb.zm, nb.zp);
    lp = lp * A.dxinv * A.beta;
    if (HAS_B) lp = lp * b[l];
    return (rhs[l] - A.alpha * a[l] * c) + lp;
  }
  // relax(X, R, 2): four colour passes, a cluster barrier before each (levelGSRB)
  __device__ void relax2(double *X, const double *R) {
    for (int pass = 0; pass < 4; pass++) {
      cl.sync();
      const int color = pass & 1;
      FOR_CELLS(m) {
        const int i = ijk[m] & 1023, j = (ijk[m] >> 10) & 1023, k = ijk[m] >> 20;
        if (((i + j + k + A.g.k0 + color) & 1) == 0) {
          const int l = li[m];
          const double c = X[l];
          const Nb nb = nbr(X, m, c);
          X[l] = gsrb_point<HAS_B>(c, nb.xm, nb.xp, nb.ym, nb.yp, nb.zm, nb.zp, a[l], HAS_B ? b[l] : 1.0, lam[l], R[l], A.alpha, A.beta,
                                   A.dxinv);
        }
      }
    }
    cl.sync();
  }
  // sums (v0, v1) over the cluster through DSMEM; one cluster barrier; identical bits in every thread of every CTA
  __device__ void reduce2(double &v0, double &v1) {
    __shared__ double sh[68];
    double *buf = red + (nred & 1) * 2;
    nred++;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { v0 += __shfl_down_sync(0xffffffffu, v0, o); v1 += __shfl_down_sync(0xffffffffu, v1, o); }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) { sh[w] = v0; sh[32 + w] = v1; }
    __syncthreads();
    if (w == 0) {
      double x0 = sh[l], x1 = sh[32 + l];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { x0 += __shfl_down_sync(0xffffffffu, x0, o); x1 += __shfl_down_sync(0xffffffffu, x1, o); }
      if (l == 0) { buf[0] = x0; buf[1] = x1; }
    }
    cl.sync();
    if (w == 0) {
      double x0 = 0.0, x1 = 0.0;
      if (l < csize) { const double *rb = cl.map_shared_rank(buf, l); x0 = rb[0]; x1 = rb[1]; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { x0 += __shfl_down_sync(0xffffffffu, x0, o); x1 += __shfl_down_sync(0xffffffffu, x1, o); }
      if (l == 0) { sh[64] = x0; sh[65] = x1; }
    }
    __syncthreads();
    v0 = sh[64]; v1 = sh[65];
  }

  __device__ void solve() {
    // load the level (this CTA's planes)
    const long long g0 = (long long)rank * A.pl * nxy;
    FOR_CELLS(m) {
      const int l = li[m];
      phi[l] = A.phi[g0 + l]; rhs[l] = A.rhs[g0 + l]; a[l] = A.a[g0 + l]; lam[l] = A.lam[g0 + l];
      if (HAS_B) b[l] = A.b[g0 + l];
    }
    cl.sync();
    double s0 = 0.0, s1 = 0.0;
    FOR_CELLS(m) {
      const int l = li[m];
      const double rv = res_point(phi, m);
      r[l] = rv; rt[l] = rv; e[l] = 0.0;
      s0 += rv * rv;
    }
    reduce2(s0, s1);
    double norm0 = sqrt(s0), norm1 = norm0;
    const double initial_norm = norm0, initial_rnorm = norm0;
    double rho1 = 0.0, rho2 = 0.0, alpha0 = 0.0, alpha1 = 0.0, beta1 = 0.0, omega0 = 0.0, omega1 = 0.0;
    bool init = true, finished = false;
    int restarts = 0, recount = 0, status = -1, it = 0;
    while ((it < A.imax && norm0 > A.eps * norm1) && (norm1 > 0)) {
      it++;
      norm1 = norm0; alpha1 = alpha0; omega1 = omega0;
      rho2 = rho1;
      s0 = 0.0; s1 = 0.0;
      FOR_CELLS(m) { const int l = li[m]; s0 += rt[l] * r[l]; }
      reduce2(s0, s1);
      rho1 = s0;
      if (rho1 == 0.0) {
        FOR_CELLS(m) { const int l = li[m]; phi[l] = phi[l] + 1.0 * e[l]; }
        status = 2; finished = true;
        break;
      }
      if (init) {
        FOR_CELLS(m) { const int l = li[m]; const double pv = r[l]; p[l] = pv; x[l] = pv * lam[l]; }
        init = false;
      } else {
        beta1 = (rho1 / rho2) * (alpha1 / omega1);
        const double c2 = -beta1 * omega1;
        FOR_CELLS(m) {
          const int l = li[m];
          double pv = p[l] * beta1;
          pv = pv + c2 * v[l];
          pv = pv + 1.0 * r[l];
          p[l] = pv;
          x[l] = pv * lam[l];
        }
      }
      relax2(x, p);                                            // p_tilde = preCond(p)
      s0 = 0.0; s1 = 0.0;
      FOR_CELLS(m) { const int l = li[m]; const double vv = op_point(x, m); v[l] = vv; s0 += rt[l] * vv; }
      reduce2(s0, s1);
      const double mm = s0;
      alpha0 = rho1 / mm;
      if (fabs(mm) > A.small * fabs(rho1)) {
        const double na = -alpha0;
        s0 = 0.0; s1 = 0.0;
        FOR_CELLS(m) {
          const int l = li[m];
          const double rv = r[l] + na * v[l];
          r[l] = rv; s0 += rv * rv;
          e[l] = e[l] + alpha0 * x[l];
        }
        reduce2(s0, s1);
        norm0 = sqrt(s0);
      } else {
        FOR_CELLS(m) { r[li[m]] = 0.0; }
        norm0 = 0.0;
      }
      if (norm0 > A.eps * initial_norm && norm0 > A.reps * initial_rnorm) {
        // (remote reads of x as p_tilde ended before the barrier inside the reduction of m)
        FOR_CELLS(m) { const int l = li[m]; x[l] = r[l] * lam[l]; }
        relax2(x, r);                                          // s_tilde = preCond(r)
        s0 = 0.0; s1 = 0.0;
        FOR_CELLS(m) { const int l = li[m]; const double tv = op_point(x, m); t[l] = tv; s0 += tv * r[l]; s1 += tv * tv; }
        reduce2(s0, s1);
        omega0 = s0 / s1;
        const double no = -omega0;
        s0 = 0.0; s1 = 0.0;
        FOR_CELLS(m) {
          const int l = li[m];
          e[l] = e[l] + omega0 * x[l];
          const double rv = r[l] + no * t[l];
          r[l] = rv; s0 += rv * rv;
        }
        reduce2(s0, s1);
        norm0 = sqrt(s0);
      }
      if (norm0 <= A.eps * initial_norm || norm0 <= A.reps * initial_rnorm) { status = 1; break; }
      if (omega0 == 0.0 || norm0 > (1 - A.hang) * norm1) {
        if (recount == 0) recount = 1;
        else {
          recount = 0;
          FOR_CELLS(m) { const int l = li[m]; phi[l] = phi[l] + 1.0 * e[l]; }
          if (restarts == A.numRestarts) { status = 3; finished = true; break; }
          cl.sync();
          s0 = 0.0; s1 = 0.0;
          FOR_CELLS(m) {
            const int l = li[m];
            const double rv = res_point(phi, m);
            r[l] = rv; rt[l] = rv; e[l] = 0.0;
            s0 += rv * rv;
          }
          reduce2(s0, s1);
          norm0 = sqrt(s0);
          rho1 = 0.0; rho2 = 0.0; alpha0 = 0.0; beta1 = 0.0; omega0 = 0.0;
          restarts++;
          init = true;
        }
      }
      // (x is rewritten only after the next iteration's rho reduction, which contains a cluster barrier)
    }
    if (!finished) FOR_CELLS(m) { const int l = li[m]; phi[l] = phi[l] + 1.0 * e[l]; }
    FOR_CELLS(m) { A.phi[g0 + li[m]] = phi[li[m]]; }
    if (rank == 0 && threadIdx.x == 0) { A.out[0] = it; A.out[1] = status; }
    cl.sync();  // no CTA may exit while a neighbour can still read its shared memory
  }
};

template <bool HAS_B>
__global__ void __launch_bounds__(NT, 1) k_bottom_dsmem(DsArgs A) {
  extern __shared__ __align__(16) double smem[];
  Ds<HAS_B> d(A, smem);
  d.solve();
}

}  // namespace

namespace mgk {

// *used = 1 if the level fits one cluster's shared memory and the kernel ran; 0: caller falls back
int bottom_bicgstab_dsmem(mgic_op *o, mgic_field *e, const mgic_field *r, int *d_out, int *used) {
  mgic_ctx *c = o->ctx;
  *used = 0;
  const Geom g = o->geom();
  const BCk bc = o->bck(true);
  for (int f = 0; f < 6; f++)
    if (bc.type[f] != MGIC_BC_DIRICHLET && bc.type[f] != MGIC_BC_NEUMANN) return MGIC_OK;
  if (g.nx > 1023 || g.ny > 1023 || g.nz > 2047) return MGIC_OK;
  const int nvec = o->b ? 12 : 11;
  void (*kern)(DsArgs) = o->b ? k_bottom_dsmem<true> : k_bottom_dsmem<false>;
  int &mc = *mgic_dev_cache(c->device, (const void *)kern, 0, -1);
  // largest cluster (<= 16, dividing nz) whose per-CTA slab fits shared memory and the per-thread cell budget
  for (int cs = 16; cs >= 1; cs >>= 1) {
    if (g.nz % cs) continue;
    const int pl = g.nz / cs;
    const long long nloc = (long long)g.nx * g.ny * pl;
    const size_t smem = (size_t)(nvec * nloc + 4) * sizeof(double);
    if (nloc > (long long)NT * MAXC || smem > 200 * 1024) continue;
    if (mc < 0) {
      mc = 8;
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) mc = 16;
      cudaGetLastError();
    }
    if (cs > mc) continue;
    MGIC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem; cfg.stream = c->stream;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = cs; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) != cudaSuccess || nclusters < 1) { cudaGetLastError(); continue; }
    DsArgs A;
    A.g = g; A.bc = bc;
    A.alpha = o->alpha; A.beta = o->beta; A.dxinv = 1.0 / (o->dx * o->dx);
    A.phi = e->p; A.rhs = r->p; A.a = o->a->p; A.b = o->b ? o->b->p : nullptr; A.lam = o->lambda->p;
    A.pl = pl;
    A.imax = 80; A.eps = 1.0e-6; A.reps = 1.0e-12; A.hang = 1.0e-8; A.small = 1.0e-30; A.numRestarts = 5;
    A.out = d_out;
    MGIC_CUDA(cudaLaunchKernelEx(&cfg, kern, A));
    c->launches++;
    *used = 1;
    return MGIC_OK;
  }
  return MGIC_OK;
}

}  // namespace mgk
