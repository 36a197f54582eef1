// bottom.cu -- the multigrid bottom solve as ONE persistent kernel (one thread-block cluster, or a cooperative grid).
//
// [Chombo 3.2] MultiGrid::cycle ends in m_bottomSolver->solve(e, r): BiCGStabSolver<LevelData<FArrayBox>> with
// its defaults (imax 80, eps 1e-6, reps 1e-12, hang 1e-8, small 1e-30, 5 restarts, normType 2), preconditioned by
// VariableCoeffPoissonOperator::preCond (phi = rhs*lambda, then two GSRB sweeps; VariableCoeffPoissonOperator.cpp:
// 72-104).  The bottom level is tiny (N/2^depth per side), so driven from the host it is pure launch + readback
// latency: ~35 launches and 6 scalar readbacks per iteration.  Here the whole solve -- residual, preconditioner
// sweeps, applyOp, dot products, norms and the convergence / restart logic -- runs in a single kernel; the thread
// blocks meet at grid-wide barriers and every thread evaluates the same control flow from the same reduced scalars.
// Point arithmetic is the shared gsrb_point()/lap7() code, so fields match the multi-launch path bit for bit given
// the same scalars; reductions use a fixed tree (deterministic, independent of timing).
#include <cooperative_groups.h>

#include "mgic_internal.h"
#include "mgic_device.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int BT = 256;   // threads per block, cooperative-grid variant
constexpr int CT = 1024;  // threads per block, cluster variant (<= 64 registers: latency, not registers, bounds this kernel)
constexpr int CSIZE = 8;  // CTAs per cluster: portable maximum; 16 is used when the device can place it

// all CTAs of the cluster: hardware barrier with release / acquire at cluster scope (orders global memory too)
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

struct BottomArgs {
  Geom g;
  BCk bc;  // homogeneous
  double alpha, beta, dxinv;
  double *phi;
  const double *rhs, *a, *b, *lam;
  double *r, *rt, *e, *p, *pt, *st, *t, *v;
  double *part;  // 2 buffers x 2 values x gridDim partials
  int imax;
  double eps, reps, hang, small;
  int numRestarts;
  int *out;  // [0] iterations, [1] exit status
};

template <bool HAS_B, bool CLUSTER>
struct Bottom {
  const BottomArgs &A;
  cg::grid_group grid;
  __device__ __forceinline__ void gsync() {
    if (CLUSTER) cluster_sync_all();
    else grid.sync();
  }
  long long n, gtid, gsize;
  int nred;
  double *sh;  // 66 doubles of shared memory

  __device__ Bottom(const BottomArgs &a_, double *sh_) : A(a_), grid(cg::this_grid()), nred(0), sh(sh_) {
    n = (long long)A.g.nx * A.g.ny * A.g.nz;
    gtid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    gsize = (long long)gridDim.x * blockDim.x;
  }

  // sums (v0, v1) over the grid; contains one grid barrier; identical result in every thread
  __device__ void reduce2(double &v0, double &v1) {
    double *buf = A.part + (size_t)(nred & 1) * 2 * gridDim.x;
    nred++;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v0 += __shfl_down_sync(0xffffffffu, v0, o);
      v1 += __shfl_down_sync(0xffffffffu, v1, o);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sh[w] = v0; sh[32 + w] = v1; }
    __syncthreads();
    if (w == 0) {
      double x0 = (l < (int)(blockDim.x >> 5)) ? sh[l] : 0.0, x1 = (l < (int)(blockDim.x >> 5)) ? sh[32 + l] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        x0 += __shfl_down_sync(0xffffffffu, x0, o);
        x1 += __shfl_down_sync(0xffffffffu, x1, o);
      }
      if (l == 0) { buf[2 * blockIdx.x] = x0; buf[2 * blockIdx.x + 1] = x1; }
    }
    gsync();
    if (w == 0) {
      double x0 = 0.0, x1 = 0.0;
      for (int q = l; q < (int)gridDim.x; q += 32) {  // fixed order: same bits in every block
        x0 += *((volatile double *)&buf[2 * q]);
        x1 += *((volatile double *)&buf[2 * q + 1]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        x0 += __shfl_down_sync(0xffffffffu, x0, o);
        x1 += __shfl_down_sync(0xffffffffu, x1, o);
      }
      if (l == 0) { sh[64] = x0; sh[65] = x1; }
    }
    __syncthreads();
    v0 = sh[64];
    v1 = sh[65];
    __syncthreads();
  }

  __device__ __forceinline__ void ijk(long long q, int &i, int &j, int &k) const {
    i = (int)(q % A.g.nx);
    const long long t = q / A.g.nx;
    j = (int)(t % A.g.ny);
    k = (int)(t / A.g.ny);
  }

  // VCCOMPUTEOP3D point value (VariableCoeffPoissonOperatorF.ChF:209-234), homogeneous BC
  __device__ __forceinline__ double op_point(const double *x, long long q) const {
    int i, j, k;
    ijk(q, i, j, k);
    const double c = x[q];
    const Nb nb = neighbours(x, q, i, j, k, A.g, A.bc, c);
    double l = lap7(c, nb.xm, nb.xp, nb.ym, nb.yp, nb.zm, nb.zp);
    l = l * A.dxinv * A.beta;
    if (HAS_B) l = l * A.b[q];
    return A.alpha * A.a[q] * c - l;
  }
  // VCCOMPUTERES3D point value (:312-336)
  __device__ __forceinline__ double res_point(const double *x, const double *rhs, long long q) const {
    int i, j, k;
    ijk(q, i, j, k);
    const double c = x[q];
    const Nb nb = neighbours(x, q, i, j, k, A.g, A.bc, c);
    double l = lap7(c, nb.xm, nb.xp, nb.ym, nb.yp, nb.zm, nb.zp);
    l = l * A.dxinv * A.beta;
    if (HAS_B) l = l * A.b[q];
    return (rhs[q] - A.alpha * A.a[q] * c) + l;
  }

  // one GSRB point: everything it reads, gathered before anything is written (lets two points overlap their loads)
  struct Pt { long long q; double c, av, bv, lv, rv; Nb nb; };
  __device__ __forceinline__ Pt gather(const double *x, const double *rhs, long long h, int color, int hx) const {
    Pt p;
    const long long row = h / hx;
    const int t = (int)(h - row * hx), j = (int)(row % A.g.ny), k = (int)(row / A.g.ny);
    const int i = 2 * t + ((j + k + A.g.k0 + color) & 1);
    p.q = i + (long long)j * A.g.sy + (long long)k * A.g.sz;
    p.c = x[p.q];
    p.nb = neighbours(x, p.q, i, j, k, A.g, A.bc, p.c);
    p.av = A.a[p.q]; p.bv = HAS_B ? A.b[p.q] : 1.0; p.lv = A.lam[p.q]; p.rv = rhs[p.q];
    return p;
  }
  __device__ __forceinline__ double point(const Pt &p) const {
    return gsrb_point<HAS_B>(p.c, p.nb.xm, p.nb.xp, p.nb.ym, p.nb.yp, p.nb.zm, p.nb.zp, p.av, p.bv, p.lv, p.rv, A.alpha, A.beta,
                             A.dxinv);
  }
  // relax(x, rhs, 2): four colour passes, a barrier before each (levelGSRB, VariableCoeffPoissonOperator.cpp:290-331).
  // Threads enumerate the cells OF THE COLOUR (i = 2t + parity, VariableCoeffPoissonOperatorF.ChF:98-106), two at a time.
  __device__ void relax2(double *x, const double *rhs) {
    const bool even = (A.g.nx & 1) == 0;
    const int hx = A.g.nx / 2;
    const long long nh = (long long)hx * A.g.ny * A.g.nz;
    for (int pass = 0; pass < 4; pass++) {
      gsync();
      const int color = pass & 1;
      if (even) {
        for (long long h0 = gtid; h0 < nh; h0 += 2 * gsize) {
          const long long h1 = h0 + gsize;
          const Pt p0 = gather(x, rhs, h0, color, hx);
          if (h1 < nh) {
            const Pt p1 = gather(x, rhs, h1, color, hx);
            x[p0.q] = point(p0);
            x[p1.q] = point(p1);
          } else {
            x[p0.q] = point(p0);
          }
        }
      } else {
        for (long long q = gtid; q < n; q += gsize) {
          int i, j, k;
          ijk(q, i, j, k);
          if (((i + j + k + A.g.k0 + color) & 1) == 0) {
            const double c = x[q];
            const Nb nb = neighbours(x, q, i, j, k, A.g, A.bc, c);
            x[q] = gsrb_point<HAS_B>(c, nb.xm, nb.xp, nb.ym, nb.yp, nb.zm, nb.zp, A.a[q], HAS_B ? A.b[q] : 1.0, A.lam[q], rhs[q],
                                     A.alpha, A.beta, A.dxinv);
          }
        }
      }
    }
    gsync();
  }

  __device__ void solve() {
    double *phi = A.phi, *r = A.r, *rt = A.rt, *e = A.e, *p = A.p, *pt = A.pt, *st = A.st, *t = A.t, *v = A.v;
    const double *rhs = A.rhs;
    double s0 = 0.0, s1 = 0.0;
    // residual(r, phi, rhs, homogeneous); r_tilde = r; e = 0
    for (long long q = gtid; q < n; q += gsize) {
      const double rv = res_point(phi, rhs, q);
      r[q] = rv; rt[q] = rv; e[q] = 0.0;
      s0 += rv * rv;
    }
    reduce2(s0, s1);
    double norm0 = sqrt(s0), norm1 = norm0;
    const double initial_norm = norm0, initial_rnorm = norm0;
    double rho1 = 0.0, rho2 = 0.0, alpha0 = 0.0, alpha1 = 0.0, beta1 = 0.0, omega0 = 0.0, omega1 = 0.0;
    bool init = true;
    int restarts = 0, recount = 0, status = -1, i = 0;
    bool finished = false;  // phi already updated (early returns of the reference code)
    while ((i < A.imax && norm0 > A.eps * norm1) && (norm1 > 0)) {
      i++;
      norm1 = norm0; alpha1 = alpha0; omega1 = omega0;
      rho2 = rho1;
      s0 = 0.0; s1 = 0.0;
      for (long long q = gtid; q < n; q += gsize) s0 += rt[q] * r[q];
      reduce2(s0, s1);
      rho1 = s0;
      if (rho1 == 0.0) {
        for (long long q = gtid; q < n; q += gsize) phi[q] = phi[q] + 1.0 * e[q];
        status = 2; finished = true;
        break;
      }
      // p update fused with the first line of preCond: p_tilde = p * lambda
      if (init) {
        for (long long q = gtid; q < n; q += gsize) { const double pv = r[q]; p[q] = pv; pt[q] = pv * A.lam[q]; }
        init = false;
      } else {
        beta1 = (rho1 / rho2) * (alpha1 / omega1);
        const double c2 = -beta1 * omega1;
        for (long long q = gtid; q < n; q += gsize) {
          double pv = p[q] * beta1;     // scale(p, beta)
          pv = pv + c2 * v[q];          // incr(p, v, -beta*omega)
          pv = pv + 1.0 * r[q];         // incr(p, r, 1)
          p[q] = pv;
          pt[q] = pv * A.lam[q];
        }
      }
      relax2(pt, p);
      s0 = 0.0; s1 = 0.0;
      for (long long q = gtid; q < n; q += gsize) { const double vv = op_point(pt, q); v[q] = vv; s0 += rt[q] * vv; }
      reduce2(s0, s1);
      const double m = s0;
      alpha0 = rho1 / m;
      if (fabs(m) > A.small * fabs(rho1)) {
        const double na = -alpha0;
        s0 = 0.0; s1 = 0.0;
        for (long long q = gtid; q < n; q += gsize) {
          const double rv = r[q] + na * v[q];
          r[q] = rv; s0 += rv * rv;
          e[q] = e[q] + alpha0 * pt[q];
        }
        reduce2(s0, s1);
        norm0 = sqrt(s0);
      } else {
        for (long long q = gtid; q < n; q += gsize) r[q] = 0.0;
        norm0 = 0.0;
      }
      if (norm0 > A.eps * initial_norm && norm0 > A.reps * initial_rnorm) {
        for (long long q = gtid; q < n; q += gsize) st[q] = r[q] * A.lam[q];
        relax2(st, r);
        s0 = 0.0; s1 = 0.0;
        for (long long q = gtid; q < n; q += gsize) {
          const double tv = op_point(st, q);
          t[q] = tv; s0 += tv * r[q]; s1 += tv * tv;
        }
        reduce2(s0, s1);
        omega0 = s0 / s1;
        const double no = -omega0;
        s0 = 0.0; s1 = 0.0;
        for (long long q = gtid; q < n; q += gsize) {
          e[q] = e[q] + omega0 * st[q];
          const double rv = r[q] + no * t[q];
          r[q] = rv; s0 += rv * rv;
        }
        reduce2(s0, s1);
        norm0 = sqrt(s0);
      }
      if (norm0 <= A.eps * initial_norm || norm0 <= A.reps * initial_rnorm) { status = 1; break; }
      if (omega0 == 0.0 || norm0 > (1 - A.hang) * norm1) {
        if (recount == 0) recount = 1;
        else {
          recount = 0;
          for (long long q = gtid; q < n; q += gsize) phi[q] = phi[q] + 1.0 * e[q];
          if (restarts == A.numRestarts) { status = 3; finished = true; break; }
          gsync();
          s0 = 0.0; s1 = 0.0;
          for (long long q = gtid; q < n; q += gsize) {
            const double rv = res_point(phi, rhs, q);
            r[q] = rv; rt[q] = rv; e[q] = 0.0;
            s0 += rv * rv;
          }
          reduce2(s0, s1);
          norm0 = sqrt(s0);
          rho1 = 0.0; rho2 = 0.0; alpha0 = 0.0; beta1 = 0.0; omega0 = 0.0;
          restarts++;
          init = true;
        }
      }
    }
    if (!finished)
      for (long long q = gtid; q < n; q += gsize) phi[q] = phi[q] + 1.0 * e[q];
    if (gtid == 0) { A.out[0] = i; A.out[1] = status; }
  }
};

template <bool HAS_B>
__global__ void __launch_bounds__(BT) k_bottom_bicgstab(BottomArgs A) {
  __shared__ double sh[66];
  Bottom<HAS_B, false> s(A, sh);
  s.solve();
}

// same solve inside one thread-block cluster: barriers are the cluster's hardware barrier (~0.3 us instead of the
// ~4 us of a 128-block cooperative grid barrier), which is what bounds this latency-dominated kernel
template <bool HAS_B>
__global__ void __launch_bounds__(CT, 1) k_bottom_bicgstab_cluster(BottomArgs A) {
  __shared__ double sh[66];
  Bottom<HAS_B, true> s(A, sh);
  s.solve();
}

}  // namespace

namespace mgk {

// e += BiCGStab(op, r) on the bottom level; work = 8 fields of the level; out = device int[2] (iterations, status)
int bottom_bicgstab(mgic_op *o, mgic_field *e, const mgic_field *r, mgic_field *const work[8], double *part, int partCap,
                    int *d_out) {
  mgic_ctx *c = o->ctx;
  BottomArgs A;
  A.g = o->geom();
  A.bc = o->bck(true);
  A.alpha = o->alpha; A.beta = o->beta; A.dxinv = 1.0 / (o->dx * o->dx);
  A.phi = e->p; A.rhs = r->p; A.a = o->a->p; A.b = o->b ? o->b->p : nullptr; A.lam = o->lambda->p;
  A.r = work[0]->p; A.rt = work[1]->p; A.e = work[2]->p; A.p = work[3]->p; A.pt = work[4]->p; A.st = work[5]->p;
  A.t = work[6]->p; A.v = work[7]->p;
  A.part = part;
  A.imax = 80; A.eps = 1.0e-6; A.reps = 1.0e-12; A.hang = 1.0e-8; A.small = 1.0e-30; A.numRestarts = 5;
  A.out = d_out;
  const long long n = (long long)A.g.nx * A.g.ny * A.g.nz;
  if (c->bottomKernel == 5) {  // cluster-sized bricks even where the level would fit one cluster (tuning)
    int used = 0;
    MGIC_TRY(bottom_bicgstab_cbrick(o, e, r, work, part, partCap, d_out, &used));
    if (used) { c->lastBottomKernel = 5; return MGIC_OK; }
  }
  if (c->bottomKernel == 1 || c->bottomKernel == 5) {  // whole level in one cluster's shared memory (bottom_dsmem.cu), if it fits
    int used = 0;
    MGIC_TRY(bottom_bicgstab_dsmem(o, e, r, d_out, &used));
    if (used) { c->lastBottomKernel = 1; return MGIC_OK; }
  }
  if (c->bottomKernel == 1) {  // bricks held by 16-CTA clusters, four grid barriers per iteration (bottom_cbrick.cu)
    int used = 0;
    MGIC_TRY(bottom_bicgstab_cbrick(o, e, r, work, part, partCap, d_out, &used));
    if (used) { c->lastBottomKernel = 5; return MGIC_OK; }
  }
  if (c->bottomKernel == 1 || c->bottomKernel == 4 || c->bottomKernel == 5) {  // brick kernel: four grid barriers per iteration (bottom_brick.cu)
    int used = 0;
    MGIC_TRY(bottom_bicgstab_brick(o, e, r, work, part, partCap, d_out, &used));
    if (used) { c->lastBottomKernel = 4; return MGIC_OK; }
  }
  if (c->bottomKernel != 3 && n <= 262144) {
    // ONE cluster; 16 CTAs (non-portable size) when the level is big enough to use them and the device can place it
    void (*ck)(BottomArgs) = o->b ? k_bottom_bicgstab_cluster<true> : k_bottom_bicgstab_cluster<false>;
    int &mc = *mgic_dev_cache(c->device, (const void *)ck, 0, 0);
    if (!mc) {
      mc = CSIZE;
      if (cudaFuncSetAttribute(ck, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
        cudaLaunchConfig_t q = {};
        q.gridDim = dim3(16); q.blockDim = dim3(CT);
        cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = 16; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        q.attrs = &at; q.numAttrs = 1;
        int nc = 0;
        if (cudaOccupancyMaxActiveClusters(&nc, ck, &q) == cudaSuccess && nc >= 1) mc = 16;
      }
      cudaGetLastError();
    }
    int cs = (n > (long long)CSIZE * CT) ? mc : CSIZE;
    if (c->bottomKernel == 2) cs = CSIZE;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs); cfg.blockDim = dim3(CT); cfg.stream = c->stream;
    cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = cs; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    MGIC_CUDA(cudaLaunchKernelEx(&cfg, ck, A));
    c->launches++;
    c->lastBottomKernel = 2;
    return MGIC_OK;
  }
  void *kern = o->b ? (void *)k_bottom_bicgstab<true> : (void *)k_bottom_bicgstab<false>;
  int &per = *mgic_dev_cache(c->device, kern, 0, 0);
  if (!per) MGIC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, BT, 0));
  long long blocks = (n + BT - 1) / BT;
  // few, fat blocks keep the grid barrier cheap; never more than can be co-resident (cooperative launch)
  long long cap = std::min<long long>((long long)per * c->numSMs, (long long)partCap / 4);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  void *args[] = {&A};
  MGIC_CUDA(cudaLaunchCooperativeKernel(kern, dim3((unsigned)blocks), dim3(BT), args, 0, c->stream));
  c->launches++;
  c->lastBottomKernel = 3;
  return MGIC_OK;
}

}  // namespace mgk
