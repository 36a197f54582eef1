// bottom_cbrick.cu -- the bottom BiCGStab for levels too big for one cluster's shared memory (the agglomerated bottom
// level of a multi-GPU run: 32x32x64 ... 64^3 cells): bottom_brick.cu's scheme with CLUSTER-sized bricks.
//
// bottom_brick.cu gives every CTA a 16x16x8 brick grown by five cells, so that a whole preconditioner application
// (lambda*r + two GSRB sweeps = four colour passes, VariableCoeffPoissonOperator::preCond) + applyOp needs no grid
// barrier; the price is that the 26x26x18 region is three times the brick, and the kernel is bound by instruction issue
// on those redundant cells (profiles/r1_bottom_brick_ncu_summary.json; 62 us per iteration on 64^3).  Here a brick is
// owned by a 16-CTA thread-block cluster: its region (e.g. 32^3 grown to 37^3 inside a 64^3 level: 1.26x redundant
// instead of 3x) lies z-plane-wise in the cluster's distributed shared memory together with its aCoef / lambda, the
// colour passes are separated by hardware cluster barriers, z-neighbours across a CTA's plane range are read from the
// neighbouring CTA's shared memory, and the coefficients are read from HBM/L2 once per solve instead of once per pass.
// Clusters meet at the same four grid-wide barriers per iteration as bottom_brick.cu.  Arithmetic and control flow are
// those of bottom.cu / the oracle (gsrb_point, lap7, the [Chombo] BiCGStabSolver restatement).
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "mgic_internal.h"
#include "mgic_device.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int HALO = 5;
constexpr int BAR_GEN = 32;  // the barrier's generation word sits 128 bytes after its arrival counter

struct CbArgs {
  Geom g;
  BCk bc;
  double alpha, beta, dxinv;
  double *phi;
  const double *rhs, *a, *b, *lam;
  double *r, *rt, *e, *p0, *p1, *v0, *v1;
  double *part;      // 2 buffers x 2 values x gridDim partials
  unsigned *bar;     // barrier of the clusters: [0] arrivals, [BAR_GEN] generation
  int nclusters;
  int nbx, nby, nbz; // bricks per direction; brick q of n cells spans [q*n/nb, (q+1)*n/nb)
  int cs, maxp;      // cluster size; most region planes a CTA holds
  unsigned stride;   // doubles per shared-memory vector
  int imax;
  double eps, reps, hang, small;
  int numRestarts;
  int *out;
  unsigned long long *dbg;  // optional phase timers of CTA 0 (MGIC_DEBUG): load, sweeps, own work, reductions
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
struct Tick {  // accumulates the time since the previous tick into slot `i`
  unsigned long long *d, t;
  __device__ Tick(unsigned long long *d_) : d(d_), t(d_ ? gtime() : 0) {}
  __device__ __forceinline__ void operator()(int i) {
    if (d && blockIdx.x == 0 && threadIdx.x == 0) { const unsigned long long n = gtime(); d[i] += n - t; t = n; }
  }
};

template <bool HAS_B, int NT>
struct Cb {
  const CbArgs &A;
  cg::cluster_group cl;
  double *S, *P, *AA, *LL, *BB, *sh;
  double *Sdn, *Sup;   // S of the CTAs holding the planes below / above this CTA's
  int nred = 0;
  int rank;
  int lo[3], hi[3], rlo[3], rhi[3], rx, ry, rz, rxy, bxl, byl;
  int z0, np, npDn;    // this CTA's region planes [z0, z0 + np) (relative to rlo[2]); plane count of the CTA below
  int kbase;           // global k of local plane 0
  int oz0, nop;        // first brick plane of this CTA, number of brick planes of this CTA
  int gsy, gsz;
  Tick *tk = nullptr;
#define TICK(i) do { if (tk) (*tk)(i); } while (0)

  __device__ Cb(const CbArgs &a_, double *smem, double *sh_) : A(a_), cl(cg::this_cluster()), sh(sh_) {
    rank = (int)cl.block_rank();
    const int b = blockIdx.x / A.cs;
    const int ib = b % A.nbx, jb = (b / A.nbx) % A.nby, kb = b / (A.nbx * A.nby);
    const int n[3] = {A.g.nx, A.g.ny, A.g.nz}, nb[3] = {A.nbx, A.nby, A.nbz}, q[3] = {ib, jb, kb};
    for (int d = 0; d < 3; d++) {
      lo[d] = (q[d] * n[d]) / nb[d]; hi[d] = ((q[d] + 1) * n[d]) / nb[d] - 1;
      rlo[d] = max(lo[d] - HALO, 0); rhi[d] = min(hi[d] + HALO, n[d] - 1);
    }
    rx = rhi[0] - rlo[0] + 1; ry = rhi[1] - rlo[1] + 1; rz = rhi[2] - rlo[2] + 1;
    rxy = rx * ry;
    bxl = hi[0] - lo[0] + 1; byl = hi[1] - lo[1] + 1;
    z0 = (rank * rz) / A.cs;
    np = ((rank + 1) * rz) / A.cs - z0;
    npDn = rank > 0 ? z0 - ((rank - 1) * rz) / A.cs : 0;
    kbase = rlo[2] + z0;
    oz0 = max(lo[2], kbase);
    const int oz1 = min(hi[2] + 1, kbase + np);
    nop = max(oz1 - oz0, 0);
    gsy = (int)A.g.sy; gsz = (int)A.g.sz;
    const size_t stride = A.stride;
    S = smem; P = S + stride; AA = P + stride; LL = AA + stride; BB = HAS_B ? LL + stride : nullptr;
    Sdn = rank > 0 ? cl.map_shared_rank(S, rank - 1) : S;
    Sup = rank + 1 < A.cs ? cl.map_shared_rank(S, rank + 1) : S;
  }

  // exact a / b for 0 <= a < 2^22, 0 < b (see bottom_brick.cu)
  static __device__ __forceinline__ int fdiv(int a, int b) { return __float2int_rz(__fdividef((float)a + 0.5f, (float)b)); }
  __device__ __forceinline__ int gidx(int i, int j, int k) const { return i + j * gsy + k * gsz; }
  __device__ __forceinline__ int sidx(int i, int j, int k) const { return (i - rlo[0]) + rx * (j - rlo[1]) + rxy * (k - kbase); }
  // All loops over cells use a flat index that is decoded per visit (two exact float divisions): measured on B200, keeping
  // every thread busy beats the cheaper plane-major walk of bottom_brick.cu here (passes 22.4 vs 24.5 us per iteration on
  // 64^3) -- with one to three visits per thread a pass is bound by the latency of its dependent FP64 chain, not by issue.
  // f(global index, region index, i, j, k) for each brick cell of this CTA that belongs to this thread
  template <class F> __device__ __forceinline__ void for_own(F f) const {
    const int nown = bxl * byl * nop;
#pragma unroll 2
    for (int q = threadIdx.x; q < nown; q += NT) {
      const int row = fdiv(q, bxl), kk = fdiv(row, byl);
      const int i = lo[0] + (q - row * bxl), j = lo[1] + (row - kk * byl), k = oz0 + kk;
      f(gidx(i, j, k), sidx(i, j, k), i, j, k);
    }
  }
  // f(global index, region index, in the brick?) for each region cell of this CTA that belongs to this thread
  template <class F> __device__ __forceinline__ void for_region(F f) const {
    const int total = np * rxy;
#pragma unroll 2
    for (int q = threadIdx.x; q < total; q += NT) {
      const int kl = fdiv(q, rxy), m = q - kl * rxy;
      const int jr = fdiv(m, rx), ir = m - jr * rx;
      const int i = rlo[0] + ir, j = rlo[1] + jr, k = kbase + kl;
      f(gidx(i, j, k), q, i >= lo[0] && i <= hi[0] && j >= lo[1] && j <= hi[1] && k >= lo[2] && k <= hi[2]);
    }
  }

  // neighbours of the region cell (i,j,k) at local index s; physical BC folded in (same rule as mgic_device.cuh);
  // the planes below / above this CTA's range live in the neighbouring CTA's shared memory
  __device__ __forceinline__ Nb nbS(int i, int j, int k, int s, double c) const {
    Nb n;
    const BCk &bc = A.bc;
    n.xm = (i > 0) ? S[s - 1] : bc.a[0] * c + bc.b[0];
    n.xp = (i < A.g.nx - 1) ? S[s + 1] : bc.a[1] * c + bc.b[1];
    n.ym = (j > 0) ? S[s - rx] : bc.a[2] * c + bc.b[2];
    n.yp = (j < A.g.ny - 1) ? S[s + rx] : bc.a[3] * c + bc.b[3];
    const int kl = k - kbase;
    if (k == 0) n.zm = bc.a[4] * c + bc.b[4];
    else if (kl > 0) n.zm = S[s - rxy];
    else n.zm = Sdn[s + rxy * (npDn - 1)];
    if (k == A.g.nz - 1) n.zp = bc.a[5] * c + bc.b[5];
    else if (kl < np - 1) n.zp = S[s + rxy];
    else n.zp = Sup[s - rxy * (np - 1)];
    return n;
  }

  // four colour passes on S (rhs P) over the brick grown by 4, 3, 2, 1 (relax(x, rhs, 2) for every cell of the brick+1),
  // each CTA on its own planes, a cluster barrier before every pass
  __device__ void sweeps() {
    const int n[3] = {A.g.nx, A.g.ny, A.g.nz};
    for (int pass = 0; pass < 4; pass++) {
      const int grow = 4 - pass, color = pass & 1;
      int slo[3], shi[3];
#pragma unroll
      for (int d = 0; d < 3; d++) { slo[d] = max(lo[d] - grow, 0); shi[d] = min(hi[d] + grow, n[d] - 1); }
      const int sxl = shi[0] - slo[0] + 1, syl = shi[1] - slo[1] + 1;
      const int hx = (sxl + 1) / 2, M = hx * syl;           // positions per plane and colour
      const int ka = max(slo[2], kbase), kb = min(shi[2], kbase + np - 1);
      const int total = max(kb - ka + 1, 0) * M;
      cl.sync();
      TICK(4);
#pragma unroll 2
      for (int q = threadIdx.x; q < total; q += NT) {
        const int pl = fdiv(q, M), m = q - pl * M;
        const int jr = fdiv(m, hx), t = m - jr * hx;
        const int j = slo[1] + jr, k = ka + pl;
        const int i = slo[0] + 2 * t + ((slo[0] + j + k + A.g.k0 + color) & 1);
        if (i > shi[0]) continue;
        const int s = sidx(i, j, k);
        const double c = S[s];
        const Nb nb = nbS(i, j, k, s, c);
        S[s] = gsrb_point<HAS_B>(c, nb.xm, nb.xp, nb.ym, nb.yp, nb.zm, nb.zp, AA[s], HAS_B ? BB[s] : 1.0, LL[s], P[s], A.alpha,
                                 A.beta, A.dxinv);
      }
      TICK(1);
    }
    cl.sync();
    TICK(4);
  }
  // VCCOMPUTEOP3D point (VariableCoeffPoissonOperatorF.ChF:209-234) from the shared-memory region
  __device__ __forceinline__ double opS(int s, int i, int j, int k) const {
    const double c = S[s];
    const Nb nb = nbS(i, j, k, s, c);
    double l = lap7(c, nb.xm, nb.xp, nb.ym, nb.yp, nb.zm, nb.zp);
    l = l * A.dxinv * A.beta;
    if (HAS_B) l = l * BB[s];
    return A.alpha * AA[s] * c - l;
  }
  // VCCOMPUTERES3D point (:312-336) straight from global memory (phi is complete when this runs)
  __device__ __forceinline__ double resG(const double *x, int q, int i, int j, int k) const {
    const double c = x[q];
    const Nb nb = neighbours(x, q, i, j, k, A.g, A.bc, c);
    double l = lap7(c, nb.xm, nb.xp, nb.ym, nb.yp, nb.zm, nb.zp);
    l = l * A.dxinv * A.beta;
    if (HAS_B) l = l * A.b[q];
    return (A.rhs[q] - A.alpha * A.a[q] * c) + l;
  }

  // grid-wide barrier (all CTAs are co-resident: the launcher checks cudaOccupancyMaxActiveClusters).  Measured on B200:
  // one thread per CTA on the global counter costs about the same ~3 us as one thread per cluster behind two
  // hardware cluster barriers, and the flat form does not add the cluster barriers' waits.
  __device__ void gsync() {
    __syncthreads();
    if (threadIdx.x == 0) {
      volatile unsigned *gen = A.bar + BAR_GEN;
      const unsigned g = *gen;
      __threadfence();
      if (atomicAdd(A.bar, 1u) == gridDim.x - 1) {
        A.bar[0] = 0;
        __threadfence();
        atomicExch(A.bar + BAR_GEN, g + 1);
      } else {
        while (*gen == g) { }
      }
      __threadfence();
    }
    __syncthreads();
  }
  // sums (v0, v1) over the grid; ONE grid barrier; identical result in every thread
  __device__ void reduce2(double &v0, double &v1) {
    double *buf = A.part + (size_t)(nred & 1) * 2 * gridDim.x;
    nred++;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { v0 += __shfl_down_sync(0xffffffffu, v0, o); v1 += __shfl_down_sync(0xffffffffu, v1, o); }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) { sh[w] = v0; sh[32 + w] = v1; }
    __syncthreads();
    if (w == 0) {
      double x0 = (l < NT / 32) ? sh[l] : 0.0, x1 = (l < NT / 32) ? sh[32 + l] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { x0 += __shfl_down_sync(0xffffffffu, x0, o); x1 += __shfl_down_sync(0xffffffffu, x1, o); }
      if (l == 0) { buf[2 * blockIdx.x] = x0; buf[2 * blockIdx.x + 1] = x1; }  // thread 0: the one that arrives at the barrier
    }
    TICK(3);
    gsync();
    TICK(6);
    if (w == 0) {
      double x0 = 0.0, x1 = 0.0;
      for (int q = l; q < (int)gridDim.x; q += 32) { x0 += *((volatile double *)&buf[2 * q]); x1 += *((volatile double *)&buf[2 * q + 1]); }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { x0 += __shfl_down_sync(0xffffffffu, x0, o); x1 += __shfl_down_sync(0xffffffffu, x1, o); }
      if (l == 0) { sh[64] = x0; sh[65] = x1; }
    }
    __syncthreads();
    v0 = sh[64]; v1 = sh[65];
  }

  // this CTA's planes of the region: S = x*lambda, P = x where x is built per cell by `make` (global index, in brick?)
  template <class F> __device__ void load_region(F make) {
    for_region([&](int g, int q, bool inBrick) {
      const double x = make(g, inBrick);
      P[q] = x;
      S[q] = x * LL[q];   // preCond: phi = rhs * lambda (VariableCoeffPoissonOperator.cpp:94-101)
    });
  }
  __device__ void load_coefs() {
    for_region([&](int g, int q, bool) {
      AA[q] = A.a[g]; LL[q] = A.lam[g];
      if (HAS_B) BB[q] = A.b[g];
    });
  }

  __device__ void solve() {
    double *phi = A.phi, *r = A.r, *rt = A.rt, *e = A.e;
    double *pin = A.p0, *pout = A.p1, *vin = A.v0, *vout = A.v1;
    Tick tick(A.dbg);
    if (A.dbg) tk = &tick;
    load_coefs();
    double s0 = 0.0, s1 = 0.0;
    // residual(r, phi, rhs, homogeneous); r_tilde = r; e = 0
    for_own([&](int q, int, int i, int j, int k) {
      const double rv = resG(phi, q, i, j, k);
      r[q] = rv; rt[q] = rv; e[q] = 0.0;
      s0 += rv * rv;
    });
    reduce2(s0, s1);
    double norm0 = sqrt(s0), norm1 = norm0;
    const double initial_norm = norm0, initial_rnorm = norm0;
    double rho1 = s0 /* dot(r_tilde, r) with r_tilde = r */, rho2 = 0.0, alpha0 = 0.0, alpha1 = 0.0, beta1 = 0.0, omega0 = 0.0, omega1 = 0.0;
    bool init = true, finished = false;
    int restarts = 0, recount = 0, status = -1, it = 0;
    while ((it < A.imax && norm0 > A.eps * norm1) && (norm1 > 0)) {
      it++;
      norm1 = norm0; alpha1 = alpha0; omega1 = omega0;
      // rho1 = dot(r_tilde, r) was reduced together with the norm of the phase that last changed r
      if (rho1 == 0.0) {
        for_own([&](int q, int, int, int, int) { phi[q] = phi[q] + 1.0 * e[q]; });
        status = 2; finished = true;
        break;
      }
      // ---- phase A: p update, p_tilde = preCond(p), v = L p_tilde, m = dot(r_tilde, v) ------------------------------
      if (init) {
        load_region([&](int q, bool own_) { const double pv = r[q]; if (own_) pout[q] = pv; return pv; });
        init = false;
      } else {
        beta1 = (rho1 / rho2) * (alpha1 / omega1);
        const double c2 = -beta1 * omega1, b1 = beta1;
        load_region([&](int q, bool own_) {
          double pv = pin[q] * b1;      // scale(p, beta)
          pv = pv + c2 * vin[q];        // incr(p, v, -beta*omega)
          pv = pv + 1.0 * r[q];         // incr(p, r, 1)
          if (own_) pout[q] = pv;
          return pv;
        });
      }
      tick(0);
      sweeps();
      tick(1);
      s0 = 0.0; s1 = 0.0;
      // (P is free once the sweeps are done: the brick cells' slots keep v, then t, for the update that follows)
      for_own([&](int q, int s, int i, int j, int k) {
        const double vv = opS(s, i, j, k);
        P[s] = vv; vout[q] = vv;
        s0 += rt[q] * vv;
      });
      tick(2);
      reduce2(s0, s1);
      tick(3);
      const double mm = s0;
      alpha0 = rho1 / mm;
      // ---- phase B: r -= alpha v, e += alpha p_tilde, |r|, next rho ------------------------------------------------
      s0 = 0.0; s1 = 0.0;
      if (fabs(mm) > A.small * fabs(rho1)) {
        const double na = -alpha0, al = alpha0;
        for_own([&](int q, int s, int, int, int) {
          const double rv = r[q] + na * P[s];
          r[q] = rv; s0 += rv * rv; s1 += rt[q] * rv;
          e[q] = e[q] + al * S[s];
        });
        tick(2);
        reduce2(s0, s1);
        tick(3);
        norm0 = sqrt(s0);
      } else {
        for_own([&](int q, int, int, int, int) { r[q] = 0.0; });
        reduce2(s0, s1);
        norm0 = 0.0;
      }
      rho2 = rho1;
      double rhoNext = s1;
      if (norm0 > A.eps * initial_norm && norm0 > A.reps * initial_rnorm) {
        // ---- phase C: s_tilde = preCond(r), t = L s_tilde, dots (t,r), (t,t) ----------------------------------------
        load_region([&](int q, bool) { return r[q]; });
        tick(0);
        sweeps();
        tick(1);
        s0 = 0.0; s1 = 0.0;
        for_own([&](int q, int s, int i, int j, int k) {
          const double tv = opS(s, i, j, k);
          P[s] = tv;
          s0 += tv * r[q]; s1 += tv * tv;
        });
        tick(2);
        reduce2(s0, s1);
        tick(3);
        omega0 = s0 / s1;
        // ---- phase D: e += omega s_tilde, r -= omega t, |r|, next rho ----------------------------------------------
        const double no = -omega0, om = omega0;
        s0 = 0.0; s1 = 0.0;
        for_own([&](int q, int s, int, int, int) {
          e[q] = e[q] + om * S[s];
          const double rv = r[q] + no * P[s];
          r[q] = rv; s0 += rv * rv; s1 += rt[q] * rv;
        });
        tick(2);
        reduce2(s0, s1);
        tick(3);
        norm0 = sqrt(s0);
        rhoNext = s1;
      }
      rho1 = rhoNext;
      { double *tq = pin; pin = pout; pout = tq; tq = vin; vin = vout; vout = tq; }
      if (norm0 <= A.eps * initial_norm || norm0 <= A.reps * initial_rnorm) { status = 1; break; }
      if (omega0 == 0.0 || norm0 > (1 - A.hang) * norm1) {
        if (recount == 0) recount = 1;
        else {
          recount = 0;
          for_own([&](int q, int, int, int, int) { phi[q] = phi[q] + 1.0 * e[q]; });
          if (restarts == A.numRestarts) { status = 3; finished = true; break; }
          gsync();
          s0 = 0.0; s1 = 0.0;
          for_own([&](int q, int, int i, int j, int k) {
            const double rv = resG(phi, q, i, j, k);
            r[q] = rv; rt[q] = rv; e[q] = 0.0;
            s0 += rv * rv;
          });
          reduce2(s0, s1);
          norm0 = sqrt(s0);
          rho1 = s0; rho2 = 0.0; alpha0 = 0.0; beta1 = 0.0; omega0 = 0.0;
          restarts++;
          init = true;
        }
      }
    }
    if (!finished) for_own([&](int q, int, int, int, int) { phi[q] = phi[q] + 1.0 * e[q]; });
    if (blockIdx.x == 0 && threadIdx.x == 0) { A.out[0] = it; A.out[1] = status; }
    cl.sync();  // no CTA may exit while a neighbour can still read its shared memory
  }
};

template <bool HAS_B, int NT>
__global__ void __launch_bounds__(NT, 1) k_bottom_cbrick(CbArgs A) {
  extern __shared__ __align__(16) double smem[];
  __shared__ double sh[72];
  Cb<HAS_B, NT> b(A, smem, sh);
  b.solve();
}

unsigned *g_bar[16] = {nullptr};  // per-device barrier words
unsigned long long *g_dbg[16] = {nullptr};

}  // namespace

namespace mgk {

// returns MGIC_OK and *used = 1 if the cluster-brick kernel ran; *used = 0 if the level does not fit it (caller falls back)
int bottom_bicgstab_cbrick(mgic_op *o, mgic_field *e, const mgic_field *r, mgic_field *const work[8], double *part, int partCap,
                           int *d_out, int *used) {
  mgic_ctx *c = o->ctx;
  *used = 0;
  const Geom g = o->geom();
  const BCk bc = o->bck(true);
  for (int f = 0; f < 6; f++)
    if (bc.type[f] != MGIC_BC_DIRICHLET && bc.type[f] != MGIC_BC_NEUMANN) return MGIC_OK;  // periodic / slab-interior: not here
  if (g.nx > 1023 || g.ny > 1023 || g.nz > 2047 || (long long)g.nx * g.ny * g.nz >= (1 << 22)) return MGIC_OK;
  if (c->device < 0 || c->device >= 16) return MGIC_OK;
  static const int NT = [] { const char *e = getenv("MGIC_CBRICK_NT"); return e && atoi(e) == 512 ? 512 : 1024; }();  // tuning; 1024 measured faster
  void (*kern)(CbArgs) = NT == 512 ? (o->b ? k_bottom_cbrick<true, 512> : k_bottom_cbrick<false, 512>)
                                   : (o->b ? k_bottom_cbrick<true, 1024> : k_bottom_cbrick<false, 1024>);
  const int nvec = o->b ? 5 : 4;
  int &npok = *mgic_dev_cache(c->device, (const void *)kern, 0, -1);
  if (npok < 0) {
    npok = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess ? 1 : 0;
    cudaGetLastError();
  }
  // Brick decomposition: any (nbx, nby, nbz) whose clusters are all co-resident (B200 places 7 clusters of 16 CTAs or 15
  // of 8), bricks of at least 8 cells per edge, every CTA of a cluster at least one plane.  A pass is bound by the cell
  // visits of one CTA, so the candidates are tried in the order of the per-CTA shared-memory vector length.
  struct Cand { int cs, nb[3], maxp, count; size_t stride; };
  std::vector<Cand> cands;
  // clusters the device can hold at once, per cluster size (1024-thread CTAs: one per SM whatever the shared memory)
  int *ma[2] = {mgic_dev_cache(c->device, (const void *)kern, 1, -1), mgic_dev_cache(c->device, (const void *)kern, 2, -1)};
  for (int ci = 0; ci < 2; ci++) {
    if (*ma[ci] >= 0) continue;
    const int cs = ci ? 8 : 16;
    *ma[ci] = 0;
    if (cs == 16 && !npok) continue;
    const size_t smem = 160 * 1024;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); continue; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(16 * cs); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem; cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess) *ma[ci] = n;
    cudaGetLastError();
  }
  const int nn[3] = {g.nx, g.ny, g.nz};
  // largest / smallest region edge over the bricks of one direction (the halo is clipped at the domain faces)
  auto rmax = [&](int d, int nb) { const int bmax = (nn[d] + nb - 1) / nb; return std::min(nn[d], bmax + (nb == 1 ? 0 : nb == 2 ? HALO : 2 * HALO)); };
  auto rmin = [&](int d, int nb) { const int bmin = nn[d] / nb; return std::min(nn[d], bmin + (nb == 1 ? 0 : HALO)); };
  for (int cs = npok ? 16 : 8; cs >= 8; cs >>= 1)
    for (int nbz = 1; nbz <= 8; nbz++)
      for (int nby = 1; nby <= 8; nby++)
        for (int nbx = 1; nbx <= 8; nbx++) {
          const int count = nbx * nby * nbz;
          if (count > *ma[cs == 16 ? 0 : 1] || nn[0] / nbx < 8 || nn[1] / nby < 8 || nn[2] / nbz < 8) continue;
          if (rmin(2, nbz) < cs) continue;
          Cand q;
          q.cs = cs; q.nb[0] = nbx; q.nb[1] = nby; q.nb[2] = nbz; q.count = count;
          q.maxp = (rmax(2, nbz) + cs - 1) / cs;
          q.stride = (size_t)q.maxp * rmax(0, nbx) * rmax(1, nby);
          if (nvec * q.stride * sizeof(double) > 200 * 1024 || 4 * count * cs > partCap) continue;
          cands.push_back(q);
        }
  std::stable_sort(cands.begin(), cands.end(), [](const Cand &x, const Cand &y) {
    return x.stride != y.stride ? x.stride < y.stride : x.count < y.count;
  });
  static const bool debug = getenv("MGIC_DEBUG") != nullptr;
  int tried = 0;
  for (const Cand &q : cands) {
    if (++tried > 12) break;  // each rejected candidate costs an occupancy query per launch
    {
      const int cs = q.cs, count = q.count, maxp = q.maxp;
      const size_t stride = q.stride, smem = nvec * stride * sizeof(double);
      const int blocks = count * cs;
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); continue; }
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)blocks); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem; cfg.stream = c->stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int nclusters = 0;
      const cudaError_t oe = cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg);
      if (debug) fprintf(stderr, "mgic cbrick: cs %d bricks %dx%dx%d smem %zu -> max active clusters %d (%s)\n", cs, q.nb[0], q.nb[1], q.nb[2], smem, nclusters, cudaGetErrorString(oe));
      if (oe != cudaSuccess || nclusters < count) { cudaGetLastError(); continue; }
      if (!g_bar[c->device]) {
        MGIC_CUDA(cudaMalloc(&g_bar[c->device], 2 * BAR_GEN * sizeof(unsigned)));
        MGIC_CUDA(cudaMemsetAsync(g_bar[c->device], 0, 2 * BAR_GEN * sizeof(unsigned), c->stream));
      }
      CbArgs A;
      A.g = g; A.bc = bc;
      A.alpha = o->alpha; A.beta = o->beta; A.dxinv = 1.0 / (o->dx * o->dx);
      A.phi = e->p; A.rhs = r->p; A.a = o->a->p; A.b = o->b ? o->b->p : nullptr; A.lam = o->lambda->p;
      A.r = work[0]->p; A.rt = work[1]->p; A.e = work[2]->p; A.p0 = work[3]->p; A.p1 = work[4]->p; A.v0 = work[5]->p; A.v1 = work[6]->p;
      A.part = part; A.bar = g_bar[c->device]; A.nclusters = count;
      A.nbx = q.nb[0]; A.nby = q.nb[1]; A.nbz = q.nb[2];
      A.cs = cs; A.maxp = maxp; A.stride = (unsigned)stride;
      A.imax = 80; A.eps = 1.0e-6; A.reps = 1.0e-12; A.hang = 1.0e-8; A.small = 1.0e-30; A.numRestarts = 5;
      A.out = d_out;
      A.dbg = nullptr;
      if (debug) {
        if (!g_dbg[c->device]) MGIC_CUDA(cudaMalloc(&g_dbg[c->device], 8 * sizeof(unsigned long long)));
        MGIC_CUDA(cudaMemsetAsync(g_dbg[c->device], 0, 8 * sizeof(unsigned long long), c->stream));
        A.dbg = g_dbg[c->device];
      }
      MGIC_CUDA(cudaLaunchKernelEx(&cfg, kern, A));
      c->launches++;
      *used = 1;
      if (debug) {
        unsigned long long h[8];
        int ho[2];
        MGIC_CUDA(cudaStreamSynchronize(c->stream));
        MGIC_CUDA(cudaMemcpy(h, A.dbg, sizeof(h), cudaMemcpyDeviceToHost));
        MGIC_CUDA(cudaMemcpy(ho, d_out, sizeof(ho), cudaMemcpyDeviceToHost));
        fprintf(stderr, "mgic cbrick: %d its; CTA 0 us: load %.1f passes %.1f (cluster-barrier waits %.1f) own %.1f block-reduce %.1f "
                        "grid-barrier %.1f\n", ho[0], h[0] * 1e-3, h[1] * 1e-3, h[4] * 1e-3, h[2] * 1e-3, h[3] * 1e-3, h[6] * 1e-3);
      }
      return MGIC_OK;
    }
  }
  return MGIC_OK;
}

}  // namespace mgk
