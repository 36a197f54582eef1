// bottom_brick.cu -- the bottom BiCGStab with FOUR grid barriers per iteration instead of fifteen.
//
// bottom.cu runs [Chombo] BiCGStabSolver::solve in one kernel, but every colour pass of the preconditioner
// (VariableCoeffPoissonOperator::preCond = lambda*r, then two GSRB sweeps) and every dot product ends in a grid-wide
// barrier, and on a level this small the barrier (2-3 us) IS the cost.  Here each CTA owns a brick of the level and
// evaluates a whole preconditioner application + applyOp without leaving the SM: it loads the brick grown by five cells
// into shared memory, performs the four colour passes on regions that shrink by one cell per pass (the ring is recomputed
// redundantly by the neighbouring CTAs -- same formula on the same inputs, so the same bits), applies the operator on
// the brick and reduces its share of the dot product.  Per iteration: {p update, preCond, applyOp, dot} | {r, e update,
// norm} | {preCond, applyOp, two dots} | {e, r update, norm, next rho}.  p_tilde, s_tilde and t never touch global memory.
// Arithmetic and control flow are those of bottom.cu / the oracle (gsrb_point, lap7, the BiCGStab restatement).
#include <cooperative_groups.h>

#include "mgic_internal.h"
#include "mgic_device.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int NT = 1024;
constexpr int HALO = 5;
constexpr int MAXOWN = 2;  // brick cells per thread

struct BrickArgs {
  Geom g;
  BCk bc;
  double alpha, beta, dxinv;
  double *phi;
  const double *rhs, *a, *b, *lam;
  double *r, *rt, *e, *p0, *p1, *v0, *v1;
  double *part;  // 2 buffers x 2 values x gridDim partials
  int bx, by, bz, nbx, nby;
  int imax;
  double eps, reps, hang, small;
  int numRestarts;
  int *out;
};

template <bool HAS_B>
struct Brick {
  const BrickArgs &A;
  cg::grid_group grid;
  double *S, *P, *sh;
  int nred = 0;
  // brick / region geometry
  int lo[3], hi[3], rlo[3], rhi[3], rx, ry, rz, rxy, bxl, byl, bzl, nown;
  bool touches;  // the brick grown by four cells reaches a physical face
  double own[MAXOWN];  // v / t of the thread's brick cells, kept across a barrier
  int ownq[MAXOWN], owns[MAXOWN], ownijk[MAXOWN];  // their global index, region index and packed (i,j,k): decoded once

  __device__ Brick(const BrickArgs &a_, double *smem, double *sh_) : A(a_), grid(cg::this_grid()), sh(sh_) {
    const int b = blockIdx.x;
    const int ib = b % A.nbx, jb = (b / A.nbx) % A.nby, kb = b / (A.nbx * A.nby);
    const int n[3] = {A.g.nx, A.g.ny, A.g.nz}, bs[3] = {A.bx, A.by, A.bz}, q[3] = {ib, jb, kb};
    for (int d = 0; d < 3; d++) {
      lo[d] = q[d] * bs[d]; hi[d] = min(lo[d] + bs[d], n[d]) - 1;
      rlo[d] = max(lo[d] - HALO, 0); rhi[d] = min(hi[d] + HALO, n[d] - 1);
    }
    rx = rhi[0] - rlo[0] + 1; ry = rhi[1] - rlo[1] + 1; rz = rhi[2] - rlo[2] + 1;
    rxy = rx * ry;
    touches = false;
    for (int d = 0; d < 3; d++) touches = touches || lo[d] - HALO < 0 || hi[d] + HALO > n[d] - 1;
    bxl = hi[0] - lo[0] + 1; byl = hi[1] - lo[1] + 1; bzl = hi[2] - lo[2] + 1;
    nown = bxl * byl * bzl;
    gsy = (int)A.g.sy; gsz = (int)A.g.sz;
    S = smem; P = smem + (size_t)rx * ry * rz;
  }

  // exact a / b for 0 <= a < 2^22, 0 < b: (a + 0.5) / b is at least 0.5/b away from an integer, far more than the error
  // of the approximate float division (integer division has no hardware instruction; it dominated this kernel)
  static __device__ __forceinline__ int fdiv(int a, int b) { return __float2int_rz(__fdividef((float)a + 0.5f, (float)b)); }
  int gsy, gsz;  // the level is tiny: 32-bit global indices
  __device__ __forceinline__ int gidx(int i, int j, int k) const { return i + j * gsy + k * gsz; }
  __device__ __forceinline__ int sidx(int i, int j, int k) const { return (i - rlo[0]) + rx * ((j - rlo[1]) + ry * (k - rlo[2])); }
  __device__ __forceinline__ void own_cell(int m, int &i, int &j, int &k) const {  // m-th brick cell of this thread
    const int q = threadIdx.x + m * NT;
    const int row = fdiv(q, bxl), kk = fdiv(row, byl);
    i = lo[0] + (q - row * bxl); j = lo[1] + (row - kk * byl); k = lo[2] + kk;
  }
  __device__ __forceinline__ int nmine() const { return (nown - (int)threadIdx.x + NT - 1) / NT; }
  __device__ void decode_own() {
#pragma unroll
    for (int m = 0; m < MAXOWN; m++) {
      int i = lo[0], j = lo[1], k = lo[2];
      if (m < nmine()) own_cell(m, i, j, k);
      ownq[m] = gidx(i, j, k); owns[m] = sidx(i, j, k); ownijk[m] = i | (j << 10) | (k << 20);
    }
  }
#define FOR_OWN(m) _Pragma("unroll") for (int m = 0; m < MAXOWN; m++) if (m < mine)

  // neighbours of (i,j,k) from the shared-memory region, physical BC folded in (same rule as mgic_device.cuh)
  __device__ __forceinline__ Nb nbS(int i, int j, int k, int s, double c) const {
    Nb n;
    const BCk &bc = A.bc;
    n.xm = (i > 0) ? S[s - 1] : bc.a[0] * c + bc.b[0];
    n.xp = (i < A.g.nx - 1) ? S[s + 1] : bc.a[1] * c + bc.b[1];
    n.ym = (j > 0) ? S[s - rx] : bc.a[2] * c + bc.b[2];
    n.yp = (j < A.g.ny - 1) ? S[s + rx] : bc.a[3] * c + bc.b[3];
    n.zm = (k > 0) ? S[s - rx * ry] : bc.a[4] * c + bc.b[4];
    n.zp = (k < A.g.nz - 1) ? S[s + rx * ry] : bc.a[5] * c + bc.b[5];
    return n;
  }

  // one GSRB point of the region; BND = the region touches a physical face (otherwise no boundary code is compiled in)
  template <bool BND>
  __device__ __forceinline__ void relax_point(int i, int j, int k, int s, int q) {
    const double c = S[s];
    Nb nb;
    if (BND) nb = nbS(i, j, k, s, c);
    else { nb.xm = S[s - 1]; nb.xp = S[s + 1]; nb.ym = S[s - rx]; nb.yp = S[s + rx]; nb.zm = S[s - rxy]; nb.zp = S[s + rxy]; }
    S[s] = gsrb_point<HAS_B>(c, nb.xm, nb.xp, nb.ym, nb.yp, nb.zm, nb.zp, __ldg(A.a + q), HAS_B ? __ldg(A.b + q) : 1.0,
                             __ldg(A.lam + q), P[s], A.alpha, A.beta, A.dxinv);
  }

  // four colour passes on S (rhs P) over the brick grown by 4, 3, 2, 1: relax(x, rhs, 2) for every cell of the brick+1.
  // Thread mapping is plane-major: a thread keeps one (x half-index, y) position of the pass's sub-region and walks
  // over z planes, so a visit costs two adds and a parity flip instead of an index decode (this kernel is bound by
  // instruction issue on one SM per brick: 177 instructions per visit before, FP64 14 % of them).
  template <bool BND>
  __device__ void sweeps_t() {
    const int n[3] = {A.g.nx, A.g.ny, A.g.nz};
    for (int pass = 0; pass < 4; pass++) {
      const int grow = 4 - pass, color = pass & 1;
      int slo[3], shi[3];
#pragma unroll
      for (int d = 0; d < 3; d++) { slo[d] = max(lo[d] - grow, 0); shi[d] = min(hi[d] + grow, n[d] - 1); }
      const int sxl = shi[0] - slo[0] + 1, syl = shi[1] - slo[1] + 1;
      const int hx = (sxl + 1) / 2, M = hx * syl;          // positions per plane and colour
      const int G = max(NT / M, 1);                         // planes swept concurrently
      __syncthreads();
      for (int m0 = 0; m0 < M; m0 += NT) {                  // (one trip unless a plane has more positions than threads)
        const int g = fdiv((int)threadIdx.x, M), m = m0 + (int)threadIdx.x - g * M;
        if (g >= G || m >= M) continue;
        const int jr = fdiv(m, hx), t = m - jr * hx;
        const int j = slo[1] + jr, i0 = slo[0] + 2 * t;
        const int par0 = (slo[0] + j + A.g.k0 + color) & 1;
        const int srow = (i0 - rlo[0]) + rx * (j - rlo[1]), qrow = i0 + j * gsy;
        for (int k = slo[2] + g; k <= shi[2]; k += G) {
          const int par = (par0 + k) & 1;
          const int i = i0 + par;
          if (i > shi[0]) continue;
          relax_point<BND>(i, j, k, srow + par + rxy * (k - rlo[2]), qrow + par + gsz * k);
        }
      }
    }
    __syncthreads();
  }
  __device__ void sweeps() {
    if (touches) sweeps_t<true>();
    else sweeps_t<false>();
  }
  // VCCOMPUTEOP3D point (VariableCoeffPoissonOperatorF.ChF:209-234) from the shared-memory region
  __device__ __forceinline__ double opS(int i, int j, int k) const {
    const int s = sidx(i, j, k);
    const int q = gidx(i, j, k);
    const double c = S[s];
    const Nb nb = nbS(i, j, k, s, c);
    double l = lap7(c, nb.xm, nb.xp, nb.ym, nb.yp, nb.zm, nb.zp);
    l = l * A.dxinv * A.beta;
    if (HAS_B) l = l * __ldg(A.b + q);
    return A.alpha * __ldg(A.a + q) * c - l;
  }
  // VCCOMPUTERES3D point (:312-336) straight from global memory (phi is complete when this runs)
  __device__ __forceinline__ double resG(const double *x, int i, int j, int k) const {
    const int q = gidx(i, j, k);
    const double c = x[q];
    const Nb nb = neighbours(x, q, i, j, k, A.g, A.bc, c);
    double l = lap7(c, nb.xm, nb.xp, nb.ym, nb.yp, nb.zm, nb.zp);
    l = l * A.dxinv * A.beta;
    if (HAS_B) l = l * A.b[q];
    return (A.rhs[q] - A.alpha * A.a[q] * c) + l;
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
      if (l == 0) { buf[2 * blockIdx.x] = x0; buf[2 * blockIdx.x + 1] = x1; }
    }
    if (gridDim.x > 1) grid.sync(); else __syncthreads();
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
  __device__ void gsync() { if (gridDim.x > 1) grid.sync(); else __syncthreads(); }

  // load the region: S = x*lambda, P = x where x is built per cell by `make` (a lambda over the global index).
  // Plane-major like the sweeps: a thread keeps one (x, y) position and walks over z.
  template <class F> __device__ void load_region(F make) {
    const int M = rxy, G = max(NT / M, 1);
    __syncthreads();
    for (int m0 = 0; m0 < M; m0 += NT) {
      const int g = fdiv((int)threadIdx.x, M), m = m0 + (int)threadIdx.x - g * M;
      if (g >= G || m >= M) continue;
      const int jr = fdiv(m, rx), ir = m - jr * rx;
      const int i = rlo[0] + ir, j = rlo[1] + jr;
      const bool mineXY = (i >= lo[0] && i <= hi[0] && j >= lo[1] && j <= hi[1]);
      const int qrow = i + j * gsy;
      for (int k = rlo[2] + g; k <= rhi[2]; k += G) {
        const int q = qrow + gsz * k, sx = m + rxy * (k - rlo[2]);
        const double x = make(q, mineXY && k >= lo[2] && k <= hi[2]);
        P[sx] = x;
        S[sx] = x * __ldg(A.lam + q);   // preCond: phi = rhs * lambda (VariableCoeffPoissonOperator.cpp:94-101)
      }
    }
  }

  __device__ void solve() {
    double *phi = A.phi, *r = A.r, *rt = A.rt, *e = A.e;
    double *pin = A.p0, *pout = A.p1, *vin = A.v0, *vout = A.v1;
    const int mine = nmine();
    decode_own();
    double s0 = 0.0, s1 = 0.0;
    // residual(r, phi, rhs, homogeneous); r_tilde = r; e = 0
    FOR_OWN(m) {
      const int q = ownq[m], i = ownijk[m] & 1023, j = (ownijk[m] >> 10) & 1023, k = ownijk[m] >> 20;
      const double rv = resG(phi, i, j, k);
      r[q] = rv; rt[q] = rv; e[q] = 0.0;
      s0 += rv * rv;
    }
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
        FOR_OWN(m) { const int q = ownq[m]; phi[q] = phi[q] + 1.0 * e[q]; }
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
      sweeps();
      s0 = 0.0; s1 = 0.0;
      FOR_OWN(m) {
        const int q = ownq[m];
        const double vv = opS(ownijk[m] & 1023, (ownijk[m] >> 10) & 1023, ownijk[m] >> 20);
        own[m] = vv; vout[q] = vv;
        s0 += rt[q] * vv;
      }
      reduce2(s0, s1);
      const double mm = s0;
      alpha0 = rho1 / mm;
      // ---- phase B: r -= alpha v, e += alpha p_tilde, |r|, next rho ------------------------------------------------
      s0 = 0.0; s1 = 0.0;
      if (fabs(mm) > A.small * fabs(rho1)) {
        const double na = -alpha0;
        FOR_OWN(m) {
          const int q = ownq[m];
          const double rv = r[q] + na * own[m];
          r[q] = rv; s0 += rv * rv; s1 += rt[q] * rv;
          e[q] = e[q] + alpha0 * S[owns[m]];
        }
        reduce2(s0, s1);
        norm0 = sqrt(s0);
      } else {
        FOR_OWN(m) { r[ownq[m]] = 0.0; }
        reduce2(s0, s1);
        norm0 = 0.0;
      }
      rho2 = rho1;
      double rhoNext = s1;
      if (norm0 > A.eps * initial_norm && norm0 > A.reps * initial_rnorm) {
        // ---- phase C: s_tilde = preCond(r), t = L s_tilde, dots (t,r), (t,t) ----------------------------------------
        load_region([&](int q, bool) { return r[q]; });
        sweeps();
        s0 = 0.0; s1 = 0.0;
        FOR_OWN(m) {
          const int q = ownq[m];
          const double tv = opS(ownijk[m] & 1023, (ownijk[m] >> 10) & 1023, ownijk[m] >> 20);
          own[m] = tv;
          s0 += tv * r[q]; s1 += tv * tv;
        }
        reduce2(s0, s1);
        omega0 = s0 / s1;
        // ---- phase D: e += omega s_tilde, r -= omega t, |r|, next rho ----------------------------------------------
        const double no = -omega0;
        s0 = 0.0; s1 = 0.0;
        FOR_OWN(m) {
          const int q = ownq[m];
          e[q] = e[q] + omega0 * S[owns[m]];
          const double rv = r[q] + no * own[m];
          r[q] = rv; s0 += rv * rv; s1 += rt[q] * rv;
        }
        reduce2(s0, s1);
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
          FOR_OWN(m) { const int q = ownq[m]; phi[q] = phi[q] + 1.0 * e[q]; }
          if (restarts == A.numRestarts) { status = 3; finished = true; break; }
          gsync();
          s0 = 0.0; s1 = 0.0;
          FOR_OWN(m) {
            const int q = ownq[m], i = ownijk[m] & 1023, j = (ownijk[m] >> 10) & 1023, k = ownijk[m] >> 20;
            const double rv = resG(phi, i, j, k);
            r[q] = rv; rt[q] = rv; e[q] = 0.0;
            s0 += rv * rv;
          }
          reduce2(s0, s1);
          norm0 = sqrt(s0);
          rho1 = s0; rho2 = 0.0; alpha0 = 0.0; beta1 = 0.0; omega0 = 0.0;
          restarts++;
          init = true;
        }
      }
    }
    if (!finished)
      FOR_OWN(m) { const int q = ownq[m]; phi[q] = phi[q] + 1.0 * e[q]; }
    if (blockIdx.x == 0 && threadIdx.x == 0) { A.out[0] = it; A.out[1] = status; }
  }
};

template <bool HAS_B>
__global__ void __launch_bounds__(NT) k_bottom_brick(BrickArgs A) {
  extern __shared__ __align__(16) double smem[];
  __shared__ double sh[66];
  Brick<HAS_B> b(A, smem, sh);
  b.solve();
}

}  // namespace

namespace mgk {

// returns MGIC_OK and *used = 1 if the brick kernel ran; *used = 0 if the level does not fit it (caller falls back)
int bottom_bicgstab_brick(mgic_op *o, mgic_field *e, const mgic_field *r, mgic_field *const work[8], double *part, int partCap,
                          int *d_out, int *used) {
  mgic_ctx *c = o->ctx;
  *used = 0;
  const Geom g = o->geom();
  const BCk bc = o->bck(true);
  for (int f = 0; f < 6; f++)
    if (bc.type[f] != MGIC_BC_DIRICHLET && bc.type[f] != MGIC_BC_NEUMANN) return MGIC_OK;  // periodic / slab-interior: not here
  // brick shape: the whole level if it is tiny, else 8^3 bricks while they all fit on the GPU at once, else 16x16x8
  int bx, by, bz;
  const long long n = (long long)g.nx * g.ny * g.nz;
  if (n <= (long long)NT * MAXOWN && (long long)g.nx * g.ny * g.nz <= 12168) { bx = g.nx; by = g.ny; bz = g.nz; }
  else {
    bx = by = bz = 8;
    auto count = [&](int x, int y, int z) { return (long long)((g.nx + x - 1) / x) * ((g.ny + y - 1) / y) * ((g.nz + z - 1) / z); };
    if (count(bx, by, bz) > 128) { bx = 16; by = 16; bz = 8; }
    if (count(bx, by, bz) > 128) return MGIC_OK;
  }
  const int nbx = (g.nx + bx - 1) / bx, nby = (g.ny + by - 1) / by, nbz = (g.nz + bz - 1) / bz;
  const int blocks = nbx * nby * nbz;
  if (4 * blocks > partCap) return MGIC_OK;
  auto ext = [](int b, int nn) { return std::min(b + 2 * HALO, nn); };
  const size_t smem = 2 * (size_t)ext(bx, g.nx) * ext(by, g.ny) * ext(bz, g.nz) * sizeof(double);
  void *kern = o->b ? (void *)k_bottom_brick<true> : (void *)k_bottom_brick<false>;
  static size_t smemSet[2] = {0, 0};
  size_t &ss = smemSet[o->b ? 1 : 0];
  if (smem > ss) {
    MGIC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ss = smem;
  }
  int per = 0;
  MGIC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, NT, smem));
  if ((long long)per * c->numSMs < blocks) return MGIC_OK;  // cooperative launch needs all bricks co-resident
  BrickArgs A;
  A.g = g; A.bc = bc;
  A.alpha = o->alpha; A.beta = o->beta; A.dxinv = 1.0 / (o->dx * o->dx);
  A.phi = e->p; A.rhs = r->p; A.a = o->a->p; A.b = o->b ? o->b->p : nullptr; A.lam = o->lambda->p;
  A.r = work[0]->p; A.rt = work[1]->p; A.e = work[2]->p; A.p0 = work[3]->p; A.p1 = work[4]->p; A.v0 = work[5]->p; A.v1 = work[6]->p;
  A.part = part;
  A.bx = bx; A.by = by; A.bz = bz; A.nbx = nbx; A.nby = nby;
  A.imax = 80; A.eps = 1.0e-6; A.reps = 1.0e-12; A.hang = 1.0e-8; A.small = 1.0e-30; A.numRestarts = 5;
  A.out = d_out;
  void *args[] = {&A};
  MGIC_CUDA(cudaLaunchCooperativeKernel(kern, dim3((unsigned)blocks), dim3(NT), args, smem, c->stream));
  c->launches++;
  *used = 1;
  return MGIC_OK;
}

}  // namespace mgk
