// restrict_tma.cu -- restrictResidual (RESTRICTRESVC3D, VariableCoeffPoissonOperatorF.ChF:379-437) as a plane-streaming
// kernel: resC = 8-cell average of rhs - L(phi), each fine array read from HBM once through TMA-staged planes.
//
// k_restrict (kernels.cu) gives one thread a coarse cell: 72 scalar loads per thread, half of each sector per load
// instruction, 62 registers at half occupancy -- 0.60 of the DRAM throughput under ncu where the fused sweep reaches 0.80
// (profiles/r2_other_kernels_ncu_summary.json).  Here a CTA owns an x-y tile of TX x TY fine cells and marches through a
// chunk of z planes like the sweep does (gsrb_fused.cu):
//   * per plane one slot = the phi plane with one halo row above / below ((TY+2) x 64 doubles) plus the TY-row planes of aCoef
//     and rhs (and bCoef), staged by cp.async.bulk.tensor and signalled through one mbarrier, a ring of five planes (k-1, k, k+1 in use, two in flight);
//   * a warp is one row, a lane one x-pair = the two fine cells of ONE coarse cell's row: the residual of both cells from
//     shared memory (y and z neighbours), a shuffle (x neighbours) and the folded boundary condition;
//   * the eight contributions of a coarse cell arrive in the Fortran loop's order -- (0,0,0), (1,0,0), (0,1,0), (1,1,0),
//     (0,0,1), ... -- in the even row's warp: its own pair, then the odd row's pair (handed over through shared memory), on
//     the even plane and again on the odd plane, starting from the zero the caller stored (Operator.cpp:177).  Same
//     operations in the same order as k_restrict: bit-identical.
// Rectangular, non-periodic levels with even nx; everything else stays with k_restrict.
#include "mgic_internal.h"
#include "mgic_device.cuh"
#include "tma.cuh"

namespace {

constexpr int RTX = 60, RRW = 64, RNS = 5;   // output cells per row, staged row width, slots in the ring

struct RestrictArgs {
  Geom g;
  BCk bc;
  double *resC;
  long long csy, csz;
  double alpha, beta, dxinv;
  int zchunk;
};

template <int TY, bool HAS_B>
struct RestrictT {
  static constexpr int RR = TY + 2, PLANE = RRW * RR, CPLANE = RRW * TY, NCOEF = HAS_B ? 3 : 2;
  static constexpr int SLOT = PLANE + NCOEF * CPLANE;                       // doubles
  static constexpr uint32_t PLANE_BYTES = PLANE * sizeof(double), CPLANE_BYTES = CPLANE * sizeof(double);
  static constexpr int NT = 32 * TY;
  static constexpr size_t XBUF = (size_t)2 * (TY / 2) * 32 * sizeof(double2);   // the odd rows' pairs, two plane parities
  static constexpr size_t SMEM = (size_t)RNS * SLOT * sizeof(double) + XBUF + RNS * sizeof(uint64_t);
  static_assert(TY % 2 == 0, "rows pair up into coarse rows");
  static_assert(PLANE_BYTES % 128 == 0 && CPLANE_BYTES % 128 == 0, "TMA destinations must stay 128-byte aligned");
};

template <int TY, bool HAS_B>
__global__ void __launch_bounds__(32 * TY, 2)
k_restrict_tma(const __grid_constant__ CUtensorMap tm_phi, const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_r,
               const __grid_constant__ CUtensorMap tm_b, const RestrictArgs A) {
  using R = RestrictT<TY, HAS_B>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *slots = reinterpret_cast<double *>(smem_raw);
  double2 *xbuf = reinterpret_cast<double2 *>(smem_raw + (size_t)RNS * R::SLOT * sizeof(double));
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)RNS * R::SLOT * sizeof(double) + R::XBUF);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int x0 = blockIdx.x * RTX, y0 = blockIdx.y * TY;
  const int zs = blockIdx.z * A.zchunk, ze = min(zs + A.zchunk, A.g.nz);
  const int pfirst = zs - 1, plast = ze;                                    // planes staged: the chunk and one plane either side
  if (tid == 0) {
    for (int q = 0; q < RNS; q++) mbar_init(&full[q], 1);
    fence_mbar_init();
  }
  __syncthreads();
  // plane p goes into slot (p - pfirst) % RNS; the halo planes carry phi only
  auto issue = [&](int p) {
    const int slot = (p - pfirst) % RNS;
    double *d = slots + (size_t)slot * R::SLOT;
    const bool coefs = p >= zs && p < ze;
    mbar_arrive_expect_tx(&full[slot], R::PLANE_BYTES + (coefs ? R::NCOEF * R::CPLANE_BYTES : 0u));
    const int z = p + MGIC_GZ;
    tma_load_3d(d, &tm_phi, &full[slot], x0 - 2, y0 - 1, z);
    if (coefs) {
      tma_load_3d(d + R::PLANE, &tm_a, &full[slot], x0 - 2, y0, z);
      tma_load_3d(d + R::PLANE + R::CPLANE, &tm_r, &full[slot], x0 - 2, y0, z);
      if (HAS_B) tma_load_3d(d + R::PLANE + 2 * R::CPLANE, &tm_b, &full[slot], x0 - 2, y0, z);
    }
  };
  if (tid == 0)
    for (int p = pfirst; p < pfirst + RNS && p <= plast; p++) issue(p);
  const int x = x0 - 2 + 2 * lane, y = y0 + w;
  const bool inDom = x >= 0 && x + 1 < A.g.nx && y < A.g.ny;
  const bool out = inDom && lane >= 1 && lane <= 30;
  const bool bx0 = (x == 0), bxn = (x == A.g.nx - 2), by0 = (y == 0), byn = (y == A.g.ny - 1);
  const bool anyxy = bx0 || bxn || by0 || byn;
  const bool zloPhys = A.bc.type[4] != MGIC_FACE_INTERIOR, zhiPhys = A.bc.type[5] != MGIC_FACE_INTERIOR;
  const int s = (w + 1) * RRW + 2 * lane, cidx = w * RRW + 2 * lane;
  const BCk &bc = A.bc;
  constexpr unsigned FULL = 0xffffffffu;
  double acc = 0.0;
  for (int k = zs; k < ze; k++) {
    const int rel = k - pfirst;                                             // >= 1
    const int sm = (rel - 1) % RNS, sc = rel % RNS, sp = (rel + 1) % RNS;
    if (k == zs) {   // the two older planes were waited for in the previous steps
      mbar_wait(&full[sm], ((rel - 1) / RNS) & 1);
      mbar_wait(&full[sc], (rel / RNS) & 1);
    }
    mbar_wait(&full[sp], ((rel + 1) / RNS) & 1);
    const double *pm = slots + (size_t)sm * R::SLOT, *pc = slots + (size_t)sc * R::SLOT, *pp = slots + (size_t)sp * R::SLOT;
    const double2 c = *reinterpret_cast<const double2 *>(pc + s);
    double2 zm = *reinterpret_cast<const double2 *>(pm + s), zp = *reinterpret_cast<const double2 *>(pp + s);
    double2 ym = *reinterpret_cast<const double2 *>(pc + s - RRW), yp = *reinterpret_cast<const double2 *>(pc + s + RRW);
    // x neighbours: inside the pair, and the neighbouring lanes' near elements
    double xm0 = __shfl_up_sync(FULL, c.y, 1), xp1 = __shfl_down_sync(FULL, c.x, 1);
    double xp0 = c.y, xm1 = c.x;
    if (anyxy) {
      if (bx0) xm0 = bc.a[0] * c.x + bc.b[0];
      if (bxn) xp1 = bc.a[1] * c.y + bc.b[1];
      if (by0) { ym.x = bc.a[2] * c.x + bc.b[2]; ym.y = bc.a[2] * c.y + bc.b[2]; }
      if (byn) { yp.x = bc.a[3] * c.x + bc.b[3]; yp.y = bc.a[3] * c.y + bc.b[3]; }
    }
    if (k == 0 && zloPhys) { zm.x = bc.a[4] * c.x + bc.b[4]; zm.y = bc.a[4] * c.y + bc.b[4]; }
    if (k == A.g.nz - 1 && zhiPhys) { zp.x = bc.a[5] * c.x + bc.b[5]; zp.y = bc.a[5] * c.y + bc.b[5]; }
    const double2 av = *reinterpret_cast<const double2 *>(pc + R::PLANE + cidx);
    const double2 rv = *reinterpret_cast<const double2 *>(pc + R::PLANE + R::CPLANE + cidx);
    double2 bv = make_double2(1.0, 1.0);
    if (HAS_B) bv = *reinterpret_cast<const double2 *>(pc + R::PLANE + 2 * R::CPLANE + cidx);
    double2 t;
    {
      double lof = A.alpha * av.x * c.x;                                    // :411-412
      double l = lap7(c.x, xm0, xp0, ym.x, yp.x, zm.x, zp.x);
      l = l * A.dxinv * A.beta;                                             // :427
      if (HAS_B) l = l * bv.x;
      lof = lof - l;                                                        // :429
      t.x = (rv.x - lof) / 8.0;                                             // :431-432
    }
    {
      double lof = A.alpha * av.y * c.y;
      double l = lap7(c.y, xm1, xp1, ym.y, yp.y, zm.y, zp.y);
      l = l * A.dxinv * A.beta;
      if (HAS_B) l = l * bv.y;
      lof = lof - l;
      t.y = (rv.y - lof) / 8.0;
    }
    double2 *xb = xbuf + (size_t)(k & 1) * (TY / 2) * 32;
    if (w & 1) xb[(w >> 1) * 32 + lane] = t;                                // the odd row hands its pair to the even row's warp
    __syncthreads();                                                        // ... and every warp is done with plane k-1's slot
    if (tid == 0 && k - 1 + RNS <= plast) issue(k - 1 + RNS);
    if (!(w & 1)) {
      const double2 u = xb[(w >> 1) * 32 + lane];
      if (!(k & 1)) acc = 0.0;
      acc = acc + t.x; acc = acc + t.y; acc = acc + u.x; acc = acc + u.y;   // (di, dj) = (0,0), (1,0), (0,1), (1,1) of plane dk
      if ((k & 1) && out) A.resC[(x >> 1) + (long long)(y >> 1) * A.csy + (long long)(k >> 1) * A.csz] = acc;
    }
  }
}

template <int TY, bool HAS_B>
int launch_restrict(mgic_op *o, const BCk &bc, mgic_field *resC, const mgic_field *phi, const mgic_field *rhs) {
  using R = RestrictT<TY, HAS_B>;
  mgic_ctx *c = o->ctx;
  auto kern = k_restrict_tma<TY, HAS_B>;
  int &resident = *mgic_dev_cache(c->device, (const void *)kern, 0, 0);   // per device: the opt-in and the occupancy
  if (!resident) {
    MGIC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)R::SMEM));
    int per = 1;
    MGIC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, R::NT, R::SMEM));
    resident = (per < 1 ? 1 : per) * c->numSMs;
  }
  RestrictArgs A;
  A.g = o->geom();
  A.bc = bc;
  A.resC = resC->p; A.csy = resC->sy; A.csz = resC->sz;
  A.alpha = o->alpha; A.beta = o->beta; A.dxinv = 1.0 / (o->dx * o->dx);
  const int np = A.g.nz + 2 * MGIC_GZ;
  const long long goff = (long long)MGIC_GZ * A.g.sz;
  CUtensorMap tp, ta, tr, tb;
  MGIC_TRY(make_tmap(&tp, phi->p - goff, A.g.nx, A.g.ny, np, RRW, R::RR));
  MGIC_TRY(make_tmap(&ta, o->a->p - goff, A.g.nx, A.g.ny, np, RRW, TY));
  MGIC_TRY(make_tmap(&tr, rhs->p - goff, A.g.nx, A.g.ny, np, RRW, TY));
  if (HAS_B) MGIC_TRY(make_tmap(&tb, o->b->p - goff, A.g.nx, A.g.ny, np, RRW, TY));
  else tb = ta;
  const int tilesX = (A.g.nx + RTX - 1) / RTX, tilesY = (A.g.ny + TY - 1) / TY;
  // z chunks of an even number of planes (a coarse plane is two fine ones): plan on plane pairs
  const Plan pl = plan_chunks(tilesX * tilesY, A.g.nz / 2, resident);
  A.zchunk = 2 * pl.zchunk;
  dim3 grd(tilesX, tilesY, pl.nch);
  kern<<<grd, R::NT, R::SMEM, c->stream>>>(tp, ta, tr, tb, A);
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { mgic_set_error("kernel restrict_tma: %s", cudaGetErrorString(e)); return MGIC_ERR_CUDA; }
  return MGIC_OK;
}

}  // namespace

namespace mgk {

bool restrict_tma_applicable(const mgic_op *o) {
  if (!o->ctx->restrictTma || o->isPatch || o->mask) return false;
  for (int d = 0; d < 3; d++)
    if (o->bc_lo[d] == MGIC_BC_PERIODIC) return false;
  // measured per level: 512^3 525 vs 688 us under ncu; 256^3: the level's restrictions 123 vs 135 us per V-cycle
  // (bench.py --n 256, profiles/r3s_*); 128^3 no gain (21 vs 18 us with the kernel's first version) -- from 256^3 on
  const long long cells = (long long)o->n[0] * o->n[1] * o->nzl;
  return !(o->n[0] & 1) && !(o->n[1] & 1) && !(o->nzl & 1) && o->n[0] >= 8 && cells >= 8 * o->ctx->fusedMinCells;
}

// resC = restriction of rhs - L(phi) (homogeneous boundary values: bc), phi's z ghost planes already exchanged
int restrict_tma(mgic_op *o, const BCk &bc, mgic_field *resC, const mgic_field *phi, const mgic_field *rhs) {
  if (o->b) return launch_restrict<8, true>(o, bc, resC, phi, rhs);   // a fourth stream: 8-row tiles keep two CTAs per SM
  return launch_restrict<12, false>(o, bc, resC, phi, rhs);
}

}  // namespace mgk
