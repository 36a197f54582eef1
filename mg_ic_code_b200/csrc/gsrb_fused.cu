// gsrb_fused.cu -- fused red+black GSRB sweep: one kernel = one full levelGSRB (both colour passes of
// VariableCoeffPoissonOperator.cpp:290-331), each array streamed from HBM once.
//
// Scheme (out of place, phi_in -> phi_out, the field and the operator's scratch array ping-pong):
//   * a CTA owns an x-y tile of TX x TY cells and marches through a chunk of z planes;
//   * halo'd planes of phi_in ((TX+4) x (TY+4) doubles) are staged into a ring of shared-memory slots by TMA
//     (cp.async.bulk.tensor.3d, zero fill outside the array) and signalled through mbarriers, NSLOT-3 planes ahead;
//   * step k: RED update of plane k in place in shared memory on the tile grown by one cell (it reads old black
//     neighbours in planes k-1, k, k+1), block barrier, then BLACK update of plane k-1 on the tile itself (it
//     reads the new red values of planes k-2, k-1, k) -- exactly the reference's ordering: every red cell sees old
//     blacks, every black cell sees new reds -- and the finished plane k-1 goes to phi_out as 16-byte stores;
//   * rhs / aCoef / lambda (/ bCoef) are read once as 16-byte pairs when the red cell of the pair is updated; the
//     black cell's half waits in registers for the next step;
//   * physical boundary ghosts are folded in as a*centre + b (Source/SetBCs.cpp:49-131), never stored.
// The halo ring (2 cells in x/y, 2 planes per z chunk) is recomputed redundantly by neighbouring CTAs; those
// re-reads hit L2.  Arithmetic is the shared gsrb_point() => bit-identical to the per-colour kernel.
#include <cuda.h>

#include <map>

#include "mgic_internal.h"
#include "mgic_device.cuh"

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded spin: a lost TMA completion traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// Geometry of one CTA: a warp is one row of the halo'd region, a lane is one x-pair (16 bytes), so the colour of
// a lane's red cell is warp-uniform and compiled in (template parameter E): no per-thread selects.
//   region  = RW x (TY+4) cells, RW = 64 (32 pairs), staged by TMA           [tile grown by 2]
//   threads = rows 1 .. TY+2 of the region (TY+2 warps)                       [tile grown by 1 = red region]
//   tile    = rows 2 .. TY+1, lanes 1 .. 30  => TX = 60 cells x TY rows of output per plane
constexpr int TX = 60, RW = 64;

template <int TY, int NSLOT, bool HAS_B>
struct Fused {
  static constexpr int RR = TY + 4, NW = TY + 2, PLANE = RW * RR, NT = 32 * NW;
  static constexpr uint32_t PLANE_BYTES = PLANE * sizeof(double);
  static constexpr size_t SMEM = (size_t)NSLOT * PLANE_BYTES + 2 * NW * 32 * sizeof(double) + NSLOT * sizeof(uint64_t);

  // per-thread state
  double2 c0, c1, c2, c3;        // the lane's pair in planes kr-2, kr-1, kr, kr+1 (reds already updated where due)
  double2 ca, cl, cr, cb;        // coefficient pairs of plane kr (aCoef, lambda, rhs, bCoef)
  double sa, sl, sr, sb;         // black-cell halves of plane kr-1
  const double *dense;           // TMA slot of plane kr
  double *redw;                  // red buffer written this step (plane kr)
  const double *redr;            // red buffer of plane kr-1
  int lane, w, s;
  bool inDom, tile, anyxy, bx0, bxn, by0, byn;

  template <int E>
  __device__ __forceinline__ void step(const Geom &g, const BCk &bc, double alpha, double beta, double dxinv, int kr, bool doRed,
                                       bool doBlack, bool zloPhys, bool zhiPhys, double *outp) {
    constexpr unsigned FULL = 0xffffffffu;
    if (doRed) {
      const double c = E ? c2.y : c2.x;
      // x neighbours: one lives in the pair, the other in the neighbouring lane's pair (old black value)
      const double xo = E ? __shfl_down_sync(FULL, c2.x, 1) : __shfl_up_sync(FULL, c2.y, 1);
      double xm = E ? c2.x : xo, xp = E ? xo : c2.y;
      double ym = dense[s + E - RW], yp = dense[s + E + RW];
      double zm = E ? c1.y : c1.x, zp = E ? c3.y : c3.x;
      if (anyxy) {
        if (E == 0 && bx0) xm = bc.a[0] * c + bc.b[0];
        if (E == 1 && bxn) xp = bc.a[1] * c + bc.b[1];
        if (by0) ym = bc.a[2] * c + bc.b[2];
        if (byn) yp = bc.a[3] * c + bc.b[3];
      }
      if (kr == 0 && zloPhys) zm = bc.a[4] * c + bc.b[4];
      if (kr == g.nz - 1 && zhiPhys) zp = bc.a[5] * c + bc.b[5];
      const bool act = inDom && !(E == 0 && lane == 0) && !(E == 1 && lane == 31);
      double nv = c;
      if (act)
        nv = gsrb_point<HAS_B>(c, xm, xp, ym, yp, zm, zp, E ? ca.y : ca.x, HAS_B ? (E ? cb.y : cb.x) : 1.0, E ? cl.y : cl.x,
                               E ? cr.y : cr.x, alpha, beta, dxinv);
      if (E) c2.y = nv; else c2.x = nv;
      redw[w * 32 + lane] = nv;
    }
    if (doBlack) {
      const int kb = kr - 1;
      const double c = E ? c1.y : c1.x;
      const double xo = E ? __shfl_down_sync(FULL, c1.x, 1) : __shfl_up_sync(FULL, c1.y, 1);  // neighbour pair's new red
      double xm = E ? c1.x : xo, xp = E ? xo : c1.y;
      if (tile) {
        double ym = redr[(w - 1) * 32 + lane], yp = redr[(w + 1) * 32 + lane];
        double zm = E ? c0.y : c0.x, zp = E ? c2.y : c2.x;
        if (anyxy) {
          if (E == 0 && bx0) xm = bc.a[0] * c + bc.b[0];
          if (E == 1 && bxn) xp = bc.a[1] * c + bc.b[1];
          if (by0) ym = bc.a[2] * c + bc.b[2];
          if (byn) yp = bc.a[3] * c + bc.b[3];
        }
        if (kb == 0 && zloPhys) zm = bc.a[4] * c + bc.b[4];
        if (kb == g.nz - 1 && zhiPhys) zp = bc.a[5] * c + bc.b[5];
        const double nv = gsrb_point<HAS_B>(c, xm, xp, ym, yp, zm, zp, sa, HAS_B ? sb : 1.0, sl, sr, alpha, beta, dxinv);
        const double2 o = E ? make_double2(c1.x, nv) : make_double2(nv, c1.y);
        *reinterpret_cast<double2 *>(outp) = o;
      }
    }
    if (doRed) {  // keep the black-cell halves of plane kr for the next step
      sa = E ? ca.x : ca.y; sl = E ? cl.x : cl.y; sr = E ? cr.x : cr.y;
      if (HAS_B) sb = E ? cb.x : cb.y;
    }
  }
};

template <int TY, int NSLOT, bool HAS_B, int MINB>
__global__ void __launch_bounds__(32 * (TY + 2), MINB)
k_gsrb_fused(const __grid_constant__ CUtensorMap tmap, Geom g, BCk bc, double *__restrict__ out, const double *__restrict__ rhs,
             const double *__restrict__ a, const double *__restrict__ b, const double *__restrict__ lam, double alpha, double beta,
             double dxinv, int zchunk, int redLo, int redHi) {
  using F = Fused<TY, NSLOT, HAS_B>;
  constexpr int PLANE = F::PLANE, NW = F::NW;
  constexpr uint32_t PLANE_BYTES = F::PLANE_BYTES;
  static_assert(PLANE_BYTES % 128 == 0, "TMA destination slots must stay 128-byte aligned");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *planes = reinterpret_cast<double *>(smem_raw);
  double *redbuf = reinterpret_cast<double *>(smem_raw + (size_t)NSLOT * PLANE_BYTES);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)NSLOT * PLANE_BYTES + 2 * NW * 32 * sizeof(double));

  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  const int zs = blockIdx.z * zchunk, ze = min(zs + zchunk, g.nz);
  const int pfirst = zs - 2, plast = ze + 1;  // planes staged

  if (tid == 0) {
    for (int s = 0; s < NSLOT; s++) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    for (int p = pfirst; p < pfirst + NSLOT && p <= plast; p++) {
      const int sl = p - pfirst;
      mbar_arrive_expect_tx(&full[sl], PLANE_BYTES);
      tma_load_3d(planes + (size_t)sl * PLANE, &tmap, &full[sl], x0 - 2, y0 - 2, p + MGIC_GZ);
    }
  }

  F st;
  st.lane = tid & 31;
  st.w = tid >> 5;
  const int rr = st.w + 1;
  const int x = x0 - 2 + 2 * st.lane, y = y0 - 2 + rr;
  st.inDom = (x >= 0 && x + 1 < g.nx && y >= 0 && y < g.ny);
  st.tile = st.inDom && st.w >= 1 && st.w <= TY && st.lane >= 1 && st.lane <= 30;
  st.bx0 = (x == 0); st.bxn = (x == g.nx - 2); st.by0 = (y == 0); st.byn = (y == g.ny - 1);
  st.anyxy = st.bx0 || st.bxn || st.by0 || st.byn;
  st.s = rr * RW + 2 * st.lane;
  const int yodd = y & 1;
  const long long gofs = x + (long long)y * g.sy;
  const bool zloPhys = bc.type[4] != MGIC_FACE_INTERIOR, zhiPhys = bc.type[5] != MGIC_FACE_INTERIOR;
  const double2 zero2 = make_double2(0.0, 0.0);
  st.c0 = zero2; st.sa = st.sl = st.sr = st.sb = 0.0;
  st.ca = st.cl = st.cr = st.cb = zero2;

  // prologue: the lane's pairs of planes zs-2 and zs-1; coefficient pairs of the first red plane
  mbar_wait(&full[0], 0);
  st.c1 = *reinterpret_cast<const double2 *>(planes + st.s);
  if (plast >= pfirst + 1) mbar_wait(&full[1 % NSLOT], 0);
  st.c2 = *reinterpret_cast<const double2 *>(planes + (size_t)(1 % NSLOT) * PLANE + st.s);
  {
    const int k = zs - 1;
    if (st.inDom && k >= redLo && k <= redHi) {
      const long long gi = gofs + (long long)k * g.sz;
      st.ca = *reinterpret_cast<const double2 *>(a + gi);
      st.cl = *reinterpret_cast<const double2 *>(lam + gi);
      st.cr = *reinterpret_cast<const double2 *>(rhs + gi);
      if (HAS_B) st.cb = *reinterpret_cast<const double2 *>(b + gi);
    }
  }
  __syncthreads();
  if (tid == 0 && pfirst + NSLOT <= plast) {  // slot of plane zs-2 is free again
    mbar_arrive_expect_tx(&full[0], PLANE_BYTES);
    tma_load_3d(planes, &tmap, &full[0], x0 - 2, y0 - 2, pfirst + NSLOT + MGIC_GZ);
  }

  for (int kr = zs - 1; kr <= ze; kr++) {
    const bool doRed = (kr >= redLo && kr <= redHi);
    const bool doBlack = (kr - 1 >= zs && kr - 1 < ze);
    // prefetch the coefficient pairs of plane kr+1 (consumed by the next step's red update)
    double2 na = zero2, nl = zero2, nr = zero2, nb = zero2;
    const bool nextRed = (kr + 1 >= redLo && kr + 1 <= redHi && kr + 1 <= ze);
    if (st.inDom && nextRed) {
      const long long gi = gofs + (long long)(kr + 1) * g.sz;
      na = *reinterpret_cast<const double2 *>(a + gi);
      nl = *reinterpret_cast<const double2 *>(lam + gi);
      nr = *reinterpret_cast<const double2 *>(rhs + gi);
      if (HAS_B) nb = *reinterpret_cast<const double2 *>(b + gi);
    }
    // the lane's pair of plane kr+1
    if (kr + 1 <= plast) {
      const int q = kr + 1 - pfirst;
      mbar_wait(&full[q % NSLOT], (q / NSLOT) & 1);
      st.c3 = *reinterpret_cast<const double2 *>(planes + (size_t)(q % NSLOT) * PLANE + st.s);
    }
    st.dense = planes + (size_t)((kr - pfirst) % NSLOT) * PLANE;
    st.redw = redbuf + (size_t)(kr & 1) * NW * 32;
    st.redr = redbuf + (size_t)((kr - 1) & 1) * NW * 32;
    double *outp = out + gofs + (long long)(kr - 1) * g.sz;
    const int e = (yodd + kr + g.k0) & 1;  // warp-uniform: element of the pair that is red in plane kr
    if (e) st.template step<1>(g, bc, alpha, beta, dxinv, kr, doRed, doBlack, zloPhys, zhiPhys, outp);
    else st.template step<0>(g, bc, alpha, beta, dxinv, kr, doRed, doBlack, zloPhys, zhiPhys, outp);
    st.c0 = st.c1; st.c1 = st.c2; st.c2 = st.c3;
    st.ca = na; st.cl = nl; st.cr = nr;
    if (HAS_B) st.cb = nb;
    __syncthreads();
    // plane kr's slot is dead (pair copied a step ago, y neighbours read by this step's red update)
    if (tid == 0) {
      const int pn = kr + NSLOT;
      if (pn <= plast) {
        const int sl = (kr - pfirst) % NSLOT;
        mbar_arrive_expect_tx(&full[sl], PLANE_BYTES);
        tma_load_3d(planes + (size_t)sl * PLANE, &tmap, &full[sl], x0 - 2, y0 - 2, pn + MGIC_GZ);
      }
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap(CUtensorMap *m, const double *base, int nx, int ny, int nplanes, int boxx, int boxy) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { mgic_set_error("cuTensorMapEncodeTiled is not available from the driver"); return MGIC_ERR_CUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nplanes};
  cuuint64_t strides[2] = {(cuuint64_t)nx * 8, (cuuint64_t)nx * ny * 8};
  cuuint32_t box[3] = {(cuuint32_t)boxx, (cuuint32_t)boxy, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { mgic_set_error("cuTensorMapEncodeTiled failed (%d) for %dx%dx%d box %dx%d", (int)r, nx, ny, nplanes, boxx, boxy); return MGIC_ERR_CUDA; }
  return MGIC_OK;
}

struct Plan { int nch, zchunk; };

// number of z chunks: fill whole waves of resident CTAs while keeping the two redundant planes per chunk cheap
Plan plan_chunks(int tiles, int nz, int resident) {
  Plan best = {1, nz};
  double bestScore = -1.0;
  for (int nch = 1; nch <= nz; nch++) {
    const int zc = (nz + nch - 1) / nch;
    if (zc < 8 && nch > 1) break;
    const int nchEff = (nz + zc - 1) / zc;
    const long long total = (long long)tiles * nchEff;
    const long long waves = (total + resident - 1) / resident;
    const double fill = (double)total / (double)(waves * resident);
    const double score = fill * zc / (zc + 2.0);
    if (score > bestScore + 1e-9) { bestScore = score; best = {nchEff, zc}; }
  }
  return best;
}

template <int TY, int NSLOT, bool HAS_B, int MINB>
int launch_cfg(mgic_op *o, const double *in, double *outp, const mgic_field *r) {
  using F = Fused<TY, NSLOT, HAS_B>;
  mgic_ctx *c = o->ctx;
  auto kern = k_gsrb_fused<TY, NSLOT, HAS_B, MINB>;
  static bool attrSet = false;
  static int resident = 1;
  if (!attrSet) {
    MGIC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F::SMEM));
    int per = 1;
    MGIC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, F::NT, F::SMEM));
    resident = (per < 1 ? 1 : per) * c->numSMs;
    attrSet = true;
  }
  const Geom g = o->geom();
  const BCk bc = o->bck(true);
  CUtensorMap tm;
  MGIC_TRY(make_tmap(&tm, in - (long long)MGIC_GZ * g.sz, g.nx, g.ny, g.nz + 2 * MGIC_GZ, RW, F::RR));
  const int tilesX = (g.nx + TX - 1) / TX, tilesY = (g.ny + TY - 1) / TY;
  const Plan pl = plan_chunks(tilesX * tilesY, g.nz, resident);
  const int redLo = (bc.type[4] == MGIC_FACE_INTERIOR) ? -1 : 0;
  const int redHi = (bc.type[5] == MGIC_FACE_INTERIOR) ? g.nz : g.nz - 1;
  dim3 grd(tilesX, tilesY, pl.nch);
  kern<<<grd, F::NT, F::SMEM, c->stream>>>(tm, g, bc, outp, r->p, o->a->p, o->b ? o->b->p : nullptr, o->lambda->p, o->alpha, o->beta,
                                           1.0 / (o->dx * o->dx), pl.zchunk, redLo, redHi);
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { mgic_set_error("kernel gsrb_fused: %s", cudaGetErrorString(e)); return MGIC_ERR_CUDA; }
  return MGIC_OK;
}

template <bool HAS_B>
int launch(mgic_op *o, const double *in, double *outp, const mgic_field *r) {
  switch (o->ctx->fusedCfg) {
    case 0: return launch_cfg<8, 4, HAS_B, 2>(o, in, outp, r);
    case 2: return launch_cfg<24, 4, HAS_B, 1>(o, in, outp, r);
    case 3: return launch_cfg<16, 6, HAS_B, 1>(o, in, outp, r);
    case 4: return launch_cfg<8, 6, HAS_B, 2>(o, in, outp, r);
    case 1: return launch_cfg<16, 4, HAS_B, 1>(o, in, outp, r);
    case 6: return launch_cfg<12, 6, HAS_B, 2>(o, in, outp, r);
    case 7: return launch_cfg<10, 4, HAS_B, 2>(o, in, outp, r);
    case 8: return launch_cfg<6, 4, HAS_B, 3>(o, in, outp, r);
    default: return launch_cfg<12, 4, HAS_B, 2>(o, in, outp, r);
  }
}

}  // namespace

namespace mgk {

// Periodic faces (TMA cannot wrap), odd nx and small (launch-latency bound) levels use the per-colour kernel, which
// computes the same bits.
bool gsrb_fused_applicable(const mgic_op *o) {
  for (int d = 0; d < 3; d++)
    if (o->bc_lo[d] == MGIC_BC_PERIODIC) return false;
  const long long cells = (long long)o->n[0] * o->n[1] * o->nzl;
  if ((o->n[0] & 1) || o->n[0] < 8 || cells < o->ctx->fusedMinCells) return false;
  if (o->ctx->nranks > 1 && o->nzl < MGIC_GZ) return false;
  return true;
}

// relax(e, r, iterations) with fused sweeps.  Multi-rank: one 2-plane halo exchange of e per sweep (the neighbour's
// first plane is updated redundantly) instead of the reference's two 1-plane exchanges, plus one of rhs per call.
int gsrb_fused(mgic_op *o, mgic_field *e, const mgic_field *r, int iterations) {
  if (!o->scratch) MGIC_TRY(mgic_field_create(o, &o->scratch));
  if (o->ctx->nranks > 1) MGIC_TRY(mgic_halo(o, const_cast<mgic_field *>(r), 1));
  for (int it = 0; it < iterations; it++) {
    if (o->ctx->nranks > 1) MGIC_TRY(mgic_halo(o, e, 2));
    {
      ProfScope ps(o->ctx, o->profTag);
      if (o->b) MGIC_TRY(launch<true>(o, e->p, o->scratch->p, r));
      else MGIC_TRY(launch<false>(o, e->p, o->scratch->p, r));
    }
    std::swap(e->base, o->scratch->base);  // ping-pong: the field handle now owns the freshly written array
    std::swap(e->p, o->scratch->p);
  }
  return MGIC_OK;
}

}  // namespace mgk
