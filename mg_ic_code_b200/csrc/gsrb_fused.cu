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
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

enum : unsigned {
  F_DOM = 1u,      // pair inside the domain
  F_ROWRED = 2u,   // row inside the red region (tile grown by one) and the domain
  F_TILE = 4u,     // pair inside the tile (black update + output)
  F_PFIRST = 8u,   // leftmost pair of the region: only its element 1 is in the red region
  F_PLAST = 16u,   // rightmost pair: only element 0
  F_X0 = 32u, F_XN = 64u, F_Y0 = 128u, F_YN = 256u,  // pair touches a physical x / y face
  F_YODD = 512u
};

template <int TX, int TY, int NT, int NSLOT, bool HAS_B>
__global__ void __launch_bounds__(NT) k_gsrb_fused(const __grid_constant__ CUtensorMap tmap, Geom g, BCk bc,
                                                   double *__restrict__ out, const double *__restrict__ rhs,
                                                   const double *__restrict__ a, const double *__restrict__ b,
                                                   const double *__restrict__ lam, double alpha, double beta, double dxinv,
                                                   int zchunk, int redLo, int redHi) {
  constexpr int PR = TX / 2 + 2, RW = 2 * PR, RR = TY + 4, PLANE = RW * RR, NI = PR * RR, CPT = (NI + NT - 1) / NT;
  constexpr uint32_t PLANE_BYTES = PLANE * sizeof(double);
  static_assert(PLANE_BYTES % 128 == 0, "TMA destination slots must stay 128-byte aligned");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *planes = reinterpret_cast<double *>(smem_raw);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)NSLOT * PLANE_BYTES);

  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  const int zs = blockIdx.z * zchunk, ze = min(zs + zchunk, g.nz);
  const int pfirst = zs - 2;                    // first plane staged
  const int plast = ze + 1;                     // last plane staged

  if (tid == 0) {
    for (int s = 0; s < NSLOT; s++) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    for (int p = pfirst; p < pfirst + NSLOT && p <= plast; p++) {
      const int s = p - pfirst;
      mbar_arrive_expect_tx(&full[s], PLANE_BYTES);
      tma_load_3d(planes + (size_t)s * PLANE, &tmap, &full[s], x0 - 2, y0 - 2, p + MGIC_GZ);
    }
  }

  // the thread's pair columns (fixed for the whole march)
  int sidx[CPT];
  long long gofs[CPT];
  unsigned flg[CPT];
#pragma unroll
  for (int m = 0; m < CPT; m++) {
    const int q = tid + m * NT;
    const int rr = q / PR, pp = q - rr * PR;
    const int x = x0 - 2 + 2 * pp, y = y0 - 2 + rr;
    unsigned f = 0;
    if (q < NI && x >= 0 && x + 1 < g.nx && y >= 0 && y < g.ny) {
      f |= F_DOM;
      if (rr >= 1 && rr <= RR - 2) f |= F_ROWRED;
      if (rr >= 2 && rr <= RR - 3 && pp >= 1 && pp <= PR - 2) f |= F_TILE;
      if (pp == 0) f |= F_PFIRST;
      if (pp == PR - 1) f |= F_PLAST;
      if (x == 0) f |= F_X0;
      if (x == g.nx - 2) f |= F_XN;
      if (y == 0) f |= F_Y0;
      if (y == g.ny - 1) f |= F_YN;
      if (y & 1) f |= F_YODD;
    }
    flg[m] = f;
    sidx[m] = rr * RW + 2 * pp;
    gofs[m] = x + (long long)y * g.sy;
  }
  // black-cell halves of the coefficient pairs of the plane whose red update ran one step earlier
  double st_a[CPT], st_l[CPT], st_r[CPT], st_b[HAS_B ? CPT : 1];

  const bool zloPhys = bc.type[4] != MGIC_FACE_INTERIOR, zhiPhys = bc.type[5] != MGIC_FACE_INTERIOR;

  for (int kr = zs - 1; kr <= ze; kr++) {
    const int kb = kr - 1;
    const bool doRed = (kr >= redLo && kr <= redHi);
    const bool doBlack = (kb >= zs && kb < ze);
    // wait for the planes this step reads for the first time
    if (kr == zs - 1) {
      for (int p = pfirst; p <= kr; p++) mbar_wait(&full[(p - pfirst) % NSLOT], ((p - pfirst) / NSLOT) & 1);
    }
    if (kr + 1 <= plast) mbar_wait(&full[(kr + 1 - pfirst) % NSLOT], ((kr + 1 - pfirst) / NSLOT) & 1);

    double *P = planes + (size_t)((kr - pfirst) % NSLOT) * PLANE;                 // plane kr
    double *Pm = planes + (size_t)((kr - 1 - pfirst + NSLOT) % NSLOT) * PLANE;    // plane kr-1
    double *Pp = planes + (size_t)((kr + 1 - pfirst) % NSLOT) * PLANE;            // plane kr+1
    const int par = (kr + g.k0) & 1;

    // ---------------- RED: plane kr, in place in shared memory ------------------------------------------------
    double nx_a[CPT], nx_l[CPT], nx_r[CPT], nx_b[HAS_B ? CPT : 1];
    if (doRed) {
#pragma unroll
      for (int m = 0; m < CPT; m++) {
        const unsigned f = flg[m];
        const int e = (((f & F_YODD) ? 1 : 0) + par) & 1;   // red element of the pair: (x + e + y + k) even, x even
        const bool act = (f & F_DOM) && (f & F_ROWRED) && !((f & F_PFIRST) && e == 0) && !((f & F_PLAST) && e == 1);
        if (f & F_DOM) {
          // coefficient pair of plane kr (kept for the black update of this plane in the next step)
          const long long gi = gofs[m] + (long long)kr * g.sz;
          if (act || (f & F_TILE)) {
            const double2 a2 = *reinterpret_cast<const double2 *>(a + gi);
            const double2 l2 = *reinterpret_cast<const double2 *>(lam + gi);
            const double2 r2 = *reinterpret_cast<const double2 *>(rhs + gi);
            double2 b2 = make_double2(1.0, 1.0);
            if (HAS_B) b2 = *reinterpret_cast<const double2 *>(b + gi);
            nx_a[m] = e ? a2.x : a2.y; nx_l[m] = e ? l2.x : l2.y; nx_r[m] = e ? r2.x : r2.y;
            if (HAS_B) nx_b[m] = e ? b2.x : b2.y;
            if (act) {
              const int s = sidx[m];
              const double2 cp = *reinterpret_cast<const double2 *>(P + s);
              const double c = e ? cp.y : cp.x;
              double xm = e ? cp.x : P[s - 1];
              double xp = e ? P[s + 2] : cp.y;
              double ym = P[s + e - RW], yp = P[s + e + RW];
              double zm = Pm[s + e], zp = Pp[s + e];
              if ((f & F_X0) && e == 0) xm = bc.a[0] * c + bc.b[0];
              if ((f & F_XN) && e == 1) xp = bc.a[1] * c + bc.b[1];
              if (f & F_Y0) ym = bc.a[2] * c + bc.b[2];
              if (f & F_YN) yp = bc.a[3] * c + bc.b[3];
              if (kr == 0 && zloPhys) zm = bc.a[4] * c + bc.b[4];
              if (kr == g.nz - 1 && zhiPhys) zp = bc.a[5] * c + bc.b[5];
              const double nv = gsrb_point<HAS_B>(c, xm, xp, ym, yp, zm, zp, e ? a2.y : a2.x, e ? b2.y : b2.x, e ? l2.y : l2.x,
                                                  e ? r2.y : r2.x, alpha, beta, dxinv);
              P[s + e] = nv;
            }
          }
        }
      }
    }
    __syncthreads();
    // the slot of plane kr-3 is dead now (its last reader was the black update of plane kr-2 in the previous step)
    if (tid == 0) {
      const int pn = kr - 3 + NSLOT;
      if (kr - 3 >= pfirst && pn <= plast) {
        const int s = (pn - pfirst) % NSLOT;
        fence_proxy_async();
        mbar_arrive_expect_tx(&full[s], PLANE_BYTES);
        tma_load_3d(planes + (size_t)s * PLANE, &tmap, &full[s], x0 - 2, y0 - 2, pn + MGIC_GZ);
      }
    }
    // ---------------- BLACK: plane kb = kr-1, result streamed to phi_out --------------------------------------
    if (doBlack) {
      double *Q = Pm;                                                               // plane kb
      double *Qm = planes + (size_t)((kb - 1 - pfirst + NSLOT) % NSLOT) * PLANE;    // plane kb-1
      double *Qp = P;                                                               // plane kb+1 = kr
#pragma unroll
      for (int m = 0; m < CPT; m++) {
        const unsigned f = flg[m];
        if (f & F_TILE) {
          const int e = (((f & F_YODD) ? 1 : 0) + par) & 1;  // black element of plane kb == red element of plane kr
          const int s = sidx[m];
          const double2 cp = *reinterpret_cast<const double2 *>(Q + s);
          const double c = e ? cp.y : cp.x;
          double xm = e ? cp.x : Q[s - 1];
          double xp = e ? Q[s + 2] : cp.y;
          double ym = Q[s + e - RW], yp = Q[s + e + RW];
          double zm = Qm[s + e], zp = Qp[s + e];
          if ((f & F_X0) && e == 0) xm = bc.a[0] * c + bc.b[0];
          if ((f & F_XN) && e == 1) xp = bc.a[1] * c + bc.b[1];
          if (f & F_Y0) ym = bc.a[2] * c + bc.b[2];
          if (f & F_YN) yp = bc.a[3] * c + bc.b[3];
          if (kb == 0 && zloPhys) zm = bc.a[4] * c + bc.b[4];
          if (kb == g.nz - 1 && zhiPhys) zp = bc.a[5] * c + bc.b[5];
          const double nv = gsrb_point<HAS_B>(c, xm, xp, ym, yp, zm, zp, st_a[m], HAS_B ? st_b[m] : 1.0, st_l[m], st_r[m], alpha,
                                              beta, dxinv);
          const double2 o = e ? make_double2(cp.x, nv) : make_double2(nv, cp.y);
          *reinterpret_cast<double2 *>(out + gofs[m] + (long long)kb * g.sz) = o;
        }
      }
    }
    if (doRed) {
#pragma unroll
      for (int m = 0; m < CPT; m++) {
        st_a[m] = nx_a[m]; st_l[m] = nx_l[m]; st_r[m] = nx_r[m];
        if (HAS_B) st_b[m] = nx_b[m];
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

template <int TX, int TY, int NT, int NSLOT, bool HAS_B>
int launch_cfg(mgic_op *o, const double *in, double *outp, const mgic_field *r) {
  constexpr int PR = TX / 2 + 2, RW = 2 * PR, RR = TY + 4;
  constexpr size_t SMEM = (size_t)NSLOT * RW * RR * 8 + NSLOT * 8;
  mgic_ctx *c = o->ctx;
  auto kern = k_gsrb_fused<TX, TY, NT, NSLOT, HAS_B>;
  static bool attrSet = false;
  static int resident = 1;
  if (!attrSet) {
    MGIC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    int per = 1;
    MGIC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, NT, SMEM));
    resident = (per < 1 ? 1 : per) * c->numSMs;
    attrSet = true;
  }
  const Geom g = o->geom();
  const BCk bc = o->bck(true);
  CUtensorMap tm;
  MGIC_TRY(make_tmap(&tm, in - (long long)MGIC_GZ * g.sz, g.nx, g.ny, g.nz + 2 * MGIC_GZ, RW, RR));
  const int tilesX = (g.nx + TX - 1) / TX, tilesY = (g.ny + TY - 1) / TY;
  const Plan pl = plan_chunks(tilesX * tilesY, g.nz, resident);
  const int redLo = (bc.type[4] == MGIC_FACE_INTERIOR) ? -1 : 0;
  const int redHi = (bc.type[5] == MGIC_FACE_INTERIOR) ? g.nz : g.nz - 1;
  dim3 grd(tilesX, tilesY, pl.nch);
  kern<<<grd, NT, SMEM, c->stream>>>(tm, g, bc, outp, r->p, o->a->p, o->b ? o->b->p : nullptr, o->lambda->p, o->alpha, o->beta,
                                     1.0 / (o->dx * o->dx), pl.zchunk, redLo, redHi);
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { mgic_set_error("kernel gsrb_fused: %s", cudaGetErrorString(e)); return MGIC_ERR_CUDA; }
  return MGIC_OK;
}

template <bool HAS_B>
int launch(mgic_op *o, const double *in, double *outp, const mgic_field *r) {
  const int cfg = o->ctx->fusedCfg;
  if (o->n[0] >= 128 && cfg == 0) return launch_cfg<128, 16, 512, 6, HAS_B>(o, in, outp, r);
  if (o->n[0] >= 64 && cfg == 2) return launch_cfg<64, 32, 512, 6, HAS_B>(o, in, outp, r);
  if (o->n[0] >= 128 && cfg == 3) return launch_cfg<128, 8, 512, 6, HAS_B>(o, in, outp, r);
  if (o->n[0] >= 64) return launch_cfg<64, 16, 256, 6, HAS_B>(o, in, outp, r);
  return launch_cfg<32, 8, 128, 6, HAS_B>(o, in, outp, r);
}

}  // namespace

namespace mgk {

// relax(e, r, iterations) with fused sweeps.  Periodic faces (TMA cannot wrap) and odd nx fall back to the
// per-colour kernel, which computes the same bits.
int gsrb_fused(mgic_op *o, mgic_field *e, const mgic_field *r, int iterations) {
  const Geom g = o->geom();
  bool periodic = false;
  for (int d = 0; d < 3; d++) periodic = periodic || o->bc_lo[d] == MGIC_BC_PERIODIC;
  const long long cells = (long long)g.nx * g.ny * g.nz;
  if (periodic || (g.nx & 1) || g.nx < 8 || cells < o->ctx->fusedMinCells) {
    const BCk bc = o->bck(true);
    for (int it = 0; it < iterations; it++)
      for (int pass = 0; pass <= 1; pass++) {
        ProfScope ps(o->ctx, o->profTag);
        MGIC_TRY(gsrb_color(o->ctx, g, bc, e->p, r->p, o->a->p, o->b ? o->b->p : nullptr, o->lambda->p, o->alpha, o->beta, o->dx, pass));
      }
    return MGIC_OK;
  }
  if (!o->scratch) MGIC_TRY(mgic_field_create(o, &o->scratch));
  for (int it = 0; it < iterations; it++) {
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
