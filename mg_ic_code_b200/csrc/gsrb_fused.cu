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

#include <cstring>

#include "mgic_internal.h"
#include "mgic_device.cuh"
#include "tma.cuh"

namespace {

// Geometry of one CTA: a warp is one row of the halo'd region, a lane is one x-pair (16 bytes), so the colour of
// a lane's red cell is warp-uniform and compiled in (template parameter E): no per-thread selects.
//   region  = RW x (TY+4) cells, RW = 64 (32 pairs), staged by TMA           [tile grown by 2]
//   threads = rows 1 .. TY+2 of the region (TY+2 warps)                       [tile grown by 1 = red region]
//   tile    = rows 2 .. TY+1, lanes 1 .. 30  => TX = 60 cells x TY rows of output per plane
// The z march is unrolled by four (= NSLOT): the four register pairs, the two coefficient sets and the TMA slots
// rotate by renaming, not by moves, and slot indices are compile-time.
constexpr int TX = 60, RW = 64, NSLOT = 4;

static_assert(mgk::FUSED_PLAIN == 0 && mgk::FUSED_FROM_ZERO == 1 && mgk::FUSED_PROLONG == 2, "mode numbering");
enum { MODE_PLAIN = 0,     // phi_out = sweep(phi_in)
       MODE_ZERO = 1,      // phi_in is identically zero (the correction of a fresh MG level): nothing is read for it
       MODE_PROLONG = 2 }; // phi_in + piecewise-constant prolongation of the coarse correction, added on the fly
                           // ([Chombo] AMRPoissonOp::prolongIncrement fused into the first post-smoothing sweep)

struct FusedArgs {
  Geom g;
  BCk bc;
  double *out;
  double alpha, beta, dxinv;
  int zchunk, redLo, redHi;
  int zbeg, zend;         // output planes of this launch: [zbeg, zend) (interior / boundary launches of the overlapped sweep)
  // multi-rank, halo folded into the sweep (mgic_internal.h MGIC_SW_*): the finished planes 0, 1 / nz-2, nz-1 also go into
  // the lo / hi neighbour's ghost planes (NVLink peer stores); ctl == null: single rank or exchange by k_halo_push
  SweepPeers sw;
  int ctasPerChunk, ctasTotal;
};

typedef unsigned long long u64;
__device__ __forceinline__ void st_release_sys(u64 *p, u64 v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ u64 ld_acquire_sys(const u64 *p) {
  u64 v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// spin until *p >= v; a neighbour that died must not leave this GPU spinning forever: trap after 60 s
__device__ __forceinline__ void sweep_wait_ge(const u64 *p, u64 v) {
  if (ld_acquire_sys(p) >= v) return;
  u64 t0;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  while (ld_acquire_sys(p) < v) {
    __nanosleep(64);
    u64 t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    if (t - t0 > 60000000000ull) { printf("mgic: fused sweep timed out waiting for a neighbour rank's boundary planes\n"); __trap(); }
  }
}

// All input streams arrive by TMA: per z plane one slot = the halo'd phi plane plus the (TY+2)-row planes of aCoef,
// lambda, rhs (and bCoef) -- and, in MODE_PROLONG, the 34 x (TY+4)/2 tile of the coarse correction under the region --
// completed through ONE mbarrier.  Threads never form a global load address.  (The coarse values used to be plain
// global loads: three dependent-latency loads per lane and plane made that sweep 36 % slower than the plain one,
// 1138 vs 835 us on 512^3, profiles/r1b_launches_bench_512.csv.)
template <int TY, bool HAS_B, int MODE, bool FOLD = false, bool PATCH = false>
struct Fused {
  static constexpr int RR = TY + 4, NW = TY + 2, PLANE = RW * RR, CPLANE = RW * NW, NT = 32 * NW, NCOEF = HAS_B ? 4 : 3;
  // coarse tile: TMA wants the innermost start coordinate on a 16-byte boundary, i.e. an even coarse x; (x0-2)/2 is odd, so
  // the tile starts one coarse cell further left and is RW/2 + 2 wide (the lane's coarse cell is column lane + 1)
  static constexpr int CROWS = RR / 2, CW = RW / 2 + 2, CBOX = (MODE == MODE_PROLONG) ? CW * CROWS : 0, CTILE = (CBOX + 15) / 16 * 16;
  static constexpr int COFF = PLANE + NCOEF * CPLANE;
  static constexpr int SLOT = PLANE + NCOEF * CPLANE + CTILE;  // doubles
  static constexpr uint32_t PLANE_BYTES = PLANE * sizeof(double), CPLANE_BYTES = CPLANE * sizeof(double), CTILE_BYTES = CTILE * sizeof(double);
  static constexpr uint32_t TX_BYTES = (MODE == MODE_ZERO ? 0 : PLANE_BYTES) + NCOEF * CPLANE_BYTES + CBOX * sizeof(double);
  static constexpr size_t SMEM = (size_t)NSLOT * SLOT * sizeof(double) + 2 * NW * 32 * sizeof(double) + NSLOT * sizeof(uint64_t);

  const FusedArgs &A;
  const CUtensorMap *tm_phi, *tm_a, *tm_l, *tm_r, *tm_b, *tm_c;
  double *slots, *redbuf;
  uint64_t *full;
  // thread constants
  int tid, lane, w, s, cidx, x0, y0, zs, ze, pfirst, plast;
  int ci, cim, cip;   // MODE_PROLONG: the lane's coarse cell in the slot's coarse tile, and those of the rows y-1 / y+1
  double *outp;   // this lane's pair in the output plane being written; advanced by one plane per step
  bool inDom, tile, anyxy, bx0, bxn, by0, byn, zloPhys, zhiPhys;
  // black-cell halves of the coefficient pairs of the plane whose red update ran one step earlier
  double sa, sl, sr, sb;

  __device__ __forceinline__ Fused(const FusedArgs &a_) : A(a_) {}

  __device__ __forceinline__ void issue(int plane, int slot) {  // thread 0: stage one plane of every stream
    double *d = slots + (size_t)slot * SLOT;
    mbar_arrive_expect_tx(&full[slot], TX_BYTES);
    const int z = plane + MGIC_GZ;
    if (MODE != MODE_ZERO) tma_load_3d(d, tm_phi, &full[slot], x0 - 2, y0 - 2, z);
    tma_load_3d(d + PLANE, tm_a, &full[slot], x0 - 2, y0 - 1, z);
    tma_load_3d(d + PLANE + CPLANE, tm_l, &full[slot], x0 - 2, y0 - 1, z);
    tma_load_3d(d + PLANE + 2 * CPLANE, tm_r, &full[slot], x0 - 2, y0 - 1, z);
    if (HAS_B) tma_load_3d(d + PLANE + 3 * CPLANE, tm_b, &full[slot], x0 - 2, y0 - 1, z);
    // coarse plane under fine plane `plane` (x0 - 2 and y0 - 2 are even; >> floors the ghost planes -2, -1 to -1)
    if (MODE == MODE_PROLONG) tma_load_3d(d + COFF, tm_c, &full[slot], ((x0 - 2) >> 1) - 1, (y0 - 2) >> 1, (plane >> 1) + MGIC_GZ);
  }

  __device__ __forceinline__ double2 load_pair(int plane, const double *slot) const {
    if (MODE == MODE_ZERO) return make_double2(0.0, 0.0);
    double2 v = *reinterpret_cast<const double2 *>(slot + s);
    if (MODE == MODE_PROLONG && inDom) {
      const double cv = slot[COFF + ci];
      v.x = v.x + cv; v.y = v.y + cv;   // phi(i,j,k) + coarse(i/2, j/2, k/2)
    }
    return v;
  }

  // the ghost value beyond face f of a cell with value c whose neighbour on the opposite side is `far`: ParseBC's a*c + b, or --
  // PATCH: a rectangular AMR patch, [Chombo] homogeneousCFInterp on a coarse-fine face -- the parabola through far, c and a
  // zero coarse value (cf_homog, as the per-colour kernel evaluates it: `far` is of the other colour, i.e. its current value)
  __device__ __forceinline__ double face_ghost(int f, double c, double far) const {
    if (PATCH && A.bc.type[f] == MGIC_FACE_CF) return cf_homog(A.bc.cf, far, c);
    return A.bc.a[f] * c + A.bc.b[f];
  }

  // one plane step: red update of plane kr (pc), black update + output of plane kr-1 (pm1)
  template <int E, int U>
  __device__ __forceinline__ void step(int kr, int it, const double2 &pm2, const double2 &pm1, double2 &pc, double2 &pn) {
    constexpr unsigned FULL = 0xffffffffu;
    const bool doRed = (kr >= A.redLo && kr <= A.redHi);
    const bool doBlack = (kr - 1 >= zs && kr - 1 < ze);
    constexpr int SN = (2 + U) & 3, SC = (1 + U) & 3;
    // the lane's pair of plane kr+1
    if (kr + 1 <= plast) mbar_wait(&full[SN], (it + ((2 + U) >> 2)) & 1);
    pn = load_pair(kr + 1, slots + SN * SLOT);
    const double *dense = slots + SC * SLOT;
    double *redw = redbuf + (kr & 1) * NW * 32;
    const double *redr = redbuf + ((kr - 1) & 1) * NW * 32;
    double2 ca, cl, cr, cb;
    if (doRed) {
      ca = *reinterpret_cast<const double2 *>(dense + PLANE + cidx);
      cl = *reinterpret_cast<const double2 *>(dense + PLANE + CPLANE + cidx);
      cr = *reinterpret_cast<const double2 *>(dense + PLANE + 2 * CPLANE + cidx);
      if (HAS_B) cb = *reinterpret_cast<const double2 *>(dense + PLANE + 3 * CPLANE + cidx);
      const double c = E ? pc.y : pc.x;
      // x neighbours: one lives in the pair, the other in the neighbouring lane's pair (old black value)
      const double xo = E ? __shfl_down_sync(FULL, pc.x, 1) : __shfl_up_sync(FULL, pc.y, 1);
      double xm = E ? pc.x : xo, xp = E ? xo : pc.y;
      double ym = 0.0, yp = 0.0;
      if (MODE != MODE_ZERO) { ym = dense[s + E - RW]; yp = dense[s + E + RW]; }
      if (MODE == MODE_PROLONG && inDom) {  // (rows outside the domain read the tile's zero fill; the BC below replaces them)
        ym = ym + dense[COFF + cim];
        yp = yp + dense[COFF + cip];
      }
      double zm = E ? pm1.y : pm1.x, zp = E ? pn.y : pn.x;
      if (anyxy) {
        if (E == 0 && bx0) xm = face_ghost(0, c, xp);
        if (E == 1 && bxn) xp = face_ghost(1, c, xm);
        if (by0) ym = face_ghost(2, c, yp);
        if (byn) yp = face_ghost(3, c, ym);
      }
      if (kr == 0 && zloPhys) zm = face_ghost(4, c, zp);
      if (kr == A.g.nz - 1 && zhiPhys) zp = face_ghost(5, c, zm);
      // lanes outside the red region / the domain compute a value nobody reads (their coefficients are zero-filled)
      const double nv = gsrb_point<HAS_B>(c, xm, xp, ym, yp, zm, zp, E ? ca.y : ca.x, HAS_B ? (E ? cb.y : cb.x) : 1.0,
                                          E ? cl.y : cl.x, E ? cr.y : cr.x, A.alpha, A.beta, A.dxinv);
      if (E) pc.y = nv; else pc.x = nv;
      redw[w * 32 + lane] = nv;
    }
    if (doBlack) {
      const int kb = kr - 1;
      const double c = E ? pm1.y : pm1.x;
      const double xo = E ? __shfl_down_sync(FULL, pm1.x, 1) : __shfl_up_sync(FULL, pm1.y, 1);  // neighbour pair's new red
      double xm = E ? pm1.x : xo, xp = E ? xo : pm1.y;
      if (tile) {
        double ym = redr[(w - 1) * 32 + lane], yp = redr[(w + 1) * 32 + lane];
        double zm = E ? pm2.y : pm2.x, zp = E ? pc.y : pc.x;
        if (anyxy) {
          if (E == 0 && bx0) xm = face_ghost(0, c, xp);
          if (E == 1 && bxn) xp = face_ghost(1, c, xm);
          if (by0) ym = face_ghost(2, c, yp);
          if (byn) yp = face_ghost(3, c, ym);
        }
        if (kb == 0 && zloPhys) zm = face_ghost(4, c, zp);
        if (kb == A.g.nz - 1 && zhiPhys) zp = face_ghost(5, c, zm);
        const double nv = gsrb_point<HAS_B>(c, xm, xp, ym, yp, zm, zp, sa, HAS_B ? sb : 1.0, sl, sr, A.alpha, A.beta, A.dxinv);
        const double2 o = E ? make_double2(pm1.x, nv) : make_double2(nv, pm1.y);
        *reinterpret_cast<double2 *>(outp) = o;
        if (FOLD) {   // the planes the z-neighbours need go straight into their ghost planes (warp-uniform tests)
          if (kb < MGIC_GZ && A.sw.peerLo) *reinterpret_cast<double2 *>(A.sw.peerLo + (outp - A.out)) = o;
          if (kb >= A.g.nz - MGIC_GZ && A.sw.peerHi) *reinterpret_cast<double2 *>(A.sw.peerHi + (outp - A.out)) = o;
        }
      }
    }
    outp += A.g.sz;
    if (doRed) {  // keep the black-cell halves of plane kr for the next step
      sa = E ? ca.x : ca.y; sl = E ? cl.x : cl.y; sr = E ? cr.x : cr.y;
      if (HAS_B) sb = E ? cb.x : cb.y;
    }
    __syncthreads();
    // plane kr's slot is dead: its phi pair was copied a step ago, its y neighbours and coefficients were read above
    if (tid == 0 && kr + NSLOT <= plast) issue(kr + NSLOT, SC);
  }

  template <int E0>
  __device__ __forceinline__ void march(double2 p1, double2 p2) {
    double2 p0 = make_double2(0.0, 0.0), p3 = p0;
    for (int it = 0;; it++) {
      const int kr = zs - 1 + 4 * it;
      if (kr > ze) break;
      step<E0, 0>(kr, it, p0, p1, p2, p3);
      if (kr + 1 > ze) break;
      step<1 - E0, 1>(kr + 1, it, p1, p2, p3, p0);
      if (kr + 2 > ze) break;
      step<E0, 2>(kr + 2, it, p2, p3, p0, p1);
      if (kr + 3 > ze) break;
      step<1 - E0, 3>(kr + 3, it, p3, p0, p1, p2);
    }
  }

  __device__ __forceinline__ void run(unsigned char *smem_raw) {
    slots = reinterpret_cast<double *>(smem_raw);
    redbuf = reinterpret_cast<double *>(smem_raw + (size_t)NSLOT * SLOT * sizeof(double));
    full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)NSLOT * SLOT * sizeof(double) + 2 * NW * 32 * sizeof(double));
    tid = threadIdx.x;
    x0 = blockIdx.x * TX; y0 = blockIdx.y * TY;
    // folded halo: the two chunks whose planes the neighbours wait for are scheduled first (blocks start in z order): chunk
    // order 0, last, 1, 2, ... -- their planes are across NVLink long before the grid drains
    int chunk = blockIdx.z;
    if (FOLD && gridDim.z > 2) chunk = (blockIdx.z == 0) ? 0 : (blockIdx.z == 1 ? (int)gridDim.z - 1 : (int)blockIdx.z - 1);
    zs = A.zbeg + chunk * A.zchunk; ze = min(zs + A.zchunk, A.zend);
    pfirst = zs - 2; plast = ze + 1;
    u64 epoch = 0;
    if (tid == 0) {
      for (int q = 0; q < NSLOT; q++) mbar_init(&full[q], 1);
      fence_mbar_init();
      if (FOLD) {
        // sweep number `epoch` of this rank.  The CTAs of the first / last z chunk read the lower / upper ghost planes and will
        // overwrite the neighbour's: the neighbour must have published its previous sweep (which filled the ghost planes read
        // here, and has finished reading the ones written here)
        epoch = *(volatile u64 *)&A.sw.ctl[MGIC_SW_EPOCH] + 1;
        if (zs == A.zbeg && A.sw.ctlLo) sweep_wait_ge(&A.sw.ctl[MGIC_SW_DATA_LO], epoch - 1);
        if (ze == A.zend && A.sw.ctlHi) sweep_wait_ge(&A.sw.ctl[MGIC_SW_DATA_HI], epoch - 1);
        asm volatile("fence.proxy.async;" ::: "memory");   // the TMA loads below read what the acquire made visible
      }
    }
    __syncthreads();
    if (tid == 0)
      for (int p = pfirst; p < pfirst + NSLOT && p <= plast; p++) issue(p, p - pfirst);
    lane = tid & 31; w = tid >> 5;
    const int rr = w + 1;
    const int x = x0 - 2 + 2 * lane, y = y0 - 2 + rr;
    inDom = (x >= 0 && x + 1 < A.g.nx && y >= 0 && y < A.g.ny);
    tile = inDom && w >= 1 && w <= TY && lane >= 1 && lane <= 30;
    bx0 = (x == 0); bxn = (x == A.g.nx - 2); by0 = (y == 0); byn = (y == A.g.ny - 1);
    anyxy = bx0 || bxn || by0 || byn;
    s = rr * RW + 2 * lane;
    cidx = w * RW + 2 * lane;
    outp = A.out + (x + (long long)y * A.g.sy + (long long)(zs - 2) * A.g.sz);  // plane kr-1 at the first step (kr = zs-1)
    ci = (rr >> 1) * CW + lane + 1; cim = ((rr - 1) >> 1) * CW + lane + 1; cip = ((rr + 1) >> 1) * CW + lane + 1;
    zloPhys = A.bc.type[4] != MGIC_FACE_INTERIOR; zhiPhys = A.bc.type[5] != MGIC_FACE_INTERIOR;
    sa = sl = sr = sb = 0.0;
    // prologue: the lane's pairs of planes zs-2 and zs-1
    mbar_wait(&full[0], 0);
    const double2 p1 = load_pair(zs - 2, slots);
    mbar_wait(&full[1], 0);
    const double2 p2 = load_pair(zs - 1, slots + SLOT);
    __syncthreads();
    if (tid == 0 && pfirst + NSLOT <= plast) issue(pfirst + NSLOT, 0);  // slot of plane zs-2 is free again
    const int e0 = ((y & 1) + zs - 1 + A.g.k0) & 1;  // warp-uniform: element of the pair that is red in plane zs-1
    if (e0) march<1>(p1, p2);
    else march<0>(p1, p2);
    if (FOLD) {
      if ((zs == A.zbeg && A.sw.ctlLo) || (ze == A.zend && A.sw.ctlHi)) {   // block-uniform: this CTA stored into a neighbour
        __threadfence_system();   // this thread's peer stores are visible before the CTA reports in
        __syncthreads();
      }
      if (tid == 0) {
        // the last CTA of the first / last chunk publishes the sweep to the lo / hi neighbour (I am its hi / lo neighbour)
        if (zs == A.zbeg && A.sw.ctlLo && atomicAdd(&A.sw.ctl[MGIC_SW_CNT_LO], 1ull) == (u64)A.ctasPerChunk - 1) {
          A.sw.ctl[MGIC_SW_CNT_LO] = 0;
          __threadfence_system();
          st_release_sys(&A.sw.ctlLo[MGIC_SW_DATA_HI], epoch);
        }
        if (ze == A.zend && A.sw.ctlHi && atomicAdd(&A.sw.ctl[MGIC_SW_CNT_HI], 1ull) == (u64)A.ctasPerChunk - 1) {
          A.sw.ctl[MGIC_SW_CNT_HI] = 0;
          __threadfence_system();
          st_release_sys(&A.sw.ctlHi[MGIC_SW_DATA_LO], epoch);
        }
        if (atomicAdd(&A.sw.ctl[MGIC_SW_CNT_ALL], 1ull) == (u64)A.ctasTotal - 1) {   // the whole grid has read the epoch: advance it
          A.sw.ctl[MGIC_SW_CNT_ALL] = 0;
          A.sw.ctl[MGIC_SW_EPOCH] = epoch;
          __threadfence();
        }
      }
    }
  }
};

template <int TY, bool HAS_B, int MODE, int MINB, bool FOLD, bool PATCH>
__global__ void __launch_bounds__(32 * (TY + 2), MINB)
k_gsrb_fused(const __grid_constant__ CUtensorMap tm_phi, const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_l,
             const __grid_constant__ CUtensorMap tm_r, const __grid_constant__ CUtensorMap tm_b, const __grid_constant__ CUtensorMap tm_c,
             const FusedArgs A) {
  using F = Fused<TY, HAS_B, MODE, FOLD, PATCH>;
  static_assert(F::PLANE_BYTES % 128 == 0 && F::CPLANE_BYTES % 128 == 0 && F::CTILE_BYTES % 128 == 0,
                "TMA destinations must stay 128-byte aligned");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  F f(A);
  f.tm_phi = &tm_phi; f.tm_a = &tm_a; f.tm_l = &tm_l; f.tm_r = &tm_r; f.tm_b = &tm_b; f.tm_c = &tm_c;
  f.run(smem_raw);
}

// ---- host side ---------------------------------------------------------------------------------------------------
template <int TY, bool HAS_B, int MODE, int MINB>
int launch_cfg(mgic_op *o, const double *in, double *outp, const mgic_field *r, const mgic_field *coarse, int zbeg, int zend,
               const SweepPeers &sw) {
  using F = Fused<TY, HAS_B, MODE>;
  mgic_ctx *c = o->ctx;
  // two builds of the kernel: the plain one (one rank, or halos exchanged by k_halo_push) and the one that stores its
  // boundary planes into the neighbours' ghost planes itself -- the plain one carries none of the other's instructions
  // ... and a third for rectangular AMR patches (coarse-fine faces by homogeneousCFInterp; patches are never z-slabs)
  auto kern = o->isPatch ? k_gsrb_fused<TY, HAS_B, MODE, MINB, false, true>
                         : (sw.ctl ? k_gsrb_fused<TY, HAS_B, MODE, MINB, true, false> : k_gsrb_fused<TY, HAS_B, MODE, MINB, false, false>);
  int &resident = *mgic_dev_cache(c->device, (const void *)kern, 0, 0);   // per device: the opt-in and the occupancy
  if (!resident) {
    MGIC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F::SMEM));
    int per = 1;
    MGIC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, F::NT, F::SMEM));
    resident = (per < 1 ? 1 : per) * c->numSMs;
  }
  FusedArgs A;
  A.g = o->geom();
  A.bc = o->bck(true);
  A.out = outp;
  A.alpha = o->alpha; A.beta = o->beta; A.dxinv = 1.0 / (o->dx * o->dx);
  const int np = A.g.nz + 2 * MGIC_GZ;
  const long long goff = (long long)MGIC_GZ * A.g.sz;
  CUtensorMap tp, ta, tl, tr, tb, tc;
  MGIC_TRY(make_tmap(&ta, o->a->p - goff, A.g.nx, A.g.ny, np, RW, F::NW));
  MGIC_TRY(make_tmap(&tl, o->lambda->p - goff, A.g.nx, A.g.ny, np, RW, F::NW));
  MGIC_TRY(make_tmap(&tr, r->p - goff, A.g.nx, A.g.ny, np, RW, F::NW));
  if (HAS_B) MGIC_TRY(make_tmap(&tb, o->b->p - goff, A.g.nx, A.g.ny, np, RW, F::NW));
  else tb = ta;
  if (MODE != MODE_ZERO) MGIC_TRY(make_tmap(&tp, in - goff, A.g.nx, A.g.ny, np, RW, F::RR));
  else tp = ta;
  if (MODE == MODE_PROLONG)
    MGIC_TRY(make_tmap(&tc, coarse->p - (long long)MGIC_GZ * coarse->sz, coarse->nx, coarse->ny, coarse->nz + 2 * MGIC_GZ, F::CW, F::CROWS));
  else tc = ta;
  const int tilesX = (A.g.nx + TX - 1) / TX, tilesY = (A.g.ny + TY - 1) / TY;
  const Plan pl = plan_chunks(tilesX * tilesY, zend - zbeg, resident);
  A.zchunk = pl.zchunk;
  A.zbeg = zbeg; A.zend = zend;
  A.redLo = (A.bc.type[4] == MGIC_FACE_INTERIOR) ? -1 : 0;
  A.redHi = (A.bc.type[5] == MGIC_FACE_INTERIOR) ? A.g.nz : A.g.nz - 1;
  dim3 grd(tilesX, tilesY, pl.nch);
  A.sw = sw;
  A.ctasPerChunk = tilesX * tilesY; A.ctasTotal = tilesX * tilesY * pl.nch;
  kern<<<grd, F::NT, F::SMEM, c->stream>>>(tp, ta, tl, tr, tb, tc, A);
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { mgic_set_error("kernel gsrb_fused: %s", cudaGetErrorString(e)); return MGIC_ERR_CUDA; }
  return MGIC_OK;
}

template <bool HAS_B, int MODE>
int launch_mode(mgic_op *o, const double *in, double *outp, const mgic_field *r, const mgic_field *coarse, int zbeg, int zend,
                const SweepPeers &sw) {
  switch (o->ctx->fusedCfg) {
    case 0: return launch_cfg<8, HAS_B, MODE, 2>(o, in, outp, r, coarse, zbeg, zend, sw);
    case 1: return launch_cfg<16, HAS_B, MODE, 1>(o, in, outp, r, coarse, zbeg, zend, sw);
    default:
      // two CTAs per SM need a slot ring of <= ~110 KB: 10 rows with three coefficient streams, 8 rows with four or with
      // the coarse tile of the prolonging sweep
      // (6 rows when both apply)
      if (HAS_B && MODE == MODE_PROLONG) return launch_cfg<6, HAS_B, MODE, 2>(o, in, outp, r, coarse, zbeg, zend, sw);
      if (HAS_B || MODE == MODE_PROLONG) return launch_cfg<8, HAS_B, MODE, 2>(o, in, outp, r, coarse, zbeg, zend, sw);
      return launch_cfg<10, HAS_B, MODE, 2>(o, in, outp, r, coarse, zbeg, zend, sw);
  }
}

int launch(mgic_op *o, int mode, const double *in, double *outp, const mgic_field *r, const mgic_field *coarse, int zbeg, int zend,
           const SweepPeers &sw = SweepPeers()) {
  if (o->b) {
    if (mode == MODE_ZERO) return launch_mode<true, MODE_ZERO>(o, in, outp, r, coarse, zbeg, zend, sw);
    if (mode == MODE_PROLONG) return launch_mode<true, MODE_PROLONG>(o, in, outp, r, coarse, zbeg, zend, sw);
    return launch_mode<true, MODE_PLAIN>(o, in, outp, r, coarse, zbeg, zend, sw);
  }
  if (mode == MODE_ZERO) return launch_mode<false, MODE_ZERO>(o, in, outp, r, coarse, zbeg, zend, sw);
  if (mode == MODE_PROLONG) return launch_mode<false, MODE_PROLONG>(o, in, outp, r, coarse, zbeg, zend, sw);
  return launch_mode<false, MODE_PLAIN>(o, in, outp, r, coarse, zbeg, zend, sw);
}

}  // namespace

namespace mgk {

// Periodic faces (TMA cannot wrap), odd nx and small (launch-latency bound) levels use the per-colour kernel, which
// computes the same bits.
bool gsrb_fused_applicable(const mgic_op *o) {
  // an AMR level that is one box: its coarse-fine ghosts are a function of the two cells inside the face, which the sweep has at
  // hand (the opposite neighbour); a union of boxes (masked level) has such faces anywhere inside its array: per-colour kernel
  if (o->isPatch && (o->mask || !o->ctx->fusedPatch)) return false;
  for (int d = 0; d < 3; d++)
    if (o->bc_lo[d] == MGIC_BC_PERIODIC) return false;
  const long long cells = (long long)o->n[0] * o->n[1] * o->nzl;
  if ((o->n[0] & 1) || o->n[0] < 8 || cells < o->ctx->fusedMinCells) return false;
  if (o->ctx->nranks > 1 && o->nzl < MGIC_GZ) return false;
  return true;
}

// relax(e, r, iterations) with fused sweeps.  Multi-rank: one 2-plane halo exchange of e per sweep (the neighbour's
// first plane is updated redundantly) instead of the reference's two 1-plane exchanges, plus one of rhs per call.
// The exchange runs on the context's communication stream WHILE the interior planes are swept; only the first and last
// OVL planes of the slab (two small launches) wait for it.
// first = FUSED_FROM_ZERO: e is known to be zero (setToZero + relax of [Chombo] MultiGrid::cycle); first =
// FUSED_PROLONG: e += prolong(coarse) is applied on the fly by the first sweep (prolongIncrement + relax).
int gsrb_fused(mgic_op *o, mgic_field *e, const mgic_field *r, int iterations, int first, const mgic_field *coarse,
               bool rhsHaloValid, bool eHaloValid, bool *pushed, bool coarseHaloValid) {
  if (pushed) *pushed = false;
  constexpr int OVL = 8;
  mgic_ctx *c = o->ctx;
  if (!o->scratch) MGIC_TRY(mgic_field_create(o, &o->scratch));
  const bool multi = c->nranks > 1 && !o->isGlobal;
  const Geom g = o->geom();
  if (first == FUSED_PROLONG && (coarse->nx & 1)) {
    // the coarse rows are not 16-byte multiples: no tensor map; prolong with the separate kernel, then plain sweeps
    MGIC_TRY(prolong(c, g, e->p, coarse->p, coarse->sy, coarse->sz));
    first = FUSED_PLAIN; coarse = nullptr; eHaloValid = false;
  }
  const bool overlap = multi && c->overlapHalo && !c->profiling && c->commStream && g.nz >= 4 * OVL;
  // halo folded into the sweep: every sweep stores its boundary planes into the neighbours' ghost planes of the array it writes
  // and synchronises through the MGIC_SW_* words, so only the FIRST sweep of a call may still need an exchange of e (when the
  // caller cannot vouch for its ghost planes).  Needs both ping-pong arrays mapped by the neighbours (collective, not while
  // capturing: vcycle_run prepares them before it captures).
  bool fold = multi && c->foldHalo && !overlap && c->sweep_peers && c->p2pHalo && g.nz >= 2 * MGIC_GZ;
  if (fold && c->array_prepare) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    MGIC_CUDA(cudaStreamIsCapturing(c->stream, &cs));
    if (cs == cudaStreamCaptureStatusNone) {
      MGIC_TRY(c->array_prepare(c, e));
      MGIC_TRY(c->array_prepare(c, o->scratch));
    }
  }
  for (int it = 0; it < iterations; it++) {
    const int mode = (it == 0) ? first : FUSED_PLAIN;
    SweepPeers sw = SweepPeers();
    if (fold) {
      SweepPeers in;   // the input array must be mapped too: the neighbours write its ghost planes in their next sweep
      MGIC_TRY(c->sweep_peers(c, e, o->scratch->base, &sw));
      MGIC_TRY(c->sweep_peers(c, e, e->base, &in));
      if (!sw.ok || !in.ok) { fold = false; sw = SweepPeers(); }
    }
    // the previous sweep of this call pushed the ghost planes of e (folded mode)
    const bool needE = multi && mode != FUSED_FROM_ZERO && !(it == 0 && eHaloValid) && !(fold && it > 0);
    const bool needR = multi && it == 0 && !rhsHaloValid;
    const bool needC = multi && mode == FUSED_PROLONG && !(fold && coarseHaloValid);   // (pushed by the coarser level's last sweep)
    if (overlap && (needE || needR || needC)) {
      // fork: the exchange depends on everything issued so far (the previous sweep wrote the planes being sent)
      MGIC_CUDA(cudaEventRecord(c->evFork, c->stream));
      MGIC_CUDA(cudaStreamWaitEvent(c->commStream, c->evFork, 0));
      c->haloStream = c->commStream;
      int rc = MGIC_OK;
      if (needR) rc = mgic_halo(o, const_cast<mgic_field *>(r), 1);
      if (rc == MGIC_OK && needC) rc = mgic_halo_shape(c, const_cast<mgic_field *>(coarse), 1);
      if (rc == MGIC_OK && needE) rc = mgic_halo(o, e, 2);
      c->haloStream = nullptr;
      MGIC_TRY(rc);
      MGIC_CUDA(cudaEventRecord(c->evJoin, c->commStream));
      MGIC_TRY(launch(o, mode, e->p, o->scratch->p, r, coarse, OVL, g.nz - OVL));     // interior: no ghost plane is read
      MGIC_CUDA(cudaStreamWaitEvent(c->stream, c->evJoin, 0));                       // join
      MGIC_TRY(launch(o, mode, e->p, o->scratch->p, r, coarse, 0, OVL));
      MGIC_TRY(launch(o, mode, e->p, o->scratch->p, r, coarse, g.nz - OVL, g.nz));
    } else {
      if (needR) MGIC_TRY(mgic_halo(o, const_cast<mgic_field *>(r), 1));
      if (needC) MGIC_TRY(mgic_halo_shape(c, const_cast<mgic_field *>(coarse), 1));
      if (needE) MGIC_TRY(mgic_halo(o, e, 2));
      ProfScope ps(c, o->profTag);
      MGIC_TRY(launch(o, mode, e->p, o->scratch->p, r, coarse, 0, g.nz, sw));
    }
    std::swap(e->base, o->scratch->base);  // ping-pong: the field handle now owns the freshly written array
    std::swap(e->p, o->scratch->p);
    if (pushed) *pushed = fold;   // the ghost planes of e are on their way from the neighbours' same sweep (sweep_wait before other readers)
  }
  return MGIC_OK;
}

}  // namespace mgk
