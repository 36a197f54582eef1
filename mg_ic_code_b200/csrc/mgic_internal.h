// mgic_internal.h -- internal structures shared by the CUDA translation units.
#ifndef MGIC_INTERNAL_H
#define MGIC_INTERNAL_H

#include <cuda_runtime.h>

#include <cstdio>
#include <string>
#include <vector>

#include "mgic.h"

// ---- error plumbing -------------------------------------------------------
void mgic_set_error(const char *fmt, ...);
#define MGIC_CUDA(call)                                                                         \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      mgic_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));    \
      return (e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver) ? MGIC_ERR_NO_DEVICE \
                                                                              : MGIC_ERR_CUDA;  \
    }                                                                                           \
  } while (0)
#define MGIC_TRY(call)          \
  do {                          \
    int r__ = (call);           \
    if (r__ != MGIC_OK) return r__; \
  } while (0)
#define MGIC_REQUIRE(cond, msg)                                  \
  do {                                                           \
    if (!(cond)) {                                               \
      mgic_set_error("%s:%d: %s", __FILE__, __LINE__, msg);      \
      return MGIC_ERR_ARG;                                       \
    }                                                            \
  } while (0)

// One-time kernel setup (cudaFuncSetAttribute opt-ins, occupancy queries) is PER DEVICE: a process may hold contexts on
// several devices.  mgic_dev_cache returns a stable int slot for (device, key), created with `init` on first use
// (mutex-protected map in capi.cu); callers fill it once per device.
int *mgic_dev_cache(int device, const void *key, int sub, int init);

// ---- kernel-side geometry ---------------------------------------------------
// Face order: 0 x-lo, 1 x-hi, 2 y-lo, 3 y-hi, 4 z-lo, 5 z-hi.
// type: 0 Dirichlet, 1 Neumann, 2 periodic, 3 interior (the neighbour plane is in memory: z-slab halo)
#define MGIC_FACE_INTERIOR 3
// 4 coarse-fine interface of an AMR patch, homogeneous form ([Chombo] homogeneousCFInterp): the ghost value is the
// parabola through the two interior cells and a zero coarse value -- a function of the current field like the physical
// ghosts, so it is evaluated in the stencil too (cf[] = m1, m2, idenom, q1, q2, x, x*x of INTERPHOMO)
#define MGIC_FACE_CF 4
// 5 coarse-fine interface with STORED ghost values ([Chombo] QuadCFInterp::coarseFineInterp: they depend on the coarser
// level's field): face[f] is a 2-D array over the face, x faces [j + ny*k], y faces [i + nx*k], z faces [i + nx*j]
#define MGIC_FACE_GHOST 5
struct BCk {
  int type[6];
  double a[6], b[6];  // ghost = a*centre + b   (Dirichlet: a=-1, b=2v; Neumann: a=+1, b=sign*dx*v)
  double cf[7];
  const double *face[6];
  // AMR level made of SEVERAL boxes (a union that is not a rectangle: touching boxes, L shapes, ...): the level lives in
  // ONE array over the union's bounding box and mask[cell] = 1 marks the cells of the level's boxes.  A neighbour that is
  // a masked-in cell is read from the array (what [Chombo] LevelData::exchange would have copied into the ghost cell); one
  // outside the domain gets the physical BC; any other is a coarse-fine ghost: homogeneousCFInterp on the fly, or -- when
  // face[f] is set -- QuadCFInterp's stored value, face[f] being a cell-indexed array (the ghost of cell idx beyond its
  // face f).  plo = the bounding box's lower corner and ndom the level's domain size, both in the level's index space.
  const unsigned char *mask;
  int plo[3], ndom[3];
};

struct Geom {
  int nx, ny, nz;     // local valid cells (z: this rank's slab)
  long long sy, sz;   // strides in doubles (sy = nx, sz = nx*ny)
  int k0;             // global k of local plane 0 (colouring uses global indices)
  int gnz;            // global nz
};

// ---- the fused sweep's own halo protocol (multi-rank, NVLink peer memory) ------------------------------------------
// Sweep number E of a context (all ranks launch the same sequence): the CTAs of the first / last z chunk store the planes
// they finish into the lo / hi neighbour's ghost planes of ITS output array, and the last of them to finish publishes E to
// that neighbour; before they touch a ghost plane they wait until the neighbour has published E-1 (its previous sweep has
// filled the ghost planes this sweep reads, and has stopped reading the ones this sweep overwrites -- the arrays ping-pong).
// Control words of a rank's IPC-exported block (comm.cu), written by the neighbours / by its own kernels:
enum { MGIC_SW_DATA_LO = 8, MGIC_SW_DATA_HI = 9,   // the lo / hi neighbour has published sweep ...
       MGIC_SW_EPOCH = 10,                         // sweeps this rank has finished
       MGIC_SW_CNT_LO = 11, MGIC_SW_CNT_HI = 12, MGIC_SW_CNT_ALL = 13 };   // CTAs done: first chunk, last chunk, grid
struct SweepPeers {
  double *peerLo, *peerHi;                       // the neighbours' arrays, offset so that "my plane 0" has my own offset arithmetic
  unsigned long long *ctl, *ctlLo, *ctlHi;
  int ok;
};

// ---- host-side objects ------------------------------------------------------
struct mgic_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool ownStream = false;
  long long launches = 0;
  int rank = 0, nranks = 1;
  int numSMs = 148;
  double *d_scal = nullptr;   // device scratch scalars (reduction results)
  double *h_scal = nullptr;   // pinned host mirror
  double *d_part = nullptr;   // reduction partials
  unsigned int *d_count = nullptr;
  size_t partCap = 0;
  // halo exchange hook (multi-GPU): set by mgic_comm; null on one GPU
  int (*halo_exchange)(mgic_ctx *, mgic_field *, int depth_planes) = nullptr;
  int (*allreduce)(mgic_ctx *, double *devvals, int n, int op /*0 sum 1 max*/) = nullptr;
  int (*allgather)(mgic_ctx *, const double *send, double *recv, size_t count) = nullptr;  // equal counts per rank
  int (*array_prepare)(mgic_ctx *, mgic_field *) = nullptr;  // collective: make the field's array exchangeable by peer stores (never inside a capture)
  int (*array_release)(mgic_ctx *, void *base) = nullptr;  // field array about to be freed; 1 = the hook frees it later
  int (*sweep_peers)(mgic_ctx *, const mgic_field *like, const double *outBase, SweepPeers *) = nullptr;  // fused sweep: where its boundary planes go
  int (*sweep_wait)(mgic_ctx *, const mgic_field *like) = nullptr;   // wait (on the stream) for the neighbours' last sweep
  int foldHalo = 0;      // 1: the fused sweep stores its boundary planes into the neighbours' ghost planes itself (measured at 2 GPUs,
                         // 512^3 per GPU: 6.38 vs 6.27 ms per V-cycle -- the exchanges cost waiting for the neighbour, not launches
                         // or data, and the wait moves into the sweep; off by default)
  void *comm = nullptr;
  int p2pHalo = 1;       // halo planes by NVLink peer stores (comm.cu k_halo_push) when the ranks could map each other
  // halo exchange overlapped with interior work: a second stream + fork/join events (capturable into a CUDA graph)
  cudaStream_t commStream = nullptr, haloStream = nullptr;  // haloStream != null: the halo hook issues on it
  cudaEvent_t evFork = nullptr, evJoin = nullptr;
  // host <-> HBM transfers overlapped with compute (mgic_field_prefetch / mgic_field_writeback): one stream per PCIe
  // direction, so that an upload, a V-cycle and a download can be in flight together
  cudaStream_t h2dStream = nullptr, d2hStream = nullptr;
  cudaEvent_t evXfer = nullptr;
  int overlapHalo = 0;   // measured on 2 and 8 GPUs: the exchange is too cheap next to a sweep for the split launches to pay
  // optional per-launch CUDA-event timing of the dominant kernel (finest-level GSRB), see mgic_ctx_profile
  // tuning knobs (mgic_ctx_set_option)
  int fusedCfg = 5;                       // tile configuration of the fused GSRB sweep (gsrb_fused.cu)
  long long fusedMinCells = 2097152;      // levels smaller than this use the per-colour kernel (launch-latency bound)
  int bottomKernel = 1;                   // bottom BiCGStab: 1 all vectors in one cluster's shared memory if the level fits
                                          // (bottom_dsmem.cu), else cluster-sized bricks (bottom_cbrick.cu), else as 4;
                                          // 5 cluster-sized bricks first; 4 brick kernel, 4 grid barriers / iteration (bottom_brick.cu);
                                          // 2 one kernel in a thread-block cluster, 3 same as a cooperative grid (bottom.cu);
                                          // 0 host-driven launches
  int lastBottomKernel = -1;               // which bottom solver ran last: 0 host, 1 dsmem, 2 cluster, 3 coop, 4 brick, 5 cluster bricks
  int fusedPatch = 1;                     // 1: rectangular AMR patches are swept by the fused kernel too (coarse-fine faces by homogeneousCFInterp in the sweep)
  int restrictTma = 1;                    // 1: restrictResidual of large rectangular levels by the plane-streaming kernel (restrict_tma.cu)
  int fusePR = 1;                         // 1: fold setToZero / prolongIncrement into the first fused sweep that follows
  long long aggloCells = 262144;          // multi-rank: depths whose slab has at most this many cells are agglomerated
  int useGraph = 1;                       // 1: replay each V-cycle as a CUDA graph
  long long cfgEpoch = 0;                 // bumped by every option / smoother change: captured graphs of older epochs are dropped
  bool profiling = false;
  struct ProfEv { cudaEvent_t a, b; int tag; };
  std::vector<ProfEv> profEvents;
};

// records a CUDA-event pair around the launches issued in its scope when profiling is armed
// tags: 0 finest-level GSRB (the roofline kernel), 1 halo exchange, 2 all-gather, 3 bottom solve, 4 restrict,
//       5 coarser-level GSRB, 6 prolong, 7 BLAS-1 / other
enum { PROF_GSRB0 = 0, PROF_HALO = 1, PROF_GATHER = 2, PROF_BOTTOM = 3, PROF_RESTRICT = 4, PROF_GSRBC = 5, PROF_PROLONG = 6, PROF_OTHER = 7 };
struct ProfScope {
  mgic_ctx *c; bool on; int tag; cudaEvent_t a = nullptr, b = nullptr;
  ProfScope(mgic_ctx *ctx, bool finest, int tg = -1) : c(ctx), on(ctx->profiling), tag(tg < 0 ? (finest ? PROF_GSRB0 : PROF_GSRBC) : tg) {
    if (on) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, c->stream); }
  }
  ~ProfScope() {
    if (on) { cudaEventRecord(b, c->stream); c->profEvents.push_back({a, b, tag}); }
  }
};

#define MGIC_GZ 2  // z ghost planes carried by every field (halo depth of one fused red+black sweep)

struct mgic_field {
  mgic_ctx *ctx = nullptr;
  int nx = 0, ny = 0, nz = 0, gz = MGIC_GZ;
  long long sy = 0, sz = 0;
  double *base = nullptr;  // first ghost plane
  double *p = nullptr;     // local cell (0,0,0)
  size_t bytes = 0;
  int k0 = 0, gnz = 0;
  bool noHalo = false;     // slab view into a whole-level array: its ghost planes are real neighbour planes
  bool zWrap = false;      // periodic in z on a multi-rank context: the slabs form a ring (rank 0's lo neighbour is the last rank)
  const unsigned char *mask = nullptr;  // masked AMR level (see BCk): cells outside the level's boxes are kept at zero
  cudaEvent_t evPending = nullptr;  // completion of the field's last prefetch / writeback (mgic_field_wait)
};

struct mgic_op {
  mgic_ctx *ctx = nullptr;
  int n[3] = {0, 0, 0};
  int k0 = 0, nzl = 0;
  double dx = 0, alpha = 0, beta = 0;
  int bc_lo[3] = {0, 0, 0}, bc_hi[3] = {0, 0, 0};
  double bc_value = 0;
  mgic_field *a = nullptr, *b = nullptr;  // shared, not owned
  mgic_field *lambda = nullptr;           // owned
  mgic_field *scratch = nullptr;          // owned; ping-pong target of the fused sweep
  bool lambdaDirty = true;
  // AMR patch (mgic_op_create_patch): n[] is the patch, cfLo/cfHi mark its coarse-fine faces, cshift keeps the GLOBAL-index
  // colouring ((lo0+lo1+lo2) & 1), dxCrse is the coarser level's spacing
  bool isPatch = false, cfLo[3] = {false, false, false}, cfHi[3] = {false, false, false};
  int cshift = 0;
  double dxCrse = 0;
  int plo[3] = {0, 0, 0}, ndom[3] = {0, 0, 0};   // the patch's lower corner and the level's domain size
  double *cfFace[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // QuadCFInterp ghost values per coarse-fine face
  double *cfCell[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // the same, cell-indexed (quad_cf_cells: psi update)
  // level = union of boxes inside the bounding box n[] (mgic_op_create_patch_boxes): device mask, host copy, valid cells
  unsigned char *mask = nullptr;
  std::vector<unsigned char> hmask;
  long long validCells = 0;
  bool isGlobal = false;                  // whole-domain operator on a multi-rank context (agglomerated level): no halos
  bool profTag = true;                    // finest-level operator: its GSRB launches are the profiled kernel
  int smoother = 1;                       // 0: one launch per colour; 1: fused red+black plane-streaming sweep
  Geom geom() const {
    Geom g; g.nx = n[0]; g.ny = n[1]; g.nz = nzl; g.sy = n[0]; g.sz = (long long)n[0] * n[1]; g.k0 = k0 + cshift; g.gnz = n[2];
    return g;
  }
  BCk bck(bool homogeneous) const;
};

struct mgic_vars {
  mgic_ctx *ctx = nullptr;
  mgic_params P;
  int n[3]; int k0 = 0, nzl = 0; double dx = 0;
  int ng = 1;
  long long sy = 0, sz = 0, sc = 0;  // padded strides
  double *d = nullptr;               // 8 components, each (nx+2)(ny+2)(nzl+2)
  // AMR level > 0 (mgic_vars_create_patch): bounding box lower corner in the level's index space (k0 = lo[2]), the level's
  // domain, the cell mask of a union of boxes (shared with the operator), psi's coarse-fine ghost values per face and cell
  bool isPatch = false;
  int lo[3] = {0, 0, 0}, ndom[3] = {0, 0, 0};
  const unsigned char *mask = nullptr;
  double *psiG[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

// ---- kernel launchers (kernels.cu) -------------------------------------------
// device allocations (capi.cu): ranges of a few large chunks per device unless `plain` (a cudaMalloc block of its own: what
// CUDA IPC can name) or large; counted and timed, see mgic_alloc_stats
cudaError_t mgic_dev_malloc_(void **p, size_t bytes, bool plain);
template <class T> inline cudaError_t mgic_dev_malloc(T **p, size_t bytes, bool plain = false) { return mgic_dev_malloc_((void **)p, bytes, plain); }
cudaError_t mgic_dev_free(void *p);

namespace mgk {
int gsrb_color(mgic_ctx *, const Geom &, const BCk &, double *phi, const double *rhs, const double *a, const double *b,
               const double *lam, double alpha, double beta, double dx, int color);
int apply_op(mgic_ctx *, const Geom &, const BCk &, double *lhs, const double *phi, const double *a, const double *b,
             double alpha, double beta, double dx);
int residual(mgic_ctx *, const Geom &, const BCk &, double *res, const double *phi, const double *rhs, const double *a,
             const double *b, double alpha, double beta, double dx);
// the same by the plane-streaming kernel (restrict_tma.cu): rectangular, non-periodic levels of at least 8 x fused_min_cells cells
bool restrict_tma_applicable(const mgic_op *);
int restrict_tma(mgic_op *, const BCk &, mgic_field *resC, const mgic_field *phi, const mgic_field *rhs);
int restrict_res(mgic_ctx *, const Geom &fine, const BCk &, double *resC, long long csy, long long csz, const double *phi,
                 const double *rhs, const double *a, const double *b, double alpha, double beta, double dx);
int prolong(mgic_ctx *, const Geom &fine, double *phi, const double *coarse, long long csy, long long csz,
            const unsigned char *fineMask = nullptr);
int apply_mask(mgic_ctx *, const Geom &, double *y, const unsigned char *mask);   // y = 0 outside the mask
int copy_box(mgic_ctx *, int nx, int ny, int nz, const double *src, long long ssy, long long ssz, double *dst, long long dsy, long long dsz);
// QuadCFInterp for every coarse-fine ghost of a masked level: face[f][cell] for the cells whose neighbour beyond face f is one
int quad_cf_masked(mgic_ctx *, const Geom &g, const unsigned char *mask, const int plo[3], const int ndom[3], double h, const double *phi,
                   const double *coarse, long long csy, long long csz, const int clo[3], double *const face[6]);
int compute_lambda(mgic_ctx *, const Geom &, double *lam, const double *a, double alpha, double beta, double dx);
int mult(mgic_ctx *, const Geom &, double *y, const double *x, const double *l);      // y = x*l
int incr(mgic_ctx *, const Geom &, double *y, const double *x, double s);             // y = y + s*x
int axby(mgic_ctx *, const Geom &, double *y, const double *x1, const double *x2, double a, double b);
int scale(mgic_ctx *, const Geom &, double *y, double s);
int assign(mgic_ctx *, const Geom &, double *y, const double *x);
int set_val(mgic_ctx *, const Geom &, double *y, double v);
int box_set_val(mgic_ctx *, const Geom &box, double *y, double v, const unsigned char *fineMask = nullptr, long long msy = 0,
                long long msz = 0);   // sub-box of a larger array (strides box.sy / box.sz); fineMask: only under masked-in fine cells
int box_set_u8(mgic_ctx *, const Geom &box, unsigned char *y, unsigned char v, const unsigned char *fineMask, long long msy, long long msz);
// QuadCFInterp on every coarse-fine face of a rectangular patch in one launch (face[f] null: no coarse-fine boundary there)
int quad_cf_faces(mgic_ctx *, const Geom &g, const int plo[3], const int ndom[3], double h, const double *phi, const double *coarse,
                  long long csy, long long csz, const int clo[3], double *const face[6]);
int jacobi_update(mgic_ctx *, const Geom &, double *phi, const double *res, const double *lam, double w);
// reductions: result left in ctx->d_scal[slot]; kind 0 max|x|, 1 sum|x|, 2 sum x^2, 3 sum x*y
int reduce(mgic_ctx *, const Geom &, const double *x, const double *y, int kind, int slot, const unsigned char *skip = nullptr);
// QuadCFInterp on one coarse-fine face (dir, side) of a patch: ghost values into `face`
int quad_cf_face(mgic_ctx *, const Geom &g, const int plo[3], const int ndom[3], double h, int dir, int side, const double *phi,
                 const double *coarse, long long csy, long long csz, const int clo[3], double *face);
int coarse_average(mgic_ctx *, const Geom &coarse, double *c, const double *fine, long long fsy, long long fsz, int nref,
                   int harmonic, const unsigned char *fineMask = nullptr);
int is_constant(mgic_ctx *, const Geom &, const double *x, double value, int slot);  // d_scal[slot] = #cells != value
// source terms (padded multigrid_vars arrays)
int init_conditions(mgic_vars *);
int set_rhs_acoef(mgic_vars *, double *rhs, double *acoef, double constant_K);
int update_psi(mgic_vars *, const Geom &, const BCk &, const double *dpsi);
int update_psi_patch(mgic_vars *, const BCk &withCfFaces, const double *dpsi);
int fill(mgic_ctx *, double *y, long long n, double value);
int output_box(mgic_vars *, const int lo[3], const int n[3], int ng, double constant_K, double *out);
int condition_box(mgic_ctx *, const mgic_params &P, double dx, const int lo[3], const int n[3], int mode, double *out);
int bottom_bicgstab(mgic_op *, mgic_field *e, const mgic_field *r, mgic_field *const work[8], double *part, int partCap,
                    int *d_out);
int bottom_bicgstab_dsmem(mgic_op *, mgic_field *e, const mgic_field *r, int *d_out, int *used);
int bottom_bicgstab_brick(mgic_op *, mgic_field *e, const mgic_field *r, mgic_field *const work[8], double *part, int partCap,
                          int *d_out, int *used);
int bottom_bicgstab_cbrick(mgic_op *, mgic_field *e, const mgic_field *r, mgic_field *const work[8], double *part, int partCap,
                           int *d_out, int *used);
// z-halo exchange of a field on the operator's level (no-op on one rank)
int mgic_halo(mgic_op *, mgic_field *, int planes);
bool gsrb_fused_applicable(const mgic_op *);
// relax(e, r, iterations) with the fused red+black sweep (gsrb_fused.cu); ping-pongs e with op->scratch
enum { FUSED_PLAIN = 0, FUSED_FROM_ZERO = 1, FUSED_PROLONG = 2 };
// *pushed (optional) = the last sweep stored e's boundary planes into the neighbours' ghost planes itself (halo folded into
// the sweep): a following non-sweep reader of e's ghost planes needs ctx->sweep_wait instead of an exchange
int gsrb_fused(mgic_op *, mgic_field *e, const mgic_field *r, int iterations, int first = FUSED_PLAIN,
               const mgic_field *coarse = nullptr, bool rhsHaloValid = false, bool eHaloValid = false, bool *pushed = nullptr,
               bool coarseHaloValid = false);
int mgic_halo_shape(mgic_ctx *, mgic_field *, int planes);  // halo exchange of a field of any level
// a ghosted FArrayBox staged in HBM (chf_abi.cu)
struct FabView { double *p; int lo[3]; long long s1, s2, sc; };
}  // namespace mgk

#endif
