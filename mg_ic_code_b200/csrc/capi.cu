// capi.cu -- the device-resident C ABI of include/mgic.h: context, level fields, the
// VariableCoeffPoissonOperator methods, the operator factory / MultiGrid hierarchy, the V-cycle with its
// bottom BiCGStab, the outer BiCGStab (f1) and the nonlinear loop of Main_PoissonSolver.cpp.
//
// Host code here only sequences kernel launches; all field data stays in HBM.  There is no CPU compute path.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>

#include "mgic_internal.h"

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[1024] = "";
void mgic_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char *mgic_last_error(void) { return g_err; }

int *mgic_dev_cache(int device, const void *key, int sub, int init) {
  static std::mutex mu;
  static std::map<std::tuple<int, const void *, int>, int> slots;   // node-based: the returned pointers stay valid
  std::lock_guard<std::mutex> lk(mu);
  auto it = slots.find(std::make_tuple(device, key, sub));
  if (it == slots.end()) it = slots.emplace(std::make_tuple(device, key, sub), init).first;
  return &it->second;
}
extern "C" const char *mgic_version(void) { return "mgic_b200 0.1 (sm_100a)"; }

// ------------------------------------------------------------------------------------------------ context
extern "C" int mgic_ctx_create(int device, mgic_ctx **out) {
  MGIC_REQUIRE(out, "out is NULL");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    mgic_set_error("no CUDA device available (%s): the B200 path has no CPU fallback",
                   e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return MGIC_ERR_NO_DEVICE;
  }
  MGIC_REQUIRE(device >= 0 && device < ndev, "bad device index");
  MGIC_CUDA(cudaSetDevice(device));
  mgic_ctx *c = new mgic_ctx;
  c->device = device;
  MGIC_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  c->ownStream = true;
  cudaDeviceProp prop;
  MGIC_CUDA(cudaGetDeviceProperties(&prop, device));
  c->numSMs = prop.multiProcessorCount;
  c->partCap = 4096;
  MGIC_CUDA(mgic_dev_malloc(&c->d_scal, 64 * sizeof(double)));
  MGIC_CUDA(cudaMemset(c->d_scal, 0, 64 * sizeof(double)));
  MGIC_CUDA(cudaMallocHost(&c->h_scal, 64 * sizeof(double)));
  MGIC_CUDA(mgic_dev_malloc(&c->d_part, c->partCap * sizeof(double)));
  MGIC_CUDA(mgic_dev_malloc(&c->d_count, sizeof(unsigned int)));
  MGIC_CUDA(cudaMemset(c->d_count, 0, sizeof(unsigned int)));
  *out = c;
  return MGIC_OK;
}

extern "C" int mgic_ctx_destroy(mgic_ctx *c) {
  if (!c) return MGIC_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  mgic_dev_free(c->d_scal);
  cudaFreeHost(c->h_scal);
  mgic_dev_free(c->d_part);
  mgic_dev_free(c->d_count);
  if (c->commStream) { cudaStreamDestroy(c->commStream); cudaEventDestroy(c->evFork); cudaEventDestroy(c->evJoin); }
  if (c->h2dStream) { cudaStreamSynchronize(c->h2dStream); cudaStreamSynchronize(c->d2hStream); cudaStreamDestroy(c->h2dStream); cudaStreamDestroy(c->d2hStream); cudaEventDestroy(c->evXfer); }
  if (c->ownStream) cudaStreamDestroy(c->stream);
  delete c;
  return MGIC_OK;
}

extern "C" int mgic_ctx_sync(mgic_ctx *c) {
  MGIC_REQUIRE(c, "ctx is NULL");
  MGIC_CUDA(cudaStreamSynchronize(c->stream));
  if (c->h2dStream) { MGIC_CUDA(cudaStreamSynchronize(c->h2dStream)); MGIC_CUDA(cudaStreamSynchronize(c->d2hStream)); }
  return MGIC_OK;
}

extern "C" int mgic_ctx_set_stream(mgic_ctx *c, void *s) {
  MGIC_REQUIRE(c, "ctx is NULL");
  MGIC_CUDA(cudaStreamSynchronize(c->stream));
  if (c->ownStream) {
    cudaStreamDestroy(c->stream);
    c->ownStream = false;
  }
  if (s) c->stream = (cudaStream_t)s;
  else {
    MGIC_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->ownStream = true;
  }
  return MGIC_OK;
}
extern "C" void *mgic_ctx_stream(mgic_ctx *c) { return c ? (void *)c->stream : nullptr; }
extern "C" long long mgic_ctx_launch_count(mgic_ctx *c) { return c ? c->launches : 0; }
extern "C" int mgic_ctx_set_rank(mgic_ctx *c, int rank, int nranks) {
  MGIC_REQUIRE(c && nranks >= 1 && rank >= 0 && rank < nranks, "bad rank/nranks");
  c->rank = rank;
  c->nranks = nranks;
  return MGIC_OK;
}

extern "C" int mgic_ctx_set_option(mgic_ctx *c, const char *name, long long value) {
  MGIC_REQUIRE(c && name, "NULL argument");
  if (!strcmp(name, "fused_cfg")) c->fusedCfg = (int)value;
  else if (!strcmp(name, "fused_min_cells")) c->fusedMinCells = value;
  else if (!strcmp(name, "bottom_kernel")) c->bottomKernel = (int)value;
  else if (!strcmp(name, "use_graph")) c->useGraph = (int)value;
  else if (!strcmp(name, "fuse_transfers")) c->fusePR = (int)value;
  else if (!strcmp(name, "fused_patch")) c->fusedPatch = (int)value;
  else if (!strcmp(name, "restrict_tma")) c->restrictTma = (int)value;
  else if (!strcmp(name, "agglo_cells")) c->aggloCells = value;
  else if (!strcmp(name, "overlap_halo")) c->overlapHalo = (int)value;
  else if (!strcmp(name, "p2p_halo")) c->p2pHalo = (int)value;
  else if (!strcmp(name, "fold_halo")) c->foldHalo = (int)value;
  else { mgic_set_error("unknown option %s", name); return MGIC_ERR_ARG; }
  c->cfgEpoch++;   // captured V-cycle graphs bake the options in: they are re-captured after any change
  return MGIC_OK;
}
extern "C" long long mgic_ctx_get_option(mgic_ctx *c, const char *name) {
  if (!c || !name) return -1;
  if (!strcmp(name, "fused_cfg")) return c->fusedCfg;
  if (!strcmp(name, "fused_min_cells")) return c->fusedMinCells;
  if (!strcmp(name, "bottom_kernel")) return c->bottomKernel;
  if (!strcmp(name, "use_graph")) return c->useGraph;
  if (!strcmp(name, "fuse_transfers")) return c->fusePR;
  if (!strcmp(name, "fused_patch")) return c->fusedPatch;
  if (!strcmp(name, "restrict_tma")) return c->restrictTma;
  if (!strcmp(name, "agglo_cells")) return c->aggloCells;
  if (!strcmp(name, "overlap_halo")) return c->overlapHalo;
  if (!strcmp(name, "p2p_halo")) return c->p2pHalo;
  if (!strcmp(name, "fold_halo")) return c->foldHalo;
  if (!strcmp(name, "last_bottom_kernel")) return c->lastBottomKernel;
  return -1;
}
extern "C" int mgic_ctx_profile(mgic_ctx *c, int enable) {
  MGIC_REQUIRE(c, "ctx is NULL");
  MGIC_CUDA(cudaStreamSynchronize(c->stream));
  for (auto &ev : c->profEvents) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
  c->profEvents.clear();
  c->profiling = enable != 0;
  return MGIC_OK;
}
extern "C" int mgic_ctx_profile_read_tag(mgic_ctx *c, int tag, long long *launches, double *total_ms) {
  MGIC_REQUIRE(c && launches && total_ms, "NULL argument");
  MGIC_CUDA(cudaStreamSynchronize(c->stream));
  double tot = 0.0;
  long long n = 0;
  for (auto &ev : c->profEvents) {
    if (ev.tag != tag) continue;
    float ms = 0.f;
    MGIC_CUDA(cudaEventElapsedTime(&ms, ev.a, ev.b));
    tot += ms;
    n++;
  }
  *launches = n;
  *total_ms = tot;
  return MGIC_OK;
}
extern "C" int mgic_ctx_profile_read(mgic_ctx *c, long long *launches, double *total_ms) {
  return mgic_ctx_profile_read_tag(c, PROF_GSRB0, launches, total_ms);
}

// read back `n` device scalars starting at slot (one sync); multi-rank: all-reduced on the device first (op 0 sum, 1 max)
static int fetch_scalars(mgic_ctx *c, int slot, int n, int op, double *out, bool collective = true) {
  if (c->nranks > 1 && collective) {
    MGIC_REQUIRE(c->allreduce, "multi-rank context without an allreduce hook (mgic_comm_init)");
    MGIC_TRY(c->allreduce(c, c->d_scal + slot, n, op));
  }
  MGIC_CUDA(cudaMemcpyAsync(c->h_scal + slot, c->d_scal + slot, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  MGIC_CUDA(cudaStreamSynchronize(c->stream));
  for (int q = 0; q < n; q++) out[q] = c->h_scal[slot + q];
  return MGIC_OK;
}

// ------------------------------------------------------------------------------------------------ operator
BCk mgic_op::bck(bool homogeneous) const {
  BCk k;
  const double v = homogeneous ? 0.0 : bc_value;
  for (int f = 0; f < 6; f++) {
    const int dir = f / 2, side = (f % 2) ? +1 : -1;
    int t = (side < 0) ? bc_lo[dir] : bc_hi[dir];
    k.type[f] = t;
    if (t == MGIC_BC_DIRICHLET) { k.a[f] = -1.0; k.b[f] = 2 * v; }               // DiriBC order 1: 2*v - near
    else if (t == MGIC_BC_NEUMANN) { k.a[f] = 1.0; k.b[f] = side * dx * v; }     // NeumBC: near + side*dx*v
    else { k.a[f] = 0.0; k.b[f] = 0.0; }
  }
  if (k0 > 0) k.type[4] = MGIC_FACE_INTERIOR;
  if (k0 + nzl < n[2]) k.type[5] = MGIC_FACE_INTERIOR;
  // periodic in z across ranks: the slabs form a ring, the wrap-around planes arrive by the halo exchange -- but a whole-level
  // (agglomerated / replicated) operator on a multi-rank context wraps its own indices like a single-GPU one
  if (ctx->nranks > 1 && !isGlobal && bc_lo[2] == MGIC_BC_PERIODIC) { k.type[4] = MGIC_FACE_INTERIOR; k.type[5] = MGIC_FACE_INTERIOR; }
  for (int q = 0; q < 7; q++) k.cf[q] = 0.0;
  for (int f = 0; f < 6; f++) k.face[f] = nullptr;
  k.mask = mask;
  for (int d = 0; d < 3; d++) { k.plo[d] = plo[d]; k.ndom[d] = isPatch ? ndom[d] : n[d]; }
  if (isPatch) {
    // INTERPHOMO's constants, computed in its order ([Chombo] AMRPoissonOpF.ChF; oracle: Op::homogeneousCFInterp)
    const double x1 = dx;
    const double x2 = 0.5 * (3. * x1 + dxCrse);
    const double denom = 1.0 - ((x1 + x2) / x1);
    const double x = 2. * x1;
    k.cf[0] = 1 / (x1 * x1);          // m1
    k.cf[1] = 1 / (x1 * (x1 - x2));   // m2
    k.cf[2] = 1 / (denom);            // idenom
    k.cf[3] = 1 / (x1 - x2);          // q1
    k.cf[4] = x1 + x2;                // q2
    k.cf[5] = x;
    k.cf[6] = x * x;                  // xsquared
    for (int d = 0; d < 3; d++) {
      if (cfLo[d]) k.type[2 * d] = MGIC_FACE_CF;
      if (cfHi[d]) k.type[2 * d + 1] = MGIC_FACE_CF;
    }
  }
  return k;
}

static int field_alloc(mgic_ctx *c, int nx, int ny, int nz, int k0, int gnz, mgic_field **out, bool exported = true) {
  mgic_field *f = new mgic_field;
  f->ctx = c;
  f->nx = nx; f->ny = ny; f->nz = nz;
  f->sy = nx; f->sz = (long long)nx * ny;
  f->k0 = k0; f->gnz = gnz;
  f->bytes = (size_t)f->sz * (nz + 2 * MGIC_GZ) * sizeof(double);
  size_t alloc = f->bytes;
  // multi-rank: the arrays of a z-slab level are exported to the z-neighbours by CUDA IPC, which names whole cudaMalloc
  // blocks; allocations of at least 2 MiB get a block of their own.  Whole-level arrays (agglomerated MG depths, replicated
  // AMR patches, staging boxes) are never exported.
  const bool plain = c->nranks > 1 && exported;
  if (plain) alloc = (std::max(alloc, (size_t)1) + ((size_t)2 << 20) - 1) / ((size_t)2 << 20) * ((size_t)2 << 20);
  cudaError_t e = mgic_dev_malloc(&f->base, alloc, plain);
  if (e != cudaSuccess) {
    mgic_set_error("cudaMalloc(%zu bytes) failed: %s", f->bytes, cudaGetErrorString(e));
    delete f;
    return MGIC_ERR_CUDA;
  }
  MGIC_CUDA(cudaMemsetAsync(f->base, 0, f->bytes, c->stream));
  f->p = f->base + (long long)MGIC_GZ * f->sz;
  *out = f;
  return MGIC_OK;
}

extern "C" int mgic_op_create(mgic_ctx *c, const int n[3], int k0, int nz_local, double dx, double alpha, double beta,
                              const int bc_lo[3], const int bc_hi[3], double bc_value, mgic_op **out) {
  MGIC_REQUIRE(c && n && out && bc_lo && bc_hi, "NULL argument");
  MGIC_REQUIRE(n[0] >= 1 && n[1] >= 1 && n[2] >= 1, "bad dims");
  MGIC_REQUIRE(k0 >= 0 && nz_local >= 1 && k0 + nz_local <= n[2], "bad slab");
  for (int d = 0; d < 3; d++) {
    if (bc_lo[d] < 0 || bc_lo[d] > 2) { mgic_set_error("bogus bc flag low side %d", bc_lo[d]); return MGIC_ERR_ARG; }    // SetBCs.cpp:94
    if (bc_hi[d] < 0 || bc_hi[d] > 2) { mgic_set_error("bogus bc flag high side %d", bc_hi[d]); return MGIC_ERR_ARG; }   // SetBCs.cpp:123
    MGIC_REQUIRE((bc_lo[d] == MGIC_BC_PERIODIC) == (bc_hi[d] == MGIC_BC_PERIODIC), "periodic must be set on both sides");
  }
  MGIC_CUDA(cudaSetDevice(c->device));
  mgic_op *o = new mgic_op;
  o->ctx = c;
  for (int d = 0; d < 3; d++) { o->n[d] = n[d]; o->bc_lo[d] = bc_lo[d]; o->bc_hi[d] = bc_hi[d]; }
  o->k0 = k0; o->nzl = nz_local; o->dx = dx; o->alpha = alpha; o->beta = beta; o->bc_value = bc_value;
  *out = o;
  return MGIC_OK;
}

// One AMR level > 0: a box [lo, hi] of the refined domain (VariableCoeffPoissonOperator.cpp:156,296: the operator's own
// coarse-fine code is homogeneousCFInterp before levelGSRB's colour passes and before restrictResidual)
extern "C" int mgic_op_create_patch(mgic_ctx *c, const int n_domain[3], const int lo[3], const int hi[3], double dx, double dx_coarse,
                                    double alpha, double beta, const int bc_lo[3], const int bc_hi[3], double bc_value, mgic_op **out) {
  MGIC_REQUIRE(c && n_domain && lo && hi && out && bc_lo && bc_hi, "NULL argument");
  int n[3];
  for (int d = 0; d < 3; d++) {
    MGIC_REQUIRE(lo[d] >= 0 && hi[d] < n_domain[d] && hi[d] - lo[d] + 1 >= 2, "patch box outside the domain or thinner than two cells");
    MGIC_REQUIRE(lo[d] % 2 == 0 && hi[d] % 2 == 1, "patch box must be coarsenable by 2");
    n[d] = hi[d] - lo[d] + 1;
  }
  MGIC_REQUIRE(dx_coarse > dx && dx > 0, "the coarser level's spacing must exceed the patch's");
  MGIC_TRY(mgic_op_create(c, n, 0, n[2], dx, alpha, beta, bc_lo, bc_hi, bc_value, out));
  mgic_op *o = *out;
  o->isPatch = true;
  // multi-rank context: the base level is cut into z-slabs, the (small) refined levels are REPLICATED -- every rank holds and
  // updates the whole patch, like the agglomerated coarse MG levels: no halo planes, no all-reduce of its reductions
  o->isGlobal = c->nranks > 1;
  o->dxCrse = dx_coarse;
  o->cshift = (lo[0] + lo[1] + lo[2]) & 1;
  o->smoother = 1;  // (the fused sweep takes rectangular patches; unions of boxes fall back to the per-colour kernel: gsrb_fused_applicable)
  for (int d = 0; d < 3; d++) {
    o->cfLo[d] = lo[d] > 0;
    o->cfHi[d] = hi[d] < n_domain[d] - 1;
    o->plo[d] = lo[d]; o->ndom[d] = n_domain[d];
  }
  return MGIC_OK;
}

// One AMR level > 0 (or one connected part of it) made of SEVERAL boxes of the refined domain -- BRMeshRefine's output:
// boxes of at most max_grid_size cells that touch and whose union is no rectangle.  B200 layout: ONE array over the
// union's bounding box plus a cell mask, instead of one FAB per box: the fine-fine ghost exchange between the level's
// boxes ([Chombo] LevelData::exchange, VariableCoeffPoissonOperator.cpp:301) becomes a plain neighbour read, coarse-fine
// ghosts are evaluated per cell (BCk::mask), and every kernel runs once per level instead of once per box.  Results do
// not depend on how the union is cut into boxes (global-index colouring, exchange before each colour pass).
// boxes = nboxes x {lo0, lo1, lo2, hi0, hi1, hi2}, inclusive, in the level's index space.
extern "C" int mgic_op_create_patch_boxes(mgic_ctx *c, const int n_domain[3], int nboxes, const int *boxes, double dx, double dx_coarse,
                                          double alpha, double beta, const int bc_lo[3], const int bc_hi[3], double bc_value, mgic_op **out) {
  MGIC_REQUIRE(c && n_domain && boxes && nboxes >= 1 && out && bc_lo && bc_hi, "bad argument");
  int lo[3] = {1 << 30, 1 << 30, 1 << 30}, hi[3] = {-1, -1, -1};
  for (int q = 0; q < nboxes; q++)
    for (int d = 0; d < 3; d++) {
      const int l = boxes[6 * q + d], h = boxes[6 * q + 3 + d];
      MGIC_REQUIRE(l >= 0 && h < n_domain[d] && h - l + 1 >= 2, "box outside the domain or thinner than two cells");
      MGIC_REQUIRE(l % 2 == 0 && h % 2 == 1, "boxes must be coarsenable by 2");
      lo[d] = std::min(lo[d], l); hi[d] = std::max(hi[d], h);
    }
  MGIC_TRY(mgic_op_create_patch(c, n_domain, lo, hi, dx, dx_coarse, alpha, beta, bc_lo, bc_hi, bc_value, out));
  mgic_op *o = *out;
  const size_t nx = o->n[0], ny = o->n[1], nz = o->n[2];
  o->hmask.assign(nx * ny * nz, 0);
  for (int q = 0; q < nboxes; q++) {
    const int *b = boxes + 6 * q;
    for (int k = b[2]; k <= b[5]; k++)
      for (int j = b[1]; j <= b[4]; j++) {
        unsigned char *row = o->hmask.data() + (size_t)(b[0] - lo[0]) + nx * ((size_t)(j - lo[1]) + ny * (size_t)(k - lo[2]));
        for (int i = b[0]; i <= b[3]; i++) {
          if (row[i - b[0]]) { mgic_set_error("boxes %d overlaps an earlier box of the level (a DisjointBoxLayout is disjoint)", q); mgic_op_destroy(o); *out = nullptr; return MGIC_ERR_ARG; }
          row[i - b[0]] = 1;
        }
      }
  }
  o->validCells = 0;
  for (unsigned char m : o->hmask) o->validCells += m;
  if ((size_t)o->validCells == nx * ny * nz) { o->hmask.clear(); return MGIC_OK; }   // the union IS the bounding box: the rectangular patch
  // every masked-in cell must have a second masked-in cell inwards of each coarse-fine face (homogeneousCFInterp / QUADINTERP
  // read two interior cells): guaranteed by boxes >= 2 cells thick, checked here for unions
  MGIC_CUDA(mgic_dev_malloc(&o->mask, o->hmask.size()));
  MGIC_CUDA(cudaMemcpyAsync(o->mask, o->hmask.data(), o->hmask.size(), cudaMemcpyHostToDevice, c->stream));
  MGIC_CUDA(cudaStreamSynchronize(c->stream));
  return MGIC_OK;
}
extern "C" long long mgic_op_valid_cells(const mgic_op *o) {
  if (!o) return 0;
  return o->mask ? o->validCells : (long long)o->n[0] * o->n[1] * o->nzl;
}
// the level's cell mask over the bounding box (1 = a cell of the level's boxes), x fastest; all ones for a rectangular level
extern "C" int mgic_op_get_mask(const mgic_op *o, unsigned char *host) {
  MGIC_REQUIRE(o && host, "NULL argument");
  const size_t n = (size_t)o->n[0] * o->n[1] * o->nzl;
  if (o->mask) memcpy(host, o->hmask.data(), n);
  else memset(host, 1, n);
  return MGIC_OK;
}

extern "C" int mgic_op_destroy(mgic_op *o) {
  if (!o) return MGIC_OK;
  mgic_field_destroy(o->lambda);
  mgic_field_destroy(o->scratch);
  for (int f = 0; f < 6; f++) { mgic_dev_free(o->cfFace[f]); mgic_dev_free(o->cfCell[f]); }
  mgic_dev_free(o->mask);
  delete o;
  return MGIC_OK;
}

extern "C" int mgic_op_dims(const mgic_op *o, int n[3], int *k0, int *nzl, double *dx) {
  MGIC_REQUIRE(o, "op is NULL");
  if (n) for (int d = 0; d < 3; d++) n[d] = o->n[d];
  if (k0) *k0 = o->k0;
  if (nzl) *nzl = o->nzl;
  if (dx) *dx = o->dx;
  return MGIC_OK;
}

static bool same_shape(const mgic_op *o, const mgic_field *f) {
  return f && f->nx == o->n[0] && f->ny == o->n[1] && f->nz == o->nzl && f->k0 == o->k0;
}
#define REQ_SHAPE(o, f) MGIC_REQUIRE(same_shape(o, f), "field " #f " does not live on this operator's level")

extern "C" int mgic_op_set_coefs(mgic_op *o, mgic_field *a, mgic_field *b, double alpha, double beta) {
  MGIC_REQUIRE(o && a, "NULL argument");
  REQ_SHAPE(o, a);
  if (b) REQ_SHAPE(o, b);
  o->a = a; o->b = b; o->alpha = alpha; o->beta = beta;
  o->lambdaDirty = true;  // VariableCoeffPoissonOperator.cpp:216-217
  return MGIC_OK;
}
extern "C" int mgic_op_set_alpha_beta(mgic_op *o, double alpha, double beta) {
  MGIC_REQUIRE(o, "op is NULL");
  o->alpha = alpha; o->beta = beta;
  o->lambdaDirty = true;  // :204-205
  return MGIC_OK;
}
extern "C" int mgic_op_reset_lambda(mgic_op *o) {
  MGIC_REQUIRE(o && o->a, "operator has no coefficients (setCoefs)");
  if (!o->lambda) {
    MGIC_TRY(field_alloc(o->ctx, o->n[0], o->n[1], o->nzl, o->k0, o->n[2], &o->lambda, !o->isGlobal));
    o->lambda->zWrap = o->ctx->nranks > 1 && !o->isGlobal && o->bc_lo[2] == MGIC_BC_PERIODIC;
  }
  if (!o->lambdaDirty) return MGIC_OK;
  MGIC_TRY(mgk::compute_lambda(o->ctx, o->geom(), o->lambda->p, o->a->p, o->alpha, o->beta, o->dx));
  if (o->ctx->nranks > 1 && !o->isGlobal) {  // the fused sweep updates the neighbour's first plane redundantly: it needs its coefficients
    MGIC_TRY(mgk::mgic_halo(o, o->a, 1));
    MGIC_TRY(mgk::mgic_halo(o, o->lambda, 1));
    if (o->b) MGIC_TRY(mgk::mgic_halo(o, o->b, 1));
  }
  o->lambdaDirty = false;
  return MGIC_OK;
}
extern "C" int mgic_op_compute_lambda(mgic_op *o) {
  MGIC_REQUIRE(o, "op is NULL");
  o->lambdaDirty = true;
  return mgic_op_reset_lambda(o);
}
extern "C" int mgic_op_get_lambda(mgic_op *o, mgic_field **l) {
  MGIC_REQUIRE(o && l, "NULL argument");
  MGIC_TRY(mgic_op_reset_lambda(o));
  *l = o->lambda;
  return MGIC_OK;
}

// ------------------------------------------------------------------------------------------------ fields
extern "C" int mgic_field_create(mgic_op *like, mgic_field **out) {
  MGIC_REQUIRE(like && out, "NULL argument");
  MGIC_CUDA(cudaSetDevice(like->ctx->device));
  MGIC_TRY(field_alloc(like->ctx, like->n[0], like->n[1], like->nzl, like->k0, like->n[2], out, !like->isGlobal));
  (*out)->mask = like->mask;
  (*out)->zWrap = like->ctx->nranks > 1 && !like->isGlobal && like->bc_lo[2] == MGIC_BC_PERIODIC;
  return MGIC_OK;
}
extern "C" int mgic_field_destroy(mgic_field *f) {
  if (!f) return MGIC_OK;
  mgic_ctx *c = f->ctx;
  if (f->evPending) { cudaEventSynchronize(f->evPending); cudaEventDestroy(f->evPending); }
  if (!(c && c->array_release && c->array_release(c, f->base))) mgic_dev_free(f->base);
  delete f;
  return MGIC_OK;
}
extern "C" int mgic_field_upload(mgic_field *f, const double *host) {
  MGIC_REQUIRE(f && host, "NULL argument");
  const size_t n = (size_t)f->sz * f->nz;
  MGIC_CUDA(cudaMemcpyAsync(f->p, host + (size_t)f->k0 * f->sz, n * sizeof(double), cudaMemcpyHostToDevice, f->ctx->stream));
  if (f->mask) {   // masked AMR level: whatever the caller holds outside the level's boxes is not data
    Geom g; g.nx = f->nx; g.ny = f->ny; g.nz = f->nz; g.sy = f->sy; g.sz = f->sz; g.k0 = 0; g.gnz = f->nz;
    MGIC_TRY(mgk::apply_mask(f->ctx, g, f->p, f->mask));
  }
  MGIC_CUDA(cudaStreamSynchronize(f->ctx->stream));
  return MGIC_OK;
}
extern "C" int mgic_field_download(const mgic_field *f, double *host) {
  MGIC_REQUIRE(f && host, "NULL argument");
  const size_t n = (size_t)f->sz * f->nz;
  MGIC_CUDA(cudaMemcpyAsync(host + (size_t)f->k0 * f->sz, f->p, n * sizeof(double), cudaMemcpyDeviceToHost, f->ctx->stream));
  MGIC_CUDA(cudaStreamSynchronize(f->ctx->stream));
  return MGIC_OK;
}
// asynchronous variants for pinned buffers (e2e timing: copies ordered on the context stream)
extern "C" int mgic_field_upload_async(mgic_field *f, const double *host) {
  MGIC_REQUIRE(f && host, "NULL argument");
  const size_t n = (size_t)f->sz * f->nz;
  MGIC_CUDA(cudaMemcpyAsync(f->p, host + (size_t)f->k0 * f->sz, n * sizeof(double), cudaMemcpyHostToDevice, f->ctx->stream));
  return MGIC_OK;
}
extern "C" int mgic_field_download_async(const mgic_field *f, double *host) {
  MGIC_REQUIRE(f && host, "NULL argument");
  const size_t n = (size_t)f->sz * f->nz;
  MGIC_CUDA(cudaMemcpyAsync(host + (size_t)f->k0 * f->sz, f->p, n * sizeof(double), cudaMemcpyDeviceToHost, f->ctx->stream));
  return MGIC_OK;
}
// transfers on their own streams (one per PCIe direction), ordered after everything issued so far on the compute stream
static int xfer_streams(mgic_ctx *c) {
  if (c->h2dStream) return MGIC_OK;
  MGIC_CUDA(cudaStreamCreateWithFlags(&c->h2dStream, cudaStreamNonBlocking));
  MGIC_CUDA(cudaStreamCreateWithFlags(&c->d2hStream, cudaStreamNonBlocking));
  MGIC_CUDA(cudaEventCreateWithFlags(&c->evXfer, cudaEventDisableTiming));
  return MGIC_OK;
}
static int xfer(mgic_field *f, cudaStream_t st, void *dst, const void *src, cudaMemcpyKind kind) {
  mgic_ctx *c = f->ctx;
  if (!f->evPending) MGIC_CUDA(cudaEventCreateWithFlags(&f->evPending, cudaEventDisableTiming));
  MGIC_CUDA(cudaEventRecord(c->evXfer, c->stream));
  MGIC_CUDA(cudaStreamWaitEvent(st, c->evXfer, 0));
  MGIC_CUDA(cudaMemcpyAsync(dst, src, (size_t)f->sz * f->nz * sizeof(double), kind, st));
  MGIC_CUDA(cudaEventRecord(f->evPending, st));
  return MGIC_OK;
}
extern "C" int mgic_field_prefetch(mgic_field *f, const double *host) {
  MGIC_REQUIRE(f && host, "NULL argument");
  MGIC_TRY(xfer_streams(f->ctx));
  return xfer(f, f->ctx->h2dStream, f->p, host + (size_t)f->k0 * f->sz, cudaMemcpyHostToDevice);
}
extern "C" int mgic_field_writeback(mgic_field *f, double *host) {
  MGIC_REQUIRE(f && host, "NULL argument");
  MGIC_TRY(xfer_streams(f->ctx));
  return xfer(f, f->ctx->d2hStream, host + (size_t)f->k0 * f->sz, f->p, cudaMemcpyDeviceToHost);
}
extern "C" int mgic_field_wait(mgic_field *f) {
  MGIC_REQUIRE(f, "field is NULL");
  if (f->evPending) MGIC_CUDA(cudaStreamWaitEvent(f->ctx->stream, f->evPending, 0));
  return MGIC_OK;
}

// FArrayBox <-> level array: copies fab ∩ region ∩ slab with cudaMemcpy3D (strided, no staging)
static int fab_copy(const mgic_field *f, double *fab, const int flo[3], const int fhi[3], const int rlo[3], const int rhi[3],
                    bool toDevice) {
  int lo[3], hi[3];
  const int dlo[3] = {0, 0, f->k0}, dhi[3] = {f->nx - 1, f->ny - 1, f->k0 + f->nz - 1};
  for (int d = 0; d < 3; d++) {
    lo[d] = std::max(std::max(flo[d], rlo[d]), dlo[d]);
    hi[d] = std::min(std::min(fhi[d], rhi[d]), dhi[d]);
    if (hi[d] < lo[d]) return MGIC_OK;
  }
  const size_t fnx = fhi[0] - flo[0] + 1, fny = fhi[1] - flo[1] + 1;
  cudaMemcpy3DParms p;
  memset(&p, 0, sizeof(p));
  cudaPitchedPtr hp = make_cudaPitchedPtr(fab, fnx * sizeof(double), fnx, fny);
  cudaPitchedPtr dp = make_cudaPitchedPtr(f->p, (size_t)f->nx * sizeof(double), f->nx, f->ny);
  cudaPos hpos = make_cudaPos((size_t)(lo[0] - flo[0]) * sizeof(double), lo[1] - flo[1], lo[2] - flo[2]);
  cudaPos dpos = make_cudaPos((size_t)lo[0] * sizeof(double), lo[1], lo[2] - f->k0);
  p.extent = make_cudaExtent((size_t)(hi[0] - lo[0] + 1) * sizeof(double), hi[1] - lo[1] + 1, hi[2] - lo[2] + 1);
  if (toDevice) { p.srcPtr = hp; p.srcPos = hpos; p.dstPtr = dp; p.dstPos = dpos; p.kind = cudaMemcpyHostToDevice; }
  else { p.srcPtr = dp; p.srcPos = dpos; p.dstPtr = hp; p.dstPos = hpos; p.kind = cudaMemcpyDeviceToHost; }
  MGIC_CUDA(cudaMemcpy3DAsync(&p, f->ctx->stream));
  return MGIC_OK;
}
extern "C" int mgic_field_upload_fab(mgic_field *f, const double *fab, const int flo[3], const int fhi[3], const int rlo[3],
                                     const int rhi[3]) {
  MGIC_REQUIRE(f && fab && flo && fhi && rlo && rhi, "NULL argument");
  return fab_copy(f, const_cast<double *>(fab), flo, fhi, rlo, rhi, true);
}
extern "C" int mgic_field_download_fab(const mgic_field *f, double *fab, const int flo[3], const int fhi[3], const int rlo[3],
                                       const int rhi[3]) {
  MGIC_REQUIRE(f && fab && flo && fhi && rlo && rhi, "NULL argument");
  return fab_copy(f, fab, flo, fhi, rlo, rhi, false);
}
extern "C" int mgic_field_sync(const mgic_field *f) {
  MGIC_REQUIRE(f, "field is NULL");
  MGIC_CUDA(cudaStreamSynchronize(f->ctx->stream));
  if (f->evPending) MGIC_CUDA(cudaEventSynchronize(f->evPending));
  return MGIC_OK;
}
extern "C" int mgic_field_devptr(const mgic_field *f, void **ptr, long long *sy, long long *sz) {
  MGIC_REQUIRE(f && ptr, "NULL argument");
  *ptr = f->p;
  if (sy) *sy = f->sy;
  if (sz) *sz = f->sz;
  return MGIC_OK;
}

// ------------------------------------------------------------------------------------------------ op methods
int mgk::mgic_halo(mgic_op *o, mgic_field *f, int planes) {
  if (o->ctx->nranks > 1 && !o->isGlobal) {
    MGIC_REQUIRE(o->ctx->halo_exchange, "multi-rank context without a halo hook (mgic_comm)");
    ProfScope ps(o->ctx, false, PROF_HALO);
    return o->ctx->halo_exchange(o->ctx, f, planes);
  }
  return MGIC_OK;
}
int mgk::mgic_halo_shape(mgic_ctx *c, mgic_field *f, int planes) {
  if (c->nranks > 1 && !f->noHalo) {
    MGIC_REQUIRE(c->halo_exchange, "multi-rank context without a halo hook (mgic_comm)");
    ProfScope ps(c, false, PROF_HALO);
    return c->halo_exchange(c, f, planes);
  }
  return MGIC_OK;
}
static int halo(mgic_op *o, mgic_field *f, int planes) { return mgk::mgic_halo(o, f, planes); }
static inline const double *bptr(const mgic_op *o) { return o->b ? o->b->p : nullptr; }

// one colour pass of levelGSRB (VariableCoeffPoissonOperator.cpp:290-331): exchange, BC (folded in), kernel
extern "C" int mgic_op_gsrb_color(mgic_op *o, mgic_field *e, const mgic_field *r, int whichPass) {
  MGIC_REQUIRE(o && e && r, "NULL argument");
  REQ_SHAPE(o, e); REQ_SHAPE(o, r);
  MGIC_TRY(mgic_op_reset_lambda(o));  // :283
  MGIC_TRY(halo(o, e, 1));            // :301
  ProfScope ps(o->ctx, o->profTag);
  return mgk::gsrb_color(o->ctx, o->geom(), o->bck(true), e->p, r->p, o->a->p, bptr(o), o->lambda->p, o->alpha, o->beta, o->dx,
                         whichPass);
}

// relax -> levelGSRB x iterations ([Chombo] AMRPoissonOp::relax, s_relaxMode == 1)
extern "C" int mgic_op_relax(mgic_op *o, mgic_field *e, const mgic_field *r, int iterations) {
  MGIC_REQUIRE(o && e && r, "NULL argument");
  REQ_SHAPE(o, e); REQ_SHAPE(o, r);
  MGIC_TRY(mgic_op_reset_lambda(o));
  if (o->smoother == 1 && mgk::gsrb_fused_applicable(o)) return mgk::gsrb_fused(o, e, r, iterations);
  for (int it = 0; it < iterations; it++)
    for (int pass = 0; pass <= 1; pass++) MGIC_TRY(mgic_op_gsrb_color(o, e, r, pass));
  return MGIC_OK;
}

extern "C" int mgic_op_residual(mgic_op *o, mgic_field *lhs, mgic_field *phi, const mgic_field *rhs, int homogeneous) {
  MGIC_REQUIRE(o && lhs && phi && rhs && o->a, "NULL argument");
  MGIC_REQUIRE(!o->isPatch, "residual / applyOp on an AMR patch need the coarse-fine ghost values of QuadCFInterp: use mgic_op_amr_residual_nf / mgic_op_amr_operator_nf with the coarser level's field");
  REQ_SHAPE(o, lhs); REQ_SHAPE(o, phi); REQ_SHAPE(o, rhs);
  MGIC_TRY(halo(o, phi, 1));  // :48
  return mgk::residual(o->ctx, o->geom(), o->bck(homogeneous != 0), lhs->p, phi->p, rhs->p, o->a->p, bptr(o), o->alpha, o->beta,
                       o->dx);
}
extern "C" int mgic_op_apply(mgic_op *o, mgic_field *lhs, mgic_field *phi, int homogeneous) {
  MGIC_REQUIRE(o && lhs && phi && o->a, "NULL argument");
  MGIC_REQUIRE(!o->isPatch, "residual / applyOp on an AMR patch need the coarse-fine ghost values of QuadCFInterp: use mgic_op_amr_residual_nf / mgic_op_amr_operator_nf with the coarser level's field");
  REQ_SHAPE(o, lhs); REQ_SHAPE(o, phi);
  MGIC_TRY(halo(o, phi, 1));  // :131
  return mgk::apply_op(o->ctx, o->geom(), o->bck(homogeneous != 0), lhs->p, phi->p, o->a->p, bptr(o), o->alpha, o->beta, o->dx);
}
// [Chombo] QuadCFInterp::coarseFineInterp(phi, phiCoarse) on an AMR patch: the ghost values of every coarse-fine face go
// into the operator's face arrays (the fields carry no x/y ghost cells); returns the kernel-side BC table that reads them
static int quad_cf_interp(mgic_op *o, const mgic_field *phi, const mgic_field *pc, const int clo[3], bool homogeneous, BCk *out) {
  MGIC_REQUIRE(o->isPatch, "coarse-fine interpolation needs an AMR patch operator (mgic_op_create_patch)");
  MGIC_REQUIRE(pc && clo, "NULL argument");
  const int n[3] = {o->n[0], o->n[1], o->n[2]};
  const int cn[3] = {pc->nx, pc->ny, pc->nz};
  for (int d = 0; d < 3; d++) {
    // coarse cells the stencils can touch: the coarsened patch grown by two, inside the coarse domain
    const int need_lo = std::max(0, (o->plo[d] >> 1) - 2), need_hi = std::min(o->ndom[d] / 2 - 1, ((o->plo[d] + n[d] - 1) >> 1) + 2);
    MGIC_REQUIRE(clo[d] <= need_lo && clo[d] + cn[d] - 1 >= need_hi,
                 "the coarse field does not cover the coarsened patch grown by two cells (proper nesting)");
  }
  BCk k = o->bck(homogeneous);
  if (o->mask) {
    const size_t cells = (size_t)n[0] * n[1] * n[2];
    for (int f = 0; f < 6; f++) {
      if (!o->cfFace[f]) MGIC_CUDA(mgic_dev_malloc(&o->cfFace[f], cells * sizeof(double)));
      k.face[f] = o->cfFace[f];
    }
    MGIC_TRY(mgk::quad_cf_masked(o->ctx, o->geom(), o->mask, o->plo, o->ndom, o->dx, phi->p, pc->p, pc->sy, pc->sz, clo, o->cfFace));
    *out = k;
    return MGIC_OK;
  }
  double *faces[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  for (int f = 0; f < 6; f++) {
    k.face[f] = nullptr;
    const int dir = f / 2, side = (f % 2) ? +1 : -1;
    if (!(side < 0 ? o->cfLo[dir] : o->cfHi[dir])) continue;
    const int ta = dir == 0 ? 1 : 0, tb = dir == 2 ? 1 : 2;
    if (!o->cfFace[f]) MGIC_CUDA(mgic_dev_malloc(&o->cfFace[f], (size_t)n[ta] * n[tb] * sizeof(double)));
    k.type[f] = MGIC_FACE_GHOST;
    k.face[f] = o->cfFace[f];
    faces[f] = o->cfFace[f];
  }
  MGIC_TRY(mgk::quad_cf_faces(o->ctx, o->geom(), o->plo, o->ndom, o->dx, phi->p, pc->p, pc->sy, pc->sz, clo, faces));   // all faces, one launch
  *out = k;
  return MGIC_OK;
}
// the same interpolation with CELL-indexed ghost arrays whatever the level's shape (face[f][cell] = the ghost beyond face f of
// `cell`): what set_update_psi0 on an AMR level needs to carry psi's coarse-fine ghosts along (source.cu)
static int quad_cf_cells(mgic_op *o, const mgic_field *phi, const mgic_field *pc, const int clo[3], bool homogeneous, BCk *out) {
  MGIC_REQUIRE(o->isPatch && pc && clo, "coarse-fine interpolation needs an AMR patch operator and the coarser field");
  const size_t cells = (size_t)o->n[0] * o->n[1] * o->n[2];
  BCk k = o->bck(homogeneous);
  for (int f = 0; f < 6; f++) {
    if (!o->cfCell[f]) MGIC_CUDA(mgic_dev_malloc(&o->cfCell[f], cells * sizeof(double)));
    k.face[f] = o->cfCell[f];
  }
  MGIC_TRY(mgk::quad_cf_masked(o->ctx, o->geom(), o->mask, o->plo, o->ndom, o->dx, phi->p, pc->p, pc->sy, pc->sz, clo, o->cfCell));
  *out = k;
  return MGIC_OK;
}
// the ghost values QuadCFInterp left on one coarse-fine face (0 x-lo ... 5 z-hi) at the last AMROperatorNF / AMRResidualNF:
// x faces [j + ny*k], y faces [i + nx*k], z faces [i + nx*j]
extern "C" int mgic_op_cf_ghosts(mgic_op *o, int face, double *host) {
  MGIC_REQUIRE(o && host && face >= 0 && face < 6, "bad argument");
  MGIC_REQUIRE(o->isPatch && o->cfFace[face], "no coarse-fine ghost values on this face yet");
  MGIC_REQUIRE(!o->mask, "masked level: the ghost values are cell-indexed (not exposed)");
  const int dir = face / 2, ta = dir == 0 ? 1 : 0, tb = dir == 2 ? 1 : 2;
  MGIC_CUDA(cudaStreamSynchronize(o->ctx->stream));
  MGIC_CUDA(cudaMemcpy(host, o->cfFace[face], (size_t)o->n[ta] * o->n[tb] * sizeof(double), cudaMemcpyDeviceToHost));
  return MGIC_OK;
}
// [Chombo] AMRPoissonOp::AMROperatorNF (the level has a coarser but no finer level): coarseFineInterp, then applyOpI
extern "C" int mgic_op_amr_operator_nf(mgic_op *o, mgic_field *lhs, mgic_field *phi, const mgic_field *phi_coarse, const int coarse_lo[3],
                                       int homogeneous) {
  MGIC_REQUIRE(o && lhs && phi && o->a, "NULL argument");
  REQ_SHAPE(o, lhs); REQ_SHAPE(o, phi);
  BCk k;
  MGIC_TRY(quad_cf_interp(o, phi, phi_coarse, coarse_lo, homogeneous != 0, &k));
  return mgk::apply_op(o->ctx, o->geom(), k, lhs->p, phi->p, o->a->p, bptr(o), o->alpha, o->beta, o->dx);
}
// [Chombo] AMRPoissonOp::AMRResidualNF: rhs - AMROperatorNF(phi)
extern "C" int mgic_op_amr_residual_nf(mgic_op *o, mgic_field *lhs, mgic_field *phi, const mgic_field *phi_coarse, const int coarse_lo[3],
                                       const mgic_field *rhs, int homogeneous) {
  MGIC_REQUIRE(o && lhs && phi && rhs && o->a, "NULL argument");
  REQ_SHAPE(o, lhs); REQ_SHAPE(o, phi); REQ_SHAPE(o, rhs);
  BCk k;
  MGIC_TRY(quad_cf_interp(o, phi, phi_coarse, coarse_lo, homogeneous != 0, &k));
  return mgk::residual(o->ctx, o->geom(), k, lhs->p, phi->p, rhs->p, o->a->p, bptr(o), o->alpha, o->beta, o->dx);
}

// applyOpNoBoundary (:123-149): the stencil on whatever the ghost cells hold.  With ghosts folded into the
// kernels the only ghost state a caller can have produced through this API is the last BC fill, which for
// every call site in the reference's solver stack is the homogeneous one.
extern "C" int mgic_op_apply_no_boundary(mgic_op *o, mgic_field *lhs, mgic_field *phi) { return mgic_op_apply(o, lhs, phi, 1); }

extern "C" int mgic_op_level_jacobi(mgic_op *o, mgic_field *e, const mgic_field *r) {
  MGIC_REQUIRE(o && e && r, "NULL argument");
  REQ_SHAPE(o, e); REQ_SHAPE(o, r);
  MGIC_TRY(mgic_op_reset_lambda(o));                                       // :367
  if (!o->scratch) MGIC_TRY(mgic_field_create(o, &o->scratch));
  MGIC_TRY(mgic_op_residual(o, o->scratch, e, r, 1));                      // :373
  MGIC_TRY(mgk::jacobi_update(o->ctx, o->geom(), e->p, o->scratch->p, o->lambda->p, 0.5));  // :376-381
  return halo(o, e, 1);                                                    // :384
}

// haloPlanes = 2: the caller will sweep phi next without changing it (the fused prolong + relax of the V-cycle), so the
// exchange the sweep would need is done here and the one-plane exchange of the restriction is saved
static int restrict_residual(mgic_op *o, mgic_field *resC, mgic_field *phi, const mgic_field *rhs, int haloPlanes) {
  MGIC_REQUIRE(o && resC && phi && rhs && o->a, "NULL argument");
  REQ_SHAPE(o, phi); REQ_SHAPE(o, rhs);
  MGIC_REQUIRE(o->n[0] % 2 == 0 && o->n[1] % 2 == 0 && o->nzl % 2 == 0 && o->k0 % 2 == 0, "level is not coarsenable by 2");
  MGIC_REQUIRE(resC->nx == o->n[0] / 2 && resC->ny == o->n[1] / 2 && resC->nz == o->nzl / 2, "coarse residual has the wrong shape");
  if (haloPlanes < 0) {   // the last sweep pushed phi's boundary planes itself: no exchange, wait for the neighbours' planes
    if (o->ctx->nranks > 1 && !o->isGlobal) {
      MGIC_REQUIRE(o->ctx->sweep_wait, "folded halo without the communication hooks");
      ProfScope ph(o->ctx, false, PROF_HALO);
      MGIC_TRY(o->ctx->sweep_wait(o->ctx, phi));
    }
  } else {
    MGIC_TRY(halo(o, phi, haloPlanes));  // :163
  }
  ProfScope ps(o->ctx, false, PROF_RESTRICT);
  if (mgk::restrict_tma_applicable(o)) return mgk::restrict_tma(o, o->bck(true), resC, phi, rhs);
  return mgk::restrict_res(o->ctx, o->geom(), o->bck(true), resC->p, resC->sy, resC->sz, phi->p, rhs->p, o->a->p, bptr(o),
                           o->alpha, o->beta, o->dx);
}
extern "C" int mgic_op_restrict_residual(mgic_op *o, mgic_field *resC, mgic_field *phi, const mgic_field *rhs) {
  return restrict_residual(o, resC, phi, rhs, 1);
}
extern "C" int mgic_op_prolong_increment(mgic_op *o, mgic_field *phi, const mgic_field *coarse) {
  MGIC_REQUIRE(o && phi && coarse, "NULL argument");
  REQ_SHAPE(o, phi);
  MGIC_REQUIRE(coarse->nx == o->n[0] / 2 && coarse->ny == o->n[1] / 2 && coarse->nz == o->nzl / 2 && o->n[0] % 2 == 0,
               "coarse correction has the wrong shape");
  ProfScope ps(o->ctx, false, PROF_PROLONG);
  return mgk::prolong(o->ctx, o->geom(), phi->p, coarse->p, coarse->sy, coarse->sz);
}
// preCond (:72-104): phi = rhs * lambda, then relax(phi, rhs, 2)
extern "C" int mgic_op_precond(mgic_op *o, mgic_field *phi, const mgic_field *rhs) {
  MGIC_REQUIRE(o && phi && rhs, "NULL argument");
  REQ_SHAPE(o, phi); REQ_SHAPE(o, rhs);
  MGIC_TRY(mgic_op_reset_lambda(o));
  MGIC_TRY(mgk::mult(o->ctx, o->geom(), phi->p, rhs->p, o->lambda->p));
  return mgic_op_relax(o, phi, rhs, 2);
}

// BLAS-1 [Chombo AMRPoissonOp]
static int local_reduce(mgic_op *o, const mgic_field *x, const mgic_field *y, int kind, double *out) {
  MGIC_TRY(mgk::reduce(o->ctx, o->geom(), x->p, y ? y->p : nullptr, kind, 0));
  return fetch_scalars(o->ctx, 0, 1, kind == 0 ? 1 : 0, out, !o->isGlobal);   // replicated level: every rank already has the whole sum
}
extern "C" int mgic_op_norm(mgic_op *o, const mgic_field *x, int ord, double *out) {
  MGIC_REQUIRE(o && x && out, "NULL argument");
  REQ_SHAPE(o, x);
  MGIC_REQUIRE(ord >= 0 && ord <= 2, "norm order must be 0, 1 or 2");
  double v;
  MGIC_TRY(local_reduce(o, x, nullptr, ord, &v));
  *out = (ord == 2) ? sqrt(v) : v;
  return MGIC_OK;
}
extern "C" int mgic_op_dot(mgic_op *o, const mgic_field *x, const mgic_field *y, double *out) {
  MGIC_REQUIRE(o && x && y && out, "NULL argument");
  REQ_SHAPE(o, x); REQ_SHAPE(o, y);
  return local_reduce(o, x, y, 3, out);
}
extern "C" int mgic_op_incr(mgic_op *o, mgic_field *y, const mgic_field *x, double s) {
  MGIC_REQUIRE(o && x && y, "NULL argument");
  REQ_SHAPE(o, x); REQ_SHAPE(o, y);
  return mgk::incr(o->ctx, o->geom(), y->p, x->p, s);
}
extern "C" int mgic_op_axby(mgic_op *o, mgic_field *y, const mgic_field *x1, const mgic_field *x2, double a, double b) {
  MGIC_REQUIRE(o && x1 && x2 && y, "NULL argument");
  REQ_SHAPE(o, x1); REQ_SHAPE(o, x2); REQ_SHAPE(o, y);
  return mgk::axby(o->ctx, o->geom(), y->p, x1->p, x2->p, a, b);
}
extern "C" int mgic_op_scale(mgic_op *o, mgic_field *y, double s) {
  MGIC_REQUIRE(o && y, "NULL argument");
  REQ_SHAPE(o, y);
  return mgk::scale(o->ctx, o->geom(), y->p, s);
}
extern "C" int mgic_op_assign(mgic_op *o, mgic_field *y, const mgic_field *x) {
  MGIC_REQUIRE(o && x && y, "NULL argument");
  REQ_SHAPE(o, x); REQ_SHAPE(o, y);
  return mgk::assign(o->ctx, o->geom(), y->p, x->p);
}
extern "C" int mgic_op_set_to_zero(mgic_op *o, mgic_field *y) { return mgic_op_set_val(o, y, 0.0); }
extern "C" int mgic_op_set_val(mgic_op *o, mgic_field *y, double v) {
  MGIC_REQUIRE(o && y, "NULL argument");
  REQ_SHAPE(o, y);
  MGIC_TRY(mgk::set_val(o->ctx, o->geom(), y->p, v));
  if (o->mask && v != 0.0) MGIC_TRY(mgk::apply_mask(o->ctx, o->geom(), y->p, o->mask));
  return MGIC_OK;
}
extern "C" int mgic_op_set_smoother(mgic_op *o, int kind) {
  MGIC_REQUIRE(o && (kind == 0 || kind == 1), "smoother kind must be 0 or 1");
  o->smoother = kind;
  o->ctx->cfgEpoch++;
  return MGIC_OK;
}

// ------------------------------------------------------------------------------------------------ BiCGStab
// [Chombo 3.2] BiCGStabSolver<T>::solve on device-resident vectors.  `Lin` supplies
// residual / applyOp / preCond; vector ops go to the level operator.
struct BiCGParams {
  int imax = 80;
  double eps = 1.0e-6, reps = 1.0e-12, hang = 1.0e-8, small = 1.0e-30;
  int numRestarts = 5, normType = 2;
  bool homogeneous = false;
};

struct LinOp {
  mgic_op *op;
  virtual int preCond(mgic_field *cor, mgic_field *res) = 0;
  virtual ~LinOp() {}
};

struct BiCGWork {
  mgic_field *v[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int alloc(mgic_op *op) {
    for (auto &f : v) if (!f) MGIC_TRY(mgic_field_create(op, &f));
    return MGIC_OK;
  }
  void release() {
    for (auto &f : v) { mgic_field_destroy(f); f = nullptr; }
  }
};

// The solver is written over a vector space S: one level (LevelSpace: the mgic_op_* vector operations) or the composite
// level vector of an AMR hierarchy (AmrSpace below: [Chombo] MultilevelLinearOp).
struct LevelSpace {
  typedef mgic_field *Vec;
  typedef const mgic_field *CVec;
  LinOp &L;
  mgic_op *op;
  explicit LevelSpace(LinOp &l) : L(l), op(l.op) {}
  int residual(Vec r, Vec phi, CVec rhs, bool homogeneous) { return mgic_op_residual(op, r, phi, rhs, homogeneous); }
  int apply(Vec lhs, Vec phi, int homogeneous) { return mgic_op_apply(op, lhs, phi, homogeneous); }
  int preCond(Vec cor, Vec res) { return L.preCond(cor, res); }
  int assign(Vec y, CVec x) { return mgic_op_assign(op, y, x); }
  int set_to_zero(Vec y) { return mgic_op_set_to_zero(op, y); }
  int norm(CVec x, int ord, double *out) { return mgic_op_norm(op, x, ord, out); }
  int dot(CVec x, CVec y, double *out) { return mgic_op_dot(op, x, y, out); }
  int scale(Vec y, double s) { return mgic_op_scale(op, y, s); }
  int incr(Vec y, CVec x, double s) { return mgic_op_incr(op, y, x, s); }
};

template <class S>
static int bicgstab_t(S &sp, typename S::Vec phi, typename S::CVec rhs, typename S::Vec const w[8], const BiCGParams &P,
                      int *iterations, int *exitStatus, double *hist, int maxHist) {
  typename S::Vec r = w[0], rt = w[1], e = w[2], p = w[3], pt = w[4], st = w[5], t = w[6], v = w[7];
  int nh = 0;
  auto push = [&](double x) { if (hist && nh < maxHist) hist[nh] = x; nh++; };
  int recount = 0;
  MGIC_TRY(sp.residual(r, phi, rhs, P.homogeneous));
  MGIC_TRY(sp.assign(rt, r));
  MGIC_TRY(sp.set_to_zero(e));
  MGIC_TRY(sp.set_to_zero(pt));
  MGIC_TRY(sp.set_to_zero(st));
  int i = 0;
  double rho[4] = {0, 0, 0, 0}, norm[2];
  MGIC_TRY(sp.norm(r, P.normType, &norm[0]));
  const double initial_norm = norm[0], initial_rnorm = norm[0];
  norm[1] = norm[0];
  double alpha[2] = {0, 0}, beta[2] = {0, 0}, omega[2] = {0, 0};
  bool init = true;
  int restarts = 0, status = -1;
  push(norm[0]);
  while ((i < P.imax && norm[0] > P.eps * norm[1]) && (norm[1] > 0)) {
    i++;
    norm[1] = norm[0]; alpha[1] = alpha[0]; beta[1] = beta[0]; omega[1] = omega[0];
    rho[3] = rho[2]; rho[2] = rho[1];
    MGIC_TRY(sp.dot(rt, r, &rho[1]));
    if (rho[1] == 0.0) {  // we are finished, we will not converge anymore
      MGIC_TRY(sp.incr(phi, e, 1.0));
      status = 2;
      if (iterations) *iterations = i;
      if (exitStatus) *exitStatus = status;
      return MGIC_OK;
    }
    if (init) {
      MGIC_TRY(sp.assign(p, r));
      init = false;
    } else {
      beta[1] = (rho[1] / rho[2]) * (alpha[1] / omega[1]);
      MGIC_TRY(sp.scale(p, beta[1]));
      MGIC_TRY(sp.incr(p, v, -beta[1] * omega[1]));
      MGIC_TRY(sp.incr(p, r, 1.0));
    }
    MGIC_TRY(sp.preCond(pt, p));
    MGIC_TRY(sp.apply(v, pt, 1));
    double m;
    MGIC_TRY(sp.dot(rt, v, &m));
    alpha[0] = rho[1] / m;
    if (fabs(m) > P.small * fabs(rho[1])) {
      MGIC_TRY(sp.incr(r, v, -alpha[0]));
      MGIC_TRY(sp.norm(r, P.normType, &norm[0]));
      MGIC_TRY(sp.incr(e, pt, alpha[0]));
    } else {
      MGIC_TRY(sp.set_to_zero(r));
      norm[0] = 0.0;
    }
    if (norm[0] > P.eps * initial_norm && norm[0] > P.reps * initial_rnorm) {
      MGIC_TRY(sp.preCond(st, r));
      MGIC_TRY(sp.apply(t, st, 1));
      double tr, tt;
      MGIC_TRY(sp.dot(t, r, &tr));
      MGIC_TRY(sp.dot(t, t, &tt));
      omega[0] = tr / tt;
      MGIC_TRY(sp.incr(e, st, omega[0]));
      MGIC_TRY(sp.incr(r, t, -omega[0]));
      MGIC_TRY(sp.norm(r, P.normType, &norm[0]));
    }
    push(norm[0]);
    if (norm[0] <= P.eps * initial_norm || norm[0] <= P.reps * initial_rnorm) {
      status = 1;
      break;
    }
    if (omega[0] == 0.0 || norm[0] > (1 - P.hang) * norm[1]) {
      if (recount == 0) recount = 1;
      else {
        recount = 0;
        MGIC_TRY(sp.incr(phi, e, 1.0));
        if (restarts == P.numRestarts) {
          status = 3;
          if (iterations) *iterations = i;
          if (exitStatus) *exitStatus = status;
          return MGIC_OK;
        }
        MGIC_TRY(sp.residual(r, phi, rhs, P.homogeneous));
        MGIC_TRY(sp.norm(r, P.normType, &norm[0]));
        rho[1] = 0.0; rho[2] = 0.0; rho[3] = 0.0;
        alpha[0] = 0; beta[0] = 0; omega[0] = 0;
        MGIC_TRY(sp.assign(rt, r));
        MGIC_TRY(sp.set_to_zero(e));
        restarts++;
        init = true;
      }
    }
  }
  MGIC_TRY(sp.incr(phi, e, 1.0));
  if (iterations) *iterations = i;
  if (exitStatus) *exitStatus = status;
  return MGIC_OK;
}

static int bicgstab(LinOp &L, BiCGWork &W, mgic_field *phi, const mgic_field *rhs, const BiCGParams &P, int *iterations,
                    int *exitStatus, double *hist, int maxHist) {
  MGIC_TRY(W.alloc(L.op));
  LevelSpace sp(L);
  return bicgstab_t(sp, phi, rhs, W.v, P, iterations, exitStatus, hist, maxHist);
}

// ------------------------------------------------------------------------------------------------ factory / MG
struct mgic_mg {
  mgic_ctx *ctx = nullptr;
  mgic_params P;
  int nd = 0;
  std::vector<mgic_op *> ops;
  std::vector<mgic_field *> e, r;        // MultiGrid m_correction / m_residual; [0] unused (caller's vectors)
  std::vector<mgic_field *> aOwn, bOwn;  // CoarseAverage'd coefficients of depth > 0
  mgic_field *a0 = nullptr, *b0 = nullptr;
  bool bIsOne = false;
  int lastBottomIters = 0;
  BiCGWork bottomWork, outerWork;
  int *d_bottomOut = nullptr;   // device {iterations, status} of the last persistent bottom solve
  bool bottomOnDevice = false;
  // V-cycle graphs, keyed by the (correction, residual) arrays they were captured for
  // ... and by everything else a capture bakes in: the option epoch of the context and the ping-pong partner of the
  // finest level (fused sweeps swap the field's array with op->scratch, so after an odd number of caller-driven sweeps
  // the same field pointer pairs with a different scratch array)
  struct VGraph { const void *e, *r; bool zero; cudaGraphExec_t exec; long long launches; long long epoch; const void *scratch0; };
  std::vector<VGraph> graphs;
  bool graphBroken = false;
  // multi-rank agglomeration: from depth dA down every rank holds the WHOLE level (all-gather of the restricted
  // residual, redundant single-GPU cycle incl. the one-kernel bottom solve, own slab of the correction prolonged back):
  // small levels need no halo traffic and the bottom solve no host round trips.  Results are bit-identical on every
  // rank and identical to the single-GPU cycle.
  int dA = 1 << 30;
  std::vector<mgic_op *> locOps;          // slab geometry of the agglomerated depths (coefficient averaging, views)
  std::vector<mgic_field *> locA, locB;   // slab-shaped staging of the coarsened coefficients
  bool warm = false;   // multi-rank: one eager V-cycle has run (NCCL connections exist) before graph capture
};

static int mg_coarsen_coefs(mgic_mg *mg) {
  const int type = mg->P.coefficient_average_type >= 0 ? mg->P.coefficient_average_type : MGIC_AVG_ARITHMETIC;  // Factory.cpp:44-46,321
  if (type != MGIC_AVG_ARITHMETIC && type != MGIC_AVG_HARMONIC) {
    mgic_set_error("MGNewOp -- bad averagetype");  // Factory.cpp:222-224
    return MGIC_ERR_ARG;
  }
  for (int d = 1; d < mg->nd; d++) {
    const int coarsening = 1 << d;
    mgic_op *o = mg->ops[d];
    // directly from the AMR-level coefficients, not recursively (Factory.cpp:208-220)
    if (d >= mg->dA) {
      // agglomerated depth: average this rank's slab, then all-gather the slabs into the whole-level array
      mgic_op *lo = mg->locOps[d];
      const size_t cnt = (size_t)lo->n[0] * lo->n[1] * lo->nzl;
      MGIC_TRY(mgk::coarse_average(mg->ctx, lo->geom(), mg->locA[d]->p, mg->a0->p, mg->a0->sy, mg->a0->sz, coarsening, type));
      MGIC_TRY(mg->ctx->allgather(mg->ctx, mg->locA[d]->p, mg->aOwn[d]->p, cnt));
      if (!mg->bIsOne) {
        MGIC_TRY(mgk::coarse_average(mg->ctx, lo->geom(), mg->locB[d]->p, mg->b0->p, mg->b0->sy, mg->b0->sz, coarsening, type));
        MGIC_TRY(mg->ctx->allgather(mg->ctx, mg->locB[d]->p, mg->bOwn[d]->p, cnt));
      }
    } else {
      MGIC_TRY(mgk::coarse_average(mg->ctx, o->geom(), mg->aOwn[d]->p, mg->a0->p, mg->a0->sy, mg->a0->sz, coarsening, type));
      if (!mg->bIsOne)
        MGIC_TRY(mgk::coarse_average(mg->ctx, o->geom(), mg->bOwn[d]->p, mg->b0->p, mg->b0->sy, mg->b0->sz, coarsening, type));
    }
    MGIC_TRY(mgic_op_compute_lambda(o));  // :229
  }
  MGIC_TRY(mgic_op_compute_lambda(mg->ops[0]));
  return MGIC_OK;
}

extern "C" int mgic_mg_create(mgic_ctx *c, const mgic_params *P, mgic_field *a0, mgic_field *b0, mgic_mg **out) {
  return mgic_mg_create_ex(c, P, a0, b0, 0, out);
}
extern "C" int mgic_mg_create_ex(mgic_ctx *c, const mgic_params *P, mgic_field *a0, mgic_field *b0, int flags, mgic_mg **out) {
  MGIC_REQUIRE(c && P && a0 && out, "NULL argument");
  MGIC_REQUIRE(P->max_level == 0, "single AMR level only (max_level = 0)");
  MGIC_REQUIRE(P->max_grid_size >= 1, "max_grid_size must be positive");
  for (int d = 0; d < 3; d++) MGIC_REQUIRE(P->N[d] % P->max_grid_size == 0, "N must be a multiple of max_grid_size (domainSplit lattice)");
  MGIC_REQUIRE(a0->nx == P->N[0] && a0->ny == P->N[1] && a0->gnz == P->N[2], "aCoef does not match params.N");
  mgic_mg *mg = new mgic_mg;
  mg->ctx = c;
  mg->P = *P;
  mg->a0 = a0;
  mg->b0 = b0;
  const int s_maxCoarse = 2;  // [Chombo] AMRPoissonOp::s_maxCoarse
  const double dx0 = P->L / P->N[0];  // PoissonParameters.cpp:82
  int bclo[3], bchi[3];
  for (int d = 0; d < 3; d++) {
    bclo[d] = P->is_periodic ? MGIC_BC_PERIODIC : P->bc_lo[d];
    bchi[d] = P->is_periodic ? MGIC_BC_PERIODIC : P->bc_hi[d];
  }
  // bCoef == 1 everywhere (set_b_coef, SetLevelData.cpp:330-340): b*x == x exactly, and every average of ones
  // is exactly one, so the b stream can be dropped from all kernels without changing a bit of the result.
  if (!b0) mg->bIsOne = true;
  else if (b0->nx != a0->nx || b0->ny != a0->ny || b0->nz != a0->nz) { mgic_set_error("bCoef does not match aCoef"); return MGIC_ERR_ARG; }
  else {
    mgic_op probe;
    probe.ctx = c; probe.n[0] = b0->nx; probe.n[1] = b0->ny; probe.n[2] = b0->gnz; probe.k0 = b0->k0; probe.nzl = b0->nz;
    MGIC_TRY(mgk::is_constant(c, probe.geom(), b0->p, 1.0, 1));
    double cnt;
    MGIC_TRY(fetch_scalars(c, 1, 1, 0, &cnt));
    mg->bIsOne = (cnt == 0.0) && !(flags & MGIC_MG_KEEP_B);
  }
  for (int depth = 0;; depth++) {
    // [Chombo 3.2] MultiGrid::define: an operator is pushed, m_depth++, and the next one is asked for only while
    // (m_depth < a_maxDepth || a_maxDepth < 0): maxDepth = D >= 1 gives D operators, D = 0 gives the one it always has
    if (P->preCondSolverDepth >= 0 && depth >= (P->preCondSolverDepth > 1 ? P->preCondSolverDepth : 1)) break;
    const int coarsening = 1 << depth;
    // Factory.cpp:168-172: boxes (the max_grid_size lattice) must be coarsenable by coarsening * s_maxCoarse
    if (coarsening > 1 && (P->max_grid_size % (coarsening * s_maxCoarse)) != 0) break;
    int n[3] = {P->N[0] / coarsening, P->N[1] / coarsening, P->N[2] / coarsening};
    if (a0->k0 % coarsening != 0 || a0->nz % coarsening != 0) {
      mgic_set_error("z-slab [%d,%d) is not coarsenable by %d: choose slabs that are multiples of max_grid_size", a0->k0,
                     a0->k0 + a0->nz, coarsening);
      return MGIC_ERR_ARG;
    }
    const int k0d = a0->k0 / coarsening, nzd = a0->nz / coarsening;
    // agglomerate this depth (and all deeper ones) on every rank once the slab is small: its sweeps are launch /
    // exchange-latency bound, not bandwidth bound
    bool agg = mg->dA <= depth;
    if (!agg && c->nranks > 1 && c->allgather && depth >= 1 && nzd * c->nranks == n[2] && k0d == c->rank * nzd &&
        (long long)n[0] * n[1] * nzd <= c->aggloCells)
      agg = true;
    if (agg && mg->dA > depth) mg->dA = depth;
    mgic_op *o = nullptr;
    MGIC_TRY(mgic_op_create(c, n, agg ? 0 : k0d, agg ? n[2] : nzd, dx0 * coarsening, P->alpha, P->beta, bclo, bchi, P->bc_value, &o));
    o->isGlobal = agg;
    o->profTag = (depth == 0);
    mg->ops.push_back(o);
    mgic_op *lo = nullptr;
    mgic_field *la = nullptr, *lb = nullptr;
    if (agg) {
      MGIC_TRY(mgic_op_create(c, n, k0d, nzd, dx0 * coarsening, P->alpha, P->beta, bclo, bchi, P->bc_value, &lo));
      MGIC_TRY(mgic_field_create(lo, &la));
      if (!mg->bIsOne) MGIC_TRY(mgic_field_create(lo, &lb));
    }
    mg->locOps.push_back(lo); mg->locA.push_back(la); mg->locB.push_back(lb);
    mgic_field *ea = nullptr, *ra = nullptr, *aa = nullptr, *ba = nullptr;
    if (depth > 0) {
      MGIC_TRY(mgic_field_create(o, &ea));
      MGIC_TRY(mgic_field_create(o, &ra));
      MGIC_TRY(mgic_field_create(o, &aa));
      if (!mg->bIsOne) MGIC_TRY(mgic_field_create(o, &ba));
      MGIC_TRY(mgic_op_set_coefs(o, aa, mg->bIsOne ? nullptr : ba, P->alpha, P->beta));
    } else {
      MGIC_TRY(mgic_op_set_coefs(o, a0, mg->bIsOne ? nullptr : b0, P->alpha, P->beta));  // :194-197
    }
    mg->e.push_back(ea); mg->r.push_back(ra); mg->aOwn.push_back(aa); mg->bOwn.push_back(ba);
  }
  mg->nd = (int)mg->ops.size();
  MGIC_TRY(mg_coarsen_coefs(mg));
  *out = mg;
  return MGIC_OK;
}

extern "C" int mgic_mg_destroy(mgic_mg *mg) {
  if (!mg) return MGIC_OK;
  cudaStreamSynchronize(mg->ctx->stream);
  mg->bottomWork.release();
  mg->outerWork.release();
  for (size_t d = 0; d < mg->locOps.size(); d++) {
    mgic_field_destroy(mg->locA[d]); mgic_field_destroy(mg->locB[d]);
    mgic_op_destroy(mg->locOps[d]);
  }
  for (auto &g : mg->graphs) cudaGraphExecDestroy(g.exec);
  mgic_dev_free(mg->d_bottomOut);
  for (int d = 0; d < mg->nd; d++) {
    mgic_field_destroy(mg->e[d]); mgic_field_destroy(mg->r[d]);
    mgic_field_destroy(mg->aOwn[d]); mgic_field_destroy(mg->bOwn[d]);
    mgic_op_destroy(mg->ops[d]);
  }
  delete mg;
  return MGIC_OK;
}
extern "C" int mgic_mg_depths(const mgic_mg *mg) { return mg ? mg->nd : 0; }
extern "C" int mgic_mg_op(mgic_mg *mg, int depth, mgic_op **op) {
  MGIC_REQUIRE(mg && op, "NULL argument");
  *op = (depth >= 0 && depth < mg->nd) ? mg->ops[depth] : nullptr;  // MGnewOp returns NULL past the limit
  return MGIC_OK;
}
extern "C" int mgic_mg_scratch(mgic_mg *mg, int depth, mgic_field **e, mgic_field **r) {
  MGIC_REQUIRE(mg && depth >= 1 && depth < mg->nd, "depth out of range (scratch exists for depth >= 1)");
  if (e) *e = mg->e[depth];
  if (r) *r = mg->r[depth];
  return MGIC_OK;
}
extern "C" int mgic_mg_refresh_coefs(mgic_mg *mg) {
  MGIC_REQUIRE(mg, "mg is NULL");
  return mg_coarsen_coefs(mg);
}
extern "C" int mgic_mg_set_smoother(mgic_mg *mg, int kind) {
  MGIC_REQUIRE(mg, "mg is NULL");
  mg->ctx->cfgEpoch++;
  for (auto *o : mg->ops) MGIC_TRY(mgic_op_set_smoother(o, kind));
  return MGIC_OK;
}
extern "C" int mgic_mg_b_is_one(const mgic_mg *mg) { return mg && mg->bIsOne; }

struct BottomLin : LinOp {
  int preCond(mgic_field *cor, mgic_field *res) override { return mgic_op_precond(op, cor, res); }
};

extern "C" int mgic_mg_bottom_solve(mgic_mg *mg, mgic_field *e, const mgic_field *r, int *iterations) {
  MGIC_REQUIRE(mg && e && r, "NULL argument");
  mgic_op *op = mg->ops.back();
  REQ_SHAPE(op, e); REQ_SHAPE(op, r);
  if (mg->ctx->bottomKernel && (mg->ctx->nranks == 1 || op->isGlobal)) {
    // one persistent cooperative kernel (bottom.cu); iteration count stays on the device until asked for
    MGIC_TRY(mg->bottomWork.alloc(op));
    MGIC_TRY(mgic_op_reset_lambda(op));
    if (!mg->d_bottomOut) MGIC_CUDA(mgic_dev_malloc(&mg->d_bottomOut, 2 * sizeof(int)));
    {
      ProfScope ps(mg->ctx, false, PROF_BOTTOM);
      MGIC_TRY(mgk::bottom_bicgstab(op, e, r, mg->bottomWork.v, mg->ctx->d_part, (int)mg->ctx->partCap, mg->d_bottomOut));
    }
    mg->bottomOnDevice = true;
    if (iterations) *iterations = mgic_mg_last_bottom_iterations(mg);
    return MGIC_OK;
  }
  BottomLin L;
  L.op = op;
  mg->ctx->lastBottomKernel = 0;
  BiCGParams bp;
  bp.homogeneous = true;  // [Chombo] MultiGrid::define: m_bottomSolver->define(op, true)
  int it = 0, st = 0;
  MGIC_TRY(bicgstab(L, mg->bottomWork, e, r, bp, &it, &st, nullptr, 0));
  mg->lastBottomIters = it;
  mg->bottomOnDevice = false;
  if (iterations) *iterations = it;
  return MGIC_OK;
}
extern "C" int mgic_mg_last_bottom_iterations(mgic_mg *mg) {
  if (!mg) return 0;
  if (mg->bottomOnDevice && mg->d_bottomOut) {
    int h[2] = {0, 0};
    if (cudaStreamSynchronize(mg->ctx->stream) != cudaSuccess) return -1;
    if (cudaMemcpy(h, mg->d_bottomOut, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    mg->lastBottomIters = h[0];
  }
  return mg->lastBottomIters;
}

// [Chombo] MultiGrid::cycle, m_cycle = 1; pre = post = bottom = numMGsmooth (Main:111-113)
// non-owning view of this rank's slab inside a whole-level (agglomerated) array
static mgic_field slab_view(const mgic_field *whole, const mgic_op *slab) {
  mgic_field v;
  v.ctx = whole->ctx;
  v.nx = whole->nx; v.ny = whole->ny; v.nz = slab->nzl;
  v.sy = whole->sy; v.sz = whole->sz;
  v.base = nullptr;
  v.p = whole->p + (long long)slab->k0 * whole->sz;
  v.bytes = 0; v.k0 = slab->k0; v.gnz = whole->gnz;
  v.noHalo = true;
  return v;
}

// relax(e, r, S) where e is known to be zero: the first fused sweep reads nothing for it (and the zero fill is skipped)
static int relax_from_zero(mgic_op *op, mgic_field *e, const mgic_field *r, int S, bool *pushed = nullptr) {
  if (pushed) *pushed = false;
  if (S >= 1 && op->smoother == 1 && mgk::gsrb_fused_applicable(op)) {
    MGIC_TRY(mgic_op_reset_lambda(op));
    return mgk::gsrb_fused(op, e, r, S, mgk::FUSED_FROM_ZERO, nullptr, false, false, pushed);
  }
  MGIC_TRY(mgic_op_set_to_zero(op, e));
  return mgic_op_relax(op, e, r, S);
}
// relax(e, r, S) that reports whether the last sweep pushed e's ghost planes itself (multi-rank, halo folded into the sweep)
static int relax_report(mgic_op *op, mgic_field *e, const mgic_field *r, int S, bool *pushed) {
  *pushed = false;
  if (S >= 1 && op->smoother == 1 && mgk::gsrb_fused_applicable(op)) {
    MGIC_TRY(mgic_op_reset_lambda(op));
    return mgk::gsrb_fused(op, e, r, S, mgk::FUSED_PLAIN, nullptr, false, false, pushed);
  }
  return mgic_op_relax(op, e, r, S);
}
// prolongIncrement(e, eCoarse) followed by relax(e, r, S): the increment is folded into the first fused sweep
static bool prolong_relax_is_fused(const mgic_op *op, int S) {
  return S >= 1 && op->smoother == 1 && mgk::gsrb_fused_applicable(op) && op->ctx->fusePR;
}
static int prolong_relax(mgic_op *op, mgic_field *e, const mgic_field *ec, const mgic_field *r, int S, bool rhsHaloValid,
                         bool eHaloValid, bool coarseHaloValid, bool *pushed) {
  *pushed = false;
  if (prolong_relax_is_fused(op, S)) {
    MGIC_TRY(mgic_op_reset_lambda(op));
    return mgk::gsrb_fused(op, e, r, S, mgk::FUSED_PROLONG, ec, rhsHaloValid, eHaloValid, pushed, coarseHaloValid);
  }
  MGIC_TRY(mgic_op_prolong_increment(op, e, ec));
  return relax_report(op, e, r, S, pushed);
}

// [Chombo] MultiGrid::cycle, m_cycle = 1; pre = post = bottom = numMGsmooth (Main:111-113)
// *ePushed: on return the ghost planes of e are those the neighbours' last sweep of this depth stored (multi-rank, folded
// halo) -- the finer depth's prolonging sweep then needs no exchange of them
static int mg_cycle(mgic_mg *mg, int depth, mgic_field *e, const mgic_field *r, bool eIsZero, bool *ePushed = nullptr) {
  mgic_op *op = mg->ops[depth];
  const int S = mg->P.numMGsmooth;
  const bool z = eIsZero && mg->ctx->fusePR;
  bool dummy = false;
  if (!ePushed) ePushed = &dummy;
  *ePushed = false;
  if (depth == mg->nd - 1) {
    const long long cells = (long long)op->n[0] * op->n[1] * op->n[2];
    if (cells == 1) return z ? relax_from_zero(op, e, r, 1) : mgic_op_relax(op, e, r, 1);
    MGIC_TRY(z ? relax_from_zero(op, e, r, S) : mgic_op_relax(op, e, r, S));
    return mgic_mg_bottom_solve(mg, e, r, nullptr);
  }
  bool pushedPre = false;
  MGIC_TRY(z ? relax_from_zero(op, e, r, S, &pushedPre) : relax_report(op, e, r, S, &pushedPre));
  // the pre-smoothing fused relax exchanged the ghost planes of r; r is not written again on this depth
  const bool preFused = S >= 1 && op->smoother == 1 && mgk::gsrb_fused_applicable(op);
  // e is not touched between the restriction and the fused prolong + relax: one two-plane exchange serves both -- or none at
  // all when the last pre-smoothing sweep pushed its boundary planes itself (then the restriction only has to wait for them)
  const bool eOnce = prolong_relax_is_fused(op, S) && MGIC_GZ >= 2 && op->nzl >= 2;
  const int rp = pushedPre ? -1 : (eOnce ? 2 : 1);
  const bool eValidAfter = pushedPre || eOnce;
  bool coarsePushed = false;
  if (depth + 1 == mg->dA) {
    // slab-distributed -> agglomerated: restrict into this rank's slab of the whole-level residual, all-gather in
    // place, run the rest of the cycle on the whole level, prolong from the slab view of the whole-level correction
    // (its ghost planes are the neighbouring planes of the same array: no exchange)
    mgic_op *lo = mg->locOps[depth + 1];
    mgic_field rv = slab_view(mg->r[depth + 1], lo);
    MGIC_TRY(restrict_residual(op, &rv, e, r, rp));
    {
      ProfScope ps(mg->ctx, false, PROF_GATHER);
      MGIC_TRY(mg->ctx->allgather(mg->ctx, rv.p, mg->r[depth + 1]->p, (size_t)lo->n[0] * lo->n[1] * lo->nzl));
    }
    if (!mg->ctx->fusePR) MGIC_TRY(mgic_op_set_to_zero(mg->ops[depth + 1], mg->e[depth + 1]));
    MGIC_TRY(mg_cycle(mg, depth + 1, mg->e[depth + 1], mg->r[depth + 1], true));
    mgic_field ev = slab_view(mg->e[depth + 1], lo);
    return prolong_relax(op, e, &ev, r, S, preFused, eValidAfter, false, ePushed);
  }
  MGIC_TRY(restrict_residual(op, mg->r[depth + 1], e, r, rp));
  if (!mg->ctx->fusePR) MGIC_TRY(mgic_op_set_to_zero(mg->ops[depth + 1], mg->e[depth + 1]));   // setToZero(e[depth+1])
  MGIC_TRY(mg_cycle(mg, depth + 1, mg->e[depth + 1], mg->r[depth + 1], true, &coarsePushed));
  return prolong_relax(op, e, mg->e[depth + 1], r, S, preFused, eValidAfter, coarsePushed, ePushed);
}

// One V-cycle, replayed as a CUDA graph when possible: the cycle is a fixed launch sequence (the bottom solve is a
// single kernel), so it is captured once per (correction, residual) pair and afterwards costs one graph launch.
static int vcycle_run(mgic_mg *mg, mgic_field *e, const mgic_field *r, bool eIsZero = false) {
  mgic_ctx *c = mg->ctx;
  const bool graphable = c->useGraph && c->bottomKernel && (c->nranks == 1 || mg->ops.back()->isGlobal) && !c->profiling && !mg->graphBroken;
  if (!graphable) return mg_cycle(mg, 0, e, r, eIsZero);
  if (c->nranks > 1 && !mg->warm) {  // NCCL sets up its connections on first use: not inside a capture
    mg->warm = true;
    return mg_cycle(mg, 0, e, r, eIsZero);
  }
  for (size_t gi = 0; gi < mg->graphs.size();) {   // stale captures (options / smoother changed since): drop them
    if (mg->graphs[gi].epoch != c->cfgEpoch) { cudaGraphExecDestroy(mg->graphs[gi].exec); mg->graphs.erase(mg->graphs.begin() + gi); }
    else gi++;
  }
  const void *scr0 = mg->ops[0]->scratch ? (const void *)mg->ops[0]->scratch->base : nullptr;
  for (auto &g : mg->graphs)
    if (g.e == e->p && g.r == r->p && g.zero == eIsZero && g.scratch0 == scr0) {
      MGIC_CUDA(cudaGraphLaunch(g.exec, c->stream));
      c->launches += g.launches;
      mg->bottomOnDevice = true;
      return MGIC_OK;
    }
  // everything the cycle allocates lazily must exist before capture
  MGIC_TRY(mg->bottomWork.alloc(mg->ops.back()));
  if (!mg->d_bottomOut) MGIC_CUDA(mgic_dev_malloc(&mg->d_bottomOut, 2 * sizeof(int)));
  for (auto *o : mg->ops) {
    MGIC_TRY(mgic_op_reset_lambda(o));
    if (!o->scratch) MGIC_TRY(mgic_field_create(o, &o->scratch));
  }
  if (c->array_prepare) {  // arrays whose halo planes travel inside the graph: map them into the neighbours now (collective)
    MGIC_TRY(c->array_prepare(c, e));
    MGIC_TRY(c->array_prepare(c, const_cast<mgic_field *>(r)));
    for (int d = 0; d < mg->nd; d++) {
      if (mg->ops[d]->isGlobal) continue;
      if (mg->e[d]) MGIC_TRY(c->array_prepare(c, mg->e[d]));
      if (mg->r[d]) MGIC_TRY(c->array_prepare(c, mg->r[d]));
      MGIC_TRY(c->array_prepare(c, mg->ops[d]->scratch));
    }
  }
  MGIC_CUDA(cudaStreamSynchronize(c->stream));
  // ping-pong state (fused sweeps swap array pointers inside the field handles) must be the same after the cycle
  std::vector<std::pair<mgic_field *, double *>> saved;
  saved.emplace_back(e, e->base);
  for (int d = 0; d < mg->nd; d++) {
    if (mg->e[d]) saved.emplace_back(mg->e[d], mg->e[d]->base);
    saved.emplace_back(mg->ops[d]->scratch, mg->ops[d]->scratch->base);
  }
  const long long l0 = c->launches;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  bool ok = cudaStreamBeginCapture(c->stream, c->nranks > 1 ? cudaStreamCaptureModeRelaxed : cudaStreamCaptureModeThreadLocal) == cudaSuccess;
  int rc = MGIC_OK;
  if (ok) {
    rc = mg_cycle(mg, 0, e, r, eIsZero);
    ok = (cudaStreamEndCapture(c->stream, &graph) == cudaSuccess) && rc == MGIC_OK && graph;
  }
  const long long nl = c->launches - l0;
  c->launches = l0;
  for (auto &sv : saved) ok = ok && (sv.first->base == sv.second);
  if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
  if (graph) cudaGraphDestroy(graph);
  if (!ok) {
    // not capturable (e.g. an odd number of fused sweeps on some level): restore the handles, run eagerly from now on
    cudaGetLastError();
    for (auto &sv : saved) {
      sv.first->base = sv.second;
      sv.first->p = sv.second + (long long)MGIC_GZ * sv.first->sz;
    }
    mg->graphBroken = true;
    return mg_cycle(mg, 0, e, r, eIsZero);
  }
  mg->graphs.push_back({e->p, r->p, eIsZero, exec, nl, c->cfgEpoch, (const void *)mg->ops[0]->scratch->base});
  MGIC_CUDA(cudaGraphLaunch(exec, c->stream));
  c->launches += nl;
  mg->bottomOnDevice = true;
  return MGIC_OK;
}

extern "C" int mgic_mg_vcycle(mgic_mg *mg, mgic_field *e, const mgic_field *r) {
  MGIC_REQUIRE(mg && e && r, "NULL argument");
  REQ_SHAPE(mg->ops[0], e); REQ_SHAPE(mg->ops[0], r);
  mg->lastBottomIters = 0;
  return vcycle_run(mg, e, r);
}
// setToZero(e) followed by one V-cycle (what [Chombo] MultilevelLinearOp::preCond does first): knowing that e is zero
// lets the first sweep skip both the fill and the read of e
extern "C" int mgic_mg_vcycle_from_zero(mgic_mg *mg, mgic_field *e, const mgic_field *r) {
  MGIC_REQUIRE(mg && e && r, "NULL argument");
  REQ_SHAPE(mg->ops[0], e); REQ_SHAPE(mg->ops[0], r);
  mg->lastBottomIters = 0;
  if (!mg->ctx->fusePR) MGIC_TRY(mgic_op_set_to_zero(mg->ops[0], e));
  return vcycle_run(mg, e, r, true);
}

// ------------------------------------------------------------------------------------------------ AMR hierarchy
// [Chombo] AMRMultiGrid::AMRVCycle and MultilevelLinearOp over a hierarchy of levels (restated from
// the published algorithm): level 0 is the MG hierarchy (one array), every finer level a list of patch operators, each
// nested with refinement ratio 2 in ONE array of the level below and not touching its siblings (no fine-fine exchange).
// A "level vector" is an array of fields in node order: [0] the base level, then the patches in creation order.
// reflux is the reference's no-op (VariableCoeffPoissonOperator.cpp:264-271), so AMROperator == AMROperatorNF and the
// coarser residual outside the patches is the caller's; only cells under a patch are overwritten by the averaged fine
// residual.  Host orchestration over the operator methods:
//   down:  corr_q = 0; relax(corr_q, res_q, pre)                       (homogeneousCFInterp)          for every patch q of level l
//          corr_P = 0 on level l-1; res_P[under q] = average(res_q - L(corr_q))      (AMRRestrictS; QuadCFInterp from 0)
//   base:  corr_0 = MultiGrid::oneCycle(res_0) from zero               (the V-cycle graph)
//   up:    corr_q += prolong(corr_P)                                   (AMRProlongS, piecewise constant)
//          res_q -= L(corr_q), coarse-fine ghosts from corr_P           (AMRUpdateResidual)
//          d = 0; relax(d, res_q, post); corr_q += d
struct AmrNode {
  mgic_op *op = nullptr;           // base level operator / patch operator (not owned)
  int level = 0, parent = -1;      // parent: node of the level below whose array contains the coarsened patch
  int lo[3] = {0, 0, 0};           // origin of the node's array in its level's index space
  int off[3] = {0, 0, 0};          // coarsened patch origin inside the parent's array
  mgic_field *corr = nullptr, *res = nullptr, *tmp = nullptr;   // owned
  // multi-rank: the parent is the z-slab-distributed base level, this patch is replicated.  What the patch reads of the
  // parent (QuadCFInterp's stencils, the prolongation) is gathered into a replicated staging box = the coarsened patch grown
  // by two cells, clipped to the coarse domain (owned); stageLo = its origin in the parent level's index space
  mgic_field *stage = nullptr;
  int stageLo[3] = {0, 0, 0};
  // one byte per cell of this node's (local) array: 1 = a finer node covers the cell ([Chombo] zeroCovered's set) -- what the
  // composite norms and dot products skip; null: nothing finer (owned)
  unsigned char *covered = nullptr;
};
struct mgic_amr {
  mgic_ctx *ctx = nullptr;
  mgic_mg *base = nullptr;
  std::vector<AmrNode> nodes;
  std::vector<int> levelStart;     // nodes of level l: [levelStart[l], levelStart[l+1])
  std::vector<mgic_field *> work;  // 8 level vectors of the outer BiCGStab + 2 for composite residual / correction, lazily
  int nlevels() const { return (int)levelStart.size() - 1; }
};

static int amr_build_covered(mgic_amr *A);
static int amr_fail(mgic_amr *A, const char *fmt, int a, int b) {
  mgic_set_error(fmt, a, b);
  delete A;
  return MGIC_ERR_ARG;
}

extern "C" int mgic_amr_create_levels(mgic_mg *base, int nfiner, const int *npatches, mgic_op *const *patches, mgic_amr **out) {
  MGIC_REQUIRE(base && out && nfiner >= 0 && (nfiner == 0 || (patches && npatches)), "bad argument");
  mgic_amr *A = new mgic_amr;
  A->ctx = base->ctx; A->base = base;
  AmrNode root;
  root.op = base->ops[0];
  A->nodes.push_back(root);
  A->levelStart.push_back(0);
  A->levelStart.push_back(1);
  int next = 0;
  for (int l = 1; l <= nfiner; l++) {
    if (npatches[l - 1] < 1) return amr_fail(A, "level %d has %d patches", l, npatches[l - 1]);
    for (int q = 0; q < npatches[l - 1]; q++, next++) {
      mgic_op *o = patches[next];
      if (!o || !o->isPatch || !o->a) return amr_fail(A, "patch %d of level %d is not a patch operator with coefficients", q, l);
      AmrNode nd;
      nd.op = o; nd.level = l;
      for (int d = 0; d < 3; d++) nd.lo[d] = o->plo[d];
      // the parent: the one array of the level below that contains the coarsened patch
      for (int P = A->levelStart[l - 1]; P < A->levelStart[l] && nd.parent < 0; P++) {
        const AmrNode &pn = A->nodes[P];
        bool in = true;
        for (int d = 0; d < 3; d++) {
          const int pdom = pn.op->isPatch ? pn.op->ndom[d] : pn.op->n[d];
          const int oc = (o->plo[d] >> 1) - pn.lo[d];
          in = in && o->ndom[d] == 2 * pdom && oc >= 0 && oc + o->n[d] / 2 <= pn.op->n[d];
        }
        if (in) nd.parent = P;
      }
      if (nd.parent < 0) return amr_fail(A, "patch %d of level %d is not nested in one box of the level below with refinement ratio 2", q, l);
      for (int d = 0; d < 3; d++) nd.off[d] = (o->plo[d] >> 1) - A->nodes[nd.parent].lo[d];
      // proper nesting against the parent's cells: every coarse cell under the patch and its six neighbours (inside the
      // domain) must be cells of the parent level (QuadCFInterp reads them)
      {
        const AmrNode &pn = A->nodes[nd.parent];
        const mgic_op *po = pn.op;
        auto pvalid = [&](int ci, int cj, int ck) {   // coarse-level index -> is a cell of the parent node
          const int li = ci - pn.lo[0], lj = cj - pn.lo[1], lk = ck - pn.lo[2];
          if (li < 0 || lj < 0 || lk < 0 || li >= po->n[0] || lj >= po->n[1] || lk >= po->n[2]) return false;
          return po->hmask.empty() || po->hmask[(size_t)li + (size_t)po->n[0] * ((size_t)lj + (size_t)po->n[1] * lk)] != 0;
        };
        // QuadCFInterp reads, around every coarse cell next to the patch, its tangential and diagonal neighbours: the whole
        // 3 x 3 x 3 neighbourhood of the cells under the patch must be cells of the parent level (or outside the domain).
        // The coarse footprint of the patch, grown by one cell axis by axis, is checked cell by cell.
        const int fn[3] = {o->n[0] / 2 + 2, o->n[1] / 2 + 2, o->n[2] / 2 + 2};   // footprint array with a margin of one
        const size_t f1 = (size_t)fn[0], f2 = (size_t)fn[0] * fn[1];
        std::vector<unsigned char> fp(f2 * fn[2], 0), tmpv;
        for (int k = 0; k < o->n[2]; k += 2)
          for (int j = 0; j < o->n[1]; j += 2)
            for (int i = 0; i < o->n[0]; i += 2)
              if (o->hmask.empty() || o->hmask[(size_t)i + (size_t)o->n[0] * ((size_t)j + (size_t)o->n[1] * k)])
                fp[(size_t)(i / 2 + 1) + f1 * (j / 2 + 1) + f2 * (k / 2 + 1)] = 1;
        const size_t fst[3] = {1, f1, f2};
        for (int ax = 0; ax < 3; ax++) {
          tmpv = fp;
          for (size_t q = fst[ax]; q + fst[ax] < fp.size(); q++)
            if (tmpv[q - fst[ax]] | tmpv[q + fst[ax]]) fp[q] = 1;   // (rows do not wrap: the margin cells of a row are never set before their own axis is grown)
        }
        bool nested = true;
        const int c0[3] = {(o->plo[0] >> 1) - 1, (o->plo[1] >> 1) - 1, (o->plo[2] >> 1) - 1};   // coarse index of footprint cell (0, 0, 0)
        for (int k = 0; k < fn[2] && nested; k++)
          for (int j = 0; j < fn[1] && nested; j++)
            for (int i = 0; i < fn[0] && nested; i++) {
              if (!fp[(size_t)i + f1 * j + f2 * k]) continue;
              const int q3[3] = {c0[0] + i, c0[1] + j, c0[2] + k};
              bool outside = false;
              for (int d = 0; d < 3; d++) outside = outside || q3[d] < 0 || q3[d] >= o->ndom[d] / 2;
              if (!outside) nested = pvalid(q3[0], q3[1], q3[2]);
            }
        if (!nested) return amr_fail(A, "patch %d of level %d is not properly nested in the cells of the level below (one coarse cell all around, diagonals included)", q, l);
      }
      // the nodes of one level must not touch: touching boxes belong into ONE node (mgic_op_create_patch_boxes: a union of
      // boxes in one masked array, where the fine-fine exchange is a neighbour read)
      for (int s2 = A->levelStart[l]; s2 < (int)A->nodes.size(); s2++) {
        const mgic_op *b = A->nodes[s2].op;
        bool apart = false;
        for (int d = 0; d < 3; d++) apart = apart || o->plo[d] > b->plo[d] + b->n[d] || b->plo[d] > o->plo[d] + o->n[d];
        if (apart) continue;
        bool touch = o->hmask.empty() && b->hmask.empty();
        if (!touch) {   // bounding boxes touch or overlap: decide on the cells (face, edge or corner contact all count)
          auto valid = [](const mgic_op *w, int gi, int gj, int gk) {
            const int li = gi - w->plo[0], lj = gj - w->plo[1], lk = gk - w->plo[2];
            if (li < 0 || lj < 0 || lk < 0 || li >= w->n[0] || lj >= w->n[1] || lk >= w->n[2]) return false;
            return w->hmask.empty() || w->hmask[(size_t)li + (size_t)w->n[0] * ((size_t)lj + (size_t)w->n[1] * lk)] != 0;
          };
          for (int k = 0; k < o->n[2] && !touch; k++)
            for (int j = 0; j < o->n[1] && !touch; j++)
              for (int i = 0; i < o->n[0] && !touch; i++) {
                if (!valid(o, o->plo[0] + i, o->plo[1] + j, o->plo[2] + k)) continue;
                for (int dk = -1; dk <= 1 && !touch; dk++)
                  for (int dj = -1; dj <= 1 && !touch; dj++)
                    for (int di = -1; di <= 1 && !touch; di++)
                      touch = valid(b, o->plo[0] + i + di, o->plo[1] + j + dj, o->plo[2] + k + dk);
              }
        }
        if (touch) return amr_fail(A, "patches %d and %d of one level touch or overlap (give them as ONE node: mgic_op_create_patch_boxes)", s2 - A->levelStart[l], q);
      }
      A->nodes.push_back(nd);
    }
    A->levelStart.push_back((int)A->nodes.size());
  }
  for (AmrNode &nd : A->nodes) {
    if (mgic_field_create(nd.op, &nd.corr) != MGIC_OK || mgic_field_create(nd.op, &nd.res) != MGIC_OK ||
        mgic_field_create(nd.op, &nd.tmp) != MGIC_OK) {
      mgic_amr_destroy(A);
      return MGIC_ERR_CUDA;
    }
    if (A->ctx->nranks > 1 && nd.parent == 0) {   // staging box for what this patch reads of the distributed base level
      int ns[3];
      for (int d = 0; d < 3; d++) {
        const int clo = std::max(0, (nd.op->plo[d] >> 1) - 2), chi = std::min(nd.op->ndom[d] / 2 - 1, ((nd.op->plo[d] + nd.op->n[d] - 1) >> 1) + 2);
        nd.stageLo[d] = clo; ns[d] = chi - clo + 1;
      }
      if (field_alloc(A->ctx, ns[0], ns[1], ns[2], 0, ns[2], &nd.stage, false) != MGIC_OK) { mgic_amr_destroy(A); return MGIC_ERR_CUDA; }
    }
  }
  if (amr_build_covered(A) != MGIC_OK) { mgic_amr_destroy(A); return MGIC_ERR_CUDA; }
  *out = A;
  return MGIC_OK;
}
// a chain: one patch per finer level
extern "C" int mgic_amr_create(mgic_mg *base, int nfiner, mgic_op *const *patches, mgic_amr **out) {
  MGIC_REQUIRE(nfiner >= 0, "bad argument");
  std::vector<int> ones((size_t)std::max(nfiner, 1), 1);
  return mgic_amr_create_levels(base, nfiner, ones.data(), patches, out);
}
extern "C" int mgic_amr_destroy(mgic_amr *A) {
  if (!A) return MGIC_OK;
  for (AmrNode &nd : A->nodes) {
    mgic_field_destroy(nd.corr); mgic_field_destroy(nd.res); mgic_field_destroy(nd.tmp); mgic_field_destroy(nd.stage);
    if (nd.covered) mgic_dev_free(nd.covered);
  }
  for (auto *f : A->work) mgic_field_destroy(f);
  delete A;
  return MGIC_OK;
}
extern "C" int mgic_amr_levels(const mgic_amr *A) { return A ? A->nlevels() : 0; }
extern "C" int mgic_amr_nodes(const mgic_amr *A) { return A ? (int)A->nodes.size() : 0; }
extern "C" int mgic_amr_node_info(const mgic_amr *A, int node, int *level, int *parent) {
  MGIC_REQUIRE(A && node >= 0 && node < (int)A->nodes.size(), "bad argument");
  if (level) *level = A->nodes[node].level;
  if (parent) *parent = A->nodes[node].parent;
  return MGIC_OK;
}

// the cells of the parent's array `below` that lie under node q's patch, and that sub-box as a geometry
// ---- what a patch READS of its parent level (QuadCFInterp's stencils, the prolongation): the parent's array, or -- multi-rank,
// parent = the z-slab-distributed base level -- the patch's replicated staging box, gathered here: every rank copies the part
// of the box that lies in its slab into the zeroed box, one all-reduce (sum; exactly one rank contributes to each cell, so
// the values arrive bit for bit) replicates it.  zero = the parent field is known to be identically zero.
struct CoarseSrc { const mgic_field *f; int lo[3]; };
static int amr_coarse_source(mgic_amr *A, const AmrNode &n, mgic_field *parentField, CoarseSrc *out, bool zero = false) {
  const AmrNode &pn = A->nodes[n.parent];
  if (!n.stage) {
    out->f = parentField;
    for (int d = 0; d < 3; d++) out->lo[d] = pn.lo[d];
    return MGIC_OK;
  }
  mgic_ctx *c = A->ctx;
  mgic_field *st = n.stage;
  MGIC_CUDA(cudaMemsetAsync(st->base, 0, st->bytes, c->stream));
  if (!zero) {
    const int za = std::max(n.stageLo[2], parentField->k0), zb = std::min(n.stageLo[2] + st->nz, parentField->k0 + parentField->nz);
    if (zb > za)
      MGIC_TRY(mgk::copy_box(c, st->nx, st->ny, zb - za,
                             parentField->p + n.stageLo[0] + (long long)n.stageLo[1] * parentField->sy + (long long)(za - parentField->k0) * parentField->sz,
                             parentField->sy, parentField->sz, st->p + (long long)(za - n.stageLo[2]) * st->sz, st->sy, st->sz));
    MGIC_REQUIRE(c->allreduce, "multi-rank hierarchy without the communication hooks (mgic_comm_init)");
    ProfScope ps(c, false, PROF_GATHER);
    MGIC_TRY(c->allreduce(c, st->p, (int)((long long)st->sz * st->nz), 0));
  }
  out->f = st;
  for (int d = 0; d < 3; d++) out->lo[d] = n.stageLo[d];
  return MGIC_OK;
}
// the cells under the patch inside a coarse source (its array starts at src.lo in the parent level's index space)
static const double *under_in(const CoarseSrc &src, const AmrNode &n) {
  const int o[3] = {(n.op->plo[0] >> 1) - src.lo[0], (n.op->plo[1] >> 1) - src.lo[1], (n.op->plo[2] >> 1) - src.lo[2]};
  return src.f->p + o[0] + (long long)o[1] * src.f->sy + (long long)o[2] * src.f->sz;
}
// ---- what a patch WRITES into its parent level (the averaged residual, zeroCovered, averageDown): the cells of `below` under
// the patch -- on a multi-rank context only the planes of this rank's slab (the patch is replicated, so every rank has the
// fine data for its own part).  fineOff = first fine plane that maps into the local part.
struct UnderLocal { Geom g; double *ptr; long long fineOff; bool empty; };
static UnderLocal under_local(const AmrNode &q, mgic_field *below) {
  UnderLocal u;
  u.g = q.op->geom();
  u.g.nx /= 2; u.g.ny /= 2; u.g.sy = below->sy; u.g.sz = below->sz;
  const int zA = q.off[2], zB = q.off[2] + q.op->n[2] / 2;                 // in the parent array's (global) plane index
  const int la = std::max(zA, below->k0), lb = std::min(zB, below->k0 + below->nz);
  u.empty = lb <= la;
  u.g.nz = u.empty ? 0 : lb - la;
  u.ptr = below->p + q.off[0] + (long long)q.off[1] * below->sy + (long long)((u.empty ? zA : la) - below->k0) * below->sz;
  u.fineOff = u.empty ? 0 : 2LL * (la - zA);
  return u;
}
static inline const unsigned char *mask_at(const mgic_op *o, long long plane) {
  return o->mask ? o->mask + plane * (long long)o->n[0] * o->n[1] : nullptr;
}
// the 8-cell average of a patch field into the parent's cells under it ([Chombo] CoarseAverage, AMRRestrictS' second half)
static int average_under(mgic_amr *A, const AmrNode &n, mgic_field *below, const mgic_field *fine) {
  const UnderLocal u = under_local(n, below);
  if (u.empty) return MGIC_OK;
  return mgk::coarse_average(A->ctx, u.g, u.ptr, fine->p + u.fineOff * fine->sz, fine->sy, fine->sz, 2, 0, mask_at(n.op, u.fineOff));
}
static int zero_under(mgic_amr *A, const AmrNode &n, mgic_field *below) {
  const UnderLocal u = under_local(n, below);
  if (u.empty) return MGIC_OK;
  return mgk::box_set_val(A->ctx, u.g, u.ptr, 0.0, mask_at(n.op, u.fineOff), n.op->n[0], (long long)n.op->n[0] * n.op->n[1]);
}

// the `covered` marks of every node that has finer nodes above it (the cells zero_under would zero)
static int amr_build_covered(mgic_amr *A) {
  for (size_t q = 1; q < A->nodes.size(); q++) {
    const AmrNode &n = A->nodes[q];
    AmrNode &pn = A->nodes[n.parent];
    const mgic_field *ref = pn.corr;   // any field of the parent: the geometry of its (local) array
    if (!pn.covered) {
      const size_t cells = (size_t)ref->sz * ref->nz;
      MGIC_CUDA(mgic_dev_malloc(&pn.covered, cells));
      MGIC_CUDA(cudaMemsetAsync(pn.covered, 0, cells, A->ctx->stream));
    }
    const UnderLocal u = under_local(n, pn.corr);
    if (u.empty) continue;
    MGIC_TRY(mgk::box_set_u8(A->ctx, u.g, pn.covered + (u.ptr - ref->p), 1, mask_at(n.op, u.fineOff), n.op->n[0], (long long)n.op->n[0] * n.op->n[1]));
  }
  return MGIC_OK;
}

static int amr_cycle(mgic_amr *A, int l) {
  const int S = A->base->P.numMGsmooth;
  if (l == 0) return vcycle_run(A->base, A->nodes[0].corr, A->nodes[0].res, true);   // corr_0 = oneCycle(res_0) from zero
  const int q0 = A->levelStart[l], q1 = A->levelStart[l + 1];
  // ---- down
  for (int q = q0; q < q1; q++) {
    AmrNode &n = A->nodes[q];
    MGIC_TRY(relax_from_zero(n.op, n.corr, n.res, S));
  }
  for (int P = A->levelStart[l - 1]; P < A->levelStart[l]; P++) MGIC_TRY(mgic_op_set_to_zero(A->nodes[P].op, A->nodes[P].corr));
  for (int q = q0; q < q1; q++) {
    AmrNode &n = A->nodes[q];
    AmrNode &pn = A->nodes[n.parent];
    CoarseSrc cs;
    MGIC_TRY(amr_coarse_source(A, n, pn.corr, &cs, true));          // the coarser correction was just zeroed
    MGIC_TRY(mgic_op_amr_residual_nf(n.op, n.tmp, n.corr, cs.f, cs.lo, n.res, 1));
    MGIC_TRY(average_under(A, n, pn.res, n.tmp));
  }
  MGIC_TRY(amr_cycle(A, l - 1));
  // ---- up
  for (int q = q0; q < q1; q++) {
    AmrNode &n = A->nodes[q];
    AmrNode &pn = A->nodes[n.parent];
    CoarseSrc cs;
    MGIC_TRY(amr_coarse_source(A, n, pn.corr, &cs));
    MGIC_TRY(mgk::prolong(A->ctx, n.op->geom(), n.corr->p, under_in(cs, n), cs.f->sy, cs.f->sz, n.op->mask));
    MGIC_TRY(mgic_op_amr_residual_nf(n.op, n.tmp, n.corr, cs.f, cs.lo, n.res, 1));
    MGIC_TRY(mgic_op_assign(n.op, n.res, n.tmp));
    MGIC_TRY(relax_from_zero(n.op, n.tmp, n.res, S));
    MGIC_TRY(mgic_op_incr(n.op, n.corr, n.tmp, 1.0));
  }
  return MGIC_OK;
}

static int amr_check_vec(const mgic_amr *A, mgic_field *const *x) {
  MGIC_REQUIRE(A && x, "NULL argument");
  for (size_t q = 0; q < A->nodes.size(); q++) {
    MGIC_REQUIRE(x[q], "NULL field in a level vector");
    REQ_SHAPE(A->nodes[q].op, x[q]);
  }
  return MGIC_OK;
}

// corr (out) = the correction of one AMR V-cycle for the residuals res (in), both level vectors.  res of a coarser node
// under a finer patch is ignored (replaced by the averaged fine residual), as in AMRVCycle.
extern "C" int mgic_amr_vcycle(mgic_amr *A, mgic_field *const *corr, mgic_field *const *res) {
  MGIC_TRY(amr_check_vec(A, corr)); MGIC_TRY(amr_check_vec(A, res));
  for (AmrNode &n : A->nodes) MGIC_TRY(mgic_op_assign(n.op, n.res, res[&n - A->nodes.data()]));
  MGIC_TRY(amr_cycle(A, A->nlevels() - 1));
  for (AmrNode &n : A->nodes) MGIC_TRY(mgic_op_assign(n.op, corr[&n - A->nodes.data()], n.corr));
  return MGIC_OK;
}

// ---- [Chombo] MultilevelLinearOp over the hierarchy: applyOp / residual per level with the coarse-fine ghosts from the
// level below (AMROperatorNF; AMROperator's reflux is the reference's no-op), norms and dot products over the valid cells
// NOT covered by a finer level (zeroCovered on a temporary, as MultilevelLinearOp::dotProduct / norm do).
extern "C" int mgic_amr_apply(mgic_amr *A, mgic_field *const *lhs, mgic_field *const *phi, int homogeneous) {
  MGIC_TRY(amr_check_vec(A, lhs)); MGIC_TRY(amr_check_vec(A, phi));
  MGIC_TRY(mgic_op_apply(A->nodes[0].op, lhs[0], phi[0], homogeneous));
  for (size_t q = 1; q < A->nodes.size(); q++) {
    const AmrNode &n = A->nodes[q];
    CoarseSrc cs;
    MGIC_TRY(amr_coarse_source(A, n, phi[n.parent], &cs));
    MGIC_TRY(mgic_op_amr_operator_nf(n.op, lhs[q], phi[q], cs.f, cs.lo, homogeneous));
  }
  return MGIC_OK;
}
extern "C" int mgic_amr_residual(mgic_amr *A, mgic_field *const *res, mgic_field *const *phi, mgic_field *const *rhs, int homogeneous) {
  MGIC_TRY(amr_check_vec(A, res)); MGIC_TRY(amr_check_vec(A, phi)); MGIC_TRY(amr_check_vec(A, rhs));
  MGIC_TRY(mgic_op_residual(A->nodes[0].op, res[0], phi[0], rhs[0], homogeneous));
  for (size_t q = 1; q < A->nodes.size(); q++) {
    const AmrNode &n = A->nodes[q];
    CoarseSrc cs;
    MGIC_TRY(amr_coarse_source(A, n, phi[n.parent], &cs));
    MGIC_TRY(mgic_op_amr_residual_nf(n.op, res[q], phi[q], cs.f, cs.lo, rhs[q], homogeneous));
  }
  return MGIC_OK;
}
// x = 0 on every cell that a finer patch covers ([Chombo] AMRPoissonOp::zeroCovered, level by level)
extern "C" int mgic_amr_zero_covered(mgic_amr *A, mgic_field *const *x) {
  MGIC_TRY(amr_check_vec(A, x));
  for (size_t q = 1; q < A->nodes.size(); q++) {
    const AmrNode &n = A->nodes[q];
    MGIC_TRY(zero_under(A, n, x[n.parent]));
  }
  return MGIC_OK;
}
// the covered cells of every coarser array = the 8-cell average of the finer patch, finest level first
// ([Chombo] CoarseAverage::averageToCoarse, what Main_PoissonSolver.cpp's output path and AMRMultiGrid do after a solve)
extern "C" int mgic_amr_average_down(mgic_amr *A, mgic_field *const *x) {
  MGIC_TRY(amr_check_vec(A, x));
  for (size_t q = A->nodes.size() - 1; q >= 1; q--) {
    const AmrNode &n = A->nodes[q];
    MGIC_TRY(average_under(A, n, x[n.parent], x[q]));
  }
  return MGIC_OK;
}
// kind-reduction of every node of a level vector, the cells a finer node covers counting as zero ([Chombo]
// MultilevelLinearOp::norm / dotProduct zero them on temporaries: same bits), one launch per node into the scalar slots and
// ONE readback for the whole hierarchy.  A z-slab-distributed node is summed over the ranks; replicated nodes are not.
static int amr_reduce(mgic_amr *A, mgic_field *const *x, mgic_field *const *y, int kind, std::vector<double> *out) {
  mgic_ctx *c = A->ctx;
  const int nn = (int)A->nodes.size(), SLOTS = 64;
  out->assign((size_t)nn, 0.0);
  for (int q0 = 0; q0 < nn; q0 += SLOTS) {
    const int cnt = std::min(SLOTS, nn - q0);
    for (int q = q0; q < q0 + cnt; q++) {
      const AmrNode &n = A->nodes[q];
      MGIC_TRY(mgk::reduce(c, n.op->geom(), x[q]->p, y ? y[q]->p : nullptr, kind, q - q0, n.covered));
      if (c->nranks > 1 && !n.op->isGlobal) {
        MGIC_REQUIRE(c->allreduce, "multi-rank context without an allreduce hook (mgic_comm_init)");
        MGIC_TRY(c->allreduce(c, c->d_scal + (q - q0), 1, kind == 0 ? 1 : 0));
      }
    }
    MGIC_TRY(fetch_scalars(c, 0, cnt, 0, out->data() + q0, false));
  }
  return MGIC_OK;
}
// ord 0: max |x| over the valid, uncovered cells (the outer solver's m_normType = 0, Main_PoissonSolver.cpp:176);
// ord 1, 2: [Chombo] computeNorm -- (sum |x|^p dx_l^3)^(1/p) over the same cells (Main_PoissonSolver.cpp:208)
extern "C" int mgic_amr_norm(mgic_amr *A, mgic_field *const *x, int ord, double *out) {
  MGIC_TRY(amr_check_vec(A, x));
  MGIC_REQUIRE(out && ord >= 0 && ord <= 2, "norm order must be 0, 1 or 2");
  std::vector<double> v;
  MGIC_TRY(amr_reduce(A, x, nullptr, ord, &v));
  double acc = 0.0;
  for (size_t q = 0; q < A->nodes.size(); q++) {
    const AmrNode &n = A->nodes[q];
    if (ord == 0) acc = std::max(acc, v[q]);
    else acc += v[q] * (n.op->dx * n.op->dx * n.op->dx);
  }
  *out = ord == 2 ? sqrt(acc) : acc;
  return MGIC_OK;
}
// sum over the valid, uncovered cells of x*y*dx_l^3 ([Chombo] MultilevelLinearOp::dotProduct: covered cells zeroed on
// temporaries, level products scaled by the cell volume)
extern "C" int mgic_amr_dot(mgic_amr *A, mgic_field *const *x, mgic_field *const *y, double *out) {
  MGIC_TRY(amr_check_vec(A, x)); MGIC_TRY(amr_check_vec(A, y));
  MGIC_REQUIRE(out, "NULL argument");
  std::vector<double> v;
  MGIC_TRY(amr_reduce(A, x, y, 3, &v));
  double acc = 0.0;
  for (size_t q = 0; q < A->nodes.size(); q++) {
    const AmrNode &n = A->nodes[q];
    acc += v[q] * (n.op->dx * n.op->dx * n.op->dx);
  }
  *out = acc;
  return MGIC_OK;
}

// level vector k (0..9) of the hierarchy's work space; all ten are allocated at the first call, so the pointers stay valid
#define MGIC_AMR_WORK 10
static int amr_work(mgic_amr *A, int k, mgic_field ***vec) {
  const size_t nn = A->nodes.size();
  if (A->work.size() < MGIC_AMR_WORK * nn) {
    A->work.reserve(MGIC_AMR_WORK * nn);
    while (A->work.size() < MGIC_AMR_WORK * nn) {
      mgic_field *f = nullptr;
      MGIC_TRY(mgic_field_create(A->nodes[A->work.size() % nn].op, &f));
      A->work.push_back(f);
    }
  }
  *vec = A->work.data() + (size_t)k * nn;
  return MGIC_OK;
}

// [Chombo] MultilevelLinearOp::preCond: cor = 0, then numMGIterations AMR V-cycles.  AMRVCycle works on the residual it is
// given, so from the second cycle on it is given the residual of the correction so far (res - L cor) and its result is
// added -- on one level this is MultiGrid::oneCycle repeated on the same (e, r) pair, which mgic_mg_outer_solve does.
extern "C" int mgic_amr_precond(mgic_amr *A, mgic_field *const *cor, mgic_field *const *res) {
  MGIC_TRY(amr_check_vec(A, cor)); MGIC_TRY(amr_check_vec(A, res));
  const int nIt = A->base->P.numMGIterations;
  if (nIt < 1) {
    for (size_t q = 0; q < A->nodes.size(); q++) MGIC_TRY(mgic_op_set_to_zero(A->nodes[q].op, cor[q]));
    return MGIC_OK;
  }
  MGIC_TRY(mgic_amr_vcycle(A, cor, res));
  mgic_field **r2 = nullptr, **c2 = nullptr;
  for (int it = 1; it < nIt; it++) {
    MGIC_TRY(amr_work(A, 8, &r2));
    MGIC_TRY(amr_work(A, 9, &c2));
    MGIC_TRY(mgic_amr_residual(A, r2, cor, res, 1));
    MGIC_TRY(mgic_amr_vcycle(A, c2, r2));
    for (size_t q = 0; q < A->nodes.size(); q++) MGIC_TRY(mgic_op_incr(A->nodes[q].op, cor[q], c2[q], 1.0));
  }
  return MGIC_OK;
}

struct AmrSpace {
  typedef mgic_field *const *Vec;
  typedef mgic_field *const *CVec;
  mgic_amr *A;
  int residual(Vec r, Vec phi, CVec rhs, bool homogeneous) { return mgic_amr_residual(A, r, phi, rhs, homogeneous); }
  int apply(Vec lhs, Vec phi, int homogeneous) { return mgic_amr_apply(A, lhs, phi, homogeneous); }
  int preCond(Vec cor, Vec res) { return mgic_amr_precond(A, cor, res); }
  int norm(CVec x, int ord, double *out) { return mgic_amr_norm(A, x, ord, out); }
  int dot(CVec x, CVec y, double *out) { return mgic_amr_dot(A, x, y, out); }
#define AMR_EACH(call) \
  for (size_t q = 0; q < A->nodes.size(); q++) MGIC_TRY(call); \
  return MGIC_OK
  int assign(Vec y, CVec x) { AMR_EACH(mgic_op_assign(A->nodes[q].op, y[q], x[q])); }
  int set_to_zero(Vec y) { AMR_EACH(mgic_op_set_to_zero(A->nodes[q].op, y[q])); }
  int scale(Vec y, double s) { AMR_EACH(mgic_op_scale(A->nodes[q].op, y[q], s)); }
  int incr(Vec y, CVec x, double s) { AMR_EACH(mgic_op_incr(A->nodes[q].op, y[q], x[q], s)); }
#undef AMR_EACH
};

// The reference's linear solve on a hierarchy (Main_PoissonSolver.cpp:169-184): BiCGStabSolver<Vector<LevelData*>> over
// MultilevelLinearOp, preconditioned by numMGIterations AMR V-cycles; max-norm, eps = tolerance, imax = max_iterations.
extern "C" int mgic_amr_outer_solve(mgic_amr *A, mgic_field *const *dpsi, mgic_field *const *rhs, int *iterations, int *exit_status,
                                    double *norms, int max_norms) {
  MGIC_TRY(amr_check_vec(A, dpsi)); MGIC_TRY(amr_check_vec(A, rhs));
  mgic_field *const *w[8];
  for (int k = 0; k < 8; k++) {
    mgic_field **v = nullptr;
    MGIC_TRY(amr_work(A, k, &v));
    w[k] = v;
  }
  AmrSpace sp{A};
  BiCGParams bp;
  bp.homogeneous = false;               // Main_PoissonSolver.cpp:172-173
  bp.normType = 0;                      // :176
  bp.eps = A->base->P.tolerance;        // :177
  bp.imax = A->base->P.max_iterations;  // :178
  return bicgstab_t(sp, dpsi, rhs, w, bp, iterations, exit_status, norms, max_norms);
}

// f1: [Chombo] MultilevelLinearOp::preCond on one AMR level = zero cor, numMGIterations V-cycles
struct OuterLin : LinOp {
  mgic_mg *mg;
  int preCond(mgic_field *cor, mgic_field *res) override {
    if (mg->P.numMGIterations < 1 || !mg->ctx->fusePR) MGIC_TRY(mgic_op_set_to_zero(op, cor));
    for (int it = 0; it < mg->P.numMGIterations; it++) MGIC_TRY(vcycle_run(mg, cor, res, it == 0));
    return MGIC_OK;
  }
};

extern "C" int mgic_mg_outer_solve(mgic_mg *mg, mgic_field *dpsi, const mgic_field *rhs, int *iterations, int *exit_status,
                                   double *norms, int max_norms) {
  MGIC_REQUIRE(mg && dpsi && rhs, "NULL argument");
  OuterLin L;
  L.op = mg->ops[0];
  L.mg = mg;
  BiCGParams bp;
  bp.homogeneous = false;         // Main_PoissonSolver.cpp:172-173
  bp.normType = 0;                // :176
  bp.eps = mg->P.tolerance;       // :177
  bp.imax = mg->P.max_iterations; // :178
  return bicgstab(L, mg->outerWork, dpsi, rhs, bp, iterations, exit_status, norms, max_norms);
}

// ------------------------------------------------------------------------------------------------ source terms
extern "C" int mgic_vars_create(mgic_ctx *c, const mgic_params *P, int k0, int nzl, mgic_vars **out) {
  MGIC_REQUIRE(c && P && out, "NULL argument");
  MGIC_REQUIRE(k0 >= 0 && nzl >= 1 && k0 + nzl <= P->N[2], "bad slab");
  MGIC_CUDA(cudaSetDevice(c->device));
  mgic_vars *v = new mgic_vars;
  v->ctx = c; v->P = *P;
  for (int d = 0; d < 3; d++) v->n[d] = P->N[d];
  v->k0 = k0; v->nzl = nzl; v->dx = P->L / P->N[0];
  v->sy = P->N[0] + 2; v->sz = v->sy * (P->N[1] + 2); v->sc = v->sz * (nzl + 2);
  MGIC_CUDA(mgic_dev_malloc(&v->d, (size_t)v->sc * 8 * sizeof(double)));
  MGIC_CUDA(cudaMemsetAsync(v->d, 0, (size_t)v->sc * 8 * sizeof(double), c->stream));
  *out = v;
  return MGIC_OK;
}
// multigrid_vars of an AMR level > 0 (Main_PoissonSolver.cpp:79-88 for ilev > 0): over the bounding box of the patch
// operator's level, padded by one ghost layer; psi's coarse-fine ghost values live in six cell-indexed arrays (initial
// value 1 like every psi ghost, SetLevelData.cpp:42-54)
extern "C" int mgic_vars_create_patch(mgic_ctx *c, const mgic_params *P, const mgic_op *patch, mgic_vars **out) {
  MGIC_REQUIRE(c && P && patch && out && patch->isPatch, "bad argument (needs a patch operator)");
  MGIC_CUDA(cudaSetDevice(c->device));
  mgic_vars *v = new mgic_vars;
  v->ctx = c; v->P = *P;
  for (int d = 0; d < 3; d++) { v->n[d] = patch->n[d]; v->lo[d] = patch->plo[d]; v->ndom[d] = patch->ndom[d]; }
  v->isPatch = true; v->mask = patch->mask;
  v->k0 = patch->plo[2]; v->nzl = patch->n[2]; v->dx = patch->dx;
  v->sy = v->n[0] + 2; v->sz = v->sy * (v->n[1] + 2); v->sc = v->sz * (v->nzl + 2);
  MGIC_CUDA(mgic_dev_malloc(&v->d, (size_t)v->sc * 8 * sizeof(double)));
  MGIC_CUDA(cudaMemsetAsync(v->d, 0, (size_t)v->sc * 8 * sizeof(double), c->stream));
  const long long cells = (long long)v->n[0] * v->n[1] * v->nzl;
  for (int f = 0; f < 6; f++) {
    MGIC_CUDA(mgic_dev_malloc(&v->psiG[f], (size_t)cells * sizeof(double)));
    MGIC_TRY(mgk::fill(c, v->psiG[f], cells, 1.0));
  }
  *out = v;
  return MGIC_OK;
}
extern "C" int mgic_vars_destroy(mgic_vars *v) {
  if (!v) return MGIC_OK;
  for (int f = 0; f < 6; f++) mgic_dev_free(v->psiG[f]);
  mgic_dev_free(v->d);
  delete v;
  return MGIC_OK;
}
static int vars_copy(const mgic_vars *v, int comp, double *host, int ghost) {
  MGIC_REQUIRE(v && host && comp >= 0 && comp < 8, "bad argument");
  cudaMemcpy3DParms p;
  memset(&p, 0, sizeof(p));
  const size_t hx = v->n[0] + 2 * ghost, hy = v->n[1] + 2 * ghost;
  p.srcPtr = make_cudaPitchedPtr(v->d + (size_t)comp * v->sc, (size_t)v->sy * sizeof(double), v->sy, v->n[1] + 2);
  p.srcPos = make_cudaPos((size_t)(1 - ghost) * sizeof(double), 1 - ghost, 1 - ghost);
  p.dstPtr = make_cudaPitchedPtr(host, hx * sizeof(double), hx, hy);
  p.dstPos = make_cudaPos(0, 0, (ghost || v->isPatch) ? 0 : v->k0);   // patches: a bounding-box-shaped host array
  p.extent = make_cudaExtent(hx * sizeof(double), hy, v->nzl + 2 * ghost);
  p.kind = cudaMemcpyDeviceToHost;
  MGIC_CUDA(cudaMemcpy3DAsync(&p, v->ctx->stream));
  MGIC_CUDA(cudaStreamSynchronize(v->ctx->stream));
  return MGIC_OK;
}
extern "C" int mgic_vars_download(const mgic_vars *v, int comp, double *host) { return vars_copy(v, comp, host, 0); }
extern "C" int mgic_vars_download_ghosted(const mgic_vars *v, int comp, double *host) { return vars_copy(v, comp, host, 1); }

extern "C" int mgic_set_initial_conditions(mgic_vars *v, mgic_field *dpsi) {
  MGIC_REQUIRE(v, "vars is NULL");
  MGIC_TRY(mgk::init_conditions(v));
  if (v->isPatch)
    for (int f = 0; f < 6; f++) MGIC_TRY(mgk::fill(v->ctx, v->psiG[f], (long long)v->n[0] * v->n[1] * v->nzl, 1.0));
  if (dpsi) MGIC_CUDA(cudaMemsetAsync(dpsi->base, 0, dpsi->bytes, v->ctx->stream));  // dpsi = 0 (SetLevelData.cpp:55)
  return MGIC_OK;
}
static bool vars_match(const mgic_vars *v, const mgic_field *f) {
  return f && f->nx == v->n[0] && f->ny == v->n[1] && f->nz == v->nzl && f->k0 == (v->isPatch ? 0 : v->k0);
}
extern "C" int mgic_set_a_coef(mgic_vars *v, mgic_field *aCoef, double constant_K) {
  MGIC_REQUIRE(v && vars_match(v, aCoef), "aCoef does not match multigrid_vars");
  return mgk::set_rhs_acoef(v, nullptr, aCoef->p, constant_K);
}
extern "C" int mgic_set_rhs(mgic_vars *v, mgic_field *rhs, double constant_K) {
  MGIC_REQUIRE(v && vars_match(v, rhs), "rhs does not match multigrid_vars");
  return mgk::set_rhs_acoef(v, rhs->p, nullptr, constant_K);
}
extern "C" int mgic_set_rhs_and_a_coef(mgic_vars *v, mgic_field *rhs, mgic_field *aCoef, double constant_K) {
  MGIC_REQUIRE(v && vars_match(v, rhs) && vars_match(v, aCoef), "fields do not match multigrid_vars");
  return mgk::set_rhs_acoef(v, rhs->p, aCoef->p, constant_K);
}
extern "C" int mgic_set_b_coef(mgic_vars *v, mgic_field *bCoef) {
  MGIC_REQUIRE(v && vars_match(v, bCoef), "bCoef does not match multigrid_vars");
  Geom g; g.nx = bCoef->nx; g.ny = bCoef->ny; g.nz = bCoef->nz; g.sy = bCoef->sy; g.sz = bCoef->sz; g.k0 = bCoef->k0; g.gnz = bCoef->gnz;
  MGIC_TRY(mgk::set_val(v->ctx, g, bCoef->p, 1.0));  // SetLevelData.cpp:338
  if (bCoef->mask) MGIC_TRY(mgk::apply_mask(v->ctx, g, bCoef->p, bCoef->mask));
  return MGIC_OK;
}

// set_update_psi0 (SetLevelData.cpp:243-263) + computeNorm(dpsi, p = 2) (Main_PoissonSolver.cpp:208)
extern "C" int mgic_update_psi0(mgic_vars *v, mgic_op *op0, mgic_field *dpsi, double *dpsi_norm) {
  MGIC_REQUIRE(v && op0 && dpsi, "NULL argument");
  REQ_SHAPE(op0, dpsi);
  MGIC_REQUIRE(vars_match(v, dpsi), "dpsi does not match multigrid_vars");
  MGIC_TRY(halo(op0, dpsi, 1));  // :249 (the reference exchanges three layers; one is read)
  // dpsi's domain-face ghost as [Chombo] BiCGStabSolver leaves it: residual(r, phi, rhs, homogeneous = false) filled it with
  // the INHOMOGENEOUS value from phi's near cell, then phi += e over the ghosted FABs with e's ghost = a*e_near (every
  // vector accumulated into e went through applyOp's homogeneous fill): a*near + b with b = 2*bc_value (Dirichlet) /
  // +-dx*bc_value (Neumann).  Identical to the homogeneous form for params.txt's bc_value = 0.
  MGIC_TRY(mgk::update_psi(v, op0->geom(), op0->bck(false), dpsi->p));
  if (dpsi_norm) {
    double s;
    MGIC_TRY(local_reduce(op0, dpsi, nullptr, 2, &s));
    const double dV = op0->dx * op0->dx * op0->dx;
    *dpsi_norm = sqrt(s * dV);  // [Chombo] computeNorm: (sum |x|^2 dx^3)^(1/2)
  }
  return MGIC_OK;
}

// set_update_psi0 on an AMR level > 0 (Main_PoissonSolver.cpp:189-205): QuadCFInterp::coarseFineInterp(dpsi, dpsi_coarse),
// the exchange between the level's boxes (a neighbour read here), psi += dpsi over the ghosted boxes -- psi's coarse-fine
// ghosts are carried in the vars' cell-indexed ghost arrays, the physical ghosts in the padded array.
extern "C" int mgic_update_psi0_patch(mgic_vars *v, mgic_op *patch, mgic_field *dpsi, const mgic_field *dpsi_coarse, const int coarse_lo[3]) {
  MGIC_REQUIRE(v && patch && dpsi && dpsi_coarse && coarse_lo && v->isPatch && patch->isPatch, "bad argument");
  REQ_SHAPE(patch, dpsi);
  MGIC_REQUIRE(vars_match(v, dpsi), "dpsi does not match multigrid_vars");
  BCk k;
  MGIC_TRY(quad_cf_cells(patch, dpsi, dpsi_coarse, coarse_lo, false, &k));
  return mgk::update_psi_patch(v, k, dpsi->p);
}

// ------------------------------------------------------------------------------------------------ the whole problem on a hierarchy
// poissonSolve (Main_PoissonSolver.cpp:45-256) for max_level > 0: per level multigrid_vars / dpsi / rhs / aCoef / bCoef
// (:79-88), set_initial_conditions (:93), and per nonlinear iteration (:131-216) the source terms of every level, the
// operator factory + MultilevelLinearOp + BiCGStab rebuilt (:163-178), solver.solve (:184), the psi update with interlevel
// and intralevel ghosts (:189-205) and the composite norm of dpsi (:208).  A level is given as its connected components
// ("nodes"), each a list of boxes in ONE masked array (mgic_op_create_patch_boxes); node 0 is the base level.
struct HierNode {
  int level = 0;
  mgic_op *op = nullptr;          // patch operator (owned); null for the base level (the MG hierarchy's, rebuilt per iteration)
  mgic_vars *vars = nullptr;
  mgic_field *dpsi = nullptr, *rhs = nullptr, *a = nullptr, *b = nullptr;
  std::vector<int> boxes;         // the node's boxes, 6 ints each (the checkpoint is written box by box)
};
struct mgic_hier {
  mgic_ctx *ctx = nullptr;
  mgic_params P;            // max_level = number of finer levels
  mgic_op *lay0 = nullptr;  // base-level geometry (field factory)
  std::vector<HierNode> nodes;
  std::vector<int> perLevel;   // nodes per finer level
  int lastIterations = 0, lastStatus = 0;
  mgic_mg *mg = nullptr;       // the solver of the current nonlinear iteration (mgic_hier_define_solver .. _release_solver)
  mgic_amr *amr = nullptr;
};
static void hier_drop_solver(mgic_hier *H);
extern "C" int mgic_hier_destroy(mgic_hier *H) {
  if (!H) return MGIC_OK;
  hier_drop_solver(H);
  for (HierNode &n : H->nodes) {
    mgic_field_destroy(n.dpsi); mgic_field_destroy(n.rhs); mgic_field_destroy(n.a); mgic_field_destroy(n.b);
    mgic_vars_destroy(n.vars);
    mgic_op_destroy(n.op);
  }
  mgic_op_destroy(H->lay0);
  delete H;
  return MGIC_OK;
}
// nfiner finer levels; nnodes[l-1] nodes on level l; nboxes[q] boxes in finer node q (flattened over levels);
// boxes = all boxes, 6 ints each {lo0,lo1,lo2,hi0,hi1,hi2} in their level's index space
extern "C" int mgic_hier_create(mgic_ctx *c, const mgic_params *P, int nfiner, const int *nnodes, const int *nboxes, const int *boxes,
                                mgic_hier **out) {
  MGIC_REQUIRE(c && P && out && nfiner >= 0 && (nfiner == 0 || (nnodes && nboxes && boxes)), "bad argument");
  // multi-rank context: the base level is cut into equal z-slabs (one per rank, multiples of max_grid_size), the refined
  // levels are replicated on every rank (mgic_op_create_patch); every rank makes the same calls
  MGIC_REQUIRE(P->N[2] % c->nranks == 0 && (P->N[2] / c->nranks) % P->max_grid_size == 0,
               "the base level's z extent must split into equal slabs that are multiples of max_grid_size");
  MGIC_REQUIRE(!P->is_periodic, "the periodic constant-K branch (Main_PoissonSolver.cpp:137-150) is not implemented on hierarchies");
  mgic_hier *H = new mgic_hier;
  H->ctx = c; H->P = *P; H->P.max_level = nfiner;
  int rc = MGIC_OK;
  int bclo[3], bchi[3];
  for (int d = 0; d < 3; d++) { bclo[d] = P->bc_lo[d]; bchi[d] = P->bc_hi[d]; }
  const double dx0 = P->L / P->N[0];
#define H_TRY(x) do { rc = (x); if (rc != MGIC_OK) { mgic_hier_destroy(H); return rc; } } while (0)
  const int nzl = P->N[2] / c->nranks, k0 = c->rank * nzl;
  H_TRY(mgic_op_create(c, P->N, k0, nzl, dx0, P->alpha, P->beta, bclo, bchi, P->bc_value, &H->lay0));
  {
    HierNode n0;
    H->nodes.push_back(n0);
    HierNode &n = H->nodes.back();
    for (int k = 0; k < P->N[2]; k += P->max_grid_size)          // the base level's boxes: domainSplit (SetGrids.cpp:54-58)
      for (int j = 0; j < P->N[1]; j += P->max_grid_size)
        for (int i = 0; i < P->N[0]; i += P->max_grid_size) {
          const int b6[6] = {i, j, k, std::min(i + P->max_grid_size, P->N[0]) - 1, std::min(j + P->max_grid_size, P->N[1]) - 1,
                             std::min(k + P->max_grid_size, P->N[2]) - 1};
          n.boxes.insert(n.boxes.end(), b6, b6 + 6);
        }
    H_TRY(mgic_vars_create(c, P, k0, nzl, &n.vars));
    H_TRY(mgic_field_create(H->lay0, &n.dpsi)); H_TRY(mgic_field_create(H->lay0, &n.rhs));
    H_TRY(mgic_field_create(H->lay0, &n.a)); H_TRY(mgic_field_create(H->lay0, &n.b));
  }
  int qn = 0, qb = 0;
  for (int l = 1; l <= nfiner; l++) {
    H->perLevel.push_back(nnodes[l - 1]);
    int ndom[3];
    for (int d = 0; d < 3; d++) ndom[d] = P->N[d] << l;                       // refRatio == 2 on every level (PoissonParameters.cpp:75-79)
    const double dxl = dx0 / (double)(1 << l);
    for (int q = 0; q < nnodes[l - 1]; q++, qn++) {
      HierNode nd;
      nd.level = l;
      H->nodes.push_back(nd);
      HierNode &n = H->nodes.back();
      H_TRY(mgic_op_create_patch_boxes(c, ndom, nboxes[qn], boxes + 6 * (size_t)qb, dxl, 2.0 * dxl, P->alpha, P->beta, bclo, bchi,
                                       P->bc_value, &n.op));
      n.boxes.assign(boxes + 6 * (size_t)qb, boxes + 6 * (size_t)(qb + nboxes[qn]));
      qb += nboxes[qn];
      H_TRY(mgic_vars_create_patch(c, P, n.op, &n.vars));
      H_TRY(mgic_field_create(n.op, &n.dpsi)); H_TRY(mgic_field_create(n.op, &n.rhs));
      H_TRY(mgic_field_create(n.op, &n.a)); H_TRY(mgic_field_create(n.op, &n.b));
    }
  }
#undef H_TRY
  *out = H;
  return MGIC_OK;
}
extern "C" int mgic_hier_nodes(const mgic_hier *H) { return H ? (int)H->nodes.size() : 0; }
extern "C" int mgic_hier_node_info(const mgic_hier *H, int q, int *level, int lo[3], int n[3], long long *valid_cells) {
  MGIC_REQUIRE(H && q >= 0 && q < (int)H->nodes.size(), "bad argument");
  const HierNode &nd = H->nodes[q];
  if (level) *level = nd.level;
  for (int d = 0; d < 3; d++) {
    if (lo) lo[d] = nd.op ? nd.op->plo[d] : 0;
    if (n) n[d] = nd.op ? nd.op->n[d] : H->P.N[d];
  }
  if (valid_cells) *valid_cells = nd.op ? mgic_op_valid_cells(nd.op) : (long long)H->P.N[0] * H->P.N[1] * H->P.N[2];
  return MGIC_OK;
}
// what: 0..7 multigrid_vars component (0 = psi), 8 dpsi, 9 rhs, 10 aCoef; host array has the node's (bounding box) shape
extern "C" int mgic_hier_download(const mgic_hier *H, int q, int what, double *host) {
  MGIC_REQUIRE(H && host && q >= 0 && q < (int)H->nodes.size() && what >= 0 && what <= 10, "bad argument");
  const HierNode &nd = H->nodes[q];
  if (what < 8) {
    MGIC_TRY(mgic_vars_download(nd.vars, what, host));
    if (nd.op && !nd.op->hmask.empty())   // the array also covers bounding-box cells outside the level's boxes: not data
      for (size_t i = 0; i < nd.op->hmask.size(); i++)
        if (!nd.op->hmask[i]) host[i] = 0.0;
    return MGIC_OK;
  }
  return mgic_field_download(what == 8 ? nd.dpsi : what == 9 ? nd.rhs : nd.a, host);
}
extern "C" int mgic_hier_get_mask(const mgic_hier *H, int q, unsigned char *host) {
  MGIC_REQUIRE(H && host && q >= 0 && q < (int)H->nodes.size(), "bad argument");
  if (q == 0) { memset(host, 1, (size_t)H->P.N[0] * H->P.N[1] * H->P.N[2]); return MGIC_OK; }
  return mgic_op_get_mask(H->nodes[q].op, host);
}
extern "C" int mgic_hier_set_initial_conditions(mgic_hier *H) {   // Main_PoissonSolver.cpp:90-96
  MGIC_REQUIRE(H, "NULL argument");
  for (HierNode &n : H->nodes) MGIC_TRY(mgic_set_initial_conditions(n.vars, n.dpsi));
  return MGIC_OK;
}
// ---- the body of the nonlinear loop (Main_PoissonSolver.cpp:131-212) step by step, as the reference's driver calls it
extern "C" int mgic_hier_set_solver_params(mgic_hier *H, int numMGsmooth, int numMGIterations, int preCondSolverDepth, double tolerance,
                                           int max_iterations) {
  MGIC_REQUIRE(H, "NULL argument");
  H->P.numMGsmooth = numMGsmooth; H->P.numMGIterations = numMGIterations; H->P.preCondSolverDepth = preCondSolverDepth;   // :108-117
  H->P.tolerance = tolerance; H->P.max_iterations = max_iterations;                                                       // :119-123
  return MGIC_OK;
}
// set_a_coef / set_b_coef / set_rhs on every level (:154-160)
extern "C" int mgic_hier_set_sources(mgic_hier *H, double constant_K) {
  MGIC_REQUIRE(H, "NULL argument");
  for (HierNode &n : H->nodes) {
    MGIC_TRY(mgic_set_rhs_and_a_coef(n.vars, n.rhs, n.a, constant_K));
    MGIC_TRY(mgic_set_b_coef(n.vars, n.b));
  }
  return MGIC_OK;
}
// The reference rebuilds factory, MultilevelLinearOp and every MG operator each nonlinear iteration (:163-170) because aCoef
// changed.  Here the objects survive: the coefficient ARRAYS are the same ones, so re-deriving the coarsened coefficients and
// lambda in place (mgic_mg_refresh_coefs, the patches' lambda) gives the operators a rebuild would give -- without freeing and
// re-allocating ~3 GB of fields and re-capturing the V-cycle graph per iteration.  They go when the hierarchy goes, or when a
// parameter that shapes them changes.
static void hier_drop_solver(mgic_hier *H) {
  mgic_amr_destroy(H->amr); H->amr = nullptr;
  mgic_mg_destroy(H->mg); H->mg = nullptr;
}
extern "C" int mgic_hier_release_solver(mgic_hier *H) {
  (void)H;
  return MGIC_OK;
}
// defineOperatorFactory + MultilevelLinearOp::define (:163-170): rebuilt every nonlinear iteration, like the reference
extern "C" int mgic_hier_define_solver(mgic_hier *H) {
  MGIC_REQUIRE(H, "NULL argument");
  std::vector<mgic_op *> patches;
  for (size_t q = 1; q < H->nodes.size(); q++) {
    HierNode &n = H->nodes[q];
    // bCoef == 1 (set_b_coef, SetLevelData.cpp:330-340): b*x == x exactly, so the patch operators drop the stream like the
    // base level does (mgic_mg_create detects it there); set_coefs marks lambda for recomputation
    MGIC_TRY(mgic_op_set_coefs(n.op, n.a, nullptr, H->P.alpha, H->P.beta));
    patches.push_back(n.op);
  }
  if (H->mg && H->amr && H->mg->P.numMGsmooth == H->P.numMGsmooth && H->mg->P.preCondSolverDepth == H->P.preCondSolverDepth &&
      H->mg->P.coefficient_average_type == H->P.coefficient_average_type)
    return mgic_mg_refresh_coefs(H->mg);                                    // same arrays, new values: coarsen + lambda in place
  hier_drop_solver(H);
  mgic_params P0 = H->P;
  P0.max_level = 0;                                                         // the base level's MG hierarchy
  MGIC_TRY(mgic_mg_create(H->ctx, &P0, H->nodes[0].a, H->nodes[0].b, &H->mg));
  return mgic_amr_create_levels(H->mg, (int)H->perLevel.size(), H->perLevel.data(), patches.data(), &H->amr);
}
// solver.solve(dpsi, rhs) (:173-184); dpsi keeps its previous value as the initial guess (:93 is its only zeroing)
extern "C" int mgic_hier_solve(mgic_hier *H, int *iterations, int *exit_status) {
  MGIC_REQUIRE(H && H->amr, "mgic_hier_define_solver first");
  const size_t nn = H->nodes.size();
  std::vector<mgic_field *> dpsi(nn), rhs(nn);
  for (size_t q = 0; q < nn; q++) { dpsi[q] = H->nodes[q].dpsi; rhs[q] = H->nodes[q].rhs; }
  // the hierarchy's solver parameters may have changed since the MG hierarchy was built
  H->mg->P.tolerance = H->P.tolerance; H->mg->P.max_iterations = H->P.max_iterations; H->mg->P.numMGIterations = H->P.numMGIterations;
  int it = 0, st = 0;
  MGIC_TRY(mgic_amr_outer_solve(H->amr, dpsi.data(), rhs.data(), &it, &st, nullptr, 0));
  H->lastIterations = it; H->lastStatus = st;
  if (iterations) *iterations = it;
  if (exit_status) *exit_status = st;
  return MGIC_OK;
}
// :189-205, coarsest level first (the finer level's coarse-fine ghosts come from the coarser level's dpsi)
extern "C" int mgic_hier_update_psi(mgic_hier *H) {
  MGIC_REQUIRE(H && H->amr, "mgic_hier_define_solver first");
  MGIC_TRY(mgic_update_psi0(H->nodes[0].vars, H->mg->ops[0], H->nodes[0].dpsi, nullptr));
  for (size_t q = 1; q < H->nodes.size(); q++) {
    int level = 0, parent = -1;
    MGIC_TRY(mgic_amr_node_info(H->amr, (int)q, &level, &parent));
    CoarseSrc cs;   // the coarser level's dpsi around the patch (gathered when that level is the distributed base level)
    MGIC_TRY(amr_coarse_source(H->amr, H->amr->nodes[q], H->nodes[parent].dpsi, &cs));
    MGIC_TRY(mgic_update_psi0_patch(H->nodes[q].vars, H->nodes[q].op, H->nodes[q].dpsi, cs.f, cs.lo));
  }
  return MGIC_OK;
}
// computeNorm(dpsi, refRatio, coarsestDx, Interval(0, 0)) (:208): p = 2 over the cells no finer level covers
extern "C" int mgic_hier_dpsi_norm(mgic_hier *H, double *out) {
  MGIC_REQUIRE(H && H->amr && out, "mgic_hier_define_solver first");
  std::vector<mgic_field *> dpsi(H->nodes.size());
  for (size_t q = 0; q < H->nodes.size(); q++) dpsi[q] = H->nodes[q].dpsi;
  return mgic_amr_norm(H->amr, dpsi.data(), 2, out);
}
// one pass of the nonlinear loop's body (Main_PoissonSolver.cpp:131-212, non-periodic: constant_K = 0)
extern "C" int mgic_hier_nl_iteration(mgic_hier *H, double *dpsi_norm, int *solver_iterations, int *solver_status) {
  MGIC_REQUIRE(H, "NULL argument");
  MGIC_TRY(mgic_hier_set_sources(H, 0.0));
  int rc = mgic_hier_define_solver(H);
  if (rc == MGIC_OK) rc = mgic_hier_solve(H, solver_iterations, solver_status);
  if (rc == MGIC_OK) rc = mgic_hier_update_psi(H);
  double nrm = 0.0;
  if (rc == MGIC_OK) rc = mgic_hier_dpsi_norm(H, &nrm);
  mgic_hier_release_solver(H);
  if (dpsi_norm) *dpsi_norm = nrm;
  return rc;
}
// output_final_data (Source/WriteOutput.H:127-227): the GRChombo checkpoint -- header ints / reals / strings, per level a
// group with ref_ratio, tag_buffer_size, dx, dt = dx / 4, time, prob_domain, is_periodic_*, the level's boxes and the 32
// GRChombo variables (set_output_data, Source/SetLevelData.cpp:343-396) with three ghost layers per box, stored as Chombo's
// write(handle, LevelData, "data") does: box after box, per box component after component, x fastest over the ghosted box.
// There is no HDF5 library in this build: the file is a self-describing container -- "MGICCHK1", a uint64 header length,
// a JSON header with everything above plus byte offsets, zero padding to 8 bytes, then the doubles -- that
// tools/mgic2hdf5.py turns into vcPoissonFinal.3d.hdf5 (dataset for dataset) wherever h5py exists.
extern "C" int mgic_hier_write_checkpoint(mgic_hier *H, const char *path, double constant_K) {
  MGIC_REQUIRE(H && path, "NULL argument");
  mgic_ctx *c = H->ctx;
  const int NV = 32, NG = 3;
  // multi-rank: the base level lives in z-slabs, its boxes' three ghost layers reach into the neighbours' planes.  psi of the
  // whole base level is gathered on every rank (interior planes by one all-gather, the two physical z-ghost planes from the
  // first / last rank) into a whole-level multigrid_vars whose other components are the analytic initial data; rank 0 writes
  // the file, the refined levels are replicated anyway.  Collective: every rank calls, only rank 0 touches `path`.
  mgic_vars *whole = nullptr;
  struct WholeGuard { mgic_vars *&v; ~WholeGuard() { mgic_vars_destroy(v); } } wholeGuard{whole};
  if (c->nranks > 1) {
    MGIC_REQUIRE(c->allgather, "multi-rank hierarchy without the communication hooks (mgic_comm_init)");
    const mgic_vars *v0 = H->nodes[0].vars;
    MGIC_REQUIRE(v0->nzl * c->nranks == H->P.N[2] && v0->k0 == c->rank * v0->nzl, "the base level is not cut into equal z-slabs");
    MGIC_TRY(mgic_vars_create(c, &H->P, 0, H->P.N[2], &whole));
    MGIC_TRY(mgk::init_conditions(whole));
    const size_t plane = (size_t)v0->sz;
    MGIC_TRY(c->allgather(c, v0->d + plane, whole->d + plane, plane * v0->nzl));          // psi is component 0; padded planes 1 .. nzl
    double *ends = nullptr;
    MGIC_CUDA(mgic_dev_malloc(&ends, (size_t)(c->nranks + 1) * 2 * plane * sizeof(double)));
    cudaError_t e1 = cudaMemcpyAsync(ends, v0->d, plane * sizeof(double), cudaMemcpyDeviceToDevice, c->stream);
    cudaError_t e2 = cudaMemcpyAsync(ends + plane, v0->d + plane * (v0->nzl + 1), plane * sizeof(double), cudaMemcpyDeviceToDevice, c->stream);
    int rc = (e1 == cudaSuccess && e2 == cudaSuccess) ? c->allgather(c, ends, ends + 2 * plane, 2 * plane) : MGIC_ERR_CUDA;
    if (rc == MGIC_OK) {
      const double *all = ends + 2 * plane;
      e1 = cudaMemcpyAsync(whole->d, all, plane * sizeof(double), cudaMemcpyDeviceToDevice, c->stream);                                 // rank 0's plane below the domain
      e2 = cudaMemcpyAsync(whole->d + plane * (H->P.N[2] + 1), all + (size_t)(c->nranks - 1) * 2 * plane + plane, plane * sizeof(double),
                           cudaMemcpyDeviceToDevice, c->stream);                                                                       // the last rank's plane above it
      if (e1 != cudaSuccess || e2 != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) rc = MGIC_ERR_CUDA;
    }
    mgic_dev_free(ends);
    if (rc != MGIC_OK) { mgic_set_error("gathering the base level for the checkpoint failed"); return rc; }
    if (c->rank != 0) return MGIC_OK;
  }
  static const char *names[32] = {"chi", "h11", "h12", "h13", "h22", "h23", "h33", "K", "A11", "A12", "A13", "A22", "A23", "A33", "Theta",
                                  "Gamma1", "Gamma2", "Gamma3", "lapse", "shift1", "shift2", "shift3", "B1", "B2", "B3", "phi", "Pi", "Ham",
                                  "Mom1", "Mom2", "Mom3", nullptr};
  const int nlev = (int)H->perLevel.size() + 1;
  // boxes per level in node order; data offsets in doubles
  struct LB { int node; const int *b; long long off; };
  std::vector<std::vector<LB>> lev(nlev);
  long long total = 0;
  for (size_t q = 0; q < H->nodes.size(); q++)
    for (size_t b = 0; b + 5 < H->nodes[q].boxes.size(); b += 6) {
      const int *bx = H->nodes[q].boxes.data() + b;
      lev[H->nodes[q].level].push_back({(int)q, bx, total});
      total += (long long)NV * (bx[3] - bx[0] + 1 + 2 * NG) * (bx[4] - bx[1] + 1 + 2 * NG) * (bx[5] - bx[2] + 1 + 2 * NG);
    }
  std::string js = "{\"format\": \"MGICCHK1: GRChombo checkpoint of MG_IC_code (Source/WriteOutput.H:127-227) without HDF5\", ";
  char buf[512];
  js += "\"filename\": \"vcPoissonFinal.3d.hdf5\", \"root\": {\"ints\": {";
  snprintf(buf, sizeof buf, "\"max_level\": %d, \"num_levels\": %d, \"iteration\": 0, \"num_components\": %d", nlev - 1, nlev, NV);
  js += buf;
  for (int l = 0; l < nlev; l++) { snprintf(buf, sizeof buf, ", \"regrid_interval_%d\": 1, \"steps_since_regrid_%d\": 0", l, l); js += buf; }
  js += "}, \"reals\": {\"time\": 0.0}, \"strings\": {";
  for (int v = 0; v < NV - 1; v++) { snprintf(buf, sizeof buf, "%s\"component_%d\": \"%s\"", v ? ", " : "", v, names[v]); js += buf; }
  js += "}}, \"note_components\": \"the reference declares 32 variables (NUM_GRCHOMBO_VARS) and names 31 (GRChomboUserVariables.hpp:55-78)\", ";
  js += "\"ghost\": [3, 3, 3], \"dtype\": \"float64 little endian\", \"levels\": [";
  const double dx0 = H->P.L / H->P.N[0];
  for (int l = 0; l < nlev; l++) {
    const double dx = dx0 / (double)(1 << l);
    snprintf(buf, sizeof buf, "%s{\"group\": \"level_%d\", \"ints\": {\"ref_ratio\": 2, \"tag_buffer_size\": 3, \"is_periodic_0\": 1, \"is_periodic_1\": 1, "
             "\"is_periodic_2\": 1}, \"reals\": {\"dx\": %.17g, \"dt\": %.17g, \"time\": 0.0}, \"prob_domain\": [0, 0, 0, %d, %d, %d], \"boxes\": [",
             l ? ", " : "", l, dx, 0.25 * dx, (H->P.N[0] << l) - 1, (H->P.N[1] << l) - 1, (H->P.N[2] << l) - 1);
    js += buf;
    for (size_t b = 0; b < lev[l].size(); b++) {
      const int *x = lev[l][b].b;
      snprintf(buf, sizeof buf, "%s[%d, %d, %d, %d, %d, %d]", b ? ", " : "", x[0], x[1], x[2], x[3], x[4], x[5]);
      js += buf;
    }
    js += "], \"offsets\": [";
    for (size_t b = 0; b < lev[l].size(); b++) { snprintf(buf, sizeof buf, "%s%lld", b ? ", " : "", lev[l][b].off - (lev[l].empty() ? 0 : lev[l][0].off)); js += buf; }
    long long end = (l + 1 < nlev && !lev[l + 1].empty()) ? lev[l + 1][0].off : total;
    snprintf(buf, sizeof buf, "%s%lld], \"data_start_double\": %lld}", lev[l].empty() ? "" : ", ", end - (lev[l].empty() ? end : lev[l][0].off),
             lev[l].empty() ? end : lev[l][0].off);
    js += buf;
  }
  js += "]}";
  FILE *fp = fopen(path, "wb");
  if (!fp) { mgic_set_error("cannot open %s for writing", path); return MGIC_ERR_ARG; }
  unsigned long long hl = js.size();
  fwrite("MGICCHK1", 1, 8, fp);
  fwrite(&hl, sizeof hl, 1, fp);
  fwrite(js.data(), 1, js.size(), fp);
  const char zeros[8] = {0};
  fwrite(zeros, 1, (8 - js.size() % 8) % 8, fp);
  int rc = MGIC_OK;
  std::vector<double> hbuf;
  double *dbuf = nullptr;
  size_t dcap = 0;
  for (int l = 0; l < nlev && rc == MGIC_OK; l++)
    for (const LB &lb : lev[l]) {
      const int lo[3] = {lb.b[0], lb.b[1], lb.b[2]}, n[3] = {lb.b[3] - lb.b[0] + 1, lb.b[4] - lb.b[1] + 1, lb.b[5] - lb.b[2] + 1};
      const size_t cnt = (size_t)NV * (n[0] + 2 * NG) * (n[1] + 2 * NG) * (n[2] + 2 * NG);
      if (cnt > dcap) {
        mgic_dev_free(dbuf);
        dbuf = nullptr;
        if (mgic_dev_malloc(&dbuf, cnt * sizeof(double)) != cudaSuccess) { mgic_set_error("cudaMalloc failed in the checkpoint writer"); rc = MGIC_ERR_CUDA; break; }
        dcap = cnt;
      }
      hbuf.resize(cnt);
      rc = mgk::output_box((lb.node == 0 && whole) ? whole : H->nodes[lb.node].vars, lo, n, NG, constant_K, dbuf);
      if (rc != MGIC_OK) break;
      if (cudaMemcpyAsync(hbuf.data(), dbuf, cnt * sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
          cudaStreamSynchronize(c->stream) != cudaSuccess) { mgic_set_error("copy failed in the checkpoint writer"); rc = MGIC_ERR_CUDA; break; }
      if (fwrite(hbuf.data(), sizeof(double), cnt, fp) != cnt) { mgic_set_error("short write to %s", path); rc = MGIC_ERR_ARG; break; }
    }
  mgic_dev_free(dbuf);
  fclose(fp);
  return rc;
}

extern "C" int mgic_hier_nl_solve(mgic_hier *H, double *dpsi_norms, int max_out, int *nl_iterations) {
  MGIC_REQUIRE(H, "NULL argument");
  MGIC_TRY(mgic_hier_set_initial_conditions(H));                            // :93
  int its = 0;
  for (int NL_iter = 0; NL_iter < H->P.max_NL_iterations; NL_iter++) {      // :131
    double nrm = 0.0;
    MGIC_TRY(mgic_hier_nl_iteration(H, &nrm, nullptr, nullptr));
    if (dpsi_norms && NL_iter < max_out) dpsi_norms[NL_iter] = nrm;
    its = NL_iter + 1;
    if (nrm < H->P.tolerance || nrm > 1e5) break;                           // :212
  }
  if (nl_iterations) *nl_iterations = its;
  return MGIC_OK;
}

// The NL loop of Main_PoissonSolver.cpp:131-216 for one AMR level, device resident.
extern "C" int mgic_nl_solve(mgic_ctx *c, const mgic_params *P, double *dpsi_norms, int max_out, int *nl_iterations,
                             double *psi_out) {
  MGIC_REQUIRE(c && P, "NULL argument");
  MGIC_REQUIRE(c->nranks == 1, "mgic_nl_solve drives a single GPU; multi-rank callers compose the pieces per rank");
  MGIC_REQUIRE(P->max_level == 0, "mgic_nl_solve: single AMR level (max_level = 0); hierarchies go through mgic_hier_nl_solve");
  // Main_PoissonSolver.cpp:137-150 recomputes constant_K from set_constant_K_integrand + computeSum every iteration on periodic
  // domains; that branch is out of scope -- refuse instead of silently solving the singular K = 0 problem
  MGIC_REQUIRE(!P->is_periodic, "mgic_nl_solve: the periodic constant-K branch (Main_PoissonSolver.cpp:137-150) is not implemented");
  mgic_vars *vars = nullptr;
  mgic_op *lay = nullptr;
  mgic_field *dpsi = nullptr, *rhs = nullptr, *aC = nullptr, *bC = nullptr;
  mgic_mg *mg = nullptr;
  int rc = MGIC_OK, its = 0;
  int bclo[3], bchi[3];
  for (int d = 0; d < 3; d++) {
    bclo[d] = P->is_periodic ? MGIC_BC_PERIODIC : P->bc_lo[d];
    bchi[d] = P->is_periodic ? MGIC_BC_PERIODIC : P->bc_hi[d];
  }
#define NL_TRY(x) do { rc = (x); if (rc != MGIC_OK) goto done; } while (0)
  NL_TRY(mgic_vars_create(c, P, 0, P->N[2], &vars));
  NL_TRY(mgic_op_create(c, P->N, 0, P->N[2], P->L / P->N[0], P->alpha, P->beta, bclo, bchi, P->bc_value, &lay));
  NL_TRY(mgic_field_create(lay, &dpsi));
  NL_TRY(mgic_field_create(lay, &rhs));
  NL_TRY(mgic_field_create(lay, &aC));
  NL_TRY(mgic_field_create(lay, &bC));
  NL_TRY(mgic_set_initial_conditions(vars, dpsi));                       // Main:93
  for (int NL_iter = 0; NL_iter < P->max_NL_iterations; NL_iter++) {     // :131
    NL_TRY(mgic_set_rhs_and_a_coef(vars, rhs, aC, 0.0));                 // :154-160 (non-periodic: constant_K = 0)
    NL_TRY(mgic_set_b_coef(vars, bC));
    // :163-170 rebuilds factory, MultilevelLinearOp and every MG operator per iteration; here the objects survive and the
    // coarsened coefficients / lambda are re-derived in place from the same coefficient arrays (identical operators, no
    // re-allocation, the V-cycle graph stays valid)
    if (!mg) NL_TRY(mgic_mg_create(c, P, aC, bC, &mg));
    else NL_TRY(mgic_mg_refresh_coefs(mg));
    int it = 0, st = 0;
    NL_TRY(mgic_mg_outer_solve(mg, dpsi, rhs, &it, &st, nullptr, 0));    // :184
    double nrm = 0.0;
    NL_TRY(mgic_update_psi0(vars, mg->ops[0], dpsi, &nrm));              // :189-208
    if (dpsi_norms && NL_iter < max_out) dpsi_norms[NL_iter] = nrm;
    its = NL_iter + 1;
    if (nrm < P->tolerance || nrm > 1e5) break;                          // :212
  }
  if (psi_out) NL_TRY(mgic_vars_download(vars, 0, psi_out));
done:
#undef NL_TRY
  if (nl_iterations) *nl_iterations = its;
  mgic_mg_destroy(mg);
  mgic_field_destroy(dpsi); mgic_field_destroy(rhs); mgic_field_destroy(aC); mgic_field_destroy(bC);
  mgic_op_destroy(lay);
  mgic_vars_destroy(vars);
  return rc;
}
