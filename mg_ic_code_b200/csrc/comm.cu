// comm.cu -- z-slab halo exchange and scalar all-reduce (include/mgic_comm.h).
//
// Scalars and the coarse-level gather go through NCCL.  Halo planes go through NVLink peer memory: every array that is
// exchanged is exported once with a CUDA IPC handle and mapped by its two z-neighbours; one kernel per exchange
// (k_halo_push) then STORES the rank's boundary planes straight into the neighbours' ghost planes and synchronises the
// three ranks with epoch flags that live in peer memory too -- no host round trip, capturable into the V-cycle graph.
// Measured motivation (profiles/, 8 GPUs, 1024^2 planes): the grouped ncclSend/ncclRecv pair moved 160 GB/s per rank
// and cost 1.6 ms of a 9.1 ms V-cycle.  ncclSend/ncclRecv remains the path for arrays that could not be mapped.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <unordered_map>

#include "mgic_comm.h"
#include "mgic_internal.h"

namespace {

struct NcclApi {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *);
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  const char *(*GetErrorString)(ncclResult_t);
};

NcclApi *api() {
  static NcclApi a;
  static bool tried = false;
  if (tried) return a.h ? &a : nullptr;
  tried = true;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    a.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (a.h) break;
  }
  if (!a.h) return nullptr;
#define SYM(field, name)                                             \
  *(void **)(&a.field) = dlsym(a.h, name);                           \
  if (!a.field) { a.h = nullptr; return nullptr; }
  SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
  SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd")
  SYM(AllReduce, "ncclAllReduce") SYM(AllGather, "ncclAllGather") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  return &a;
}

// ---- peer-memory halo exchange ---------------------------------------------------------------------------------------
typedef unsigned long long u64;
// control words of a rank (one IPC-exported block, written by the neighbours):
enum { CTL_FREE_LO = 0, CTL_FREE_HI = 1,   // "my ghost planes may be overwritten", posted by the lo / hi neighbour
       CTL_DATA_LO = 2, CTL_DATA_HI = 3,   // "your ghost planes are filled", posted by the lo / hi neighbour
       CTL_EPOCH = 4, CTL_COUNT = 5, CTL_WORDS = 8 };
// words 8.. belong to the fused sweep's own protocol (MGIC_SW_* in mgic_internal.h): the sweep kernel stores its boundary
// planes into the neighbours' ghost planes itself
constexpr size_t IPC_GRAIN = (size_t)2 << 20;  // cudaMalloc gives allocations of >= 2 MiB a block of their own

struct PeerArr {
  bool ok = false;
  double *lo = nullptr, *hi = nullptr;  // the z-neighbours' copies of this array (their `base`), mapped here
  int nzLo = 0, nzHi = 0;               // their slab thickness
};

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  long long haloBytes = 0;
  // peer-memory path
  bool p2p = false;
  u64 *ctl = nullptr, *ctlLo = nullptr, *ctlHi = nullptr;
  unsigned char *d_stage = nullptr;      // handle exchange staging (device), nranks records
  std::unordered_map<const void *, PeerArr> reg;
  std::vector<void *> graveyard;         // exported arrays whose cudaFree waits until the neighbours have unmapped them
  long long p2pExchanges = 0, ncclExchanges = 0;
};

// z-neighbours of a field's slab: the ranks below / above, on a ring when the level is periodic in z
inline bool has_lo(const mgic_field *f) { return f->zWrap || f->k0 > 0; }
inline bool has_hi(const mgic_field *f) { return f->zWrap || f->k0 + f->nz < f->gnz; }

struct PushArgs {
  const double *srcLo, *srcHi;  // this rank's lowest / highest `planes` valid planes
  double *dstLo, *dstHi;        // the lo neighbour's upper ghost planes, the hi neighbour's lower ghost planes (null: none)
  size_t n;                     // doubles per side
  int vec;                      // 1: n is even and all four pointers are 16-byte aligned
  u64 *ctl, *ctlLo, *ctlHi;
};

__device__ __forceinline__ void st_release_sys(u64 *p, u64 v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ u64 ld_acquire_sys(const u64 *p) {
  u64 v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ u64 now_ns() {
  u64 t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// spin until *p >= v; a neighbour that died must not leave this GPU spinning forever: trap after 60 s
__device__ __forceinline__ void wait_ge(const u64 *p, u64 v) {
  if (ld_acquire_sys(p) >= v) return;
  const u64 t0 = now_ns();
  while (ld_acquire_sys(p) < v) {
    __nanosleep(100);
    if (now_ns() - t0 > 60000000000ull) { printf("mgic: halo exchange timed out waiting for a neighbour rank\n"); __trap(); }
  }
}

// One halo exchange of one array.  Protocol per epoch E (the epoch counter lives in device memory, so a replayed graph
// keeps counting): (1) tell both neighbours "my ghost planes are free" -- everything this rank launched before this
// kernel, i.e. every reader of the previous ghost values, has completed by stream order; (2) wait for the neighbours'
// "free"; (3) store the boundary planes into their ghost planes; (4) the last block to finish publishes "data E" to
// both neighbours and waits for theirs, so that when the kernel ends this rank's ghost planes are valid.
// No block waits for another block of the same grid, so the grid needs no co-residency.
__global__ void __launch_bounds__(512, 2) k_halo_push(PushArgs A) {
  const u64 ep = *(volatile u64 *)&A.ctl[CTL_EPOCH] + 1;
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) {
      if (A.ctlLo) st_release_sys(&A.ctlLo[CTL_FREE_HI], ep);  // I am my lo neighbour's hi neighbour
      if (A.ctlHi) st_release_sys(&A.ctlHi[CTL_FREE_LO], ep);
    }
    if (A.ctlLo) wait_ge(&A.ctl[CTL_FREE_LO], ep);
    if (A.ctlHi) wait_ge(&A.ctl[CTL_FREE_HI], ep);
  }
  __syncthreads();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int side = 0; side < 2; side++) {
    double *dstS = side ? A.dstHi : A.dstLo;
    const double *srcS = side ? A.srcHi : A.srcLo;
    if (!dstS) continue;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (A.vec) {
      double2 *dst = reinterpret_cast<double2 *>(dstS);
      const double2 *src = reinterpret_cast<const double2 *>(srcS);
      const size_t n2 = A.n / 2;
      for (; i + 3 * stride < n2; i += 4 * stride) {  // four independent 16-byte loads in flight per thread
        const double2 v0 = src[i], v1 = src[i + stride], v2 = src[i + 2 * stride], v3 = src[i + 3 * stride];
        dst[i] = v0; dst[i + stride] = v1; dst[i + 2 * stride] = v2; dst[i + 3 * stride] = v3;
      }
      for (; i < n2; i += stride) dst[i] = src[i];
    } else {
      for (; i < A.n; i += stride) dstS[i] = srcS[i];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const u64 done = atomicAdd(&A.ctl[CTL_COUNT], 1ull);
    if (done == gridDim.x - 1) {
      A.ctl[CTL_COUNT] = 0;
      __threadfence_system();
      if (A.ctlLo) st_release_sys(&A.ctlLo[CTL_DATA_HI], ep);
      if (A.ctlHi) st_release_sys(&A.ctlHi[CTL_DATA_LO], ep);
      if (A.ctlLo) wait_ge(&A.ctl[CTL_DATA_LO], ep);
      if (A.ctlHi) wait_ge(&A.ctl[CTL_DATA_HI], ep);
      A.ctl[CTL_EPOCH] = ep;
      __threadfence();
    }
  }
}

#define NCCL_TRY(call)                                                                   \
  do {                                                                                   \
    ncclResult_t r__ = (call);                                                           \
    if (r__ != ncclSuccess) {                                                            \
      mgic_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, A->GetErrorString(r__)); \
      return MGIC_ERR_CUDA;                                                              \
    }                                                                                    \
  } while (0)

int halo_hook(mgic_ctx *c, mgic_field *f, int planes) { return mgic_comm_halo_exchange(c, f, planes); }

// the fused sweep pushes its boundary planes itself (gsrb_fused.cu): where do they go, and which control words synchronise it
int sweep_peers_hook(mgic_ctx *c, const mgic_field *like, const double *outBase, SweepPeers *sp) {
  memset(sp, 0, sizeof(*sp));
  Comm *cm = (Comm *)c->comm;
  if (!cm || !cm->p2p || !c->p2pHalo) return MGIC_OK;
  auto it = cm->reg.find(outBase);
  if (it == cm->reg.end() || !it->second.ok) return MGIC_OK;
  const PeerArr &pa = it->second;
  const bool hasLo = has_lo(like), hasHi = has_hi(like);
  // my plane k (k = 0, 1) is the lo neighbour's upper ghost plane k; my plane nz-2+k is the hi neighbour's lower ghost plane k:
  // both as "address of MY plane 0" in the neighbour's array, so that the kernel adds the same offset it uses for its own store
  sp->peerLo = hasLo ? pa.lo + (long long)(MGIC_GZ + pa.nzLo) * like->sz : nullptr;
  sp->peerHi = hasHi ? pa.hi - (long long)(like->nz - MGIC_GZ) * like->sz : nullptr;
  sp->ctl = cm->ctl; sp->ctlLo = hasLo ? cm->ctlLo : nullptr; sp->ctlHi = hasHi ? cm->ctlHi : nullptr;
  sp->ok = 1;
  return MGIC_OK;
}

// a reader that is not a fused sweep (residual + restrict ...) of an array whose ghost planes the neighbours' last sweep
// filled: wait until both neighbours have published that sweep
__global__ void k_sweep_wait(u64 *ctl, int hasLo, int hasHi) {
  const u64 ep = *(volatile u64 *)&ctl[MGIC_SW_EPOCH];
  if (hasLo) wait_ge(&ctl[MGIC_SW_DATA_LO], ep);
  if (hasHi) wait_ge(&ctl[MGIC_SW_DATA_HI], ep);
}
int sweep_wait_hook(mgic_ctx *c, const mgic_field *like) {
  Comm *cm = (Comm *)c->comm;
  if (!cm || !cm->p2p) { mgic_set_error("sweep_wait without peer mapping"); return MGIC_ERR_STATE; }
  k_sweep_wait<<<1, 1, 0, c->haloStream ? c->haloStream : c->stream>>>(cm->ctl, has_lo(like), has_hi(like));
  MGIC_CUDA(cudaGetLastError());
  c->launches++;
  return MGIC_OK;
}

int allreduce_hook(mgic_ctx *c, double *dev, int n, int op) {
  NcclApi *A = api();
  Comm *cm = (Comm *)c->comm;
  if (!A || !cm) { mgic_set_error("NCCL communicator not initialised"); return MGIC_ERR_STATE; }
  NCCL_TRY(A->AllReduce(dev, dev, (size_t)n, ncclDouble, op == 1 ? ncclMax : ncclSum, cm->comm, c->stream));
  return MGIC_OK;
}

int allgather_hook(mgic_ctx *c, const double *send, double *recv, size_t count) {
  NcclApi *A = api();
  Comm *cm = (Comm *)c->comm;
  if (!A || !cm) { mgic_set_error("NCCL communicator not initialised"); return MGIC_ERR_STATE; }
  NCCL_TRY(A->AllGather(send, recv, count, ncclDouble, cm->comm, c->stream));
  return MGIC_OK;
}

// ---- IPC bookkeeping (host) ----------------------------------------------------------------------------------------
struct IpcRec {
  cudaIpcMemHandle_t h;
  int ok, nz;
  long long sz;
};

// every rank contributes one record and gets everybody's (collective, synchronous: never call while capturing)
int gather_recs(mgic_ctx *c, Comm *cm, NcclApi *A, const IpcRec &mine, std::vector<IpcRec> &all) {
  const size_t R = sizeof(IpcRec);
  unsigned char *send = cm->d_stage + (size_t)cm->nranks * R;
  MGIC_CUDA(cudaMemcpyAsync(send, &mine, R, cudaMemcpyHostToDevice, c->stream));
  NCCL_TRY(A->AllGather(send, cm->d_stage, R, ncclChar, cm->comm, c->stream));
  all.resize(cm->nranks);
  MGIC_CUDA(cudaMemcpyAsync(all.data(), cm->d_stage, (size_t)cm->nranks * R, cudaMemcpyDeviceToHost, c->stream));
  MGIC_CUDA(cudaStreamSynchronize(c->stream));
  return MGIC_OK;
}

// is `base` the start of its own cudaMalloc block?  (an IPC handle names the whole block, so a sub-allocated pointer
// would map at the wrong address on the other side)
bool owns_its_block(const void *base) {
  typedef int (*Fn)(unsigned long long *, size_t *, unsigned long long);
  static Fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *q = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &q, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      fn = (Fn)q;
    cudaGetLastError();
  }
  if (!fn) return false;
  unsigned long long b = 0;
  size_t sz = 0;
  if (fn(&b, &sz, (unsigned long long)base) != 0) return false;
  return b == (unsigned long long)base;
}

// export `base`, map the neighbours' counterparts; all ranks agree on the outcome.  lo/hi = mapped pointers (or null)
int map_neighbours(mgic_ctx *c, Comm *cm, NcclApi *A, void *base, int nz, long long sz, bool hasLo, bool hasHi, bool want,
                   void **lo, void **hi, int *nzLo, int *nzHi, bool *ok) {
  IpcRec mine;
  memset(&mine, 0, sizeof(mine));
  mine.nz = nz; mine.sz = sz;
  mine.ok = want && owns_its_block(base) && cudaIpcGetMemHandle(&mine.h, base) == cudaSuccess;
  cudaGetLastError();
  std::vector<IpcRec> all;
  MGIC_TRY(gather_recs(c, cm, A, mine, all));
  bool allOk = true;
  for (const IpcRec &r : all) allOk = allOk && r.ok;
  *lo = *hi = nullptr;
  int opened = allOk ? 1 : 0;
  const int loRank = (cm->rank - 1 + cm->nranks) % cm->nranks, hiRank = (cm->rank + 1) % cm->nranks;   // a ring when z is periodic
  if (allOk && hasLo) {
    const IpcRec &r = all[loRank];
    if (r.sz != sz || cudaIpcOpenMemHandle(lo, r.h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { opened = 0; *lo = nullptr; }
    *nzLo = r.nz;
  }
  if (allOk && hasHi) {
    const IpcRec &r = all[hiRank];
    if (hasLo && hiRank == loRank) *hi = *lo;   // two ranks on a ring: one peer, one mapping
    else if (r.sz != sz || cudaIpcOpenMemHandle(hi, r.h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { opened = 0; *hi = nullptr; }
    *nzHi = r.nz;
  }
  cudaGetLastError();
  // second round: did every rank manage to map?
  IpcRec res;
  memset(&res, 0, sizeof(res));
  res.ok = opened;
  MGIC_TRY(gather_recs(c, cm, A, res, all));
  bool good = true;
  for (const IpcRec &r : all) good = good && r.ok;
  if (!good) {
    if (*lo) cudaIpcCloseMemHandle(*lo);
    if (*hi && *hi != *lo) cudaIpcCloseMemHandle(*hi);
    *lo = *hi = nullptr;
    cudaGetLastError();
  }
  *ok = good;
  return MGIC_OK;
}

int p2p_setup(mgic_ctx *c, Comm *cm, NcclApi *A) {
  if (cm->nranks < 2) return MGIC_OK;
  const char *env = getenv("MGIC_P2P_HALO");
  const bool want = !(env && atoi(env) == 0);
  MGIC_CUDA(cudaMalloc(&cm->d_stage, (size_t)(cm->nranks + 1) * sizeof(IpcRec)));
  MGIC_CUDA(cudaMalloc(&cm->ctl, IPC_GRAIN));
  MGIC_CUDA(cudaMemset(cm->ctl, 0, IPC_GRAIN));
  void *lo = nullptr, *hi = nullptr;
  int a = 0, b = 0;
  bool ok = false;
  // both ring neighbours (rank 0's lo neighbour is the last rank): levels that are periodic in z exchange around the ring
  MGIC_TRY(map_neighbours(c, cm, A, cm->ctl, 0, 0, true, true, want, &lo, &hi, &a, &b, &ok));
  cm->p2p = ok;
  cm->ctlLo = (u64 *)lo; cm->ctlHi = (u64 *)hi;
  if (!ok && want && cm->rank == 0)
    fprintf(stderr, "mgic: CUDA IPC peer mapping unavailable; halo planes go through ncclSend/ncclRecv\n");
  return MGIC_OK;
}

// collective: called by every rank for the same array (same call sequence on every rank)
int register_array(mgic_ctx *c, Comm *cm, NcclApi *A, mgic_field *f) {
  const size_t buried = cm->graveyard.size();
  const bool hasLo = has_lo(f), hasHi = has_hi(f);
  PeerArr pa;
  void *lo = nullptr, *hi = nullptr;
  MGIC_TRY(map_neighbours(c, cm, A, f->base, f->nz, f->sz, hasLo, hasHi, true, &lo, &hi, &pa.nzLo, &pa.nzHi, &pa.ok));
  pa.lo = (double *)lo; pa.hi = (double *)hi;
  cm->reg[f->base] = pa;
  // every rank has passed the destroy calls that precede this registration in program order, i.e. has unmapped the
  // arrays buried before it: their memory can go now
  for (size_t i = 0; i < buried; i++) mgic_dev_free(cm->graveyard[i]);
  cm->graveyard.erase(cm->graveyard.begin(), cm->graveyard.begin() + buried);
  return MGIC_OK;
}

int prepare_hook(mgic_ctx *c, mgic_field *f) {
  NcclApi *A = api();
  Comm *cm = (Comm *)c->comm;
  if (!A || !cm || !cm->p2p || !c->p2pHalo || f->noHalo || cm->reg.count(f->base)) return MGIC_OK;
  return register_array(c, cm, A, f);
}

// mgic_field_destroy of an exported array: unmap the neighbours' counterparts now, free later (the CUDA IPC contract:
// the exporter must not cudaFree while an importer still has the block mapped).  Returns 1 if the free is deferred.
int release_hook(mgic_ctx *c, void *base) {
  Comm *cm = (Comm *)c->comm;
  if (!cm) return 0;
  auto it = cm->reg.find(base);
  if (it == cm->reg.end()) return 0;
  const PeerArr pa = it->second;
  cm->reg.erase(it);
  if (!pa.ok) return 0;
  cudaStreamSynchronize(c->stream);
  if (c->commStream) cudaStreamSynchronize(c->commStream);
  if (pa.lo) cudaIpcCloseMemHandle(pa.lo);
  if (pa.hi && pa.hi != pa.lo) cudaIpcCloseMemHandle(pa.hi);
  cudaGetLastError();
  cm->graveyard.push_back(base);
  return 1;
}

}  // namespace

extern "C" int mgic_comm_unique_id(unsigned char id[MGIC_NCCL_ID_BYTES]) {
  NcclApi *A = api();
  if (!A) { mgic_set_error("libnccl.so.2 could not be loaded: %s", dlerror()); return MGIC_ERR_STATE; }
  static_assert(sizeof(ncclUniqueId) == MGIC_NCCL_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId u;
  NCCL_TRY(A->GetUniqueId(&u));
  memcpy(id, &u, MGIC_NCCL_ID_BYTES);
  return MGIC_OK;
}

extern "C" int mgic_comm_init(mgic_ctx *c, const unsigned char id[MGIC_NCCL_ID_BYTES], int rank, int nranks) {
  MGIC_REQUIRE(c && id && nranks >= 1 && rank >= 0 && rank < nranks, "bad argument");
  NcclApi *A = api();
  if (!A) { mgic_set_error("libnccl.so.2 could not be loaded: %s", dlerror()); return MGIC_ERR_STATE; }
  MGIC_CUDA(cudaSetDevice(c->device));
  Comm *cm = new Comm;
  cm->rank = rank; cm->nranks = nranks;
  ncclUniqueId u;
  memcpy(&u, id, MGIC_NCCL_ID_BYTES);
  NCCL_TRY(A->CommInitRank(&cm->comm, nranks, u, rank));
  c->comm = cm;
  c->rank = rank; c->nranks = nranks;
  c->halo_exchange = halo_hook;
  c->allreduce = allreduce_hook;
  c->allgather = allgather_hook;
  if (!c->commStream) {
    // highest priority: the exchange kernels must get an SM slot as soon as one CTA of the concurrently running sweep
    // retires, instead of queueing behind the rest of its grid
    int prLo = 0, prHi = 0;
    MGIC_CUDA(cudaDeviceGetStreamPriorityRange(&prLo, &prHi));
    MGIC_CUDA(cudaStreamCreateWithPriority(&c->commStream, cudaStreamNonBlocking, prHi));
    MGIC_CUDA(cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming));
    MGIC_CUDA(cudaEventCreateWithFlags(&c->evJoin, cudaEventDisableTiming));
  }
  MGIC_TRY(p2p_setup(c, cm, A));
  c->array_release = release_hook;
  c->array_prepare = prepare_hook;
  c->sweep_peers = sweep_peers_hook;
  c->sweep_wait = sweep_wait_hook;
  return MGIC_OK;
}

extern "C" int mgic_comm_destroy(mgic_ctx *c) {
  if (!c || !c->comm) return MGIC_OK;
  NcclApi *A = api();
  Comm *cm = (Comm *)c->comm;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->commStream) cudaStreamSynchronize(c->commStream);
  // unmap everything this rank imported; once every rank has done so (barrier) the exported blocks may be freed
  for (auto &kv : cm->reg) {
    if (kv.second.lo) cudaIpcCloseMemHandle(kv.second.lo);
    if (kv.second.hi && kv.second.hi != kv.second.lo) cudaIpcCloseMemHandle(kv.second.hi);
  }
  cm->reg.clear();
  if (cm->ctlLo) cudaIpcCloseMemHandle(cm->ctlLo);
  if (cm->ctlHi && cm->ctlHi != cm->ctlLo) cudaIpcCloseMemHandle(cm->ctlHi);
  cudaGetLastError();
  if (A && cm->comm && cm->d_stage) {
    A->AllReduce(cm->d_stage, cm->d_stage, 1, ncclChar, ncclSum, cm->comm, c->stream);
    cudaStreamSynchronize(c->stream);
  }
  for (void *q : cm->graveyard) mgic_dev_free(q);
  cudaFree(cm->ctl);
  cudaFree(cm->d_stage);
  if (A && cm->comm) A->CommDestroy(cm->comm);
  delete cm;
  c->sweep_peers = nullptr; c->sweep_wait = nullptr;
  c->comm = nullptr; c->halo_exchange = nullptr; c->allreduce = nullptr; c->allgather = nullptr; c->array_release = nullptr; c->array_prepare = nullptr;
  return MGIC_OK;
}

// Rank r sends its lowest `planes` valid planes to r-1's upper ghost planes and its highest ones to r+1's lower ghost
// planes (LevelData::exchange restricted to the z faces; x/y neighbours are in the same array).
extern "C" int mgic_comm_halo_exchange(mgic_ctx *c, mgic_field *f, int planes) {
  MGIC_REQUIRE(c && f, "NULL argument");
  MGIC_REQUIRE(planes >= 1 && planes <= MGIC_GZ && planes <= f->nz, "halo depth must be 1..MGIC_GZ and fit the slab");
  if (c->nranks == 1) return MGIC_OK;
  NcclApi *A = api();
  Comm *cm = (Comm *)c->comm;
  if (!A || !cm) { mgic_set_error("NCCL communicator not initialised"); return MGIC_ERR_STATE; }
  const bool hasLo = has_lo(f), hasHi = has_hi(f);
  const int loRank = (c->rank - 1 + c->nranks) % c->nranks, hiRank = (c->rank + 1) % c->nranks;
  const size_t cnt = (size_t)planes * f->sz;
  cudaStream_t st = c->haloStream ? c->haloStream : c->stream;
  if (cm->p2p && c->p2pHalo) {
    auto it = cm->reg.find(f->base);
    if (it == cm->reg.end()) {
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      MGIC_CUDA(cudaStreamIsCapturing(c->stream, &cs));
      if (cs == cudaStreamCaptureStatusNone) {  // first exchange of this array: export / map it (collective)
        MGIC_TRY(register_array(c, cm, A, f));
        it = cm->reg.find(f->base);
      }
    }
    if (it != cm->reg.end() && it->second.ok) {
      const PeerArr &pa = it->second;
      PushArgs P;
      P.srcLo = f->p;
      P.srcHi = f->p + (long long)(f->nz - planes) * f->sz;
      P.dstLo = hasLo ? pa.lo + (long long)(MGIC_GZ + pa.nzLo) * f->sz : nullptr;   // its upper ghost planes
      P.dstHi = hasHi ? pa.hi + (long long)(MGIC_GZ - planes) * f->sz : nullptr;    // its lower ghost planes
      P.n = cnt;
      P.vec = (f->sz % 2 == 0) ? 1 : 0;
      P.ctl = cm->ctl; P.ctlLo = hasLo ? cm->ctlLo : nullptr; P.ctlHi = hasHi ? cm->ctlHi : nullptr;
      const size_t per = 512 * 8;  // doubles per block and trip
      static const int maxBlocks = [] { const char *e = getenv("MGIC_P2P_BLOCKS"); return e ? atoi(e) : 0; }();  // tuning
      const size_t cap = maxBlocks > 0 ? (size_t)maxBlocks : (size_t)2 * c->numSMs;
      const int blocks = (int)std::min<size_t>(cap, std::max<size_t>(1, (cnt + per - 1) / per));
      k_halo_push<<<blocks, 512, 0, st>>>(P);
      MGIC_CUDA(cudaGetLastError());
      c->launches++;
      cm->haloBytes += (long long)cnt * 8 * ((hasLo ? 1 : 0) + (hasHi ? 1 : 0));
      cm->p2pExchanges++;
      return MGIC_OK;
    }
  }
  cm->ncclExchanges++;
  NCCL_TRY(A->GroupStart());
  // issue order matters when both neighbours are the same rank (two ranks on a ring): NCCL pairs the sends and receives of
  // one peer in order, and my low planes must land in its UPPER ghost planes
  if (hasLo) { NCCL_TRY(A->Send(f->p, cnt, ncclDouble, loRank, cm->comm, st)); cm->haloBytes += (long long)cnt * 8; }
  if (hasHi) NCCL_TRY(A->Recv(f->p + (long long)f->nz * f->sz, cnt, ncclDouble, hiRank, cm->comm, st));
  if (hasHi) { NCCL_TRY(A->Send(f->p + (long long)(f->nz - planes) * f->sz, cnt, ncclDouble, hiRank, cm->comm, st)); cm->haloBytes += (long long)cnt * 8; }
  if (hasLo) NCCL_TRY(A->Recv(f->p - (long long)planes * f->sz, cnt, ncclDouble, loRank, cm->comm, st));
  NCCL_TRY(A->GroupEnd());
  return MGIC_OK;
}

extern "C" long long mgic_comm_halo_bytes(mgic_ctx *c) { return (c && c->comm) ? ((Comm *)c->comm)->haloBytes : 0; }

extern "C" int mgic_comm_halo_stats(mgic_ctx *c, long long *p2p_exchanges, long long *nccl_exchanges, int *p2p_available) {
  MGIC_REQUIRE(c, "ctx is NULL");
  Comm *cm = (Comm *)c->comm;
  if (p2p_exchanges) *p2p_exchanges = cm ? cm->p2pExchanges : 0;
  if (nccl_exchanges) *nccl_exchanges = cm ? cm->ncclExchanges : 0;
  if (p2p_available) *p2p_available = (cm && cm->p2p) ? 1 : 0;
  return MGIC_OK;
}
