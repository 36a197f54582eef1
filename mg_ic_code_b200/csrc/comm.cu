// comm.cu -- z-slab halo exchange and scalar all-reduce over NCCL (include/mgic_comm.h).
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

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

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  long long haloBytes = 0;
};

#define NCCL_TRY(call)                                                                   \
  do {                                                                                   \
    ncclResult_t r__ = (call);                                                           \
    if (r__ != ncclSuccess) {                                                            \
      mgic_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, A->GetErrorString(r__)); \
      return MGIC_ERR_CUDA;                                                              \
    }                                                                                    \
  } while (0)

int halo_hook(mgic_ctx *c, mgic_field *f, int planes) { return mgic_comm_halo_exchange(c, f, planes); }

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
  return MGIC_OK;
}

extern "C" int mgic_comm_destroy(mgic_ctx *c) {
  if (!c || !c->comm) return MGIC_OK;
  NcclApi *A = api();
  Comm *cm = (Comm *)c->comm;
  cudaStreamSynchronize(c->stream);
  if (A && cm->comm) A->CommDestroy(cm->comm);
  delete cm;
  c->comm = nullptr; c->halo_exchange = nullptr; c->allreduce = nullptr; c->allgather = nullptr;
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
  const bool hasLo = f->k0 > 0, hasHi = f->k0 + f->nz < f->gnz;
  const size_t cnt = (size_t)planes * f->sz;
  cudaStream_t st = c->haloStream ? c->haloStream : c->stream;
  NCCL_TRY(A->GroupStart());
  if (hasLo) {
    NCCL_TRY(A->Send(f->p, cnt, ncclDouble, c->rank - 1, cm->comm, st));
    NCCL_TRY(A->Recv(f->p - (long long)planes * f->sz, cnt, ncclDouble, c->rank - 1, cm->comm, st));
    cm->haloBytes += (long long)cnt * 8;
  }
  if (hasHi) {
    NCCL_TRY(A->Send(f->p + (long long)(f->nz - planes) * f->sz, cnt, ncclDouble, c->rank + 1, cm->comm, st));
    NCCL_TRY(A->Recv(f->p + (long long)f->nz * f->sz, cnt, ncclDouble, c->rank + 1, cm->comm, st));
    cm->haloBytes += (long long)cnt * 8;
  }
  NCCL_TRY(A->GroupEnd());
  return MGIC_OK;
}

extern "C" long long mgic_comm_halo_bytes(mgic_ctx *c) { return (c && c->comm) ? ((Comm *)c->comm)->haloBytes : 0; }
