"""Multi-GPU plumbing (one process per GPU): slab partition + NCCL communicator bootstrap over torch.distributed.

torch.distributed is used only to broadcast the 128-byte NCCL unique id; halo planes then go through the library's own
peer-store kernel over CUDA-IPC mappings (or ncclSend / ncclRecv) and scalars through ncclAllReduce, all on the context
stream (include/mgic_comm.h)."""
import ctypes as C

from ._capi import MgicError, check, lib


def slab_partition(nz, nranks, multiple=1):
    """Split nz planes into `nranks` contiguous slabs whose sizes are multiples of `multiple` (max_grid_size, so that
    every MG depth the box size allows stays coarsenable on every rank, Factory.cpp:168-172).  Returns [(k0, nz_local)]."""
    if nz % multiple:
        raise MgicError(f"nz={nz} is not a multiple of {multiple}")
    units = nz // multiple
    if units < nranks:
        raise MgicError(f"cannot cut {nz} planes into {nranks} slabs of multiples of {multiple}")
    base, rem = divmod(units, nranks)
    out, k0 = [], 0
    for r in range(nranks):
        n = (base + (1 if r < rem else 0)) * multiple
        out.append((k0, n))
        k0 += n
    return out


def attach(ctx, dist):
    """Join the NCCL communicator of the job: rank 0 creates the id, torch.distributed broadcasts it."""
    import torch
    L = lib()
    L.mgic_comm_unique_id.argtypes = [C.c_void_p]
    L.mgic_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.mgic_comm_destroy.argtypes = [C.c_void_p]
    L.mgic_comm_halo_exchange.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.mgic_comm_halo_bytes.argtypes = [C.c_void_p]
    L.mgic_comm_halo_bytes.restype = C.c_longlong
    rank, world = dist.get_rank(), dist.get_world_size()
    buf = (C.c_ubyte * 128)()
    if rank == 0:
        check(L.mgic_comm_unique_id(buf))
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
    dist.broadcast(t, 0)
    ids = (C.c_ubyte * 128)(*t.cpu().tolist())
    check(L.mgic_comm_init(ctx.h, ids, rank, world))
    ctx.rank, ctx.nranks = rank, world
    return ctx


def halo_bytes(ctx):
    return lib().mgic_comm_halo_bytes(ctx.h)


def halo_stats(ctx):
    """(exchanges by NVLink peer stores, exchanges by ncclSend/ncclRecv, peer mapping available) since attach()"""
    L = lib()
    a, b, ok = C.c_longlong(0), C.c_longlong(0), C.c_int(0)
    L.mgic_comm_halo_stats.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_int)]
    check(L.mgic_comm_halo_stats(ctx.h, C.byref(a), C.byref(b), C.byref(ok)))
    return a.value, b.value, bool(ok.value)


def detach(ctx):
    lib().mgic_comm_destroy(ctx.h)
