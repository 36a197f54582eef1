"""world_size-2 gloo test (CPU) of the N>1 host-side logic: the z-slab partition (mg_ic_code_b200.comm.slab_partition)
and the decomposition invariants the multi-GPU path relies on -- global-index colouring (k0 offset), one halo plane per
colour pass for the per-colour smoother, TWO planes per sweep for the fused smoother (the neighbour's first plane is
updated redundantly), local restriction on even slabs.  The numpy twin plays the kernels; gloo send/recv plays NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import np_twin as T
from mg_ic_code_b200.comm import slab_partition


def test_slab_partition():
    assert slab_partition(512, 8, 32) == [(64 * r, 64) for r in range(8)]
    assert slab_partition(96, 2, 32) == [(0, 64), (64, 32)]
    assert sum(n for _, n in slab_partition(4096, 8, 32)) == 4096
    with pytest.raises(Exception):
        slab_partition(64, 4, 32)
    with pytest.raises(Exception):
        slab_partition(100, 2, 32)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _exchange(rank, world, slab, depth):
    """slab: array with `depth` ghost planes on both sides; fills them from the z-neighbours."""
    reqs = []
    if rank > 0:
        reqs.append(dist.isend(torch.from_numpy(slab[depth:2 * depth].copy()), rank - 1))
    if rank < world - 1:
        reqs.append(dist.isend(torch.from_numpy(slab[-2 * depth:-depth].copy()), rank + 1))
    if rank > 0:
        t = torch.empty(slab[:depth].shape, dtype=torch.float64)
        dist.recv(t, rank - 1)
        slab[:depth] = t.numpy()
    if rank < world - 1:
        t = torch.empty(slab[-depth:].shape, dtype=torch.float64)
        dist.recv(t, rank + 1)
        slab[-depth:] = t.numpy()
    for r in reqs:
        r.wait()


def _colour_pass_slab(phi_g, rhs, a, b, lam, dx, colour, k0, lo_phys, hi_phys):
    """one colour pass on a slab whose z ghosts (1 plane) are either neighbour data or a physical face"""
    g = T.ghosted(phi_g[1:-1], dx=dx)            # x/y physical ghosts (+ z physical ghosts, overwritten below if interior)
    if not lo_phys:
        g[0, 1:-1, 1:-1] = phi_g[0]
    if not hi_phys:
        g[-1, 1:-1, 1:-1] = phi_g[-1]
    phi = phi_g[1:-1]
    lof = 1.0 * a * phi
    l = T.lap7(g) * (1.0 / (dx * dx)) * b
    lof = lof - (-1.0) * l
    new = phi - lam * (lof - rhs)
    k, j, i = np.meshgrid(np.arange(phi.shape[0]) + k0, np.arange(phi.shape[1]), np.arange(phi.shape[2]), indexing="ij")
    return np.where(((i + j + k + colour) % 2) == 0, new, phi)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, dx = (16, 12, 32), 0.5
    rng = np.random.default_rng(5)
    phi0 = rng.standard_normal((n[2], n[1], n[0])); rhs = rng.standard_normal(phi0.shape)
    a = rng.standard_normal(phi0.shape) * 0.1; b = np.ones_like(phi0)
    lam = T.compute_lambda(a, 1.0, -1.0, dx)
    ref = T.relax(phi0, rhs, a, b, lam, 1.0, -1.0, dx, 2)
    k0, nzl = slab_partition(n[2], world, 8)[rank]
    sl = slice(k0, k0 + nzl)
    lo_phys, hi_phys = rank == 0, rank == world - 1
    # (1) per-colour smoother: exchange one plane before every colour pass (VariableCoeffPoissonOperator.cpp:301)
    slab = np.zeros((nzl + 2, n[1], n[0])); slab[1:-1] = phi0[sl]
    for _ in range(2):
        for colour in (0, 1):
            _exchange(rank, world, slab, 1)
            slab[1:-1] = _colour_pass_slab(slab, rhs[sl], a[sl], b[sl], lam[sl], dx, colour, k0, lo_phys, hi_phys)
    ok1 = np.array_equal(slab[1:-1], ref[sl])
    # (2) fused smoother: exchange two planes once per sweep, update red on the slab grown by one plane, then black
    slab2 = np.zeros((nzl + 4, n[1], n[0])); slab2[2:-2] = phi0[sl]
    for _ in range(2):
        _exchange(rank, world, slab2, 2)
        glo, ghi = (0 if lo_phys else 1), (0 if hi_phys else 1)          # planes of the neighbour updated redundantly
        lo, hi = 2 - glo, 2 + nzl + ghi
        ext = slice(k0 - glo, k0 + nzl + ghi)
        red = _colour_pass_slab(slab2[lo - 1:hi + 1], rhs[ext], a[ext], b[ext], lam[ext], dx, 0, k0 - glo, lo_phys, hi_phys)
        slab2[lo:hi] = red
        slab2[2:-2] = _colour_pass_slab(slab2[1:-1], rhs[sl], a[sl], b[sl], lam[sl], dx, 1, k0, lo_phys, hi_phys)
    ok2 = np.array_equal(slab2[2:-2], ref[sl])
    # (3) restriction is slab-local when k0 and nz_local are even
    resc = T.restrict_residual(phi0, rhs, a, b, 1.0, -1.0, dx)
    g1 = np.zeros((nzl + 2, n[1], n[0])); g1[1:-1] = phi0[sl]
    _exchange(rank, world, g1, 1)
    gg = T.ghosted(phi0[sl], dx=dx)
    if not lo_phys:
        gg[0, 1:-1, 1:-1] = g1[0]
    if not hi_phys:
        gg[-1, 1:-1, 1:-1] = g1[-1]
    l = T.lap7(gg) * (1.0 / (dx * dx)) * (-1.0) * b[sl]
    q = (rhs[sl] - (1.0 * a[sl] * phi0[sl] - l)) / 8.0
    loc = np.zeros((nzl // 2, n[1] // 2, n[0] // 2))
    for dk in (0, 1):
        for dj in (0, 1):
            for di in (0, 1):
                loc = loc + q[dk::2, dj::2, di::2]
    ok3 = np.array_equal(loc, resc[k0 // 2:(k0 + nzl) // 2])
    flags = torch.tensor([int(ok1), int(ok2), int(ok3)])
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(flags.tolist())
    dist.destroy_process_group()


def test_two_rank_slab_decomposition_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [1, 1, 1], res
