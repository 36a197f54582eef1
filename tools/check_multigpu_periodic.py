"""torchrun --nproc-per-node N tools/check_multigpu_periodic.py [n]: is_periodic = 1 across ranks -- the z-slabs form a ring
(rank 0's lower neighbour is the last rank; with two ranks both neighbours are the same peer) -- against the same periodic
problem on one GPU: colour sweeps, residual, restrictResidual / prolongIncrement bitwise, by NVLink peer stores and by
ncclSend/ncclRecv; one V-cycle to 1e-6 (with K = 0 the periodic problem is singular and the bottom solver's hang / restart
path amplifies the summation order of its dot products, as on one GPU: tests/test_gpu_parity.py::test_vcycle_periodic)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mg_ic_code_b200 as m
from mg_ic_code_b200 import comm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = m.Context(local, rank=rank, nranks=world)
comm.attach(ctx, dist)
box = 16
P = m.make_params(dict(m.DEFAULTS, N=(n, n, n), L=20.0, max_grid_size=box, numMGsmooth=2, is_periodic=1))
k0, nzl = comm.slab_partition(n, world, box)[rank]
rng = np.random.default_rng(7)
e0, r0, a0 = rng.standard_normal((n, n, n)), rng.standard_normal((n, n, n)), -0.5 + 0.1 * rng.standard_normal((n, n, n))


def gather(field, shape):
    full = np.zeros(shape)
    field.download(full)
    t = torch.from_numpy(full).cuda()
    dist.all_reduce(t)
    return t.cpu().numpy()


def run(c, k0, nzl, multi):
    lvl = m.level_op_from_params(c, P, k0, nzl)
    a, b, e, r, t = (lvl.create() for _ in range(5))
    a.upload(a0); b.upload(np.ones((n, n, n))); e.upload(e0); r.upload(r0)
    f = m.VariableCoeffPoissonOperatorFactory(c, P, a, b)
    op = f.MGnewOp(0)
    out = {}
    get = (lambda fld, shape=(n, n, n): gather(fld, shape)) if multi else (lambda fld, shape=(n, n, n): fld.download())
    op.relax(e, r, 3)
    out["relax"] = get(e)
    op.residual(t, e, r, True)
    out["residual"] = get(t)
    # restrictResidual into a coarse slab field of our own (the hierarchy's own coarse level may be agglomerated)
    cop = m.VariableCoeffPoissonOperator(c, (n // 2,) * 3, 2 * (P.L / n), bc_lo=(2, 2, 2), bc_hi=(2, 2, 2), k0=k0 // 2, nz_local=nzl // 2)
    rc = cop.create()
    op.restrictResidual(rc, e, r)
    out["restrict"] = get(rc, (n // 2,) * 3)
    ec = cop.create()
    ec.upload(e0[::2, ::2, ::2].copy())
    op.prolongIncrement(e, ec)
    out["prolong"] = get(e)
    op.setToZero(e)
    f.vcycle(e, r)
    out["vcycle"] = get(e)
    return out


res = {"p2p": run(ctx, k0, nzl, True)}
ctx.set_option("p2p_halo", 0)
res["nccl"] = run(ctx, k0, nzl, True)
ctx.set_option("p2p_halo", 1)
stats = comm.halo_stats(ctx)
ok = True
if rank == 0:
    one = run(m.Context(local), 0, n, False)
    for name in ("p2p", "nccl"):
        for k in ("relax", "residual"):
            same = np.array_equal(res[name][k], one[k])
            print(f"{name} {k}: bitwise", "OK" if same else "MISMATCH")
            ok &= same
        for k in ("restrict", "prolong"):
            same = np.array_equal(res[name][k], one[k])
            print(f"{name} {k}: bitwise", "OK" if same else "MISMATCH")
            ok &= same
        err = np.abs(res[name]["vcycle"] - one["vcycle"]).max() / np.abs(one["vcycle"]).max()
        print(f"{name} V-cycle: rel err", err)
        ok &= err < 1e-6
    print("halo exchanges (peer stores, nccl, peer mapping available):", stats)
    ok &= stats[1] > 0 and (not stats[2] or stats[0] > 0)
viol, _ = m.Context.guard_check()   # guard bands around the device arrays (MGIC_ARENA_GUARD, set by the test suite): none damaged
if viol:
    print(f"rank {rank}: {viol} arrays with damaged guard bands")
ok = bool(ok) and viol == 0
flag = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTIGPU PERIODIC CHECK", "PASSED" if flag.item() else "FAILED", f"({world} ranks, {n}^3)")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
