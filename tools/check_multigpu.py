"""torchrun --nproc-per-node N tools/check_multigpu.py [n]: z-slab decomposed hot path vs the same problem on one GPU.

Every rank owns a slab of an n x n x n domain (n planes split evenly); rank 0 also solves the whole domain on its own
GPU with a single-rank context.  Checked: fused sweeps, per-colour sweeps, residual, restrictResidual bitwise; norms and
dot products to 1e-13; V-cycles to 1e-10 relative (the bottom solver's reductions differ in summation order)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mg_ic_code_b200 as m
from mg_ic_code_b200 import comm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = m.Context(local)
comm.attach(ctx, dist)
ctx.set_option("fused_min_cells", 0)
box = 32
P = m.make_params(dict(m.DEFAULTS, N=(n, n, n), max_grid_size=box, numMGsmooth=2))
k0, nzl = comm.slab_partition(n, world, box)[rank]


def build(c, k0, nzl):
    lvl = m.level_op_from_params(c, P, k0, nzl)
    v = m.MultigridVars(c, P, k0, nzl)
    dpsi, rhs, a, b = lvl.create(), lvl.create(), lvl.create(), lvl.create()
    v.set_initial_conditions(dpsi)
    v.set_rhs_and_a_coef(rhs, a)
    v.set_b_coef(b)
    f = m.VariableCoeffPoissonOperatorFactory(c, P, a, b)
    op = f.MGnewOp(0)
    return dict(f=f, op=op, rhs=rhs, a=a, b=b, e=op.create(), t=op.create(), keep=(lvl, v, dpsi))


def gather(field):
    """global array on rank 0 from the ranks' slabs"""
    full = np.zeros((n, n, n))
    field.download(full)  # fills this rank's planes
    t = torch.from_numpy(full).cuda()
    dist.reduce(t, 0, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


D = build(ctx, k0, nzl)
rng = np.random.default_rng(3)
e0 = rng.standard_normal((n, n, n))
results = {}
for name, smoother in (("fused", 1), ("colour", 0)):
    D["f"].set_smoother(smoother)
    D["e"].upload(e0)
    D["op"].relax(D["e"], D["rhs"], 3)
    results["relax_" + name] = gather(D["e"])
D["e"].upload(e0)
D["op"].residual(D["t"], D["e"], D["rhs"], True)
results["residual"] = gather(D["t"])
nrm = D["op"].norm(D["t"], 0), D["op"].norm(D["t"], 2), D["op"].dotProduct(D["t"], D["e"])
D["f"].set_smoother(1)
D["op"].setToZero(D["e"])
hist = []
for _ in range(3):
    D["f"].vcycle(D["e"], D["rhs"])
    D["op"].residual(D["t"], D["e"], D["rhs"], True)
    hist.append(D["op"].norm(D["t"], 0))
results["vcycle_e"] = gather(D["e"])
# same V-cycles with the halo exchange overlapped with the interior planes (off by default)
ctx.set_option("overlap_halo", 1)
D2 = build(ctx, k0, nzl)
D2["op"].setToZero(D2["e"])
for _ in range(3):
    D2["f"].vcycle(D2["e"], D2["rhs"])
results["vcycle_e_overlap"] = gather(D2["e"])
ctx.set_option("overlap_halo", 0)
# and with the halo planes sent by ncclSend/ncclRecv instead of peer stores
stats_p2p = comm.halo_stats(ctx)
ctx.set_option("p2p_halo", 0)
D3 = build(ctx, k0, nzl)
D3["op"].setToZero(D3["e"])
for _ in range(3):
    D3["f"].vcycle(D3["e"], D3["rhs"])
results["vcycle_e_nccl"] = gather(D3["e"])
ctx.set_option("p2p_halo", 1)
stats_all = comm.halo_stats(ctx)
# and with the halo folded into the sweep: every sweep stores its boundary planes into the neighbours' ghost planes itself
# (fold_halo = 1; off by default) instead of a k_halo_push exchange before every sweep
ctx.set_option("fold_halo", 1)
D4 = build(ctx, k0, nzl)
D4["op"].setToZero(D4["e"])
for _ in range(3):
    D4["f"].vcycle(D4["e"], D4["rhs"])
results["vcycle_e_unfolded"] = gather(D4["e"])
D4["e"].upload(e0)
D4["op"].relax(D4["e"], D4["rhs"], 3)
results["relax_unfolded"] = gather(D4["e"])
stats_unfolded = comm.halo_stats(ctx)
ctx.set_option("fold_halo", 0)
ok = True
if rank == 0:
    c1 = m.Context(local)
    c1.set_option("fused_min_cells", 0)
    S = build(c1, 0, n)
    for name, smoother in (("fused", 1), ("colour", 0)):
        S["f"].set_smoother(smoother)
        S["e"].upload(e0)
        S["op"].relax(S["e"], S["rhs"], 3)
        same = np.array_equal(S["e"].download(), results["relax_" + name])
        print(f"relax {name}: bitwise {'OK' if same else 'MISMATCH'}")
        ok &= same
    S["e"].upload(e0)
    S["op"].residual(S["t"], S["e"], S["rhs"], True)
    same = np.array_equal(S["t"].download(), results["residual"])
    print("residual: bitwise", "OK" if same else "MISMATCH")
    ok &= same
    n1 = S["op"].norm(S["t"], 0), S["op"].norm(S["t"], 2), S["op"].dotProduct(S["t"], S["e"])
    close = nrm[0] == n1[0] and abs(nrm[1] - n1[1]) <= 1e-13 * n1[1] and abs(nrm[2] - n1[2]) <= 1e-11 * abs(n1[2]) + 1e-9
    print("norms/dot:", nrm, n1, "OK" if close else "MISMATCH")
    ok &= close
    S["f"].set_smoother(1)
    S["op"].setToZero(S["e"])
    h1 = []
    for _ in range(3):
        S["f"].vcycle(S["e"], S["rhs"])
        S["op"].residual(S["t"], S["e"], S["rhs"], True)
        h1.append(S["op"].norm(S["t"], 0))
    es = S["e"].download()
    err = np.abs(es - results["vcycle_e"]).max() / np.abs(es).max()
    print("vcycle residual history", hist, h1, "rel err of e", err)
    ok &= err < 1e-10 and all(abs(x - y) <= 1e-9 * y + 1e-18 for x, y in zip(hist, h1))
    same = np.array_equal(results["vcycle_e_overlap"], results["vcycle_e"])
    print("vcycle with overlapped halo exchange: bitwise", "OK" if same else "MISMATCH")
    ok &= same
    same = np.array_equal(results["vcycle_e_nccl"], results["vcycle_e"])
    print("vcycle with ncclSend/ncclRecv halos: bitwise", "OK" if same else "MISMATCH")
    ok &= same
    same = np.array_equal(results["vcycle_e_unfolded"], results["vcycle_e"]) and np.array_equal(results["relax_unfolded"], results["relax_fused"])
    print("vcycle / sweeps with the halo stored by the sweep itself vs exchanged before every sweep: bitwise", "OK" if same else "MISMATCH",
          "(exchange kernels left in the folded pass:", stats_unfolded[0] - stats_all[0], ")")
    ok &= same
    print("halo exchanges (peer stores, nccl, peer mapping available): before the nccl pass", stats_p2p, "after", stats_all)
    ok &= stats_all[1] > 0 and (not stats_all[2] or (stats_p2p[0] > 0 and stats_p2p[1] == 0))
viol, _ = m.Context.guard_check()   # guard bands around the device arrays (MGIC_ARENA_GUARD, set by the test suite): none damaged
if viol:
    print(f"rank {rank}: {viol} arrays with damaged guard bands")
ok = bool(ok) and viol == 0
flag = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTIGPU CHECK", "PASSED" if flag.item() else "FAILED", f"({world} ranks, {n}^3, halo bytes sent by rank 0: {comm.halo_bytes(ctx)})")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
