"""torchrun --nproc-per-node N tools/bench_halo.py: cost of the halo exchange in the fused sweeps, with and without
overlap, eager and as a replayed V-cycle graph (512^3 cells per rank, weak-scaling shape of bench.py)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mg_ic_code_b200 as m
from mg_ic_code_b200 import comm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = m.Context(local)
comm.attach(ctx, dist)
n = 512
mult = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}[world]
N = (n * mult[0], n * mult[1], n * mult[2])
P = m.make_params(dict(m.DEFAULTS, N=N, L=100.0 * mult[0], max_grid_size=32, numMGsmooth=2))
nzl = N[2] // world
k0 = rank * nzl
lvl = m.level_op_from_params(ctx, P, k0, nzl)
v = m.MultigridVars(ctx, P, k0, nzl)
dpsi, rhs, a, b = lvl.create(), lvl.create(), lvl.create(), lvl.create()
v.set_initial_conditions(dpsi); v.set_rhs_and_a_coef(rhs, a); v.set_b_coef(b); v.close()
f = m.VariableCoeffPoissonOperatorFactory(ctx, P, a, b)
op = f.MGnewOp(0)
e = op.create()
stream = torch.cuda.ExternalStream(ctx.stream)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    ctx.sync(); dist.barrier(); torch.cuda.synchronize()
    ev0.record(stream)
    for _ in range(reps):
        fn()
    ev1.record(stream)
    ctx.sync(); dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([ev0.elapsed_time(ev1) / reps], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


res = {}
for p2p in (1, 0):
    ctx.set_option("p2p_halo", p2p)
    tr = "p2p" if p2p else "nccl"
    for ov in (0, 1):
        ctx.set_option("overlap_halo", ov)
        res[f"{tr}_relax4_overlap{ov}"] = timeit(lambda: op.relax(e, rhs, 4))
        for g in ((1,) if ov else (0, 1)):
            ctx.set_option("use_graph", g)
            f2 = m.VariableCoeffPoissonOperatorFactory(ctx, P, a, b)   # fresh graph cache
            o2 = f2.MGnewOp(0)
            e2 = o2.create()
            res[f"{tr}_vcycle_overlap{ov}_graph{g}"] = timeit(lambda: f2.vcycle_from_zero(e2, rhs))
            f2.close()
ctx.set_option("use_graph", 1)
ctx.set_option("overlap_halo", 0)
ctx.set_option("p2p_halo", 1)
if rank == 0:
    print("HALO BENCH", world, "ranks", N, {k: round(x, 3) for k, x in res.items()}, "halo bytes rank0", comm.halo_bytes(ctx),
          "exchanges (p2p, nccl, p2p available)", comm.halo_stats(ctx))
dist.barrier()
dist.destroy_process_group()
