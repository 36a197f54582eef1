"""torchrun --nproc-per-node N tools/bench_exchange.py: the halo exchange alone (back-to-back exchanges of one field, no
sweeps in between), NVLink peer stores vs ncclSend/ncclRecv, for the plane sizes of the 512^3-per-GPU V-cycle."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mg_ic_code_b200 as m
from mg_ic_code_b200 import comm
from mg_ic_code_b200._capi import check, lib

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = m.Context(local, rank=rank, nranks=world)
comm.attach(ctx, dist)
L = lib()
stream = torch.cuda.ExternalStream(ctx.stream)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
out = {}
for side in (128, 256, 512, 1024):
    nzl = 16
    P = m.make_params(dict(m.DEFAULTS, N=(side, side, nzl * world), max_grid_size=16))
    lvl = m.level_op_from_params(ctx, P, rank * nzl, nzl)
    f = lvl.create()
    for tr in ("p2p", "nccl"):
        ctx.set_option("p2p_halo", 1 if tr == "p2p" else 0)
        for planes in (1, 2):
            for _ in range(5):
                check(L.mgic_comm_halo_exchange(ctx.h, f.h, planes))
            ctx.sync(); dist.barrier(); torch.cuda.synchronize()
            reps = 50
            ev0.record(stream)
            for _ in range(reps):
                check(L.mgic_comm_halo_exchange(ctx.h, f.h, planes))
            ev1.record(stream)
            ctx.sync(); dist.barrier(); torch.cuda.synchronize()
            t = torch.tensor([ev0.elapsed_time(ev1) / reps * 1e3], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            nb = 2 if world > 2 else 1
            mb = side * side * planes * 8 * nb / 1e6
            out[f"{side}^2 x{planes} {tr}"] = f"{t.item():.1f} us ({mb:.1f} MB out per inner rank, {mb / t.item() * 1e3:.0f} GB/s)"
    ctx.set_option("p2p_halo", 1)
    f.close(); lvl.close()
if rank == 0:
    print("EXCHANGE BENCH", world, "ranks; blocks env", os.environ.get("MGIC_P2P_BLOCKS"))
    for k, v in out.items():
        print("  ", k, v)
    print("   stats", comm.halo_stats(ctx))
dist.barrier()
dist.destroy_process_group()
