"""torchrun --nproc-per-node N tools/check_multigpu_amr.py [n]: poissonSolve on an AMR hierarchy (config C4's shape) with the
base level cut into z-slabs over the ranks and the refined levels replicated, against the same hierarchy on one GPU.

Every rank also runs the whole problem on its own GPU with a single-rank context.  Checked: dpsi norms of three nonlinear
iterations, BiCGStab iteration counts, psi on every node to 1e-10 relative in max-norm (the base level's reductions are summed in
a different order across ranks; everything else is the same arithmetic), and the GRChombo checkpoint rank 0 writes from the
distributed hierarchy (the base level gathered for it) against the one-GPU file: same header, data to 1e-10."""
import json
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mg_ic_code_b200 as m
from mg_ic_code_b200 import comm
from tools.bench_amr import c4_boxes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = m.Context(local, rank=rank, nranks=world)
comm.attach(ctx, dist)
P = m.make_params(dict(m.DEFAULTS, N=(n, n, n), L=100.0, max_grid_size=n // world if n // world < 32 else 32, numMGsmooth=2, numMGIterations=2,
                       max_NL_iterations=3, max_level=2))
l1, l2 = c4_boxes(n)
levels = [[l1], [b for b in l2]]


tmp = tempfile.mkdtemp(prefix="mgic_chk_")


def run(c, chk):
    H = m.Hierarchy(c, P, levels)
    H.set_initial_conditions()
    rows = [H.nl_iteration() for _ in range(3)]
    psi = [H.download(q, "psi") for q in range(H.nodes)]
    H.write_checkpoint(chk)          # collective on a multi-rank context; rank 0 writes
    H.close()
    return rows, psi


rows_m, psi_m = run(ctx, os.path.join(tmp, "ranks.mgic"))
t = torch.from_numpy(psi_m[0]).cuda()          # node 0: each rank filled its own planes of the global array
dist.all_reduce(t)
psi_m[0] = t.cpu().numpy()
c1 = m.Context(local)
rows_1, psi_1 = run(c1, os.path.join(tmp, f"one_{rank}.mgic"))
ok = True
for a, b in zip(rows_m, rows_1):
    ok &= a[1:] == b[1:] and abs(a[0] - b[0]) <= 1e-9 * b[0]
errs = [float(np.abs(x - y).max() / np.abs(y).max()) for x, y in zip(psi_m, psi_1)]
ok &= all(e < 1e-10 for e in errs)
if rank == 0:
    from mg_ic_code_b200 import checkpoint
    ha, da = checkpoint.read(os.path.join(tmp, "ranks.mgic"))
    hb, db = checkpoint.read(os.path.join(tmp, "one_0.mgic"))
    same_header = json.dumps(ha, sort_keys=True) == json.dumps(hb, sort_keys=True)
    chk_err = max(float(np.abs(x - y).max() / max(np.abs(y).max(), 1e-300)) for la, lb in zip(da, db) for x, y in zip(la, lb))
    print("checkpoint written from the ranks vs one GPU: header", "equal" if same_header else "DIFFERENT", "data rel err", chk_err)
    ok &= same_header and chk_err < 1e-10
    print("nonlinear iterations (dpsi norm, BiCGStab iterations, status):", rows_m, "one GPU:", rows_1)
    print("psi per node, relative max-norm difference to one GPU:", errs)
viol, _ = m.Context.guard_check()   # guard bands around the device arrays (MGIC_ARENA_GUARD, set by the test suite): none damaged
if viol:
    print(f"rank {rank}: {viol} arrays with damaged guard bands")
ok = bool(ok) and viol == 0
flag = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTIGPU AMR CHECK", "PASSED" if flag.item() else "FAILED", f"({world} ranks, base {n}^3, {len(psi_m)} nodes)")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
