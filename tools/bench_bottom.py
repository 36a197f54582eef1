"""python tools/bench_bottom.py: the bottom BiCGStab alone on the level sizes that occur as (agglomerated) bottom levels,
one kernel variant after the other (1 default, 4 per-CTA bricks, 5 cluster-held bricks, 3 cooperative grid)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mg_ic_code_b200 as m

ctx = m.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
shapes = ((32, 32, 32), (32, 32, 64), (32, 64, 64), (64, 64, 64))
if len(sys.argv) > 1:
    shapes = tuple(tuple(int(x) for x in a.split("x")) for a in sys.argv[1].split(","))
kernels = tuple(int(x) for x in sys.argv[2].split(",")) if len(sys.argv) > 2 else (1, 4, 5, 3)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
for shape in shapes:
    P = m.make_params(dict(m.DEFAULTS, N=shape, L=100.0 * shape[0] / 64, max_grid_size=2))
    lvl = m.level_op_from_params(ctx, P)
    v = m.MultigridVars(ctx, P)
    dpsi, rhs, a, b = lvl.create(), lvl.create(), lvl.create(), lvl.create()
    v.set_initial_conditions(dpsi); v.set_rhs_and_a_coef(rhs, a); v.set_b_coef(b)
    f = m.VariableCoeffPoissonOperatorFactory(ctx, P, a, b)
    op = f.MGnewOp(0)
    r = np.random.default_rng(5).standard_normal((shape[2], shape[1], shape[0]))
    res, e = op.create(), op.create()
    res.upload(r)
    line = []
    for kern in kernels:
        ctx.set_option("bottom_kernel", kern)
        for _ in range(min(3, reps)):
            op.setToZero(e)
            its = f.bottom_solve(e, res)
        ctx.sync()
        t = 0.0
        for _ in range(reps):
            op.setToZero(e)
            ctx.sync()
            ev0.record(stream)
            m._capi.check(ctx.L.mgic_mg_bottom_solve(f.h, e.h, res.h, None))
            ev1.record(stream)
            ctx.sync()
            t += ev0.elapsed_time(ev1)
        line.append(f"kernel {kern} (ran {ctx.get_option('last_bottom_kernel')}): {t / reps * 1e3:.0f} us, {its} its, {t / reps * 1e3 / its:.1f} us/it")
    print("BOTTOM", shape, " | ".join(line), flush=True)
    ctx.set_option("bottom_kernel", 1)
