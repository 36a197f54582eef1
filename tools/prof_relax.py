"""Profiling driver: a few finest-level relax() calls (and optionally a V-cycle) at n^3, nothing else.
   python tools/prof_relax.py [n] [sweeps] [smoother] [fused_cfg] [vcycles] [use_graph]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mg_ic_code_b200 as m

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
smoother = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cfg = int(sys.argv[4]) if len(sys.argv) > 4 else -1
vcycles = int(sys.argv[5]) if len(sys.argv) > 5 else 0
graph = int(sys.argv[6]) if len(sys.argv) > 6 else 1
ctx = m.Context(0)
ctx.set_option("use_graph", graph)
if cfg >= 0:
    ctx.set_option("fused_cfg", cfg)
P = m.make_params(dict(m.DEFAULTS, N=(n, n, n), max_grid_size=32, numMGsmooth=2))
lvl = m.level_op_from_params(ctx, P)
v = m.MultigridVars(ctx, P)
dpsi, rhs, a, b = lvl.create(), lvl.create(), lvl.create(), lvl.create()
v.set_initial_conditions(dpsi)
v.set_rhs_and_a_coef(rhs, a)
v.set_b_coef(b)
v.close()
f = m.VariableCoeffPoissonOperatorFactory(ctx, P, a, b)
f.set_smoother(smoother)
op = f.MGnewOp(0)
e = op.create()
op.relax(e, rhs, sweeps)
for _ in range(vcycles):
    op.setToZero(e)
    f.vcycle(e, rhs)
ctx.sync()
print("done", ctx.launch_count)
