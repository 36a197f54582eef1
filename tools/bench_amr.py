"""python tools/bench_amr.py [--n 256] [--cycles 10] [--smooth 2]: config C4 on one GPU -- base level n^3,
level 1 = the ONE box that the two 64^3-coarse-cell cubes around the punctures merge into, level 2 = two disjoint
32^3-level-1-cell cubes -- timed as AMR V(smooth, smooth) cycles through mgic_amr_vcycle (CUDA events on the library's
stream), with the composite residual history of an AMRMultiGrid-style iteration (residual, cycle, phi += correction) as
the correctness check: it has to fall by about an order of magnitude per cycle, like tests/test_amr_hierarchy.py's.

NOT YET RUN ON A GPU: written after round 1's GPU budget was spent (the entry points it calls are covered at small size by
tests/test_gpu_amr_hierarchy.py).  Coefficients and right-hand sides of the refined levels are the base level's, injected
piecewise-constantly -- enough for a timing; a level-resolution source evaluation on patches does not exist yet.

Prints one JSON line: cells per level, ms per AMR V-cycle, GDOF/s over the composite (uncovered) cells, residual history."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def c4_boxes(n, L=100.0, offset=10.0, block=8):
    """level-1 and level-2 boxes (lo, hi inclusive, in their level's index space), snapped to `block` coarse cells"""
    def cube(centre_x, half, nlev):
        c = [(centre_x + L / 2) / L * nlev, nlev / 2, nlev / 2]                 # the puncture in this level's cell units
        lo = [int(round((x - half) / block)) * block for x in c]
        return tuple(lo), tuple(l + 2 * half - 1 for l in lo)
    # level 1: two cubes of 64 coarse cells (scaled with n / 256) around the punctures; they overlap, so one merged box
    half0 = max(block, 32 * n // 256)
    a, b = cube(-offset, half0, n), cube(+offset, half0, n)
    lo0 = tuple(min(a[0][d], b[0][d]) for d in range(3))
    hi0 = tuple(max(a[1][d], b[1][d]) for d in range(3))
    l1 = (tuple(2 * x for x in lo0), tuple(2 * x + 1 for x in hi0))
    # level 2: two cubes of 32 level-1 cells around the punctures, kept two level-1 cells inside the level-1 box
    half1 = max(block, 16 * n // 256)
    l2 = []
    for off in (-offset, +offset):
        lo, hi = cube(off, half1, 2 * n)
        lo = tuple(max(lo[d], l1[0][d] + 2 * block) for d in range(3))
        hi = tuple(min(hi[d], l1[1][d] - 2 * block) for d in range(3))
        if any(hi[d] < lo[d] for d in range(3)):
            raise ValueError(f"n = {n} is too small for config C4's nesting (use n >= 128)")
        l2.append((tuple(2 * x for x in lo), tuple(2 * x + 1 for x in hi)))
    return l1, l2


def rep2(x):
    return np.repeat(np.repeat(np.repeat(x, 2, 0), 2, 1), 2, 2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--cycles", type=int, default=10)
    ap.add_argument("--smooth", type=int, default=2)
    ap.add_argument("--box", type=int, default=32)
    args = ap.parse_args()
    import torch
    import mg_ic_code_b200 as m
    n, L = args.n, 100.0
    ctx = m.Context(0)
    P = m.make_params(dict(m.DEFAULTS, N=(n, n, n), L=L, max_grid_size=args.box, numMGsmooth=args.smooth))
    lvl = m.level_op_from_params(ctx, P)
    v = m.MultigridVars(ctx, P)
    dpsi, rhs0, a0, b0 = lvl.create(), lvl.create(), lvl.create(), lvl.create()
    v.set_initial_conditions(dpsi); v.set_rhs_and_a_coef(rhs0, a0); v.set_b_coef(b0)
    f = m.VariableCoeffPoissonOperatorFactory(ctx, P, a0, b0, keep_b=True)
    l1, l2 = c4_boxes(n, L)
    host = {0: dict(a=a0.download(), b=b0.download(), r=rhs0.download())}
    origin = {0: (0, 0, 0)}
    levels, keep, rhs = [], [], [rhs0]

    def make_patch(level, lo, hi, parent):
        op = m.VariableCoeffPoissonOperator.patch(ctx, (n << level,) * 3, lo, hi, L / n / (1 << level))
        sl = tuple(slice(lo[d] // 2 - origin[parent][d], hi[d] // 2 - origin[parent][d] + 1) for d in (2, 1, 0))
        arrays = {k: rep2(x[sl]) for k, x in host[parent].items()}
        fa, fb, fr = op.create(), op.create(), op.create()
        fa.upload(arrays["a"]); fb.upload(arrays["b"]); fr.upload(arrays["r"])
        op.setCoefs(fa, fb, 1.0, -1.0)
        keep.extend([fa, fb])
        rhs.append(fr)
        return op, arrays

    op1, arr1 = make_patch(1, l1[0], l1[1], 0)
    host[1], origin[1] = arr1, l1[0]
    levels.append([op1])
    levels.append([make_patch(2, lo, hi, 1)[0] for lo, hi in l2])
    amr = m.AMRHierarchy(f, levels)
    phi, res, corr = amr.create(), amr.create(), amr.create()
    for x in phi:
        x.upload(np.zeros(x.shape))
    cells = [int(np.prod(x.shape)) for x in phi]
    covered = [cells[1] // 8, (cells[2] + cells[3]) // 8, 0, 0]
    composite = sum(cells) - sum(covered)
    stream = torch.cuda.ExternalStream(ctx.stream)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hist, ms = [], []
    ops = [f.MGnewOp(0)] + amr.patches
    for _ in range(args.cycles):
        amr.residual(res, phi, rhs, False)
        hist.append(amr.norm(res, 0))
        ctx.sync()
        ev0.record(stream)
        amr.vcycle(corr, res)
        ev1.record(stream)
        ctx.sync()
        ms.append(ev0.elapsed_time(ev1))
        for op, p, c in zip(ops, phi, corr):
            op.incr(p, c, 1.0)
    amr.residual(res, phi, rhs, False)
    hist.append(amr.norm(res, 0))
    t = float(np.median(ms[min(3, len(ms) - 1):]))
    print(json.dumps({"tool": "bench_amr", "config": "C4 shape: base n^3 + one level-1 box + two level-2 boxes, one GPU",
                      "n": n, "boxes": {"level1": l1, "level2": l2}, "cells_per_array": cells, "composite_cells": composite,
                      "smooth": args.smooth, "ms_per_amr_vcycle": t, "gdof_per_s_composite": composite / t / 1e6,
                      "residual_history": hist, "launches": ctx.launch_count}))
    assert hist[-1] < 1e-3 * hist[0], "the AMR V-cycle iteration does not converge"


if __name__ == "__main__":
    main()
