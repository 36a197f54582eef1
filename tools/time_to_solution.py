"""python tools/time_to_solution.py [--n 64] [--max-level 6] [--smooth 4] [--nl 4] [--repeat 2]: wall-clock phases of the
reference's own configuration (params.txt: N = 64, max_level = 6, V(4,4), refine_threshold 0.1, fill_ratio 0.5, block_factor 8,
max_grid_size 16) from process start to the converged nonlinear loop: CUDA context, set_grids, hierarchy, every step of every
nonlinear iteration.  --repeat 2 builds a second, fresh hierarchy in the same process: what its first iteration no longer
pays (kernel module loading, the allocator's first growth) is the process's one-time cost, what it still pays belongs to
the hierarchy.  One JSON line."""
import argparse
import json
import os
import sys
import time

T0 = time.perf_counter()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--max-level", type=int, default=6)
    ap.add_argument("--smooth", type=int, default=4)
    ap.add_argument("--nl", type=int, default=4)
    ap.add_argument("--repeat", type=int, default=2)
    args = ap.parse_args()
    import mg_ic_code_b200 as m
    t_import = time.perf_counter()
    ctx = m.Context(0)
    ctx.sync()
    t_ctx = time.perf_counter()
    P = m.make_params(dict(m.DEFAULTS, N=(args.n,) * 3, L=100.0, numMGsmooth=args.smooth, numMGIterations=2, max_NL_iterations=args.nl,
                           max_grid_size=16, block_factor=8, max_level=args.max_level))
    runs = []
    for rep in range(args.repeat):
        w = {}
        st0 = m.Context.alloc_stats()
        a = time.perf_counter()
        g = m.Grids.generate(ctx, P, 0.1, 0.5)
        ctx.sync()
        b = time.perf_counter()
        w["set_grids_s"] = b - a
        st1 = m.Context.alloc_stats()
        w["set_grids_alloc"] = {"mallocs": st1[0] - st0[0], "malloc_s": st1[1] - st0[1], "frees": st1[2] - st0[2], "free_s": st1[3] - st0[3]}
        H = m.Hierarchy.from_grids(ctx, P, g)
        H.set_initial_conditions()
        ctx.sync()
        c = time.perf_counter()
        w["hierarchy_and_initial_conditions_s"] = c - b
        st2 = m.Context.alloc_stats()
        w["hierarchy_alloc"] = {"mallocs": st2[0] - st1[0], "malloc_s": st2[1] - st1[1], "frees": st2[2] - st1[2], "free_s": st2[3] - st1[3]}
        its = []
        for it in range(args.nl):
            marks = []

            def timer(name):
                ctx.sync()
                marks.append((name, time.perf_counter()))
            s = time.perf_counter()
            sa = m.Context.alloc_stats()
            nrm, nit, st = H.nl_iteration_steps(timer=timer)
            ctx.sync()
            e = time.perf_counter()
            sb = m.Context.alloc_stats()
            its.append({"s": e - s, "dpsi_norm": nrm, "bicgstab_iterations": nit,
                        "alloc": {"mallocs": sb[0] - sa[0], "malloc_s": round(sb[1] - sa[1], 4), "frees": sb[2] - sa[2], "free_s": round(sb[3] - sa[3], 4)},
                        "steps_ms": {x[0]: round((y[1] - x[1]) * 1e3, 3) for x, y in zip(marks, marks[1:])}})
            if nrm < P.tolerance:
                break
        w["nonlinear_iterations"] = its
        w["total_s"] = time.perf_counter() - a
        w["levels"] = g.levels
        w["boxes"] = sum(len(g.boxes(l)) for l in range(g.levels))
        runs.append(w)
        del H, g
    print(json.dumps({"tool": "time_to_solution", "config": "params.txt: %d^3 base, max_level %d, V(%d,%d)" % (args.n, args.max_level, args.smooth, args.smooth),
                      "import_s": t_import - T0, "cuda_context_s": t_ctx - t_import, "runs": runs, "process_total_s": time.perf_counter() - T0}))


if __name__ == "__main__":
    main()
