"""CPU tests of the AMR hierarchy algorithms (tests/amr_twin.py) over the oracle: config C4's shape -- a Bowen-York base
level, ONE refined box around both punctures, TWO disjoint refined boxes on level 2 -- at a size the oracle finishes in
seconds.  The GPU tests (test_gpu_amr.py) check the library's mgic_amr_* entry points against the same twin."""
import numpy as np
import pytest

from amr_twin import AmrTwin, OracleBackend, hierarchy_nl_solve, rep2
from oracle import Oracle, OraclePatch

# boxes [lo, hi] in their level's index space for a base level of N^3 cells, N = 32 (scaled with N // 32)
C4_BOXES = {1: [((16, 24, 24), (47, 39, 39))],
            2: [((44, 56, 56), (59, 71, 71)), ((68, 56, 56), (83, 71, 71))]}


# the same three levels with levels made of TOUCHING boxes whose union is no rectangle (what BRMeshRefine produces): a node is
# a list of boxes = one connected component of its level, held by the library in one masked array
MASKED_BOXES = {1: [[((16, 24, 24), (31, 39, 39)), ((32, 24, 24), (47, 31, 39))]],
                2: [[((36, 52, 52), (51, 67, 67)), ((52, 52, 52), (59, 59, 67))], ((68, 52, 52), (83, 59, 67))]}


def c4_hierarchy(N=32, L=100.0, smooth=2, mg_iterations=2, box=16, boxes=None):
    """(base Oracle, patches [[level 1], [level 2]], rhs level vector): coefficients and right-hand sides of every level
    are the Bowen-York source terms evaluated AT THAT LEVEL'S resolution (a whole-domain Oracle of the refined grid, cut
    to the box).  boxes: {level: [node, ...]}, node = (lo, hi) or a list of (lo, hi) that touch."""
    s = N // 32
    C4_BOXES = boxes if boxes is not None else globals()["C4_BOXES"]
    o = Oracle(N=(N, N, N), max_grid_size=box, numMGsmooth=smooth, numMGIterations=mg_iterations, L=L)
    o.setup()
    rhs, patches = [o.get("RHS")], []
    for l in (1, 2):
        fine = Oracle(N=(N << l,) * 3, max_grid_size=box << l, numMGsmooth=smooth, L=L)
        fine.set_initial_conditions()
        fine.set_coefs_and_rhs()
        a, b, r = fine.get("A"), fine.get("B"), fine.get("RHS")
        fine.close()
        lv = []
        for node in C4_BOXES[l]:
            scale = lambda lo, hi: (tuple(x * s for x in lo), tuple((x + 1) * s - 1 for x in hi))
            if isinstance(node, list):
                P = OraclePatch((N << l,) * 3, None, None, L / N / (1 << l), max_grid_size=box, boxes=[scale(*b) for b in node])
            else:
                P = OraclePatch((N << l,) * 3, *scale(*node), L / N / (1 << l), max_grid_size=box)
            lo, hi = P.lo, P.hi
            sl = tuple(slice(lo[d], hi[d] + 1) for d in (2, 1, 0))
            P.set("A", a[sl]); P.set("B", b[sl])
            rhs.append(r[sl] * P.mask())
            lv.append(P)
        patches.append(lv)
    return o, patches, rhs


@pytest.fixture(scope="module")
def c4():
    o, patches, rhs = c4_hierarchy()
    return AmrTwin(OracleBackend(o, patches)), rhs


def test_hierarchy_shape(c4):
    tw, rhs = c4
    assert tw.n == 4 and tw.levels == 3
    assert tw.b.parent == [-1, 0, 1, 1] and tw.b.level == [0, 1, 2, 2]
    assert tw.b.shape[1] == (16, 16, 32) and tw.b.shape[2] == tw.b.shape[3] == (16, 16, 16)
    assert tw.under[1] == (slice(12, 20), slice(12, 20), slice(8, 24))
    assert tw.under[2] == (slice(4, 12), slice(4, 12), slice(6, 14)) and tw.under[3][2] == slice(18, 26)


def test_masked_norms_and_dot(c4):
    tw, _ = c4
    rng = np.random.default_rng(3)
    x = [rng.standard_normal(s) for s in tw.b.shape]
    y = [rng.standard_normal(s) for s in tw.b.shape]
    xm = tw.zero_covered(x)
    assert np.all(xm[0][tw.under[1]] == 0) and np.all(xm[1][tw.under[2]] == 0) and np.all(xm[1][tw.under[3]] == 0)
    assert np.array_equal(xm[2], x[2]) and np.count_nonzero(xm[0]) == 32 ** 3 - 16 * 8 * 8
    big = [a.copy() for a in x]
    big[0][tw.under[1]] = 1e9                      # covered cells do not count
    assert tw.norm(big, 0) == tw.norm(x, 0) == max(np.abs(a).max() for a in xm)
    vol = sum(np.count_nonzero(a) * tw.b.dx[q] ** 3 for q, a in enumerate(xm))
    assert abs(vol - 100.0 ** 3) < 1e-6            # the uncovered cells tile the domain exactly once
    assert abs(tw.dot(x, y) - tw.dot(y, x)) < 1e-9 * abs(tw.dot(x, x))
    assert abs(tw.norm(x, 2) ** 2 - tw.dot(x, x)) < 1e-9 * tw.dot(x, x)
    ad = tw.average_down(x)
    assert np.allclose(ad[1][tw.under[2]], x[2].reshape(8, 2, 8, 2, 8, 2).mean(axis=(1, 3, 5)))
    assert np.allclose(ad[0][tw.under[1]], ad[1].reshape(8, 2, 8, 2, 16, 2).mean(axis=(1, 3, 5)))   # level 1 AFTER level 2 came down


def test_amr_vcycles_converge_on_the_composite_residual(c4):
    """AMRMultiGrid::solveNoInit's iteration -- composite residual, one AMR V(2,2) cycle, phi += correction -- on the
    three-level hierarchy: the composite residual (covered cells excluded) falls by about an order of magnitude per cycle."""
    tw, rhs = c4
    phi = tw.zeros()
    hist = []
    for _ in range(7):
        r = tw.residual(phi, rhs, False)
        hist.append(tw.norm(r, 0))
        phi = [a + c for a, c in zip(phi, tw.vcycle(r))]
    assert all(hist[i + 1] < 0.2 * hist[i] for i in range(len(hist) - 1)), hist
    assert hist[-1] < 1e-6 * hist[0]


def test_outer_bicgstab_on_the_hierarchy(c4):
    """solver.solve(dpsi, rhs) on the hierarchy: BiCGStab over the level vectors, two AMR V-cycles as preconditioner.
    Converges to the tolerance in a handful of iterations and lands on the fixed point of the V-cycle iteration."""
    tw, rhs = c4
    phi = tw.zeros()
    its, status, hist = tw.bicgstab(phi, rhs, eps=1e-10, imax=100)
    assert status == 1 and its <= 8, (its, status, hist)
    assert hist[-1] <= 1e-10 * hist[0]
    assert tw.norm(tw.residual(phi, rhs, False), 0) < 1e-9 * hist[0]
    ref = tw.zeros()
    for _ in range(14):
        ref = [a + c for a, c in zip(ref, tw.vcycle(tw.residual(ref, rhs, False)))]
    scale = max(np.abs(a).max() for a in ref)
    ym, rm = tw.zero_covered(phi), tw.zero_covered(ref)      # covered cells are not part of the composite system
    assert max(np.abs(a - c).max() for a, c in zip(ym, rm)) < 1e-8 * scale


def test_refinement_improves_the_solution_near_the_punctures():
    """The composite solution on the refined boxes is closer to a uniformly refined solve than the base level alone:
    the coarse-fine interpolation and the composite operator do what refinement is for."""
    o, patches, rhs = c4_hierarchy(mg_iterations=1)
    tw = AmrTwin(OracleBackend(o, patches))
    phi = tw.zeros()
    tw.bicgstab(phi, rhs, eps=1e-10, imax=100)
    coarse_only = Oracle(N=(32, 32, 32), max_grid_size=16, numMGsmooth=2, L=100.0)
    coarse_only.setup(); coarse_only.load_rhs_zero_e()
    coarse_only.outer_solve()
    c0 = coarse_only.get("DPSI")
    uni = Oracle(N=(64, 64, 64), max_grid_size=16, numMGsmooth=2, L=100.0)    # level 1's resolution everywhere
    uni.setup(); uni.load_rhs_zero_e()
    uni.outer_solve()
    u = uni.get("DPSI")
    lo, sh = tw.b.lo[1], tw.b.shape[1]
    u1 = u[lo[2]:lo[2] + sh[0], lo[1]:lo[1] + sh[1], lo[0]:lo[0] + sh[2]]
    err_amr = np.abs(phi[1] - u1).max()
    err_coarse = np.abs(rep2(c0[tw.under[1]]) - u1).max()
    assert err_amr < 0.5 * err_coarse, (err_amr, err_coarse)


def test_c4_boxes_of_the_timing_tool_make_a_converging_hierarchy():
    """tools/bench_amr.py's config-C4 geometry (merged level-1 box around both punctures, two disjoint level-2
    cubes, snapped to block_factor 8, nested by two coarse cells) at half size (128^3 base): a valid hierarchy on which the
    AMR V(2,2) iteration converges -- with the tool's own piecewise-constant injection of coefficients and right-hand sides"""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "bench_amr", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "bench_amr.py"))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    with pytest.raises(ValueError):
        tool.c4_boxes(64)
    l1, l2 = tool.c4_boxes(256)
    assert [l1[1][d] - l1[0][d] + 1 for d in range(3)] == [224, 128, 128]           # one box of 112 x 64 x 64 coarse cells
    assert all([hi[d] - lo[d] + 1 for d in range(3)] == [64, 64, 64] for lo, hi in l2)
    N, L = 128, 100.0
    l1, l2 = tool.c4_boxes(N)
    o = Oracle(N=(N, N, N), max_grid_size=32, numMGsmooth=2, L=L)
    o.setup()
    host = {0: dict(a=o.get("A"), b=o.get("B"), r=o.get("RHS"))}
    origin = {0: (0, 0, 0)}
    patches, rhs = [[], []], [host[0]["r"]]
    for level, boxes, parent in ((1, [l1], 0), (2, l2, 1)):
        for lo, hi in boxes:
            P = OraclePatch((N << level,) * 3, lo, hi, L / N / (1 << level), max_grid_size=32)
            sl = tuple(slice(lo[d] // 2 - origin[parent][d], hi[d] // 2 - origin[parent][d] + 1) for d in (2, 1, 0))
            arr = {k: rep2(x[sl]) for k, x in host[parent].items()}
            P.set("A", arr["a"]); P.set("B", arr["b"])
            rhs.append(arr["r"])
            patches[level - 1].append(P)
            if level == 1:
                host[1], origin[1] = arr, lo
    tw = AmrTwin(OracleBackend(o, patches))
    assert tw.b.parent == [-1, 0, 1, 1]
    phi, hist = tw.zeros(), []
    for _ in range(4):
        r = tw.residual(phi, rhs, False)
        hist.append(tw.norm(r, 0))
        phi = [a + c for a, c in zip(phi, tw.vcycle(r))]
    assert all(hist[i + 1] < 0.25 * hist[i] for i in range(3)), hist


def test_touching_boxes_hierarchy_converges():
    """the same three levels with nodes that are unions of touching boxes (L shapes): AMR V-cycles still contract the
    composite residual, and the uncovered cells tile the domain exactly once"""
    o, patches, rhs = c4_hierarchy(boxes=MASKED_BOXES)
    tw = AmrTwin(OracleBackend(o, patches))
    assert tw.n == 4 and tw.fmask[1] is not None and tw.fmask[2] is not None and tw.fmask[3] is None
    assert patches[0][0].num_boxes == 2 and tw.b.shape[1] == (16, 16, 32)
    ones = [np.ones(s) if tw.fmask[q] is None else tw.fmask[q].copy() for q, s in enumerate(tw.b.shape)]
    vol = sum(np.count_nonzero(a) * tw.b.dx[q] ** 3 for q, a in enumerate(tw.zero_covered(ones)))
    assert abs(vol - 100.0 ** 3) < 1e-6
    phi = tw.zeros()
    r0 = tw.norm(tw.residual(phi, rhs, True), 0)
    hist = [r0]
    for _ in range(5):
        res = tw.residual(phi, rhs, True)
        cor = tw.vcycle(res)
        phi = [a + c for a, c in zip(phi, cor)]
        hist.append(tw.norm(tw.residual(phi, rhs, True), 0))
    assert hist[-1] < 1e-4 * hist[0] and all(b < 0.5 * a for a, b in zip(hist, hist[1:])), hist
    # nothing ever lands outside the level's boxes
    for q in (1, 2):
        assert np.all(phi[q][tw.fmask[q] == 0] == 0)


def bare_hierarchy(boxes, N=32, L=100.0, smooth=2, mg_iterations=2, box=16, **over):
    """base Oracle (not set up) + OraclePatch objects without coefficients, for the nonlinear loop"""
    o = Oracle(N=(N, N, N), max_grid_size=box, numMGsmooth=smooth, numMGIterations=mg_iterations, L=L, **over)
    patches = []
    for l in sorted(boxes):
        lv = []
        for node in boxes[l]:
            if isinstance(node, list):
                lv.append(OraclePatch((N << l,) * 3, None, None, L / N / (1 << l), max_grid_size=box, boxes=node))
            else:
                lv.append(OraclePatch((N << l,) * 3, node[0], node[1], L / N / (1 << l), max_grid_size=box))
        patches.append(lv)
    return o, patches


@pytest.mark.parametrize("boxes", [C4_BOXES, MASKED_BOXES], ids=["c4", "touching_boxes"])
def test_nonlinear_loop_on_a_hierarchy(boxes):
    """Main_PoissonSolver.cpp:131-216 with max_level = 2: sources on every level, multilevel BiCGStab, psi update with
    QuadCFInterp'd ghosts -- the Newton-like iteration converges like the single-level one (5e-2, 2e-5, 1e-8)
    and the first iteration's sources are the level-resolution Bowen-York terms c4_hierarchy evaluates independently."""
    o, patches = bare_hierarchy(boxes)
    log = []
    norms, psi = hierarchy_nl_solve(o, patches, max_nl=3, log=log)
    assert len(norms) == 3 and norms[0] > 1e-2 and norms[1] < 1e-3 * norms[0] and norms[2] < 1e-2 * norms[1], norms
    assert all(st == 1 for _, st, _ in log)
    # psi stays 1 + O(1e-3) and is smooth across levels: the fine solution averaged down is the coarse one to discretisation error
    tw = AmrTwin(OracleBackend(o, patches))
    assert all(np.abs(p[p != 0] - 1).max() < 0.05 for p in psi)
    down = tw.average_down(psi)
    assert np.abs(down[0] - psi[0]).max() < 2e-3
    # iteration 1's rhs on the patches == the independently evaluated level-resolution sources
    o2, patches2 = bare_hierarchy(boxes)
    _, _, rhs2 = c4_hierarchy(boxes=boxes)
    o2.set_initial_conditions()
    for lv in patches2:
        for q in lv:
            q.set_initial_conditions(o2.params); q.set_coefs_and_rhs()
    flat = [q for lv in patches2 for q in lv]
    for n, q in enumerate(flat, start=1):
        assert np.array_equal(q.var(8), rhs2[n])
