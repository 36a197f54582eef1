"""CPU tests of the AMR hierarchy algorithms (tests/amr_twin.py) over the oracle: config C4's shape -- a Bowen-York base
level, ONE refined box around both punctures, TWO disjoint refined boxes on level 2 -- at a size the oracle finishes in
seconds.  The GPU tests (test_gpu_amr.py) check the library's mgic_amr_* entry points against the same twin."""
import numpy as np
import pytest

from amr_twin import AmrTwin, OracleBackend, rep2
from oracle import Oracle, OraclePatch

# boxes [lo, hi] in their level's index space for a base level of N^3 cells, N = 32 (scaled with N // 32)
C4_BOXES = {1: [((16, 24, 24), (47, 39, 39))],
            2: [((44, 56, 56), (59, 71, 71)), ((68, 56, 56), (83, 71, 71))]}


def c4_hierarchy(N=32, L=100.0, smooth=2, mg_iterations=2, box=16):
    """(base Oracle, patches [[level 1], [level 2]], rhs level vector): coefficients and right-hand sides of every level
    are the Bowen-York source terms evaluated AT THAT LEVEL'S resolution (a whole-domain Oracle of the refined grid, cut
    to the box)."""
    s = N // 32
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
        for lo, hi in C4_BOXES[l]:
            lo, hi = tuple(x * s for x in lo), tuple((x + 1) * s - 1 for x in hi)
            P = OraclePatch((N << l,) * 3, lo, hi, L / N / (1 << l), max_grid_size=box)
            sl = tuple(slice(lo[d], hi[d] + 1) for d in (2, 1, 0))
            P.set("A", a[sl]); P.set("B", b[sl])
            rhs.append(r[sl].copy())
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
