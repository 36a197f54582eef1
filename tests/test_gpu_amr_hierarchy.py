"""GPU tests of the AMR hierarchy entry points (mgic_amr_*: AMRVCycle, MultilevelLinearOp, outer BiCGStab on a hierarchy)
through the C ABI on config C4's shape -- base level, one refined box, two disjoint refined boxes on level 2 -- against
tests/amr_twin.py:
  * over the library's own operator-level primitives (pinned to the oracle bit for bit by test_gpu_parity.py): the
    orchestration alone, identical bits;
  * over the CPU oracle: composite residual / applyOp identical bits, reductions 1e-13, preconditioner and solve 1e-10."""
import numpy as np
import pytest

import mg_ic_code_b200 as m
from amr_twin import AmrTwin, OracleBackend, hierarchy_nl_solve
from test_amr_hierarchy import C4_BOXES, MASKED_BOXES, bare_hierarchy, c4_hierarchy

pytestmark = pytest.mark.gpu


def relerr(x, y):
    d = np.abs(np.asarray(x) - np.asarray(y)).max()
    s = np.abs(np.asarray(y)).max()
    return d / s if s > 0 else d


class GpuPrimBackend:
    """amr_twin backend over the library's operator-level entry points (one call per primitive)."""

    def __init__(self, h, ob):
        self.h = h
        for k in ("n_nodes", "level", "parent", "lo", "shape", "dx", "smooth", "mg_iterations", "fmask"):
            setattr(self, k, getattr(ob, k))
        self.F = [dict(E=op.create(), R=op.create(), T=op.create(), C=op.create()) for op in h.ops]

    def _load(self, q, phi=None, rhs=None, coarse=None):
        if phi is not None:
            self.F[q]["E"].upload(phi)
        if rhs is not None:
            self.F[q]["R"].upload(rhs)
        if coarse is not None:
            self.F[self.parent[q]]["C"].upload(coarse)

    def residual0(self, phi, rhs, homog):
        self._load(0, phi, rhs)
        self.h.ops[0].residual(self.F[0]["T"], self.F[0]["E"], self.F[0]["R"], homog)
        return self.F[0]["T"].download()

    def apply0(self, phi, homog):
        self._load(0, phi)
        self.h.ops[0].applyOp(self.F[0]["T"], self.F[0]["E"], homog)
        return self.F[0]["T"].download()

    def vcycle0(self, res):
        self._load(0, rhs=res)
        self.h.f.vcycle_from_zero(self.F[0]["E"], self.F[0]["R"])
        return self.F[0]["E"].download()

    def residual_nf(self, q, phi, coarse, rhs, homog):
        self._load(q, phi, rhs, coarse)
        p, F = self.parent[q], self.F[q]
        self.h.ops[q].AMRResidualNF(F["T"], F["E"], self.F[p]["C"], F["R"], coarse_lo=self.lo[p], homogeneous=homog)
        return F["T"].download()

    def apply_nf(self, q, phi, coarse, homog):
        self._load(q, phi, coarse=coarse)
        p, F = self.parent[q], self.F[q]
        self.h.ops[q].AMROperatorNF(F["T"], F["E"], self.F[p]["C"], coarse_lo=self.lo[p], homogeneous=homog)
        return F["T"].download()

    def relax0(self, q, res, n):
        self._load(q, np.zeros(self.shape[q]), res)
        self.h.ops[q].relax(self.F[q]["E"], self.F[q]["R"], n)
        return self.F[q]["E"].download()


class C4:
    """the oracle's C4-shaped hierarchy and the same hierarchy on the GPU, fed with the oracle's coefficients"""

    def __init__(self, ctx, mg_iterations=2, boxes=None):
        self.o, self.patches, self.rhs = c4_hierarchy(mg_iterations=mg_iterations, boxes=boxes)
        self.ob = OracleBackend(self.o, self.patches)
        self.P = m.make_params(self.o.params)
        lvl = m.level_op_from_params(ctx, self.P)
        self.coef = [(lvl.create(), lvl.create())]
        self.coef[0][0].upload(self.o.get("A")); self.coef[0][1].upload(self.o.get("B"))
        self.f = m.VariableCoeffPoissonOperatorFactory(ctx, self.P, self.coef[0][0], self.coef[0][1], keep_b=True)
        self.f.set_smoother(1)
        self.ops = [self.f.MGnewOp(0)]
        N, L = self.o.params["N"][0], self.o.params["L"]
        levels = []
        for l, lv in enumerate(self.patches, start=1):
            cur = []
            for Pq in lv:
                if Pq.boxes:      # a union of touching boxes: one masked array
                    op = m.VariableCoeffPoissonOperator.patch_boxes(ctx, (N << l,) * 3, Pq.boxes, L / N / (1 << l))
                    assert np.array_equal(op.mask(), Pq.mask()) and op.valid_cells == int(Pq.mask().sum())
                else:
                    op = m.VariableCoeffPoissonOperator.patch(ctx, (N << l,) * 3, Pq.lo, Pq.hi, L / N / (1 << l))
                a, b = op.create(), op.create()
                a.upload(Pq.get("A")); b.upload(Pq.get("B"))
                op.setCoefs(a, b, 1.0, -1.0)
                self.coef.append((a, b))
                self.ops.append(op)
                cur.append(op)
            levels.append(cur)
        self.amr = m.AMRHierarchy(self.f, levels)

    def vec(self, arrays=None):
        v = self.amr.create()
        if arrays is not None:
            for f, a in zip(v, arrays):
                f.upload(a)
        return v

    def rand(self, seed, scale=1.0):
        rng = np.random.default_rng(seed)
        out = [scale * rng.standard_normal(s) for s in self.ob.shape]
        return [a if f is None else a * f for a, f in zip(out, self.ob.fmask)]   # nothing outside a level's boxes


def test_proper_nesting_is_enforced(ctx):
    """a finer node needs one cell of its parent level all around the cells under it, diagonals included (QuadCFInterp's
    stencils): a level-2 patch flush with a face of level 1, or with only an edge neighbour missing, is refused"""
    base = C4(ctx)
    N, L = 32, 100.0

    def patch(dims, lo, hi, dx):
        op = m.VariableCoeffPoissonOperator.patch(ctx, (dims,) * 3, lo, hi, dx)
        a = op.create()
        a.upload(np.ones(a.shape))
        op.setCoefs(a, a, 1.0, -1.0)
        return op, a
    l1, keep1 = patch(2 * N, (16, 16, 16), (47, 47, 47), L / N / 2)
    ok, keep2 = patch(4 * N, (34, 40, 40), (65, 71, 71), L / N / 4)               # level-1 cells 17..32: one cell inside the face x = 16
    assert m.AMRHierarchy(base.f, [[l1], [ok]]).nodes == 3
    flush, keep3 = patch(4 * N, (32, 40, 40), (63, 71, 71), L / N / 4)            # level-1 cells 16..31: flush with the face
    with pytest.raises(m.MgicError, match="properly nested"):
        m.AMRHierarchy(base.f, [[l1], [flush]])
    # an L-shaped level 1 (two boxes in one node) and a level-2 patch in the corner of the L whose diagonal neighbour is missing
    lshape = m.VariableCoeffPoissonOperator.patch_boxes(ctx, (2 * N,) * 3, [((16, 16, 16), (47, 31, 47)), ((16, 32, 16), (31, 47, 47))], L / N / 2)
    al = lshape.create()
    al.upload(np.ones(al.shape))
    lshape.setCoefs(al, al, 1.0, -1.0)
    corner, keep4 = patch(4 * N, (34, 34, 40), (61, 61, 71), L / N / 4)           # level-1 cells 17..30 in x and y: the cell (31, 31) is there, (32, 32) is not
    edge, keep5 = patch(4 * N, (34, 34, 40), (63, 63, 71), L / N / 4)             # level-1 cells 17..31: needs (32, 32), the hole of the L
    assert m.AMRHierarchy(base.f, [[lshape], [corner]]).nodes == 3
    with pytest.raises(m.MgicError, match="properly nested"):
        m.AMRHierarchy(base.f, [[lshape], [edge]])


@pytest.fixture(scope="module", params=["c4", "touching_boxes"])
def c4(ctx, request):
    """config C4's shape with rectangular nodes, and the same three levels with nodes that are unions of touching boxes"""
    return C4(ctx, boxes=MASKED_BOXES if request.param == "touching_boxes" else None)


def test_hierarchy_layout_and_rejections(ctx, c4):
    assert c4.amr.nodes == 4 and c4.amr.levels == 3
    assert [c4.amr.node_info(q) for q in range(4)] == [(0, -1), (1, 0), (2, 1), (2, 1)]
    with pytest.raises(m.MgicError):                       # a level vector has one field per node
        c4.amr.norm(c4.vec()[:3])
    with pytest.raises(m.MgicError):                       # level-2 boxes without their level-1 box: not nested in the base level
        m.AMRHierarchy(c4.f, [[c4.ops[2], c4.ops[3]]])
    # two boxes of one level that touch need a fine-fine exchange: rejected
    N, L = 32, 100.0
    halves = []
    for lo, hi in (((16, 24, 24), (31, 39, 39)), ((32, 24, 24), (47, 39, 39))):
        op = m.VariableCoeffPoissonOperator.patch(ctx, (2 * N,) * 3, lo, hi, L / N / 2)
        a = op.create()
        a.upload(np.ones(a.shape))
        op.setCoefs(a, a, 1.0, -1.0)
        halves.append((op, a))
    with pytest.raises(m.MgicError):
        m.AMRHierarchy(c4.f, [[halves[0][0], halves[1][0]]])
    # ... as ONE node (a union of boxes) they are a level
    both = m.VariableCoeffPoissonOperator.patch_boxes(ctx, (2 * N,) * 3, [((16, 24, 24), (31, 39, 39)), ((32, 24, 24), (47, 39, 39))], L / N / 2)
    assert both.valid_cells == 32 * 16 * 16 and both.n == (32, 16, 16)       # the union fills its bounding box: a rectangular patch
    a2 = both.create()
    a2.upload(np.ones(a2.shape))
    both.setCoefs(a2, a2, 1.0, -1.0)
    h2 = m.AMRHierarchy(c4.f, [[both]])
    assert h2.nodes == 2
    h2.close()


def test_c4_vcycle_is_the_orchestrated_cycle_bit_for_bit(c4):
    """mgic_amr_vcycle on the three-level, four-array hierarchy == the same cycle driven call by call from the host over the
    operator-level primitives; one more cycle on the updated composite residual keeps converging."""
    tw = AmrTwin(GpuPrimBackend(c4, c4.ob))
    res = [r + 1e-6 * n for r, n in zip(c4.rhs, c4.rand(12))]
    want = tw.vcycle(res)
    corr, rin = c4.vec(), c4.vec(res)
    c4.amr.vcycle(corr, rin)
    for q in range(4):
        assert np.array_equal(corr[q].download(), want[q]), q
    assert all(np.array_equal(rin[q].download(), tw._masked(q, res[q])) for q in range(4))     # the caller's residuals are not modified


def test_c4_composite_operators_against_the_oracle(c4):
    tw = AmrTwin(c4.ob)
    phi, rhs = c4.rand(5), c4.rand(6)
    gphi, grhs, out = c4.vec(phi), c4.vec(rhs), c4.vec()
    for homog in (True, False):
        c4.amr.residual(out, gphi, grhs, homog)
        for q, w in enumerate(tw.residual(phi, rhs, homog)):
            assert np.array_equal(out[q].download(), w), ("residual", homog, q)
        c4.amr.applyOp(out, gphi, homog)
        for q, w in enumerate(tw.apply(phi, homog)):
            assert np.array_equal(out[q].download(), w), ("applyOp", homog, q)
    # reductions over the uncovered cells
    big = [a.copy() for a in phi]
    tw._put_under(big, 1, 1e9)
    tw._put_under(big, 3, -1e9)
    gbig = c4.vec(big)
    assert c4.amr.norm(gbig, 0) == tw.norm(phi, 0)
    for ord_ in (1, 2):
        assert abs(c4.amr.norm(gbig, ord_) - tw.norm(phi, ord_)) <= 1e-13 * tw.norm(phi, ord_)
    d = tw.dot(phi, rhs)
    assert abs(c4.amr.dotProduct(gbig, grhs) - d) <= 1e-12 * np.sqrt(tw.dot(phi, phi) * tw.dot(rhs, rhs))
    assert all(np.array_equal(gbig[q].download(), tw._masked(q, big[q])) for q in range(4))    # norm / dot leave their arguments alone
    c4.amr.zeroCovered(gbig)
    for q, w in enumerate(tw.zero_covered(big)):
        assert np.array_equal(gbig[q].download(), w), ("zeroCovered", q)
    c4.amr.averageDown(gphi)
    for q, w in enumerate(tw.average_down(phi)):
        assert np.array_equal(gphi[q].download(), w), ("averageDown", q)


def test_c4_preconditioner_and_outer_solve_against_the_oracle(c4):
    """MultilevelLinearOp::preCond (two AMR V-cycles) and solver.solve(dpsi, rhs) on the hierarchy: north-star tolerance
    1e-10 relative in max-norm against the oracle-backed twin, same iteration count and exit status."""
    tw = AmrTwin(c4.ob)
    cor, res = c4.vec(), c4.vec(c4.rhs)
    c4.amr.preCond(cor, res)
    want = tw.precond(c4.rhs)
    for q in range(4):
        assert relerr(cor[q].download(), want[q]) < 1e-10, q
    phi = tw.zeros()
    its, status, hist = tw.bicgstab(phi, c4.rhs, eps=c4.o.params["tolerance"], imax=c4.o.params["max_iterations"])
    dpsi = c4.vec(tw.zeros())
    g_its, g_status, g_hist = c4.amr.solve(dpsi, res)
    assert (g_its, g_status) == (its, status) and status == 1
    for q in range(4):
        assert relerr(dpsi[q].download(), phi[q]) < 1e-10, q
    for a, b in zip(g_hist, hist):
        if b > 1e-9 * hist[0]:                      # while the residual is far above the rounding floor
            assert abs(a - b) <= 1e-6 * b
    r = c4.vec()
    c4.amr.residual(r, dpsi, res, False)
    assert c4.amr.norm(r, 0) <= 1e-9 * hist[0]


@pytest.mark.parametrize("boxes", [C4_BOXES, MASKED_BOXES], ids=["c4", "touching_boxes"])
def test_nonlinear_solve_on_a_hierarchy(ctx, boxes):
    """poissonSolve with max_level = 2 (Main_PoissonSolver.cpp:45-216) through mgic_hier_*: sources on every level with the
    level's dx, multilevel BiCGStab preconditioned by AMR V-cycles, QuadCFInterp + psi update with carried coarse-fine ghosts,
    composite norm -- against the oracle-backed twin: same BiCGStab iteration counts, dpsi norms to 1e-7, psi on every
    node to 1e-10 relative in max-norm (north-star tolerance), iteration 1's sources to 1e-13."""
    o, patches = bare_hierarchy(boxes)
    log = []
    norms_o, psi_o = hierarchy_nl_solve(o, patches, max_nl=3, log=log)
    levels = [[(q.boxes if q.boxes else (q.lo, q.hi)) for q in lv] for lv in patches]
    P = m.make_params(dict(o.params, max_NL_iterations=3))
    H = m.Hierarchy(ctx, P, levels)
    assert H.nodes == 1 + sum(len(lv) for lv in patches)
    flat = [q for lv in patches for q in lv]
    for n, q in enumerate(flat, start=1):
        lvl, lo, nn, cells = H.node_info(n)
        assert lo == tuple(q.lo) and nn == q.shape[::-1] and cells == int(q.mask().sum())
        assert np.array_equal(H.mask(n), q.mask())
    H.set_initial_conditions()
    got = []
    for it in range(3):
        nrm, its, st = H.nl_iteration()
        got.append(nrm)
        assert (its, st) == (log[it][0], log[it][1]), (it, its, st, log[it])
        if it == 0:   # sources of the first iteration: level-resolution Bowen-York terms
            o2, p2 = bare_hierarchy(boxes)
            o2.set_initial_conditions(); o2.set_coefs_and_rhs()
            assert relerr(H.download(0, "rhs"), o2.get("RHS")) < 1e-13
            for n, q in enumerate([q for lv in p2 for q in lv], start=1):
                q.set_initial_conditions(o2.params); q.set_coefs_and_rhs()
                assert relerr(H.download(n, "rhs"), q.var(8)) < 1e-13, n
                assert relerr(H.download(n, "aCoef"), q.get("A")) < 1e-13, n
    assert np.allclose(got, norms_o, rtol=1e-7), (got, norms_o)
    for n in range(H.nodes):
        assert relerr(H.download(n, "psi"), psi_o[n]) < 1e-10, n
    # the one-call form gives the same history
    H2 = m.Hierarchy(ctx, P, levels)
    assert np.allclose(H2.nl_solve(), got, rtol=1e-12)
    H.close(); H2.close()


def test_config_c4_at_its_full_size(ctx):
    """BASELINE config C4 as named -- 256^3 base, level 1 = 224 x 128 x 128 around both punctures, two 64^3 level-2 boxes
    (tools/bench_amr.py's hierarchy, 20.4 M composite cells) -- two nonlinear iterations against the oracle-backed twin at the
    same size: BiCGStab iteration counts, dpsi norms to 1e-7, psi on every node to 1e-10.  At this size level 1 is swept by the
    fused kernel's PATCH build and the base level by its plain build (>= fused_min_cells cells), which the small hierarchies
    above never reach.  About a minute of CPU for the twin."""
    from oracle import use_all_host_cores
    from tools.bench_amr import c4_boxes
    use_all_host_cores()
    n = 256
    l1, l2 = c4_boxes(n)
    boxes = {1: [l1], 2: [b for b in l2]}
    o, patches = bare_hierarchy(boxes, N=n, box=32)
    log = []
    norms_o, psi_o = hierarchy_nl_solve(o, patches, max_nl=2, log=log)
    P = m.make_params(dict(o.params, max_NL_iterations=2))
    H = m.Hierarchy(ctx, P, [[l1], [b for b in l2]])
    assert [H.node_info(q)[3] for q in range(H.nodes)] == [256 ** 3, 224 * 128 * 128, 64 ** 3, 64 ** 3]
    H.set_initial_conditions()
    got = []
    for it in range(2):
        nrm, its, st = H.nl_iteration()
        got.append(nrm)
        assert (its, st) == (log[it][0], log[it][1]), (it, its, st, log[it])
    assert np.allclose(got, norms_o, rtol=1e-7), (got, norms_o)
    psi = [H.download(q, "psi") for q in range(H.nodes)]
    for q in range(H.nodes):
        assert relerr(psi[q], psi_o[q]) < 1e-10, q
    H.close()
    o.close()
    # the same with level 1 swept by the per-colour kernel: identical bits
    ctx.set_option("fused_patch", 0)
    try:
        H = m.Hierarchy(ctx, P, [[l1], [b for b in l2]])
        H.set_initial_conditions()
        assert [H.nl_iteration()[0] for _ in range(2)] == got
        for q in range(H.nodes):
            assert np.array_equal(H.download(q, "psi"), psi[q]), q
        H.close()
    finally:
        ctx.set_option("fused_patch", 1)
