"""CPU tests that pin the ORACLE (the reference ships no tests / golden vectors -- SURVEY.md 4, 8c):
known-answer tests, the independent numpy twin, decomposition invariance, and the committed golden fixtures."""
import os

import numpy as np
import pytest

import np_twin as T
from oracle import Oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_trivial_kat_rhs_and_acoef_vanish():
    # no momentum, no spin, no scalar field => Aij = 0, rho_grad = 0, m = 0, lap(psi=1) = 0 => rhs = aCoef = 0
    o = Oracle(N=(16, 16, 16), max_grid_size=8, bh1_momentum=0.0, bh2_momentum=0.0, bh1_spin=0.0, bh2_spin=0.0,
               phi_amplitude=0.0)
    o.setup()
    assert np.all(o.get("RHS") == 0.0)
    assert np.all(o.get("A") == 0.0)
    o.load_rhs_zero_e()
    o.vcycle()
    assert np.all(o.get("E") == 0.0)


def test_trace_free_kat():
    o = Oracle(N=(32, 32, 32), max_grid_size=16)
    o.set_initial_conditions()
    tr = o.get("MGVAR0", comp=1) + o.get("MGVAR0", comp=4) + o.get("MGVAR0", comp=6)
    amax = max(np.abs(o.get("MGVAR0", comp=c)).max() for c in (1, 4, 6))
    assert np.abs(tr).max() < 1e-14 * max(amax, 1.0)


def test_survey_anchor_values_64():
    # SURVEY.md App. D (values of the survey's independent throwaway numpy restatement)
    o = Oracle(N=(64, 64, 64), numMGsmooth=4)
    assert o.setup() == 4  # MG levels 64, 32, 16, 8 with 16^3 boxes
    rhs, a = o.get("RHS"), o.get("A")
    assert np.isclose(np.abs(rhs).max(), 5.756623e-04, rtol=1e-6)
    assert np.isclose(a.min(), -3.126775e-03, rtol=1e-6) and np.isclose(a.max(), 2.433068e-04, rtol=1e-6)
    assert int((a > 0).sum()) == 32
    o.load_rhs_zero_e()
    hist = []
    for _ in range(3):
        o.vcycle()
        hist.append(np.abs(o.residual(0, True)).max())
    assert np.allclose(hist, [3.374e-07, 3.895e-09, 2.676e-11], rtol=2e-3)


def test_manufactured_solution_second_order():
    # dpsi = prod sin(pi x_d / L) vanishes on the Dirichlet boundary; L dpsi = (a + lap) dpsi
    errs = []
    for n in (16, 32):
        o = Oracle(N=(n, n, n), max_grid_size=8, L=1.0)
        o.setup()
        dx = 1.0 / n
        x = (np.arange(n) + 0.5) * dx
        Z, Y, X = np.meshgrid(x, x, x, indexing="ij")
        u = np.sin(np.pi * X) * np.sin(np.pi * Y) * np.sin(np.pi * Z)
        a = 0.3 * np.ones_like(u)
        o.set("A", a)
        o.set("E", u)
        lu = o.apply(0, True)
        exact = (0.3 - 3 * np.pi ** 2) * u
        errs.append(np.abs(lu - exact).max())
    assert errs[0] / errs[1] > 3.5  # second order: ratio -> 4


@pytest.mark.parametrize("bc", [dict(bc_lo=(0, 0, 0), bc_hi=(0, 0, 0)), dict(bc_lo=(1, 0, 1), bc_hi=(0, 1, 0))])
def test_decomposition_invariance_and_numpy_twin(bc):
    n = (32, 16, 24)
    rng = np.random.default_rng(7)
    e = rng.standard_normal((n[2], n[1], n[0]))
    r = rng.standard_normal((n[2], n[1], n[0]))
    results = []
    for box in (8, 4):
        o = Oracle(N=n, max_grid_size=box, L=50.0, bc_value=0.0, coefficient_average_type=0, **bc)
        o.setup()
        o.set("E", e); o.set("R", r)
        a, b, lam, dx = o.get("A"), o.get("B"), o.get("LAMBDA"), o.dims(0)[1]
        res = o.residual(0, True)
        app = o.apply(0, True)
        o.restrict(0)
        rc = o.get("R", 1)
        o.relax(0, 2)
        results.append((res, app, rc, o.get("E")))
    for x, y in zip(*results):
        assert np.array_equal(x, y)
    kw = dict(bc_lo=bc["bc_lo"], bc_hi=bc["bc_hi"])
    assert np.array_equal(results[0][0], T.residual(e, r, a, b, 1.0, -1.0, dx, **kw))
    assert np.array_equal(results[0][1], T.apply_op(e, a, b, 1.0, -1.0, dx, **kw))
    assert np.array_equal(results[0][2], T.restrict_residual(e, r, a, b, 1.0, -1.0, dx, **kw))
    assert np.array_equal(results[0][3], T.relax(e, r, a, b, lam, 1.0, -1.0, dx, 2, **kw))


def test_inhomogeneous_bc_twin():
    n = (16, 16, 16)
    rng = np.random.default_rng(3)
    e = rng.standard_normal((16, 16, 16)); r = rng.standard_normal((16, 16, 16))
    o = Oracle(N=n, max_grid_size=8, L=20.0, bc_value=0.75, bc_lo=(0, 1, 0), bc_hi=(1, 0, 0))
    o.setup()
    o.set("E", e); o.set("R", r)
    a, b, dx = o.get("A"), o.get("B"), o.dims(0)[1]
    kw = dict(bc_lo=(0, 1, 0), bc_hi=(1, 0, 0), value=0.75, homogeneous=False)
    assert np.array_equal(o.residual(0, False), T.residual(e, r, a, b, 1.0, -1.0, dx, **kw))
    assert np.array_equal(o.apply(0, False), T.apply_op(e, a, b, 1.0, -1.0, dx, **kw))


def test_periodic_exchange_matches_numpy_roll():
    # is_periodic = 1: ParseBC does nothing (Source/SetBCs.cpp:63) and the exchange copier wraps around the domain
    n = (16, 8, 24)
    o = Oracle(N=n, max_grid_size=8, L=20.0, is_periodic=1)
    o.setup()
    rng = np.random.default_rng(2)
    e = rng.standard_normal((n[2], n[1], n[0])); r = rng.standard_normal(e.shape)
    o.set("E", e); o.set("R", r)
    a, dx = o.get("A"), o.dims(0)[1]
    t = 2 * e
    lap = ((np.roll(e, -1, 2) + np.roll(e, 1, 2)) - t) + ((np.roll(e, -1, 1) + np.roll(e, 1, 1)) - t) + \
          ((np.roll(e, -1, 0) + np.roll(e, 1, 0)) - t)
    assert np.array_equal(o.residual(0, True), (r - 1.0 * a * e) + lap * (1.0 / (dx * dx)) * (-1.0) * 1.0)
    assert np.array_equal(o.apply(0, True), 1.0 * a * e - lap * (1.0 / (dx * dx)) * (-1.0) * 1.0)


def test_mg_depth_follows_box_size():
    # Factory.cpp:168-172: depth limit = boxes coarsenable by 2^depth * s_maxCoarse
    for box, depths in ((8, 3), (16, 4), (32, 5)):
        o = Oracle(N=(32, 32, 32), max_grid_size=box)
        assert o.setup() == depths
    # [Chombo 3.2] MultiGrid::define: maxDepth = D >= 1 builds D operators, D = 0 the finest one alone
    for D, depths in ((0, 1), (1, 1), (2, 2), (3, 3), (9, 4)):
        o = Oracle(N=(32, 32, 32), max_grid_size=16, preCondSolverDepth=D)
        assert o.setup() == depths


def test_coarse_average_matches_twin():
    for typ in (0, 1):
        o = Oracle(N=(32, 32, 32), max_grid_size=16, coefficient_average_type=typ)
        nd = o.setup()
        a = o.get("A")
        for d in range(1, nd):
            assert np.array_equal(o.get("A", d), T.coarse_average(a, 2 ** d, typ == 1))
            assert np.all(o.get("B", d) == 1.0)


@pytest.mark.parametrize("name,over", [
    ("kernels_16_dirichlet", dict(N=(16, 16, 16), max_grid_size=8, L=40.0)),
    ("kernels_24x16x32_neumann", dict(N=(24, 16, 32), max_grid_size=8, L=60.0, bc_lo=(1, 0, 1), bc_hi=(0, 1, 1),
                                      bc_value=0.25, coefficient_average_type=0)),
])
def test_golden_kernels(name, over):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    o = Oracle(**over)
    assert o.setup() == int(g["depths"])
    assert np.array_equal(o.get("RHS"), g["rhs"]) and np.array_equal(o.get("A"), g["a"])
    o.set("E", g["e"]); o.set("R", g["r"])
    assert np.array_equal(o.residual(0, True), g["residual_h"])
    assert np.array_equal(o.residual(0, False), g["residual_i"])
    assert np.array_equal(o.apply(0, True), g["apply_h"])
    o.restrict(0)
    assert np.array_equal(o.get("R", 1), g["restrict"])
    o.relax(0, 4)
    assert np.array_equal(o.get("E"), g["relax4"])


@pytest.mark.parametrize("name,cycles,over", [
    ("solver_32_v22", 5, dict(N=(32, 32, 32), max_grid_size=16, numMGsmooth=2)),
    ("solver_32_v44_box8", 4, dict(N=(32, 32, 32), max_grid_size=8, numMGsmooth=4)),
])
def test_golden_solver(name, cycles, over):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    o = Oracle(**over)
    o.setup()
    o.load_rhs_zero_e()
    hist = [o.norm(0, "R", 0)]
    for _ in range(cycles):
        o.vcycle()
        hist.append(float(np.abs(o.residual(0, True)).max()))
    assert np.allclose(hist, g["vcycle_resnorm"], rtol=1e-9, atol=1e-18)
    o2 = Oracle(**over)
    o2.set_initial_conditions()
    nl = o2.nl_solve()
    assert len(nl) == len(g["nl_dpsi_norms"])
    assert np.allclose(nl[:2], g["nl_dpsi_norms"][:2], rtol=1e-8)
    assert nl[-1] < o2.params["tolerance"]


# ---- AMR level > 0: the operator's own coarse-fine code (homogeneousCFInterp, Operator.cpp:156,296) ----------------------
PATCHES = {
    "interior": dict(n=(32, 32, 32), lo=(8, 8, 8), hi=(23, 23, 23), bc_lo=(0, 0, 0), bc_hi=(0, 0, 0)),
    "on_faces": dict(n=(32, 32, 48), lo=(0, 8, 16), hi=(15, 31, 47), bc_lo=(1, 0, 0), bc_hi=(0, 1, 0)),
    "slab": dict(n=(40, 24, 24), lo=(10, 0, 4), hi=(25, 7, 19), bc_lo=(0, 1, 0), bc_hi=(0, 0, 0)),
}


def test_homogeneous_cf_interp_is_the_parabola_with_a_zero_coarse_value():
    """ghost = 2/3 near - 1/5 far at refinement ratio 2 ([Chombo] INTERPHOMO restated); exact for any parabola that
    vanishes at the coarse cell centre."""
    from oracle import interp_homo
    for dx in (0.1, 0.390625, 1.0):
        assert abs(interp_homo(dx, 2 * dx, 0.0, 1.0) - 2.0 / 3.0) < 4e-15
        assert abs(interp_homo(dx, 2 * dx, 1.0, 0.0) + 1.0 / 5.0) < 4e-15
        # far cell at 0, near at dx, ghost at 2dx, coarse centre at 2.5dx
        q = lambda t: (t - 2.5 * dx) * (0.3 * t + 1.7)
        assert abs(interp_homo(dx, 2 * dx, q(0.0), q(dx)) - q(2 * dx)) < 1e-13 * (1 + abs(q(2 * dx)))


@pytest.mark.parametrize("name", sorted(PATCHES))
@pytest.mark.parametrize("mgs", [8, 16])
def test_patch_level_matches_numpy_twin(name, mgs):
    """One AMR patch: the boxed oracle (explicit homogeneousCFInterp ghost fill, then exchange, then BC, per colour pass)
    equals the single-array twin that evaluates the same ghost values on the fly -- bit for bit, for any box size,
    with physical and coarse-fine faces mixed."""
    from oracle import OraclePatch
    c = PATCHES[name]
    dx = 0.5
    P = OraclePatch(c["n"], c["lo"], c["hi"], dx, max_grid_size=mgs, bc_lo=c["bc_lo"], bc_hi=c["bc_hi"])
    rng = np.random.default_rng(4)
    e, r = rng.standard_normal(P.shape), rng.standard_normal(P.shape)
    a, b = 0.1 * rng.standard_normal(P.shape) - 0.5, 1 + 0.1 * rng.standard_normal(P.shape)
    for f, x in (("E", e), ("R", r), ("A", a), ("B", b)):
        P.set(f, x)
    kw = dict(bc_lo=c["bc_lo"], bc_hi=c["bc_hi"], cf_lo=tuple(c["lo"][d] > 0 for d in range(3)),
              cf_hi=tuple(c["hi"][d] < c["n"][d] - 1 for d in range(3)), dx_crse=2 * dx)
    lam = T.compute_lambda(a, 1.0, -1.0, dx)
    assert np.array_equal(lam, P.get("LAMBDA"))
    t = e.copy()
    for _ in range(2):
        for colour in (0, 1):
            P.gsrb_color(colour)
            t = T.gsrb_colour(t, r, a, b, lam, 1.0, -1.0, dx, colour, origin=c["lo"], **kw)
            assert np.array_equal(t, P.get("E"))
    assert np.array_equal(P.restrict(), T.restrict_residual(t, r, a, b, 1.0, -1.0, dx, **kw))
    P.set("E", e)
    P.precond()
    assert np.array_equal(P.get("E"), T.relax(r * lam, r, a, b, lam, 1.0, -1.0, dx, 2, origin=c["lo"], **kw))


def _quadratic(x, y, z):
    return 0.3 + 0.7 * x - 0.2 * y + 0.5 * z + 0.11 * x * x - 0.23 * y * y + 0.05 * z * z + 0.4 * x * y - 0.31 * y * z + 0.17 * x * z


@pytest.mark.parametrize("lo,hi", [((8, 8, 8), (23, 23, 23)), ((2, 2, 2), (29, 29, 29))])
def test_quad_cf_interp_is_exact_for_quadratics(lo, hi):
    """[Chombo] QuadCFInterp restated (tangential second-order Taylor expansion of the coarse field incl. the mixed term,
    one-sided next to domain faces, then the normal parabola): with ghost values from it the 7-point operator of ANY
    quadratic polynomial is exact in every cell of the patch, also next to the coarse-fine faces, edges and corners."""
    from oracle import OraclePatch
    n, dx = (32, 32, 32), 0.25
    P = OraclePatch(n, lo, hi, dx, max_grid_size=8)
    k, j, i = np.meshgrid(*(np.arange(lo[d], hi[d] + 1) for d in (2, 1, 0)), indexing="ij")
    K, J, I = np.meshgrid(*(np.arange(n[d] // 2) for d in (2, 1, 0)), indexing="ij")
    P.set("E", _quadratic((i + .5) * dx, (j + .5) * dx, (k + .5) * dx))
    P.set_coarse(_quadratic((I + .5) * 2 * dx, (J + .5) * 2 * dx, (K + .5) * 2 * dx))
    P.set("A", np.zeros(P.shape)); P.set("B", np.ones(P.shape)); P.set("R", np.zeros(P.shape))
    lap = 2 * (0.11 - 0.23 + 0.05)
    assert np.abs(P.amr_operator_nf(True) - lap).max() < 1e-11      # alpha*a = 0, beta = -1: L(phi) = laplacian
    assert np.abs(P.amr_residual_nf(True) + lap).max() < 1e-11


@pytest.mark.parametrize("name", sorted(PATCHES))
def test_patch_amr_operator_matches_numpy_twin(name):
    """AMROperatorNF / AMRResidualNF (QuadCFInterp from the coarser level, then applyOpI / residualI): boxed oracle ==
    single-array twin, bit for bit, homogeneous and inhomogeneous physical BC."""
    from oracle import OraclePatch
    c = PATCHES[name]
    dx, val = 0.25, 0.3
    P = OraclePatch(c["n"], c["lo"], c["hi"], dx, max_grid_size=8, bc_lo=c["bc_lo"], bc_hi=c["bc_hi"], bc_value=val)
    rng = np.random.default_rng(8)
    e, r = rng.standard_normal(P.shape), rng.standard_normal(P.shape)
    a, b = 0.1 * rng.standard_normal(P.shape) - 0.5, 1 + 0.1 * rng.standard_normal(P.shape)
    crse = rng.standard_normal(tuple(x // 2 for x in c["n"][::-1]))
    for f, x in (("E", e), ("R", r), ("A", a), ("B", b)):
        P.set(f, x)
    P.set_coarse(crse)
    kw = dict(bc_lo=c["bc_lo"], bc_hi=c["bc_hi"], value=val)
    for homog in (True, False):
        assert np.array_equal(P.amr_operator_nf(homog),
                              T.amr_operator_nf(e, crse, a, b, 1.0, -1.0, dx, c["lo"], c["hi"], c["n"], homogeneous=homog, **kw))
        assert np.array_equal(P.amr_residual_nf(homog),
                              T.amr_residual_nf(e, crse, r, a, b, 1.0, -1.0, dx, c["lo"], c["hi"], c["n"], homogeneous=homog, **kw))


def two_level_amr_vcycles(level0_residual, level0_vcycle, patch_relax, patch_residual_nf, a0_shape, shape, sl, rhs0, rhs1, cycles):
    """A two-level AMR V-cycle iteration in the structure of [Chombo] AMRMultiGrid::AMRVCycle, written
    over the operator-level primitives only: relax on the patch (homogeneousCFInterp), AMRResidualNF (QuadCFInterp),
    averaging to the covered coarse cells, MultiGrid::oneCycle on the base level, piecewise-constant prolongation,
    AMRUpdateResidual, post-relaxation.  Returns the history of (max|r_fine|, max|r_coarse|) and the two fields."""
    rep = lambda x: np.repeat(np.repeat(np.repeat(x[sl], 2, 0), 2, 1), 2, 2)
    phi0, phi1 = np.zeros(a0_shape), np.zeros(shape)
    z0, z1 = np.zeros(a0_shape), np.zeros(shape)
    hist = []
    for _ in range(cycles):
        r1 = patch_residual_nf(phi1, phi0, rhs1, False)
        r0 = level0_residual(phi0, rhs0)
        r0[sl] = T.coarse_average(r1, 2, False)          # covered cells: the averaged fine residual
        hist.append((np.abs(r1).max(), np.abs(r0).max()))
        e1 = patch_relax(z1, r1)                         # down: pre-relaxation of a zero correction
        rc = r0.copy()
        rc[sl] = T.coarse_average(patch_residual_nf(e1, z0, r1, True), 2, False)   # AMRRestrictS
        e0 = level0_vcycle(rc)                           # base level: MultiGrid::oneCycle
        e1 = e1 + rep(e0)                                # up: AMRProlongS
        e1 = e1 + patch_relax(z1, patch_residual_nf(e1, e0, r1, True))   # AMRUpdateResidual + post-relaxation
        phi1, phi0 = phi1 + e1, phi0 + e0
    return hist, phi0, phi1


def test_two_level_amr_vcycle_converges():
    """The operator-level AMR pieces work together the way AMRVCycle uses them: on a Bowen-York base
    level with one refined box the composite residual falls by more than an order of magnitude per V(2,2) cycle."""
    from oracle import OraclePatch
    N, L = 32, 100.0
    o = Oracle(N=(N, N, N), max_grid_size=16, numMGsmooth=2, L=L)
    o.setup()
    a0, b0, rhs0 = o.get("A"), o.get("B"), o.get("RHS")
    clo, chi = (8, 8, 8), (23, 23, 23)
    lo, hi = tuple(2 * x for x in clo), tuple(2 * x + 1 for x in chi)
    P = OraclePatch((2 * N,) * 3, lo, hi, L / N / 2, max_grid_size=16)
    sl = tuple(slice(clo[d], chi[d] + 1) for d in (2, 1, 0))
    rep = lambda x: np.repeat(np.repeat(np.repeat(x[sl], 2, 0), 2, 1), 2, 2)
    P.set("A", rep(a0)); P.set("B", rep(b0))

    def patch_residual_nf(phi, coarse, rhs, homog):
        P.set("E", phi); P.set("R", rhs); P.set_coarse(coarse)
        return P.amr_residual_nf(homog)

    def patch_relax(e, r):
        P.set("E", e); P.set("R", r); P.relax(2)
        return P.get("E")

    def level0_residual(phi, rhs):
        o.set("E", phi); o.set("R", rhs)
        return o.residual(0, False)

    def level0_vcycle(r):
        o.set("R", r); o.set("E", np.zeros_like(r)); o.vcycle()
        return o.get("E")

    hist, _, _ = two_level_amr_vcycles(level0_residual, level0_vcycle, patch_relax, patch_residual_nf, a0.shape, P.shape, sl,
                                       rhs0, rep(rhs0), 6)
    tot = [max(h) for h in hist]
    assert all(tot[i + 1] < 0.1 * tot[i] for i in range(len(tot) - 1)), tot
    assert tot[-1] < 1e-6 * tot[0]
