"""Pins of the oracle against the REFERENCE's own code (the source terms).

oracle/_ref/libmgic_ref.so is the reference's Source/SetLevelData.cpp + Source/SetBinaryBH.H + MyPhiFunction.H compiled
unmodified (recipe: oracle/Makefile target `ref`, loader oracle/pyref.py).  Where /root/reference exists (the build
container) the live tests run the reference beside the oracle; everywhere, the oracle is held to the committed outputs of
the reference (tests/golden/reference_sources_16.npz, made by tests/golden/make_reference_golden.py).  Bar: bit-exact."""
import json
import os
import sys

import numpy as np
import pytest

from oracle import Oracle, pyref

GOLD = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLD)
live = pytest.mark.skipif(not pyref.available(), reason="no /root/reference and no prebuilt oracle/_ref/libmgic_ref.so")


def gold():
    g = np.load(os.path.join(GOLD, "reference_sources_16.npz"))
    p = json.loads(str(g["params"]))
    over = {k: (tuple(v) if isinstance(v, list) else v) for k, v in p.items()}
    return g, over


def test_oracle_reproduces_the_reference_golden_bit_for_bit():
    g, over = gold()
    o = Oracle(**over)
    o.set_initial_conditions()
    o.set_coefs_and_rhs()
    for c in range(8):
        assert np.array_equal(o.get_ghosted("MGVAR0", 3, comp=c), g["mgvars_ghost3"][c]), c
    assert np.array_equal(o.get("RHS"), g["rhs"]) and np.array_equal(o.get("A"), g["acoef"])
    assert np.array_equal(o.get("B"), g["bcoef"]) and np.all(g["bcoef"] == 1.0)
    o.set_coefs_and_rhs(float(g["constant_K"]))
    assert np.array_equal(o.get("RHS"), g["rhs_K"]) and np.array_equal(o.get("A"), g["acoef_K"])
    assert float(g["m_K"]) == (2.0 / 3.0) * (float(g["constant_K"]) ** 2)      # set_m_value, rho = 0
    # state 2: psi += dpsi.  The oracle takes dpsi without ghost cells (its domain-face ghosts are whatever the solver's last
    # BC fill left, here nothing), so this compares what does not read them: psi and aCoef everywhere, rhs (which holds the
    # Laplacian of psi) one cell inside the domain.  Ghost-dependent values are compared in the live test below, after a
    # real solve, and on the GPU (tests/test_gpu_reference_golden.py), whose update applies the fill itself.
    o.set_coefs_and_rhs()
    o.set("DPSI", g["dpsi_ghost3"][3:-3, 3:-3, 3:-3])
    o.update_psi0()
    o.set_coefs_and_rhs()
    assert np.array_equal(o.get("MGVAR0", comp=0), g["psi_after_ghost3"][3:-3, 3:-3, 3:-3])
    assert np.array_equal(o.get("A"), g["acoef_after"])
    inner = (slice(1, -1),) * 3
    assert np.array_equal(o.get("RHS")[inner], g["rhs_after"][inner])


def test_golden_point_values_have_the_bowen_york_structure():
    """sanity of the fixture itself: trace-free A_ij, psi_bh = m1/r1 + m2/r2, Gaussian phi (note: divided by the wavelength,
    not its square -- MyPhiFunction.H:15)"""
    g, over = gold()
    for loc, A, psi_bh, phi in zip(g["point_locs"], g["point_Aij"], g["point_psi_bh"], g["point_phi"]):
        assert abs(A[0] + A[3] + A[5]) < 1e-17 + 1e-15 * np.abs(A).max()
        r1 = np.sqrt((loc[0] - over["bh1_offset"]) ** 2 + loc[1] ** 2 + loc[2] ** 2)
        r2 = np.sqrt((loc[0] - over["bh2_offset"]) ** 2 + loc[1] ** 2 + loc[2] ** 2)
        assert psi_bh == over["bh1_bare_mass"] / r1 + over["bh2_bare_mass"] / r2
        assert np.isclose(phi, over["phi_amplitude"] * np.exp(-(loc @ loc) / over["phi_wavelength"]), rtol=1e-12, atol=0)


@live
def test_golden_is_what_the_reference_computes():
    import make_reference_golden as mk
    g, _ = gold()
    fresh = mk.generate()
    for k, v in fresh.items():
        if k != "params":
            assert np.array_equal(np.asarray(v), g[k]), k


@live
@pytest.mark.parametrize("over", [
    dict(N=(16, 16, 16), max_grid_size=8, L=40.0),
    dict(N=(24, 16, 32), max_grid_size=8, L=60.0),                       # non-cubic domain, cubic cells (PoissonParameters.cpp:70-85)
    dict(N=(32, 32, 32), max_grid_size=16),                              # params.txt's L = 100
    dict(N=(16, 16, 16), max_grid_size=8, L=40.0, bh1_spin=0.0, bh2_spin=0.0, bh1_momentum=0.0, bh2_momentum=0.0,
         phi_amplitude=0.0),                                             # the trivial known-answer case
    dict(N=(16, 16, 16), max_grid_size=4, L=64.0, bh1_bare_mass=0.7, bh1_spin=-0.2, bh1_momentum=0.11, bh1_offset=7.0,
         bh2_bare_mass=0.3, bh2_spin=0.4, bh2_momentum=-0.02, bh2_offset=-13.0, phi_amplitude=0.02, phi_wavelength=90.0,
         G_Newton=0.5),
], ids=["c16", "noncubic", "c32_L100", "trivial", "other_physics"])
def test_oracle_source_terms_equal_the_reference(over):
    """initial data, rhs, aCoef, bCoef at NL iteration 1 and again after one real linear solve + set_update_psi0:
    the boxed oracle (any box size) against the reference's one-box run, bit for bit, ghost cells included"""
    o = Oracle(**over)
    o.setup()
    mg, rhs, a, b = pyref.set_level_data(o.params)
    for c in range(8):
        assert np.array_equal(o.get_ghosted("MGVAR0", 3, comp=c), mg[c]), c
    assert np.array_equal(o.get("RHS"), rhs) and np.array_equal(o.get("A"), a) and np.array_equal(o.get("B"), b)
    _, rhsK, aK, _ = pyref.set_level_data(o.params, constant_K=0.7)
    o.set_coefs_and_rhs(0.7)
    assert np.array_equal(o.get("RHS"), rhsK) and np.array_equal(o.get("A"), aK)
    o.set_coefs_and_rhs()
    o.define_solver()
    o.load_rhs_zero_e()
    o.outer_solve()
    d = o.get_ghosted("DPSI", 3)
    o.update_psi0()
    o.set_coefs_and_rhs()
    mg2, rhs2, a2, _ = pyref.set_level_data(o.params, dpsi_ghosted=d)
    assert np.array_equal(o.get_ghosted("MGVAR0", 3, comp=0), mg2[0])
    assert np.array_equal(o.get("RHS"), rhs2) and np.array_equal(o.get("A"), a2)


# ---- getPoissonParameters (Source/PoissonParameters.cpp:26-131, compiled unmodified) against the two readers of this repo
REF_PARAMS = os.path.join(pyref.REFERENCE, "params.txt")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PARAM_CASES = {
    "params.txt": (),
    "single_level_512": ("max_level = 0", "N = 512 512 512", "max_grid_size = 32"),
    "noncubic": ("N = 32 48 64", "L = 50.0", "max_level = 2", "coefficient_average_type = arithmetic"),
    "periodic": ("is_periodic = 1", "verbosity = 5"),
}


def host_mirror_params(path, overrides, tmp_path_factory):
    """the C++ host mirror's getPoissonParameters (host/PoissonParameters.H) through a small probe program"""
    import subprocess
    exe = str(tmp_path_factory.getbasetemp() / "params_probe")
    if not os.path.exists(exe):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O1", "-std=c++17", "-I" + os.path.join(ROOT, "include"),
                               "-I" + os.path.join(ROOT, "mg_ic_code_b200", "host"), os.path.join(HERE, "data", "params_probe.cpp"),
                               "-L" + os.path.join(ROOT, "mg_ic_code_b200", "lib"), "-lmgic_b200",
                               "-Wl,-rpath," + os.path.join(ROOT, "mg_ic_code_b200", "lib"), "-o", exe])
    r = subprocess.run([exe, path] + [o.replace(" = ", "=", 1) for o in overrides], capture_output=True, text=True)
    return r.returncode, r.stdout, r.stderr


@live
@live
@pytest.mark.parametrize("over,K", [(dict(N=(12, 8, 10), L=30.0), -0.3), (dict(N=(16, 16, 16), L=100.0, phi_amplitude=0.3, bh1_spin=-0.4), 0.0)])
def test_output_data_equals_the_reference(over, K):
    """set_output_data (Source/SetLevelData.cpp:343-396; what output_final_data writes into the GRChombo checkpoint with three
    ghost layers, Source/WriteOutput.H:180): chi = psi_0^-4, A~_ij = A-_ij chi^(3/2), phi, K, h_ii = lapse = 1, the rest 0 --
    the oracle's restatement against the reference's own function, all 32 components, bit for bit, ghost cells included"""
    from oracle import default_params, output_box
    P = default_params(**over)
    N = P["N"]
    d = 0.01 * np.random.default_rng(2).standard_normal((N[2] + 6, N[1] + 6, N[0] + 6))
    ref = pyref.output_data(P, constant_K=K, dpsi_ghosted=d)
    got = output_box(P, P["L"] / N[0], (-3, -3, -3), (N[0] + 2, N[1] + 2, N[2] + 2), 1.0 + d, K)
    assert np.array_equal(ref, got)
    assert np.all(got[1] == 1) and np.all(got[4] == 1) and np.all(got[6] == 1) and np.all(got[18] == 1) and np.all(got[7] == K)
    assert not got[[2, 3, 5, 14, 15, 16, 17, 19, 20, 21, 22, 23, 24, 26, 27, 28, 29, 30, 31]].any()


@live
@pytest.mark.parametrize("over", [
    dict(N=(16, 16, 16), L=40.0), dict(N=(24, 16, 32), L=100.0),
    dict(N=(16, 16, 16), L=6.0, phi_amplitude=0.7),          # rho_grad of order one: |rho| is not its integer part
    dict(N=(16, 16, 16), L=64.0, bh1_bare_mass=0.7, bh1_spin=-0.2, bh1_momentum=0.11, bh1_offset=7.0, bh2_bare_mass=0.3,
         bh2_spin=0.4, bh2_momentum=-0.02, bh2_offset=-13.0, phi_amplitude=0.02, phi_wavelength=90.0, G_Newton=0.5),
], ids=["c16", "noncubic", "strong_field", "other_physics"])
def test_regrid_condition_and_constant_K_integrand_equal_the_reference(over):
    """set_regrid_condition (Source/SetLevelData.cpp:188-240, what set_grids tags on) and set_constant_K_integrand (:128-186)
    on freshly initialised data: the oracle's restatement against the reference's own functions, bit for bit -- on the
    whole level, and on sub-boxes / a refined level's index box (the values depend on the cell position alone)"""
    from oracle import condition_box, default_params
    P = default_params(**over)
    N, dx = P["N"], P["L"] / P["N"][0]
    for mode in (0, 1):
        ref = pyref.condition(P, mode)
        assert np.array_equal(condition_box(P, dx, (0, 0, 0), (N[0] - 1, N[1] - 1, N[2] - 1), mode), ref), mode
        sub = condition_box(P, dx, (3, 2, 5), (10, 9, 12), mode)
        assert np.array_equal(sub, ref[5:13, 2:10, 3:11]), mode
    # the level-1 index space: the reference run on the refined grid (dx / 2, 2 N cells)
    P2 = dict(P, N=tuple(2 * x for x in N))
    ref2 = pyref.condition(P2, 0, dx=dx / 2)
    fine = condition_box(P, dx / 2, (8, 4, 6), (19, 17, 15), 0)
    assert np.allclose(fine, ref2[6:16, 4:18, 8:20], rtol=1e-13, atol=0)   # L / N[0] / 2 vs (L / 2N[0]): the same number


@pytest.mark.skipif(not os.path.exists(REF_PARAMS), reason="needs the reference's params.txt")
@pytest.mark.parametrize("case", list(PARAM_CASES))
def test_parameter_readers_agree_with_the_reference(case, tmp_path_factory):
    import mg_ic_code_b200 as m
    over = PARAM_CASES[case]
    ref = pyref.get_poisson_parameters(REF_PARAMS, over)
    # what the reference derives (PoissonParameters.cpp:64-85,110-128)
    assert ref["numLevels"] == ref["maxLevel"] + 1 and ref["refRatio0"] == ref["refRatioLast"] == 2
    assert ref["domainLo"] == [0, 0, 0] and ref["domainHi"] == [n - 1 for n in ref["nCells"]]
    assert ref["domainLength"] == [ref["coarsestDx"] * n for n in ref["nCells"]] and ref["probHi"] == ref["domainLength"]
    # the C++ host mirror (what the driver poisson_solver_b200 runs)
    rc, out, err = host_mirror_params(REF_PARAMS, over, tmp_path_factory)
    assert rc == 0, err
    host = json.loads(out.strip().splitlines()[-1])
    for k, v in host.items():
        if k == "nRefRatio":
            assert v == ref["numLevels"]
        elif k == "periodic":
            assert [v] * 3 == ref["periodic"] == ref["domainPeriodic"]
        else:
            assert v == ref[k], k
    # the Python reader used by the tests and bench.py
    P = m.read_params(REF_PARAMS, [o.replace(" = ", "=", 1) for o in over])
    assert list(P.N) == ref["nCells"] and P.L / P.N[0] == ref["coarsestDx"]
    assert (P.max_level, P.block_factor, P.max_grid_size, P.is_periodic) == (ref["maxLevel"], ref["blockFactor"], ref["maxGridSize"],
                                                                             ref["periodic"][0])
    assert P.coefficient_average_type == ref["coefficient_average_type"] and P.verbosity == ref["verbosity"]
    for k in ("alpha", "beta", "G_Newton", "phi_amplitude", "phi_wavelength", "bh1_bare_mass", "bh2_bare_mass", "bh1_spin",
              "bh2_spin", "bh1_momentum", "bh2_momentum", "bh1_offset", "bh2_offset"):
        assert getattr(P, k) == ref[k], k


@live
@pytest.mark.skipif(not os.path.exists(REF_PARAMS), reason="needs the reference's params.txt")
def test_parameter_errors_are_the_references(tmp_path, tmp_path_factory):
    """a missing required key and a bad coefficient_average_type stop the reference (MayDay::Error / ParmParse) and the
    host mirror alike; an absent coefficient_average_type is the reference's "bogus default" -1"""
    import mg_ic_code_b200 as m
    with pytest.raises(RuntimeError, match="bad coefficient_average_type in input"):
        pyref.get_poisson_parameters(REF_PARAMS, ("coefficient_average_type = geometric",))
    rc, _, err = host_mirror_params(REF_PARAMS, ("coefficient_average_type = geometric",), tmp_path_factory)
    assert rc != 0 and "bad coefficient_average_type in input" in err
    with pytest.raises(m.MgicError, match="bad coefficient_average_type in input"):
        m.read_params(REF_PARAMS, ["coefficient_average_type=geometric"])
    text = open(REF_PARAMS).read()
    for key in ("alpha", "bh2_offset", "max_level", "L", "fill_ratio", "buffer_size", "refine_threshold", "is_periodic"):
        cut = tmp_path / f"no_{key}.txt"
        cut.write_text("\n".join(l for l in text.splitlines() if not l.split("=")[0].strip() == key))
        with pytest.raises(RuntimeError, match=key):
            pyref.get_poisson_parameters(cut)
        rc, _, err = host_mirror_params(str(cut), (), tmp_path_factory)
        assert rc != 0 and key in err, key
        with pytest.raises(m.MgicError, match=key):
            m.read_params(str(cut), strict=True)
    assert m.read_params(REF_PARAMS, strict=True).numMGsmooth == 4 and m.read_params(REF_PARAMS, strict=True).verbosity == 2
    bare = tmp_path / "no_solver_keys.txt"
    solver_keys = ("numMGsmooth", "numMGIterations", "max_iterations", "max_NL_iterations", "tolerance", "verbosity")
    bare.write_text("\n".join(l for l in text.splitlines() if l.split("=")[0].strip() not in solver_keys))
    P = m.read_params(str(bare), strict=True)       # the in-code defaults of Main_PoissonSolver.cpp:106-126 / PoissonParameters.cpp:59
    assert (P.numMGIterations, P.numMGsmooth, P.preCondSolverDepth, P.max_iterations, P.max_NL_iterations, P.verbosity) == (1, 4, -1, 10, 4, 3)
    assert P.tolerance == 1.0e-7 and pyref.get_poisson_parameters(bare)["verbosity"] == 3
    none = tmp_path / "no_avg.txt"
    none.write_text("\n".join(l for l in text.splitlines() if not l.startswith("coefficient_average_type")))
    assert pyref.get_poisson_parameters(none)["coefficient_average_type"] == -1
    rc, out, _ = host_mirror_params(str(none), (), tmp_path_factory)
    assert rc == 0 and json.loads(out.strip().splitlines()[-1])["coefficient_average_type"] == -1
    assert m.read_params(str(none)).coefficient_average_type == -1


# ---- the operator class: Source/VariableCoeffPoissonOperator.cpp + Source/SetBCs.cpp, compiled unmodified (rows a6 - a12, a14).
# Pinned: the reference's own orchestration -- which ghost cells ParseBC fills with what, the exchange, the per-colour
# sequence of levelGSRB, residualI / applyOpI / restrictResidual (shifted bounds, zeroed coarse residual) / preCond /
# levelJacobi, the lambda formula.  NOT pinned by this: the .ChF kernels (the C restatement is linked in their place) and
# Chombo's DiriBC / NeumBC / exchange (restated in the stand-in).
OP_CASES = {
    "c16": dict(N=(16, 16, 16), max_grid_size=8, L=40.0),
    "one_box": dict(N=(8, 8, 8), max_grid_size=8, L=10.0),
    "neumann_inhomogeneous": dict(N=(24, 16, 32), max_grid_size=8, L=60.0, bc_lo=(1, 0, 1), bc_hi=(0, 1, 1), bc_value=0.25),
    "all_neumann": dict(N=(16, 16, 16), max_grid_size=4, L=40.0, bc_lo=(1, 1, 1), bc_hi=(1, 1, 1), bc_value=-0.5),
    "wide_boxes16": dict(N=(48, 16, 32), max_grid_size=16, L=100.0, alpha=0.7, beta=-1.3),
}


@live
@pytest.mark.parametrize("case", list(OP_CASES))
def test_oracle_operator_equals_the_reference_class(case):
    o = Oracle(**OP_CASES[case])
    o.setup()
    R = pyref.ReferenceOperator(o.params)
    n = o.params["N"]
    assert R.num_boxes == int(np.prod([-(-n[d] // o.params["max_grid_size"]) for d in range(3)]))
    rng = np.random.default_rng(7)
    e, r = rng.standard_normal(R.shape), rng.standard_normal(R.shape)
    b = 1.0 + 0.3 * rng.random(R.shape)                 # a genuinely variable bCoef (the reference's is 1)
    for bcoef in (o.get("B"), b):
        R.set("A", o.get("A")); R.set("B", bcoef)
        o.set("B", bcoef)
        assert np.array_equal(R.get("LAMBDA"), o.get("LAMBDA"))          # resetLambda (:220-249): bCoef does not enter
        for homog in (True, False):
            o.set("E", e); o.set("R", r); R.set("E", e); R.set("R", r)
            assert np.array_equal(R.residual(homog), o.residual(0, homog)), ("residualI", homog)
            assert np.array_equal(R.apply(homog), o.apply(0, homog)), ("applyOpI", homog)
        o.relax(0, 1); R.relax(1)
        assert np.array_equal(R.get("E"), o.get("E")), "levelGSRB"
        o.relax(0, 3); R.relax(3)
        assert np.array_equal(R.get("E"), o.get("E")), "relax(3)"
        o.restrict(0)
        assert np.array_equal(R.restrict(), o.get("R", 1)), "restrictResidual"
        o.precond(0); R.precond()
        assert np.array_equal(R.get("E"), o.get("E")), "preCond"
        # levelJacobi (:360-385), reached through AMRPoissonOp::relax with s_relaxMode = 4: the expression the GPU test of
        # mgic_op_level_jacobi checks against (test_gpu_parity.py::test_level_jacobi)
        o.set("E", e); o.set("R", r); R.set("E", e); R.set("R", r)
        R.relax(1, mode=4)
        assert np.array_equal(R.get("E"), e + 0.5 * (o.residual(0, True) * o.get("LAMBDA"))), "levelJacobi"


@live
def test_reference_operator_aborts_like_the_reference():
    """the relaxation modes the reference does not implement stop with MayDay::Abort (VariableCoeffPoissonOperator.cpp:334-358)"""
    import ctypes as C
    o = Oracle(**OP_CASES["one_box"])
    o.setup()
    R = pyref.ReferenceOperator(o.params)
    assert R.L.ref_op_relax_status(R.h, 1, 1) == 0
    for mode in (0, 2, 3, 5):
        msg = C.create_string_buffer(256)
        assert R.L.ref_op_relax_status_msg(R.h, 1, mode, msg, 256) == 1
        assert b"Not implemented" in msg.value


# ---- the factory: Source/VariableCoeffPoissonOperatorFactory.cpp, compiled unmodified, through defineOperatorFactory (row a17)
FACTORY_CASES = {
    "params.txt_64_box16": dict(N=(64, 64, 64), max_grid_size=16),
    "arithmetic": dict(N=(32, 32, 32), max_grid_size=8, coefficient_average_type=0),
    "noncubic": dict(N=(24, 16, 32), max_grid_size=8, L=60.0),
    "ragged_48_box16": dict(N=(48, 48, 48), max_grid_size=16, L=30.0),
    "one_box_8": dict(N=(8, 8, 8), max_grid_size=8, L=10.0),
    "box32": dict(N=(64, 64, 64), max_grid_size=32),
}


@live
@pytest.mark.parametrize("case", list(FACTORY_CASES))
def test_oracle_mg_hierarchy_equals_the_reference_factory(case):
    """how deep MGnewOp goes before it returns NULL (coarsenable(2^depth * s_maxCoarse), Factory.cpp:168-172), and dx, the
    coefficients (CoarseAverage straight from the AMR level's, arithmetic / harmonic) and lambda of every depth"""
    o = Oracle(**FACTORY_CASES[case])
    nd = o.setup()
    F = pyref.ReferenceFactory(o.params, o.get("A"), o.get("B"))
    assert F.depths == nd and F.average_type == o.params["coefficient_average_type"]
    for d in range(nd):
        n, dx, _ = F.level(d)
        assert (n, dx) == o.dims(d)
        for f in ("A", "B", "LAMBDA"):
            assert np.array_equal(F.get(f, d), o.get(f, d)), (f, d)


@live
def test_absent_average_type_means_arithmetic():
    """coefficient_average_type absent from the input = -1 in PoissonParameters (PoissonParameters.cpp:98): defineOperatorFactory
    then leaves the factory's default, arithmetic (Factory.cpp:43-45,321)"""
    o = Oracle(N=(32, 32, 32), max_grid_size=8, coefficient_average_type=0)
    o.setup()
    F = pyref.ReferenceFactory(o.params, o.get("A"), o.get("B"), coefficient_average_type=-1)
    assert F.average_type == 0
    assert np.array_equal(F.get("A", 2), o.get("A", 2))
    h = Oracle(N=(32, 32, 32), max_grid_size=8, coefficient_average_type=1)
    h.setup()
    assert not np.array_equal(h.get("A", 2), o.get("A", 2))


# ---- the six .ChF kernels themselves: oracle/chf2c.py's mechanical translation of the reference's Fortran (built into
# oracle/_ref under the Fortran symbol names) against the oracle's hand restatement, called with the same Fortran-style
# argument lists on random FABs whose boxes do not start at 0, with ghost cells, both colours, shifted bounds.
import ctypes as C  # noqa: E402


class Fab:
    """a Fortran-ordered array over [lo, hi] with ncomp components + its argument list (ptr, lo0..2, hi0..2[, ncomp])"""

    def __init__(self, lo, hi, ncomp=1, rng=None, positive=False):
        self.lo, self.hi, self.nc = lo, hi, ncomp
        shape = (ncomp,) + tuple(hi[d] - lo[d] + 1 for d in (2, 1, 0))
        self.a = rng.standard_normal(shape) if rng is not None else np.zeros(shape)
        if positive:
            self.a = 0.5 + np.abs(self.a)
        self.ints = [C.c_int(v) for v in tuple(lo) + tuple(hi)] + [C.c_int(ncomp)]

    def args(self, comp=True):
        return [self.a.ctypes.data_as(C.c_void_p)] + [C.byref(i) for i in (self.ints if comp else self.ints[:6])]

    def copy(self):
        f = Fab(self.lo, self.hi, self.nc)
        f.a = self.a.copy()
        return f


def box_args(lo, hi, keep):
    ints = [C.c_int(v) for v in tuple(lo) + tuple(hi)]
    keep.append(ints)
    return [C.byref(i) for i in ints]


@live
@pytest.mark.parametrize("lo,n", [((0, 0, 0), (8, 8, 8)), ((4, -6, 10), (9, 6, 7)), ((-8, -8, -8), (5, 4, 3))])
def test_translated_chf_kernels_equal_the_oracle_kernels(lo, n):
    from oracle import lib as oracle_lib
    O, R = oracle_lib(), pyref.lib()
    rng = np.random.default_rng(abs(hash((lo, n))) % 2 ** 32)
    hi = tuple(lo[d] + n[d] - 1 for d in range(3))
    glo, ghi = tuple(x - 1 for x in lo), tuple(x + 1 for x in hi)
    keep = []
    dx, alpha, beta = C.c_double(0.37), C.c_double(1.1), C.c_double(-0.9)
    phi = Fab(glo, ghi, rng=rng)
    rhs, a, b = Fab(lo, hi, rng=rng), Fab(lo, hi, rng=rng), Fab(lo, hi, rng=rng, positive=True)
    lam = Fab(lo, hi, rng=rng)
    region = box_args(lo, hi, keep)
    # GSRBHELMHOLTZVC3D, red then black
    p1, p2 = phi.copy(), phi.copy()
    for colour in (0, 1, 1, 0):
        c = C.c_int(colour)
        for L, f, p in ((O, "orc_gsrbhelmholtzvc3d", p1), (R, "gsrbhelmholtzvc3d_", p2)):
            getattr(L, f).restype = None
            getattr(L, f)(*p.args(), *rhs.args(), *region, C.byref(dx), C.byref(alpha), *a.args(), C.byref(beta), *b.args(), *lam.args(),
                          C.byref(c))
        assert np.array_equal(p1.a, p2.a) and not np.array_equal(p1.a, phi.a)
    # VCCOMPUTEOP3D / VCCOMPUTERES3D
    o1, o2 = Fab(lo, hi), Fab(lo, hi)
    for L, f, out in ((O, "orc_vccomputeop3d", o1), (R, "vccomputeop3d_", o2)):
        getattr(L, f).restype = None
        getattr(L, f)(*out.args(), *phi.args(), C.byref(alpha), *a.args(), C.byref(beta), *b.args(), *region, C.byref(dx))
    assert np.array_equal(o1.a, o2.a) and np.abs(o1.a).max() > 0
    for L, f, out in ((O, "orc_vccomputeres3d", o1), (R, "vccomputeres3d_", o2)):
        getattr(L, f).restype = None
        getattr(L, f)(*out.args(), *phi.args(), *rhs.args(), C.byref(alpha), *a.args(), C.byref(beta), *b.args(), *region, C.byref(dx))
    assert np.array_equal(o1.a, o2.a)
    # RESTRICTRESVC3D: bounds shifted so that the region starts at 0 (CHF_*_SHIFT, VariableCoeffPoissonOperator.cpp:173-192)
    if all(x % 2 == 0 for x in n):
        sh = lambda f: Fab(tuple(f.lo[d] - lo[d] for d in range(3)), tuple(f.hi[d] - lo[d] for d in range(3)))
        sphi, srhs, sa, sb = sh(phi), sh(rhs), sh(a), sh(b)
        sphi.a, srhs.a, sa.a, sb.a = phi.a, rhs.a, a.a, b.a
        chi = tuple(n[d] // 2 - 1 for d in range(3))
        c1, c2 = Fab((0, 0, 0), chi), Fab((0, 0, 0), chi)
        sregion = box_args((0, 0, 0), tuple(x - 1 for x in n), keep)
        for L, f, out in ((O, "orc_restrictresvc3d", c1), (R, "restrictresvc3d_", c2)):
            getattr(L, f).restype = None
            getattr(L, f)(*out.args(), *sphi.args(), *srhs.args(), C.byref(alpha), *sa.args(), C.byref(beta), *sb.args(), *sregion, C.byref(dx))
        assert np.array_equal(c1.a, c2.a) and np.abs(c1.a).max() > 0
    # GETLAPLACIANPSIF / GETRHOGRADPHIF (single-component arrays: no ncomp argument)
    for fo, fr in (("orc_getlaplacianpsif", "getlaplacianpsif_"), ("orc_getrhogradphif", "getrhogradphif_")):
        for L, f, out in ((O, fo, o1), (R, fr, o2)):
            getattr(L, f).restype = None
            getattr(L, f)(*out.args(False), *phi.args(False), C.byref(dx), *region)
        assert np.array_equal(o1.a, o2.a) and np.abs(o1.a).max() > 0


def test_chf_translator_refuses_what_it_does_not_know():
    """chf2c.py rewrites syntax only and must stop on anything outside the dialect of the reference's two files"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import chf2c
    ok = chf2c.translate("      subroutine F(CHF_FRA1[a], CHF_BOX[b])\n      integer CHF_AUTODECL[i]\n      CHF_AUTOMULTIDO[b;i]\n"
                         "        a(CHF_AUTOIX[i]) = 1.0/3.0 *\n     &     2.0\n      CHF_ENDDO\n      return\n      end\n")
    assert "a(i0,i1,i2) = 1.0/3.0 * 2.0;" in ok and 'extern "C" void f_(' in ok
    for bad in ("      subroutine F(CHF_FRA1[a])\n      a(1,1,1) = 2.0**3\n      end\n",
                "      subroutine F(CHF_FRA1[a])\n      goto 10\n      end\n",
                "      subroutine F(CHF_VR[a])\n      end\n",
                "      subroutine F(CHF_FRA1[a])\n      do i = 10, 1, -1\n      enddo\n      end\n"):
        with pytest.raises(chf2c.ChfError):
            chf2c.translate(bad)


@live
@pytest.mark.parametrize("name,over", [
    ("kernels_16_dirichlet", dict(N=(16, 16, 16), max_grid_size=8, L=40.0)),
    ("kernels_24x16x32_neumann", dict(N=(24, 16, 32), max_grid_size=8, L=60.0, bc_lo=(1, 0, 1), bc_hi=(0, 1, 1),
                                      bc_value=0.25, coefficient_average_type=0)),
])
def test_kernel_goldens_are_what_the_reference_computes(name, over):
    """tests/golden/kernels_*.npz were frozen from the oracle (make_golden.py); the reference's own source terms, operator
    class, factory and (translated) kernels reproduce every array in them that they can compute, bit for bit -- so the GPU
    tests that use these fixtures compare against the reference's numbers"""
    from oracle.pyoracle import default_params
    g = np.load(os.path.join(GOLD, name + ".npz"))
    params = default_params(**over)
    mg, rhs, a, b = pyref.set_level_data(params)
    for c in range(8):
        assert np.array_equal(mg[c][3:-3, 3:-3, 3:-3], g[f"mgvar{c}"]), c
    assert np.array_equal(rhs, g["rhs"]) and np.array_equal(a, g["a"]) and np.array_equal(b, g["b"])
    F = pyref.ReferenceFactory(params, a, b)
    assert F.depths == int(g["depths"]) and np.array_equal(F.get("LAMBDA", 0), g["lam"])
    for d in range(1, F.depths):
        assert np.array_equal(F.get("A", d), g[f"a{d}"]) and np.array_equal(F.get("LAMBDA", d), g[f"lam{d}"])
    R = pyref.ReferenceOperator(params)
    R.set("A", a); R.set("B", b); R.set("E", g["e"]); R.set("R", g["r"])
    assert np.array_equal(R.residual(True), g["residual_h"]) and np.array_equal(R.residual(False), g["residual_i"])
    assert np.array_equal(R.apply(True), g["apply_h"])
    assert np.array_equal(R.restrict(), g["restrict"])
    R.set("E", g["e"])
    R.relax(4)
    assert np.array_equal(R.get("E"), g["relax4"])
    R.set("E", g["e"])
    R.precond()
    assert np.array_equal(R.get("E"), g["precond"])


# ---- INTEGRATION.md section A for real: the reference's own operator class and source-term code with their .ChF symbols
# resolved by the PRODUCT's link-time drop-ins (include/mgic_chf.h; CUDA kernels behind the Fortran ABI) -- oracle/Makefile
# target ref_cuda links them with --no-undefined.
@live
def test_reference_classes_link_against_the_products_fortran_symbols():
    import subprocess
    if not os.path.exists(pyref.B200_SO):
        pytest.skip("libmgic_b200.so is not built")
    pyref.build()
    assert os.path.exists(pyref.SO_CUDA)
    nm = subprocess.run(["nm", "-D", pyref.SO_CUDA], capture_output=True, text=True).stdout
    undefined = {l.split()[-1] for l in nm.splitlines() if " U " in l}
    chf = {"gsrbhelmholtzvc3d_", "vccomputeop3d_", "vccomputeres3d_", "restrictresvc3d_", "getlaplacianpsif_", "getrhogradphif_"}
    assert chf <= undefined                                   # not defined inside: they come from the product library
    exported = subprocess.run(["nm", "-D", "--defined-only", pyref.B200_SO], capture_output=True, text=True).stdout
    assert all(f" T {s}" in exported for s in chf)
    assert "libmgic_b200.so" in subprocess.run(["ldd", pyref.SO_CUDA], capture_output=True, text=True).stdout
    pyref.lib(cuda=True)                                      # loads (no compute call: those need a GPU and abort() without one)
