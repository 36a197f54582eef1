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
