"""The CUDA source-term kernels (csrc/source.cu, through the C ABI) against committed OUTPUTS OF THE REFERENCE'S OWN CODE:
tests/golden/reference_sources_16.npz was written by the reference's Source/SetLevelData.cpp / SetBinaryBH.H /
MyPhiFunction.H compiled unmodified (tests/golden/make_reference_golden.py, oracle/pyref.py).  /root/reference does not
exist on the GPU box; the fixture travels.  Bars as in test_gpu_parity.py::test_source_terms: A_ij bit-exact
(+, -, *, /, sqrt only), exp() and the psi_0 powers (multiplications here, pow() in the reference) 1e-13 relative."""
import json
import os

import numpy as np
import pytest

import mg_ic_code_b200 as m

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
FACES = ((0, slice(1, -1), slice(1, -1)), (-1, slice(1, -1), slice(1, -1)), (slice(1, -1), 0, slice(1, -1)),
         (slice(1, -1), -1, slice(1, -1)), (slice(1, -1), slice(1, -1), 0), (slice(1, -1), slice(1, -1), -1))
INNER = (slice(1, -1),) * 3


def relerr(x, y):
    d = np.abs(np.asarray(x) - np.asarray(y)).max()
    s = np.abs(np.asarray(y)).max()
    return d / s if s > 0 else d


def test_source_kernels_against_the_reference_outputs(ctx):
    g = np.load(os.path.join(GOLD, "reference_sources_16.npz"))
    over = {k: (tuple(v) if isinstance(v, list) else v) for k, v in json.loads(str(g["params"])).items()}
    P = m.make_params(over)
    lvl = m.level_op_from_params(ctx, P)
    v = m.MultigridVars(ctx, P)
    dpsi, rhs, a, b = lvl.create(), lvl.create(), lvl.create(), lvl.create()
    # ---- set_initial_conditions (Source/SetLevelData.cpp:32-71): valid cells and the face ghosts the stencils read
    v.set_initial_conditions(dpsi)
    assert np.all(dpsi.download() == 0.0)
    for c in range(8):
        ref3 = g["mgvars_ghost3"][c]
        got, gh, ref1 = v.download(c), v.download(c, ghosted=True), ref3[2:-2, 2:-2, 2:-2]
        if c < 7:
            assert np.array_equal(got, ref3[3:-3, 3:-3, 3:-3]), m.MultigridVars.NAMES[c]
        else:
            assert relerr(got, ref3[3:-3, 3:-3, 3:-3]) < 1e-14          # exp()
        assert relerr(gh[INNER], ref1[INNER]) < 1e-14
        for f in FACES:
            assert relerr(gh[f], ref1[f]) < 1e-13, (m.MultigridVars.NAMES[c], f)
    # ---- set_rhs / set_a_coef / set_b_coef at psi = 1, without and with constant_K
    v.set_rhs(rhs); v.set_a_coef(a); v.set_b_coef(b)
    assert relerr(rhs.download(), g["rhs"]) < 1e-13 and relerr(a.download(), g["acoef"]) < 1e-13
    assert np.array_equal(b.download(), g["bcoef"])
    K = float(g["constant_K"])
    v.set_rhs(rhs, K); v.set_a_coef(a, K)
    assert relerr(rhs.download(), g["rhs_K"]) < 1e-13 and relerr(a.download(), g["acoef_K"]) < 1e-13
    # ---- set_update_psi0 (:243-263), then the sources again: psi now has a Laplacian
    v.set_a_coef(a); v.set_b_coef(b)
    f = m.VariableCoeffPoissonOperatorFactory(ctx, P, a, b)
    dpsi.upload(g["dpsi_ghost3"][3:-3, 3:-3, 3:-3])
    v.set_update_psi0(f.MGnewOp(0), dpsi)          # dpsi's domain-face ghosts: the homogeneous Dirichlet fill, -near
    psi, ref = v.download(0, ghosted=True), g["psi_after_ghost3"][2:-2, 2:-2, 2:-2]
    assert relerr(psi[INNER], ref[INNER]) < 1e-14
    for fsl in FACES:
        assert relerr(psi[fsl], ref[fsl]) < 1e-14, fsl
    r2, a2 = lvl.create(), lvl.create()
    v.set_rhs_and_a_coef(r2, a2)
    assert relerr(r2.download(), g["rhs_after"]) < 1e-13 and relerr(a2.download(), g["acoef_after"]) < 1e-13
