"""The Fortran-symbol drop-ins (include/mgic_chf.h) against the oracle's kernels with the same argument lists:
ghosted FABs, arbitrary bounds, region sub-boxes, CHF_FRA_SHIFT-style shifted indices.  Bit-exact."""
import ctypes as C
import os

import numpy as np
import pytest

import mg_ic_code_b200 as m
import oracle

pytestmark = pytest.mark.gpu


def ip(v):
    return C.byref(C.c_int(v))


def dp(v):
    return C.byref(C.c_double(v))


def fra(arr, lo, hi, nc=True):
    a = [arr.ctypes.data_as(C.c_void_p)] + [ip(x) for x in lo] + [ip(x) for x in hi]
    if nc:
        a.append(ip(arr.shape[0] if arr.ndim == 4 else 1))
    return a


def box(lo, hi):
    return [ip(x) for x in lo] + [ip(x) for x in hi]


def mk(rng, lo, hi):
    shape = (hi[2] - lo[2] + 1, hi[1] - lo[1] + 1, hi[0] - lo[0] + 1)
    return rng.standard_normal(shape)


@pytest.mark.parametrize("lo,hi,ng", [((0, 0, 0), (15, 15, 15), 1), ((16, -8, 32), (31, 7, 39), 3), ((1, 2, 3), (9, 7, 12), 1)])
def test_operator_kernels(ctx, lo, hi, ng):
    G, O = m.lib(), oracle.lib()
    rng = np.random.default_rng(11)
    glo, ghi = tuple(x - ng for x in lo), tuple(x + ng for x in hi)
    dpsi = mk(rng, glo, ghi)
    rhs, a, b, lam = (mk(rng, lo, hi) for _ in range(4))
    dx, alpha, beta = 0.37, 1.0, -1.0
    for rb in (0, 1):
        d1, d2 = dpsi.copy(), dpsi.copy()
        for lib_, fn, d in ((G, "gsrbhelmholtzvc3d_", d1), (O, "orc_gsrbhelmholtzvc3d", d2)):
            getattr(lib_, fn)(*fra(d, glo, ghi), *fra(rhs, lo, hi), *box(lo, hi), dp(dx), dp(alpha), *fra(a, lo, hi), dp(beta),
                              *fra(b, lo, hi), *fra(lam, lo, hi), ip(rb))
        assert np.array_equal(d1, d2)
    for fn_g, fn_o, with_rhs in (("vccomputeop3d_", "orc_vccomputeop3d", False), ("vccomputeres3d_", "orc_vccomputeres3d", True)):
        o1, o2 = np.zeros_like(rhs), np.zeros_like(rhs)
        for lib_, fn, out in ((G, fn_g, o1), (O, fn_o, o2)):
            args = fra(out, lo, hi) + fra(dpsi, glo, ghi) + (fra(rhs, lo, hi) if with_rhs else []) + [dp(alpha)] + fra(a, lo, hi) + \
                [dp(beta)] + fra(b, lo, hi) + box(lo, hi) + [dp(dx)]
            getattr(lib_, fn)(*args)
        assert np.array_equal(o1, o2) and o1.any()


def test_restrict_and_prolong_shifted(ctx):
    G, O = m.lib(), oracle.lib()
    rng = np.random.default_rng(12)
    # CHF_FRA_SHIFT: box lo shifted to 0 (VariableCoeffPoissonOperator.cpp:173-192)
    lo, hi = (0, 0, 0), (15, 7, 11)
    glo, ghi = (-1, -1, -1), (16, 8, 12)
    clo, chi = (0, 0, 0), (7, 3, 5)
    dpsi = mk(rng, glo, ghi)
    rhs, a, b = (mk(rng, lo, hi) for _ in range(3))
    r1, r2 = np.zeros((6, 4, 8)), np.zeros((6, 4, 8))
    for lib_, fn, out in ((G, "restrictresvc3d_", r1), (O, "orc_restrictresvc3d", r2)):
        getattr(lib_, fn)(*fra(out, clo, chi), *fra(dpsi, glo, ghi), *fra(rhs, lo, hi), dp(1.0), *fra(a, lo, hi), dp(-1.0),
                          *fra(b, lo, hi), *box(lo, hi), dp(0.5))
    assert np.array_equal(r1, r2) and r1.any()
    cg_lo, cg_hi = (-1, -1, -1), (8, 4, 6)
    coarse = mk(rng, cg_lo, cg_hi)
    p1, p2 = dpsi.copy(), dpsi.copy()
    for lib_, fn, out in ((G, "prolong_", p1), (O, "orc_prolong", p2)):
        getattr(lib_, fn)(*fra(out, glo, ghi), *fra(coarse, cg_lo, cg_hi), *box(lo, hi), ip(2))
    assert np.array_equal(p1, p2) and not np.array_equal(p1, dpsi)


def test_source_kernels(ctx):
    G, O = m.lib(), oracle.lib()
    rng = np.random.default_rng(13)
    lo, hi = (8, 0, -4), (15, 7, 3)
    glo, ghi = tuple(x - 3 for x in lo), tuple(x + 3 for x in hi)
    psi = mk(rng, glo, ghi)
    for fg, fo in (("getlaplacianpsif_", "orc_getlaplacianpsif"), ("getrhogradphif_", "orc_getrhogradphif")):
        o1, o2 = np.zeros((8, 8, 8)), np.zeros((8, 8, 8))
        for lib_, fn, out in ((G, fg, o1), (O, fo, o2)):
            getattr(lib_, fn)(*fra(out, lo, hi, nc=False), *fra(psi, glo, ghi, nc=False), dp(0.21), *box(lo, hi))
        assert np.array_equal(o1, o2) and o1.any()


REFERENCE_ON_CUDA = r"""
import sys
import numpy as np
from oracle import Oracle, pyref
over = {"c16": dict(N=(16, 16, 16), max_grid_size=8, L=40.0),
        "neumann_inhomogeneous": dict(N=(24, 16, 32), max_grid_size=8, L=60.0, bc_lo=(1, 0, 1), bc_hi=(0, 1, 1), bc_value=0.25)}[sys.argv[1]]
o = Oracle(**over)
o.setup()
mg, rhs, a, b = pyref.set_level_data(o.params, cuda=True)              # set_rhs / set_a_coef through the CUDA stencils
assert np.array_equal(rhs, o.get("RHS")) and np.array_equal(a, o.get("A")), "source terms"
R = pyref.ReferenceOperator(o.params, cuda=True)
rng = np.random.default_rng(3)
e, r = rng.standard_normal(R.shape), rng.standard_normal(R.shape)
R.set("A", o.get("A")); R.set("B", o.get("B"))
for homog in (True, False):
    o.set("E", e); o.set("R", r); R.set("E", e); R.set("R", r)
    assert np.array_equal(R.residual(homog), o.residual(0, homog)), "residualI"
    assert np.array_equal(R.apply(homog), o.apply(0, homog)), "applyOpI"
o.relax(0, 2); R.relax(2)
assert np.array_equal(R.get("E"), o.get("E")), "levelGSRB"
o.restrict(0)
assert np.array_equal(R.restrict(), o.get("R", 1)), "restrictResidual"
o.precond(0); R.precond()
assert np.array_equal(R.get("E"), o.get("E")), "preCond"
print("reference classes on the CUDA drop-ins: identical to the oracle")
"""


@pytest.mark.parametrize("case", ["c16", "neumann_inhomogeneous"])
def test_reference_operator_class_on_the_cuda_drop_ins(case):
    """INTEGRATION.md section A end to end: the reference's own VariableCoeffPoissonOperator.cpp / SetBCs.cpp / SetLevelData.cpp
    (compiled unmodified into oracle/_ref/libmgic_ref_cuda.so, built where /root/reference exists and shipped with the
    snapshot) call gsrbhelmholtzvc3d_ ... getrhogradphif_ per box -- and those symbols are libmgic_b200.so's CUDA kernels
    behind the Fortran ABI.  Results: the oracle's, bit for bit.  Runs in a child process: the Fortran ABI has no error
    return, it abort()s like MAYDAYERROR."""
    import subprocess
    import sys
    from oracle import pyref
    if not os.path.exists(pyref.SO_CUDA):
        pytest.skip("oracle/_ref/libmgic_ref_cuda.so was not built (needs /root/reference at build time)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", REFERENCE_ON_CUDA, case], capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
