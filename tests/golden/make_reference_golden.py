"""Generates tests/golden/reference_sources_16.npz from the REFERENCE's own code (python tests/golden/make_reference_golden.py,
in a container that has /root/reference).

Unlike make_golden.py (which freezes the oracle), these arrays come out of the reference's Source/SetLevelData.cpp,
Source/SetBinaryBH.H and MyPhiFunction.H, compiled unmodified by `make -C oracle ref` (oracle/pyref.py;
oracle/ref_shim/chombo_standin.H stands in for the Chombo containers; the two Fortran stencils inside set_rhs / set_a_coef
come from the C restatement because there is no Fortran compiler).  They pin SURVEY rows a18 / a19 -- the Bowen-York
A_ij, the scalar-field profile, psi_0 = psi + m1/r1 + m2/r2, the psi^5 / psi^-7 right-hand side, aCoef, bCoef,
set_update_psi0 -- for the oracle on CPU and for the CUDA source kernels on the GPU box, where /root/reference is absent.

Case: params.txt physics (params.txt:12-16,43-84), N = 16^3, L = 40 (punctures at x = -10, +10 inside [-20, 20], never on a
cell centre), Dirichlet dpsi = 0.  State 1: psi = 1 (NL iteration 1).  State 2: after set_update_psi0 with a smooth dpsi whose
first ghost layer is what the solver's homogeneous Dirichlet fill leaves there (ghost = -near); layers 2-3 are zero."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyref  # noqa: E402
from oracle.pyoracle import default_params  # noqa: E402

CASE = dict(N=(16, 16, 16), L=40.0, max_grid_size=8)
CONSTANT_K = 0.3


def dpsi_ghosted(N):
    """(N+6)^3 array [k, j, i]: interior 1e-3 * product of cosines; ghost layer 1 = -near on every face; the rest 0"""
    nx, ny, nz = N
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    inner = 1e-3 * np.cos(0.37 * i + 0.1) * np.cos(0.23 * j - 0.2) * np.cos(0.31 * k + 0.3)
    d = np.zeros((nz + 6, ny + 6, nx + 6))
    d[3:-3, 3:-3, 3:-3] = inner
    d[2, 3:-3, 3:-3], d[-3, 3:-3, 3:-3] = -inner[0], -inner[-1]
    d[3:-3, 2, 3:-3], d[3:-3, -3, 3:-3] = -inner[:, 0], -inner[:, -1]
    d[3:-3, 3:-3, 2], d[3:-3, 3:-3, -3] = -inner[:, :, 0], -inner[:, :, -1]
    return d


def generate():
    params = default_params(**CASE)
    mg, rhs, a, b = pyref.set_level_data(params)
    d = dpsi_ghosted(params["N"])
    mg2, rhs2, a2, _ = pyref.set_level_data(params, dpsi_ghosted=d)
    _, rhsK, aK, _ = pyref.set_level_data(params, constant_K=CONSTANT_K)
    locs = np.array([[0.3, -1.7, 2.9], [-9.0, 0.5, 0.25], [12.5, -3.0, 7.0], [-19.5, 19.5, -19.5]])
    pv = [pyref.point_values(params, loc) for loc in locs]
    return dict(
        params=json.dumps({k: (list(v) if isinstance(v, tuple) else v) for k, v in params.items()}),
        mgvars_ghost3=mg, rhs=rhs, acoef=a, bcoef=b,
        dpsi_ghost3=d, psi_after_ghost3=mg2[0], rhs_after=rhs2, acoef_after=a2,
        constant_K=CONSTANT_K, rhs_K=rhsK, acoef_K=aK, m_K=pyref.m_value(params, 0.05, CONSTANT_K),
        point_locs=locs, point_Aij=np.array([p[0] for p in pv]), point_psi_bh=np.array([p[1] for p in pv]),
        point_phi=np.array([p[2] for p in pv]))


if __name__ == "__main__":
    out = os.path.join(HERE, "reference_sources_16.npz")
    np.savez_compressed(out, **generate())
    print(out, os.path.getsize(out), "bytes")
