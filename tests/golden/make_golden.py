"""Generates tests/golden/kernels_*.npz and solver_*.npz from the CPU oracle (python tests/golden/make_golden.py).

The reference ships NO golden vectors and cannot be built as a whole here (needs Chombo 3.2 + Fortran + MPI), so these
fixtures freeze the oracle's output so that (a) an accidental change of the oracle is caught on CPU and (b) the CUDA
path is checked against the same numbers on the GPU box, where /root/reference does not exist.  Provenance: every array
of the kernels_*.npz files that the reference's own code can compute -- source terms, lambda and coefficients of every
MG depth, residual, applyOp, restricted residual, relaxed and preconditioned fields -- is reproduced bit for bit by the
reference's C++ compiled unmodified plus its translated .ChF kernels (oracle/_ref;
tests/test_reference_pins.py::test_kernel_goldens_are_what_the_reference_computes).  The solver_*.npz files (V-cycles,
BiCGStab, nonlinear loop) rest on the oracle's restatement of Chombo's algorithms and have no upstream pin.
See make_reference_golden.py for the fixture written by the reference's code itself.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import Oracle  # noqa: E402


def kernels_case(name, **over):
    o = Oracle(**over)
    nd = o.setup()
    n = o.params["N"]
    rng = np.random.default_rng(12345)
    e = rng.standard_normal((n[2], n[1], n[0]))
    r = rng.standard_normal((n[2], n[1], n[0]))
    out = dict(e=e, r=r, depths=nd, rhs=o.get("RHS"), a=o.get("A"), b=o.get("B"), lam=o.get("LAMBDA"))
    for d in range(1, nd):
        out[f"a{d}"] = o.get("A", d)
        out[f"lam{d}"] = o.get("LAMBDA", d)
    o.set("E", e); o.set("R", r)
    out["residual_h"] = o.residual(0, True)
    out["residual_i"] = o.residual(0, False)
    out["apply_h"] = o.apply(0, True)
    o.restrict(0)
    out["restrict"] = o.get("R", 1)
    o.gsrb_color(0, 0)
    out["gsrb_red"] = o.get("E")
    o.gsrb_color(0, 1)
    out["gsrb_redblack"] = o.get("E")
    o.relax(0, 3)
    out["relax4"] = o.get("E")
    c = rng.standard_normal((n[2] // 2, n[1] // 2, n[0] // 2))
    o.set("E", c, 1)
    o.prolong(0)
    out["coarse"] = c
    out["prolong"] = o.get("E")
    o.set("E", e); o.precond(0)
    out["precond"] = o.get("E")
    out["norm0"] = o.norm(0, "R", 0); out["norm2"] = o.norm(0, "R", 2); out["dot"] = o.dot(0, "E", "R")
    for c in range(8):
        out[f"mgvar{c}"] = o.get("MGVAR0", comp=c)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    o.close()


def solver_case(name, cycles, **over):
    o = Oracle(**over)
    o.setup()
    o.load_rhs_zero_e()
    hist, its = [o.norm(0, "R", 0)], []
    for _ in range(cycles):
        its.append(o.vcycle())
        hist.append(float(np.abs(o.residual(0, True)).max()))
    out = dict(vcycle_resnorm=np.array(hist), bottom_iters=np.array(its), e_after=o.get("E"))
    o2 = Oracle(**over)
    o2.set_initial_conditions()
    nl = o2.nl_solve()
    out["nl_dpsi_norms"] = nl
    out["psi"] = o2.get("MGVAR0", comp=0)
    o3 = Oracle(**over)
    o3.setup()
    it, st, fn, norms = o3.outer_solve()
    out["outer_iters"] = it; out["outer_status"] = st; out["outer_norms"] = norms; out["dpsi"] = o3.get("DPSI")
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    for x in (o, o2, o3):
        x.close()


if __name__ == "__main__":
    kernels_case("kernels_16_dirichlet", N=(16, 16, 16), max_grid_size=8, L=40.0)
    kernels_case("kernels_24x16x32_neumann", N=(24, 16, 32), max_grid_size=8, L=60.0, bc_lo=(1, 0, 1), bc_hi=(0, 1, 1),
                 bc_value=0.25, coefficient_average_type=0)
    solver_case("solver_32_v22", 5, N=(32, 32, 32), max_grid_size=16, numMGsmooth=2)
    solver_case("solver_32_v44_box8", 4, N=(32, 32, 32), max_grid_size=8, numMGsmooth=4)
    print("golden fixtures written to", HERE)
