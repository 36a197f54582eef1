"""The product's host-side solver templates (mg_ic_code_b200/host/ChomboSolvers.H: BiCGStabSolver<T>, MultiGrid<T> -- the code
behind poisson_solver_b200 --host-vcycle and the bottom / outer solves of the C++ mirror) run on the CPU over a mock
operator (tests/data/host_solvers_probe.cpp: 1-D variable-coefficient Helmholtz on std::vector<double>) and are compared
with the numpy restatement of the same [Chombo] algorithms in tests/amr_twin.py, which the GPU tests hold the library's
device-side BiCGStab to.  No GPU, no oracle: pure control-flow parity of three independently written solvers."""
import json
import os
import subprocess

import numpy as np
import pytest

from amr_twin import AmrTwin

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


class Helm1D:
    """numpy mirror of the probe's operator: L phi = a phi - (phi[i-1] - 2 phi[i] + phi[i+1]) / h^2, ghost = -near"""

    def __init__(self, a, h):
        self.a, self.h, self.n = np.asarray(a, float), h, len(a)
        self.lam = 1.0 / (self.a + 2.0 / (h * h))

    def L(self, x):
        g = np.concatenate(([-x[0]], x, [-x[-1]]))
        return self.a * x - ((g[:-2] - 2.0 * x) + g[2:]) / (self.h * self.h)

    def relax(self, e, r, iterations):
        e = e.copy()
        for _ in range(iterations):
            for colour in (0, 1):
                e[colour::2] = (e - self.lam * (self.L(e) - r))[colour::2]
        return e

    def precond(self, r):
        return self.relax(r * self.lam, r, 2)

    def restrict(self, phi, rhs):
        res = rhs - self.L(phi)
        return 0.5 * (res[0::2] + res[1::2])


class OneLevelBackend:
    """amr_twin backend of one node: the preconditioner is the operator's preCond"""

    def __init__(self, op):
        self.op = op
        self.n_nodes, self.level, self.parent, self.lo, self.shape, self.dx = 1, [0], [-1], [(0, 0, 0)], [(op.n,)], [1.0]
        self.smooth, self.mg_iterations = 2, 1

    def residual0(self, phi, rhs, homog):
        return rhs - self.op.L(phi)

    def apply0(self, phi, homog):
        return self.op.L(phi)

    def vcycle0(self, res):
        return self.op.precond(res)


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("probe") / "host_solvers_probe")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    lib = os.path.join(ROOT, "mg_ic_code_b200", "lib")
    subprocess.check_call([cxx, "-O1", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "mg_ic_code_b200", "host"),
                           os.path.join(HERE, "data", "host_solvers_probe.cpp"), "-L" + lib, "-lmgic_b200", "-Wl,-rpath," + lib, "-o", exe])
    return {k: (np.array(v) if isinstance(v, list) else v) for k, v in json.loads(subprocess.check_output([exe], text=True)).items()}


def test_host_bicgstab_template_follows_the_same_iteration_as_the_twin(probe):
    op = Helm1D(probe["a"], 1.0 / len(probe["a"]))
    tw = AmrTwin(OneLevelBackend(op))
    phi = tw.zeros()
    its, status, hist = tw.bicgstab(phi, [probe["rhs"]], eps=1e-10, imax=100, norm_type=0)
    assert (its, status) == (probe["bicgstab_iterations"], probe["bicgstab_status"]) and status == 1
    h = probe["bicgstab_history"]
    assert len(hist) == len(h)
    for x, y in zip(hist, h):
        if y > 1e-8 * h[0]:                               # above the rounding floor the two histories are the same numbers
            assert abs(x - y) <= 1e-6 * y
    assert np.abs(phi[0] - probe["bicgstab_phi"]).max() <= 1e-9 * np.abs(phi[0]).max()
    assert np.abs(probe["rhs"] - op.L(probe["bicgstab_phi"])).max() <= 1e-9 * np.abs(probe["rhs"]).max()


def test_host_bicgstab_template_gives_up_after_its_restarts(probe):
    """an unreachable tolerance: hang detection (two consecutive stalls -> restart), m_numRestarts = 5 restarts, exit status 3"""
    op = Helm1D(probe["a"], 1.0 / len(probe["a"]))
    tw = AmrTwin(OneLevelBackend(op))
    _, status, _ = tw.bicgstab(tw.zeros(), [probe["rhs"]], eps=1e-30, imax=400, norm_type=0, reps=1e-30)
    assert status == probe["giveup_status"] == 3
    assert probe["giveup_iterations"] < 400


def test_host_multigrid_template_runs_the_v_cycle_of_the_twin(probe):
    """MultiGrid<T>::define asks the factory for depths until it returns NULL; oneCycle = relax(pre), restrictResidual, zero,
    recurse, prolongIncrement, relax(post); at the bottom relax(bottom) + BiCGStab (defaults: eps 1e-6, L2 norm, 80 iterations)"""
    n = len(probe["a"])
    ops = []
    for d in range(8):
        c = 1 << d
        if n % c or n // c < 4:
            break
        ops.append(Helm1D(probe["a"].reshape(-1, c).mean(axis=1), c / n))
    assert probe["mg_depth"] == len(ops) - 1 == 4

    def cycle(d, e, r):
        op = ops[d]
        if d == len(ops) - 1:
            e = op.relax(e, r, 2)
            tw = AmrTwin(OneLevelBackend(op))
            phi = [e]
            tw.bicgstab(phi, [r], eps=1e-6, imax=80, norm_type=2, homog=True)
            return phi[0]
        e = op.relax(e, r, 2)
        ec = cycle(d + 1, np.zeros(op.n // 2), op.restrict(e, r))
        e = e + np.repeat(ec, 2)
        return op.relax(e, r, 2)

    e, hist = np.zeros(n), []
    for _ in range(6):
        hist.append(np.abs(probe["rhs"] - ops[0].L(e)).max())
        e = cycle(0, e, probe["rhs"])
    hist.append(np.abs(probe["rhs"] - ops[0].L(e)).max())
    assert np.allclose(hist, probe["mg_history"], rtol=1e-6)
    assert np.abs(e - probe["mg_e"]).max() <= 1e-8 * np.abs(e).max()
    assert all(hist[i + 1] < hist[i] for i in range(6))
