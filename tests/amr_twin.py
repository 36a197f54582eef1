"""Test infrastructure: the AMR hierarchy algorithms -- [Chombo] AMRMultiGrid::AMRVCycle, MultilevelLinearOp, the outer
BiCGStabSolver (restated from the published algorithms, unpinned upstream) -- written in numpy
over operator-level primitives that a BACKEND supplies.  With the CPU oracle as backend (OracleBackend) this is the
checker for the library's mgic_amr_* entry points; the GPU tests also run it over the library's own operator-level
primitives (which are pinned to the oracle bit for bit) to check the orchestration alone.

A hierarchy is a list of nodes: node 0 = the base level (whole domain), then the patches, level 1's first.  A level vector
is a list of arrays [k, j, i], one per node.

Backend interface: attributes n_nodes, level[q], parent[q], lo[q] (i, j, k: origin of the node's array in its level's
index space), shape[q] (k, j, i), dx[q], smooth, mg_iterations; methods residual0(phi, rhs, homog), apply0(phi, homog),
vcycle0(res) (MultiGrid::oneCycle from a zero correction), residual_nf(q, phi, coarse, rhs, homog),
apply_nf(q, phi, coarse, homog) (coarse = the parent node's array), relax0(q, res, n) (n sweeps from a zero correction)."""
import numpy as np

import np_twin as T


def rep2(x):
    return np.repeat(np.repeat(np.repeat(x, 2, 0), 2, 1), 2, 2)


class AmrTwin:
    def __init__(self, backend):
        self.b = backend
        self.n = backend.n_nodes
        self.levels = max(backend.level) + 1
        # the cells of the parent's array under node q
        self.under = [None] * self.n
        # nodes that are unions of boxes in a bounding-box array: fmask[q] = 1 on the node's cells, cmask[q] = the same
        # coarsened by 2 (boxes are coarsenable: all eight cells or none); None for rectangular nodes
        self.fmask = list(getattr(backend, "fmask", [None] * self.n))
        self.cmask = [None if f is None else f[::2, ::2, ::2].astype(bool) for f in self.fmask]
        for q in range(1, self.n):
            p = backend.parent[q]
            off = [backend.lo[q][d] // 2 - backend.lo[p][d] for d in range(3)]
            cs = [s // 2 for s in backend.shape[q]]          # k, j, i
            self.under[q] = tuple(slice(off[d], off[d] + cs[2 - d]) for d in (2, 1, 0))

    def nodes_of(self, level):
        return [q for q in range(self.n) if self.b.level[q] == level]

    def zeros(self):
        return [np.zeros(self.b.shape[q]) for q in range(self.n)]

    # ---- MultilevelLinearOp
    def residual(self, phi, rhs, homog=False):
        out = [self.b.residual0(phi[0], rhs[0], homog)]
        for q in range(1, self.n):
            out.append(self.b.residual_nf(q, phi[q], phi[self.b.parent[q]], rhs[q], homog))
        return out

    def apply(self, phi, homog=False):
        out = [self.b.apply0(phi[0], homog)]
        for q in range(1, self.n):
            out.append(self.b.apply_nf(q, phi[q], phi[self.b.parent[q]], homog))
        return out

    def zero_covered(self, x):
        y = [a.copy() for a in x]
        for q in range(1, self.n):
            self._put_under(y, q, 0.0)
        return y

    def _put_under(self, y, q, values):
        """the parent's cells under node q := values (only under the node's own cells when it is a masked union)"""
        sub = y[self.b.parent[q]][self.under[q]]          # a view
        if self.cmask[q] is None:
            sub[...] = values
        else:
            sub[self.cmask[q]] = values[self.cmask[q]] if isinstance(values, np.ndarray) else values

    def _masked(self, q, x):
        return x if self.fmask[q] is None else x * self.fmask[q]

    def average_down(self, x):
        y = [a.copy() for a in x]
        for q in range(self.n - 1, 0, -1):
            self._put_under(y, q, T.coarse_average(y[q], 2, False))
        return y

    def norm(self, x, ord=0):
        y = self.zero_covered(x)
        if ord == 0:
            return max(np.abs(a).max() for a in y)
        s = sum((np.abs(a) ** ord).sum() * self.b.dx[q] ** 3 for q, a in enumerate(y))
        return s ** (1.0 / ord)

    def dot(self, x, y):
        xm = self.zero_covered(x)
        return sum((a * c).sum() * self.b.dx[q] ** 3 for q, (a, c) in enumerate(zip(xm, y)))

    # ---- AMRVCycle (residuals in, corrections out)
    def vcycle(self, res):
        res = [r.copy() for r in res]
        corr = self.zeros()
        self._cycle(self.levels - 1, res, corr)
        return corr

    def _cycle(self, l, res, corr):
        b, S = self.b, self.b.smooth
        if l == 0:
            corr[0] = b.vcycle0(res[0])
            return
        mine = self.nodes_of(l)
        for q in mine:
            corr[q] = b.relax0(q, res[q], S)
        for p in self.nodes_of(l - 1):
            corr[p] = np.zeros(b.shape[p])
        for q in mine:
            p = b.parent[q]
            self._put_under(res, q, T.coarse_average(b.residual_nf(q, corr[q], corr[p], res[q], True), 2, False))
        self._cycle(l - 1, res, corr)
        for q in mine:
            p = b.parent[q]
            corr[q] = corr[q] + self._masked(q, rep2(corr[p][self.under[q]]))
            res[q] = b.residual_nf(q, corr[q], corr[p], res[q], True)
            corr[q] = corr[q] + b.relax0(q, res[q], S)

    # ---- MultilevelLinearOp::preCond
    def precond(self, res):
        n_it = self.b.mg_iterations
        if n_it < 1:
            return self.zeros()
        cor = self.vcycle(res)
        for _ in range(1, n_it):
            c2 = self.vcycle(self.residual(cor, res, True))
            cor = [a + c for a, c in zip(cor, c2)]
        return cor

    # ---- BiCGStabSolver<Vector<LevelData*>>::solve: phi updated in place; returns (iterations, status, norms)
    def bicgstab(self, phi, rhs, eps, imax, norm_type=0, reps=1e-12, hang=1e-8, small=1e-30, num_restarts=5, homog=False):
        axpy = lambda y, x, s: [a + s * c for a, c in zip(y, x)]
        r = self.residual(phi, rhs, homog)
        rt = [a.copy() for a in r]
        e = self.zeros()
        pt, st = self.zeros(), self.zeros()
        p = v = None
        i, recount, restarts, status = 0, 0, 0, -1
        rho = [0.0] * 4
        norm = [self.norm(r, norm_type)] * 2
        initial_norm = initial_rnorm = norm[0]
        alpha, beta, omega = [0.0, 0.0], [0.0, 0.0], [0.0, 0.0]
        init = True
        hist = [norm[0]]
        while i < imax and norm[0] > eps * norm[1] and norm[1] > 0:
            i += 1
            norm[1] = norm[0]; alpha[1] = alpha[0]; beta[1] = beta[0]; omega[1] = omega[0]
            rho[3] = rho[2]; rho[2] = rho[1]
            rho[1] = self.dot(rt, r)
            if rho[1] == 0.0:
                for q in range(self.n):
                    phi[q] += e[q]
                return i, 2, hist
            if init:
                p = [a.copy() for a in r]
                init = False
            else:
                beta[1] = (rho[1] / rho[2]) * (alpha[1] / omega[1])
                p = [a * beta[1] for a in p]
                p = axpy(p, v, -beta[1] * omega[1])
                p = axpy(p, r, 1.0)
            pt = self.precond(p)
            v = self.apply(pt, True)
            m = self.dot(rt, v)
            alpha[0] = rho[1] / m
            if abs(m) > small * abs(rho[1]):
                r = axpy(r, v, -alpha[0])
                norm[0] = self.norm(r, norm_type)
                e = axpy(e, pt, alpha[0])
            else:
                r = self.zeros()
                norm[0] = 0.0
            if norm[0] > eps * initial_norm and norm[0] > reps * initial_rnorm:
                st = self.precond(r)
                t = self.apply(st, True)
                omega[0] = self.dot(t, r) / self.dot(t, t)
                e = axpy(e, st, omega[0])
                r = axpy(r, t, -omega[0])
                norm[0] = self.norm(r, norm_type)
            hist.append(norm[0])
            if norm[0] <= eps * initial_norm or norm[0] <= reps * initial_rnorm:
                status = 1
                break
            if omega[0] == 0.0 or norm[0] > (1 - hang) * norm[1]:
                if recount == 0:
                    recount = 1
                else:
                    recount = 0
                    for q in range(self.n):
                        phi[q] += e[q]
                    if restarts == num_restarts:
                        return i, 3, hist
                    r = self.residual(phi, rhs, homog)
                    norm[0] = self.norm(r, norm_type)
                    rho[1] = rho[2] = rho[3] = 0.0
                    alpha[0] = beta[0] = omega[0] = 0.0
                    rt = [a.copy() for a in r]
                    e = self.zeros()
                    restarts += 1
                    init = True
        for q in range(self.n):
            phi[q] += e[q]
        return i, status, hist


class OracleBackend:
    """The CPU oracle as backend: an Oracle (set up: coefficients on every MG depth) for the base level, one OraclePatch
    (coefficients set) per patch.  patches = list of lists, level 1's patches first."""

    def __init__(self, oracle, patches):
        self.o = oracle
        self.P = [None] + [q for lv in patches for q in lv]
        self.n_nodes = len(self.P)
        self.level = [0] + [l + 1 for l, lv in enumerate(patches) for _ in lv]
        nz, ny, nx = oracle.get("A").shape
        self.shape = [(nz, ny, nx)] + [p.shape for p in self.P[1:]]
        self.lo = [(0, 0, 0)] + [p.lo for p in self.P[1:]]
        self.fmask = [None] + [(p.mask().astype(np.float64) if getattr(p, "boxes", None) else None) for p in self.P[1:]]
        self.dx = [oracle.params["L"] / oracle.params["N"][0]]
        self.parent = [-1]
        for q in range(1, self.n_nodes):
            self.dx.append(self.dx[0] / (1 << self.level[q]))
            cand = [p for p in range(self.n_nodes) if self.level[p] == self.level[q] - 1 and all(
                0 <= self.lo[q][d] // 2 - self.lo[p][d] and
                self.lo[q][d] // 2 - self.lo[p][d] + self.shape[q][2 - d] // 2 <= self.shape[p][2 - d] for d in range(3))]
            self.parent.append(cand[0])
        self.smooth = oracle.params["numMGsmooth"]
        self.mg_iterations = oracle.params["numMGIterations"]

    def residual0(self, phi, rhs, homog):
        self.o.set("E", phi); self.o.set("R", rhs)
        return self.o.residual(0, homog)

    def apply0(self, phi, homog):
        self.o.set("E", phi)
        return self.o.apply(0, homog)

    def vcycle0(self, res):
        self.o.set("R", res); self.o.set("E", np.zeros_like(res)); self.o.vcycle()
        return self.o.get("E")

    def _coarse_domain(self, q, coarse):
        """OraclePatch takes the coarser level's field over that level's whole domain"""
        p = self.parent[q]
        if p == 0:
            return coarse
        nd = [self.shape[0][2 - d] << self.level[p] for d in range(3)]      # i, j, k
        full = np.zeros((nd[2], nd[1], nd[0]))
        lo, sh = self.lo[p], self.shape[p]
        full[lo[2]:lo[2] + sh[0], lo[1]:lo[1] + sh[1], lo[0]:lo[0] + sh[2]] = coarse
        return full

    def residual_nf(self, q, phi, coarse, rhs, homog):
        P = self.P[q]
        P.set("E", phi); P.set("R", rhs); P.set_coarse(self._coarse_domain(q, coarse))
        return P.amr_residual_nf(homog)

    def apply_nf(self, q, phi, coarse, homog):
        P = self.P[q]
        P.set("E", phi); P.set_coarse(self._coarse_domain(q, coarse))
        return P.amr_operator_nf(homog)

    def relax0(self, q, res, n):
        P = self.P[q]
        P.set("E", np.zeros(self.shape[q])); P.set("R", res); P.relax(n)
        return P.get("E")


def hierarchy_nl_solve(oracle, patches, max_nl=None, log=None):
    """poissonSolve's nonlinear loop (Main_PoissonSolver.cpp:93, 131-216) on a hierarchy over the oracle: `oracle` = the base
    level (an Oracle, not yet set up), patches = [[OraclePatch, ...] for level 1, ...].  Returns (dpsi norms per NL iteration,
    psi per node).  dpsi carries over between iterations as the initial guess (:93 is its only zeroing)."""
    P = oracle.params
    oracle.set_initial_conditions()
    flat = [q for lv in patches for q in lv]
    for q in flat:
        q.set_initial_conditions(P)
    dpsi, norms, tw = None, [], None
    for nl in range(max_nl if max_nl is not None else P["max_NL_iterations"]):
        oracle.set_coefs_and_rhs()
        oracle.define_solver()
        for q in flat:
            q.set_coefs_and_rhs()
        tw = AmrTwin(OracleBackend(oracle, patches))
        rhs = [oracle.get("RHS")] + [q.var(8) for q in flat]
        if dpsi is None:
            dpsi = tw.zeros()
        its, status, hist = tw.bicgstab(dpsi, rhs, eps=P["tolerance"], imax=P["max_iterations"], homog=False)
        oracle.update_psi0(dpsi[0])
        for n, q in enumerate(flat, start=1):
            p = tw.b.parent[n]
            q.update_psi(dpsi[n], tw.b._coarse_domain(n, dpsi[p]))
        nrm = tw.norm(dpsi, 2)
        norms.append(nrm)
        if log is not None:
            log.append((its, status, hist[-1]))
        if nrm < P["tolerance"] or nrm > 1e5:
            break
    psi = [oracle.get("MGVAR0", comp=0)] + [q.var(0) for q in flat]
    return np.array(norms), psi
