"""Independent numpy twin of the operator kernels on ONE global array (no boxes), used to pin the boxed C++
oracle: GSRB / residual / applyOp / restrict are partition invariant (global-index colouring, exchange before
every colour), so oracle(any box size) must equal this twin bit for bit.  Arrays are indexed [k, j, i].

Follows Source/VariableCoeffPoissonOperatorF.ChF:56-139,181-237,283-339,379-437 and
Source/SetBCs.cpp:49-131 ([Chombo] DiriBC order 1 / NeumBC).  Test infrastructure only.
"""
import numpy as np


def interp_homo(pa, pb, dx, dx_crse):
    """[Chombo] INTERPHOMO: value at 2*dx of the parabola through pa (second interior cell, at 0), pb (first interior
    cell, at dx) and zero at (3*dx + dx_crse)/2 -- same operation order as the oracle."""
    x1 = dx
    x2 = 0.5 * (3. * x1 + dx_crse)
    denom = 1.0 - ((x1 + x2) / x1)
    idenom = 1 / (denom)
    x = 2. * x1
    xsquared = x * x
    m1 = 1 / (x1 * x1)
    m2 = 1 / (x1 * (x1 - x2))
    q1 = 1 / (x1 - x2)
    q2 = x1 + x2
    a = ((pb - pa) * m1 - (pb) * m2) * idenom
    b = (pb) * q1 - a * q2
    return a * xsquared + b * x + pa


def ghosted(phi, bc_lo=(0, 0, 0), bc_hi=(0, 0, 0), value=0.0, dx=1.0, homogeneous=True, cf_lo=(False,) * 3,
            cf_hi=(False,) * 3, dx_crse=None, origin=(0, 0, 0)):
    """phi with one ghost layer filled by the physical BC (0 Dirichlet: 2v - near, 1 Neumann: near + side*dx*v) or, on
    the coarse-fine faces of an AMR patch (cf_lo / cf_hi), by homogeneousCFInterp from the two interior cells."""
    v = 0.0 if homogeneous else value
    g = np.zeros(tuple(s + 2 for s in phi.shape))
    g[1:-1, 1:-1, 1:-1] = phi
    for d in range(3):           # d = 0 -> x (last axis)
        ax = 2 - d
        lo = [slice(1, -1)] * 3
        hi = [slice(1, -1)] * 3
        nlo = [slice(1, -1)] * 3
        nhi = [slice(1, -1)] * 3
        lo[ax], nlo[ax] = 0, 1
        hi[ax], nhi[ax] = -1, -2
        flo = [slice(1, -1)] * 3
        fhi = [slice(1, -1)] * 3
        flo[ax], fhi[ax] = 2, -3
        if cf_lo[d]:
            g[tuple(lo)] = interp_homo(g[tuple(flo)], g[tuple(nlo)], dx, dx_crse)
        else:
            g[tuple(lo)] = (2 * v - g[tuple(nlo)]) if bc_lo[d] == 0 else (g[tuple(nlo)] + (-1) * dx * v)
        if cf_hi[d]:
            g[tuple(hi)] = interp_homo(g[tuple(fhi)], g[tuple(nhi)], dx, dx_crse)
        else:
            g[tuple(hi)] = (2 * v - g[tuple(nhi)]) if bc_hi[d] == 0 else (g[tuple(nhi)] + (+1) * dx * v)
    return g


def lap7(g):
    c = g[1:-1, 1:-1, 1:-1]
    t = 2.0 * c
    return (((g[1:-1, 1:-1, 2:] + g[1:-1, 1:-1, :-2]) - t) + ((g[1:-1, 2:, 1:-1] + g[1:-1, :-2, 1:-1]) - t)
            + ((g[2:, 1:-1, 1:-1] + g[:-2, 1:-1, 1:-1]) - t))


def colour_mask(shape, red_black):
    k, j, i = np.meshgrid(np.arange(shape[0]), np.arange(shape[1]), np.arange(shape[2]), indexing="ij")
    return ((i + j + k + red_black) % 2) == 0


def gsrb_colour(phi, rhs, a, b, lam, alpha, beta, dx, red_black, **bc):
    red_black = red_black + sum(bc.get("origin", (0, 0, 0)))   # colouring by GLOBAL index
    g = ghosted(phi, dx=dx, homogeneous=True, **bc)
    dxinv = 1.0 / (dx * dx)
    lof = alpha * a * phi
    l = lap7(g)
    l = l * dxinv * b
    lof = lof - beta * l
    new = phi - lam * (lof - rhs)
    return np.where(colour_mask(phi.shape, red_black), new, phi)


def relax(phi, rhs, a, b, lam, alpha, beta, dx, iterations, **bc):
    for _ in range(iterations):
        for p in (0, 1):
            phi = gsrb_colour(phi, rhs, a, b, lam, alpha, beta, dx, p, **bc)
    return phi


def apply_op(phi, a, b, alpha, beta, dx, homogeneous=True, value=0.0, **bc):
    g = ghosted(phi, dx=dx, homogeneous=homogeneous, value=value, **bc)
    l = lap7(g) * (1.0 / (dx * dx)) * beta * b
    return alpha * a * phi - l


def residual(phi, rhs, a, b, alpha, beta, dx, homogeneous=True, value=0.0, **bc):
    g = ghosted(phi, dx=dx, homogeneous=homogeneous, value=value, **bc)
    l = lap7(g) * (1.0 / (dx * dx)) * beta * b
    return (rhs - alpha * a * phi) + l


def restrict_residual(phi, rhs, a, b, alpha, beta, dx, **bc):
    g = ghosted(phi, dx=dx, homogeneous=True, **bc)
    l = lap7(g) * (1.0 / (dx * dx)) * beta * b
    lof = alpha * a * phi - l
    q = (rhs - lof) / 8.0
    out = np.zeros(tuple(s // 2 for s in phi.shape))
    for dk in (0, 1):
        for dj in (0, 1):
            for di in (0, 1):
                out = out + q[dk::2, dj::2, di::2]
    return out


def prolong_increment(phi, coarse):
    return phi + np.repeat(np.repeat(np.repeat(coarse, 2, axis=0), 2, axis=1), 2, axis=2)


def compute_lambda(a, alpha, beta, dx):
    return 1.0 / (a * alpha + 2.0 * 3 * beta / (dx * dx))


def coarse_average(f, nref, harmonic):
    nz, ny, nx = (s // nref for s in f.shape)
    s = np.zeros((nz, ny, nx))
    for kk in range(nref):
        for jj in range(nref):
            for ii in range(nref):
                fv = f[kk::nref, jj::nref, ii::nref]
                s = s + (1.0 / fv if harmonic else fv)
    scale = 1.0 / float(nref ** 3)
    return 1.0 / (s * scale) if harmonic else s * scale


def quad_cf_face(phi, coarse, lo, hi, n_domain, dx, d, side):
    """[Chombo] QuadCFInterp on one coarse-fine face (direction d, side -1/+1) of the patch [lo, hi]: the ghost values as a
    2-D array indexed [tb, ta] (tangential directions ta < tb).  Same operation order as the oracle's Op::quadCFInterp."""
    nref = 2
    h, H = dx, 2.0 * dx
    ta, tb = sorted(t for t in range(3) if t != d)
    gi = lo[d] - 1 if side < 0 else hi[d] + 1
    fa = np.arange(lo[ta], hi[ta] + 1)[None, :]
    fb = np.arange(lo[tb], hi[tb] + 1)[:, None]
    ca, cb, cn = fa // nref, fb // nref, gi // nref
    cdom = [n // nref for n in n_domain]

    def C(oa, ob):
        idx = [None, None, None]
        idx[d] = np.full(np.broadcast(ca, cb).shape, cn)
        idx[ta] = np.broadcast_to(np.clip(ca + oa, 0, cdom[ta] - 1), idx[d].shape)
        idx[tb] = np.broadcast_to(np.clip(cb + ob, 0, cdom[tb] - 1), idx[d].shape)
        return coarse[idx[2], idx[1], idx[0]]

    c0 = C(0, 0)
    phistar = c0
    xs = []
    for w, (t, ft, ct) in enumerate(((ta, fa, ca), (tb, fb, cb))):
        x = (ft + 0.5) * h - (ct + 0.5) * H
        x = np.broadcast_to(x, c0.shape)
        xs.append(x)
        Ct = (lambda o: C(0, o)) if w else (lambda o: C(o, 0))
        has_lo = np.broadcast_to(ct - 1 >= 0, c0.shape)
        has_hi = np.broadcast_to(ct + 1 <= cdom[t] - 1, c0.shape)
        d1c = (Ct(1) - Ct(-1)) / (2.0 * H)
        d2c = ((Ct(1) - 2.0 * c0) + Ct(-1)) / (H * H)
        d1f = ((4.0 * Ct(1) - 3.0 * c0) - Ct(2)) / (2.0 * H)
        d2f = ((c0 - 2.0 * Ct(1)) + Ct(2)) / (H * H)
        d1b = ((3.0 * c0 - 4.0 * Ct(-1)) + Ct(-2)) / (2.0 * H)
        d2b = ((c0 - 2.0 * Ct(-1)) + Ct(-2)) / (H * H)
        d1 = np.where(has_lo & has_hi, d1c, np.where(has_hi, d1f, d1b))
        d2 = np.where(has_lo & has_hi, d2c, np.where(has_hi, d2f, d2b))
        phistar = phistar + (d1 * x + 0.5 * d2 * x * x)
    corners = np.broadcast_to((ca - 1 >= 0) & (ca + 1 <= cdom[ta] - 1) & (cb - 1 >= 0) & (cb + 1 <= cdom[tb] - 1), c0.shape)
    mixed = (((C(1, 1) - C(1, -1)) - C(-1, 1)) + C(-1, -1)) / (4.0 * H * H)
    phistar = np.where(corners, phistar + mixed * xs[0] * xs[1], phistar)
    # the two interior cells along the normal
    sl_near, sl_far = [slice(None)] * 3, [slice(None)] * 3
    ax = 2 - d
    sl_near[ax] = 0 if side < 0 else -1
    sl_far[ax] = 1 if side < 0 else -2
    pb, pa = phi[tuple(sl_near)], phi[tuple(sl_far)]    # 2-D, indexed by the remaining axes in [k, j, i] order = [tb, ta]
    x = 2.0 * h
    a = (2.0 / h / h) * ((2.0 * phistar + pa * (nref + 1.0)) - pb * (nref + 3.0)) / (nref * nref + 4.0 * nref + 3.0)
    b = (pb - pa) / h - a * h
    return (pa + b * x) + a * x * x


def ghosted_amr(phi, coarse, lo, hi, n_domain, dx, bc_lo=(0, 0, 0), bc_hi=(0, 0, 0), value=0.0, homogeneous=True):
    """phi of the patch [lo, hi] with one ghost layer: QuadCFInterp from `coarse` on the coarse-fine faces, the physical
    BC on domain faces (what AMROperatorNF hands to applyOpI)."""
    g = ghosted(phi, bc_lo=bc_lo, bc_hi=bc_hi, value=value, dx=dx, homogeneous=homogeneous)
    for d in range(3):
        ax = 2 - d
        for side, is_cf in ((-1, lo[d] > 0), (+1, hi[d] < n_domain[d] - 1)):
            if not is_cf:
                continue
            sl = [slice(1, -1)] * 3
            sl[ax] = 0 if side < 0 else -1
            g[tuple(sl)] = quad_cf_face(phi, coarse, lo, hi, n_domain, dx, d, side)
    return g


def amr_operator_nf(phi, coarse, a, b, alpha, beta, dx, lo, hi, n_domain, homogeneous=True, value=0.0, **bc):
    g = ghosted_amr(phi, coarse, lo, hi, n_domain, dx, homogeneous=homogeneous, value=value, **bc)
    l = lap7(g) * (1.0 / (dx * dx)) * beta * b
    return alpha * a * phi - l


def amr_residual_nf(phi, coarse, rhs, a, b, alpha, beta, dx, lo, hi, n_domain, homogeneous=True, value=0.0, **bc):
    g = ghosted_amr(phi, coarse, lo, hi, n_domain, dx, homogeneous=homogeneous, value=value, **bc)
    l = lap7(g) * (1.0 / (dx * dx)) * beta * b
    return (rhs - alpha * a * phi) + l
