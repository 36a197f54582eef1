"""Host-side mirror of the reference's operator interface on top of the C ABI.

Names and argument meaning follow Source/VariableCoeffPoissonOperator.H:40-168,
Source/VariableCoeffPoissonOperatorFactory.H:57-92 and the [Chombo] AMRLevelOp / MultiGrid methods the
reference inherits.  A LevelField is the device-resident stand-in for LevelData<FArrayBox> (one component);
its numpy views are global ghost-free arrays indexed [k, j, i].  Errors raise MgicError where the reference
calls MayDay::Error / MayDay::Abort.  Everything computes on the GPU; nothing here touches the oracle.
"""
import ctypes as C

import numpy as np

from ._capi import MgicError, check, lib
from .params import make_params


class Context:
    """One GPU (+ stream).  Multi-GPU: one Context per process/rank (see mg_ic_code_b200.comm)."""

    def __init__(self, device=0, rank=0, nranks=1):
        self.L = lib()
        h = C.c_void_p()
        check(self.L.mgic_ctx_create(device, C.byref(h)))
        self.h = h
        self.rank, self.nranks = rank, nranks
        if nranks > 1:
            check(self.L.mgic_ctx_set_rank(h, rank, nranks))

    def sync(self):
        check(self.L.mgic_ctx_sync(self.h))

    def set_stream(self, cuda_stream_ptr):
        check(self.L.mgic_ctx_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def set_option(self, name, value):
        check(self.L.mgic_ctx_set_option(self.h, name.encode(), int(value)))

    def get_option(self, name):
        self.L.mgic_ctx_get_option.restype = C.c_longlong
        self.L.mgic_ctx_get_option.argtypes = [C.c_void_p, C.c_char_p]
        return self.L.mgic_ctx_get_option(self.h, name.encode())

    def profile(self, enable=True):
        """arm / disarm CUDA-event timing of the finest-level GSRB launches"""
        check(self.L.mgic_ctx_profile(self.h, int(enable)))

    def profile_read(self):
        n, ms = C.c_longlong(), C.c_double()
        check(self.L.mgic_ctx_profile_read(self.h, C.byref(n), C.byref(ms)))
        return n.value, ms.value

    PROFILE_TAGS = ("gsrb_finest", "halo_exchange", "all_gather", "bottom_solve", "restrict", "gsrb_coarser", "prolong", "other")

    def profile_breakdown(self):
        """{category: (launches, total ms)} of the armed profiling pass"""
        out = {}
        for tag, name in enumerate(self.PROFILE_TAGS):
            n, ms = C.c_longlong(), C.c_double()
            check(self.L.mgic_ctx_profile_read_tag(self.h, tag, C.byref(n), C.byref(ms)))
            if n.value:
                out[name] = (n.value, ms.value)
        return out

    @property
    def stream(self):
        return self.L.mgic_ctx_stream(self.h)

    @staticmethod
    def alloc_stats():
        """(cudaMalloc calls, seconds inside them, cudaFree calls, seconds inside them) of the library in this process"""
        a, f, sa, sf = C.c_longlong(), C.c_longlong(), C.c_double(), C.c_double()
        check(lib().mgic_alloc_stats(C.byref(a), C.byref(sa), C.byref(f), C.byref(sf)))
        return a.value, sa.value, f.value, sf.value

    @staticmethod
    def guard_check():
        """(arrays found with a damaged guard band so far in this process, live arrays checked now); bands exist when
        MGIC_ARENA_GUARD=<bytes> is in the environment (the GPU test suite sets it)"""
        v, n = C.c_longlong(), C.c_longlong()
        check(lib().mgic_arena_guard_check(C.byref(v), C.byref(n)))
        return v.value, n.value

    @property
    def launch_count(self):
        return self.L.mgic_ctx_launch_count(self.h)

    def close(self):
        if self.h:
            self.L.mgic_ctx_destroy(self.h)
            self.h = None


class LevelField:
    """Device-resident LevelData<FArrayBox> (1 component) on the level of `op`."""

    def __init__(self, op, handle=None, owned=True):
        self.op, self.L = op, op.L
        self.owned = owned
        if handle is None:
            handle = C.c_void_p()
            check(self.L.mgic_field_create(op.h, C.byref(handle)))
        self.h = handle

    @property
    def shape(self):
        n = self.op.n
        return (n[2], n[1], n[0])

    def upload(self, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        if a.shape != self.shape:
            raise MgicError(f"array shape {a.shape} != level shape {self.shape}")
        check(self.L.mgic_field_upload(self.h, a))
        return self

    def download(self, out=None):
        if out is None:
            out = np.zeros(self.shape, dtype=np.float64)
        check(self.L.mgic_field_download(self.h, out))
        return out

    def upload_fab(self, fab, lo, hi, region_lo=None, region_hi=None):
        i3 = C.c_int * 3
        rl, rh = (lo if region_lo is None else region_lo), (hi if region_hi is None else region_hi)
        check(self.L.mgic_field_upload_fab(self.h, np.ascontiguousarray(fab, dtype=np.float64), i3(*lo), i3(*hi), i3(*rl), i3(*rh)))

    def download_fab(self, fab, lo, hi, region_lo=None, region_hi=None):
        i3 = C.c_int * 3
        rl, rh = (lo if region_lo is None else region_lo), (hi if region_hi is None else region_hi)
        check(self.L.mgic_field_download_fab(self.h, fab, i3(*lo), i3(*hi), i3(*rl), i3(*rh)))
        self.op.ctx.sync()

    def close(self):
        if self.h and self.owned:
            self.L.mgic_field_destroy(self.h)
        self.h = None


class VariableCoeffPoissonOperator:
    """L = alpha*aCoef*I - beta*bCoef*Laplacian on one level (VariableCoeffPoissonOperator.H:25)."""

    def __init__(self, ctx, n, dx, alpha=1.0, beta=-1.0, bc_lo=(0, 0, 0), bc_hi=(0, 0, 0), bc_value=0.0, k0=0,
                 nz_local=None, handle=None):
        self.ctx, self.L = ctx, ctx.L
        self.owned = handle is None
        if handle is None:
            i3 = C.c_int * 3
            handle = C.c_void_p()
            nzl = n[2] - k0 if nz_local is None else nz_local
            check(self.L.mgic_op_create(ctx.h, i3(*n), k0, nzl, dx, alpha, beta, i3(*bc_lo), i3(*bc_hi), bc_value,
                                        C.byref(handle)))
        self.h = handle
        n3 = (C.c_int * 3)()
        k0_, nzl_, dx_ = C.c_int(), C.c_int(), C.c_double()
        check(self.L.mgic_op_dims(self.h, n3, C.byref(k0_), C.byref(nzl_), C.byref(dx_)))
        self.n, self.k0, self.nz_local, self.dx = (n3[0], n3[1], n3[2]), k0_.value, nzl_.value, dx_.value

    @classmethod
    def patch(cls, ctx, n_domain, lo, hi, dx, dx_coarse=None, alpha=1.0, beta=-1.0, bc_lo=(0, 0, 0), bc_hi=(0, 0, 0), bc_value=0.0):
        """The operator on one box [lo, hi] of an AMR level > 0: coarse-fine faces get [Chombo] homogeneousCFInterp in
        relax / restrictResidual / preCond (VariableCoeffPoissonOperator.cpp:156,296)."""
        i3 = C.c_int * 3
        h = C.c_void_p()
        check(ctx.L.mgic_op_create_patch(ctx.h, i3(*n_domain), i3(*lo), i3(*hi), dx, 2 * dx if dx_coarse is None else dx_coarse,
                                         alpha, beta, i3(*bc_lo), i3(*bc_hi), bc_value, C.byref(h)))
        op = cls(ctx, None, None, handle=h)
        op.owned = True
        return op

    @classmethod
    def patch_boxes(cls, ctx, n_domain, boxes, dx, dx_coarse=None, alpha=1.0, beta=-1.0, bc_lo=(0, 0, 0), bc_hi=(0, 0, 0), bc_value=0.0):
        """The operator on an AMR level (or one connected part of it) made of several boxes [(lo, hi), ...] that may touch
        and whose union need not be a rectangle: one array over the bounding box plus a cell mask."""
        i3 = C.c_int * 3
        flat = []
        for lo, hi in boxes:
            flat += list(lo) + list(hi)
        arr = (C.c_int * len(flat))(*flat)
        h = C.c_void_p()
        check(ctx.L.mgic_op_create_patch_boxes(ctx.h, i3(*n_domain), len(boxes), arr, dx, 2 * dx if dx_coarse is None else dx_coarse,
                                               alpha, beta, i3(*bc_lo), i3(*bc_hi), bc_value, C.byref(h)))
        op = cls(ctx, None, None, handle=h)
        op.owned = True
        return op

    @property
    def valid_cells(self):
        return self.L.mgic_op_valid_cells(self.h)

    def mask(self):
        """1 = cell of the level's boxes, over the bounding box, [k, j, i]"""
        out = np.zeros((self.n[2], self.n[1], self.n[0]), dtype=np.uint8)
        check(self.L.mgic_op_get_mask(self.h, out))
        return out

    # AMRLevelOp::create
    def create(self):
        return LevelField(self)

    def setCoefs(self, aCoef, bCoef, alpha, beta):
        self._a, self._b = aCoef, bCoef
        check(self.L.mgic_op_set_coefs(self.h, aCoef.h, bCoef.h if bCoef is not None else None, alpha, beta))

    def setAlphaAndBeta(self, alpha, beta):
        check(self.L.mgic_op_set_alpha_beta(self.h, alpha, beta))

    def resetLambda(self):
        check(self.L.mgic_op_reset_lambda(self.h))

    def computeLambda(self):
        check(self.L.mgic_op_compute_lambda(self.h))

    def lambda_field(self):
        h = C.c_void_p()
        check(self.L.mgic_op_get_lambda(self.h, C.byref(h)))
        return LevelField(self, h, owned=False)

    def set_smoother(self, kind):
        check(self.L.mgic_op_set_smoother(self.h, kind))

    def relax(self, e, residual, iterations):
        check(self.L.mgic_op_relax(self.h, e.h, residual.h, iterations))

    def levelGSRB(self, e, residual):
        check(self.L.mgic_op_relax(self.h, e.h, residual.h, 1))

    def gsrb_color(self, e, residual, whichPass):
        check(self.L.mgic_op_gsrb_color(self.h, e.h, residual.h, whichPass))

    def levelJacobi(self, e, residual):
        check(self.L.mgic_op_level_jacobi(self.h, e.h, residual.h))

    def residual(self, lhs, phi, rhs, homogeneous=False):
        check(self.L.mgic_op_residual(self.h, lhs.h, phi.h, rhs.h, int(homogeneous)))

    def applyOp(self, lhs, phi, homogeneous=False):
        check(self.L.mgic_op_apply(self.h, lhs.h, phi.h, int(homogeneous)))

    def applyOpNoBoundary(self, lhs, phi):
        check(self.L.mgic_op_apply_no_boundary(self.h, lhs.h, phi.h))

    # [Chombo] AMRPoissonOp::AMROperatorNF / AMRResidualNF on a patch (QuadCFInterp from the coarser level's field)
    def AMROperatorNF(self, lhs, phi, phiCoarse, coarse_lo=(0, 0, 0), homogeneous=False):
        check(self.L.mgic_op_amr_operator_nf(self.h, lhs.h, phi.h, phiCoarse.h, (C.c_int * 3)(*coarse_lo), int(homogeneous)))

    def AMRResidualNF(self, lhs, phi, phiCoarse, rhs, coarse_lo=(0, 0, 0), homogeneous=False):
        check(self.L.mgic_op_amr_residual_nf(self.h, lhs.h, phi.h, phiCoarse.h, (C.c_int * 3)(*coarse_lo), rhs.h, int(homogeneous)))

    def cf_ghosts(self, face):
        """QuadCFInterp's ghost values on one coarse-fine face (2-D array, slow axis first) after AMROperatorNF / AMRResidualNF"""
        d = face // 2
        ta, tb = (1, 2) if d == 0 else ((0, 2) if d == 1 else (0, 1))
        out = np.empty((self.n[tb], self.n[ta]))
        check(self.L.mgic_op_cf_ghosts(self.h, face, out))
        return out

    def restrictResidual(self, resCoarse, phiFine, rhsFine):
        check(self.L.mgic_op_restrict_residual(self.h, resCoarse.h, phiFine.h, rhsFine.h))

    def prolongIncrement(self, phiThisLevel, correctCoarse):
        check(self.L.mgic_op_prolong_increment(self.h, phiThisLevel.h, correctCoarse.h))

    def preCond(self, phi, rhs):
        check(self.L.mgic_op_precond(self.h, phi.h, rhs.h))

    def norm(self, x, ord=0):
        out = C.c_double()
        check(self.L.mgic_op_norm(self.h, x.h, ord, C.byref(out)))
        return out.value

    def dotProduct(self, x, y):
        out = C.c_double()
        check(self.L.mgic_op_dot(self.h, x.h, y.h, C.byref(out)))
        return out.value

    def incr(self, y, x, scale):
        check(self.L.mgic_op_incr(self.h, y.h, x.h, scale))

    def axby(self, y, x1, x2, a, b):
        check(self.L.mgic_op_axby(self.h, y.h, x1.h, x2.h, a, b))

    def scale(self, y, s):
        check(self.L.mgic_op_scale(self.h, y.h, s))

    def assign(self, y, x):
        check(self.L.mgic_op_assign(self.h, y.h, x.h))

    assignLocal = assign

    def setToZero(self, y):
        check(self.L.mgic_op_set_to_zero(self.h, y.h))

    def setVal(self, y, v):
        check(self.L.mgic_op_set_val(self.h, y.h, v))

    def close(self):
        if self.h and self.owned:
            self.L.mgic_op_destroy(self.h)
        self.h = None


class VariableCoeffPoissonOperatorFactory:
    """define() + MGnewOp(depth) until NULL (VariableCoeffPoissonOperatorFactory.cpp:59-106,139-234), i.e. the
    hierarchy [Chombo] MultiGrid::define builds, plus MultiGrid::oneCycle and the solvers that drive it."""

    def __init__(self, ctx, params, aCoef, bCoef, keep_b=False):
        self.ctx, self.L = ctx, ctx.L
        self.params = params if not isinstance(params, dict) else make_params(params)
        self._a, self._b = aCoef, bCoef
        h = C.c_void_p()
        check(self.L.mgic_mg_create_ex(ctx.h, C.byref(self.params), aCoef.h, bCoef.h if bCoef is not None else None,
                                       1 if keep_b else 0, C.byref(h)))
        self.h = h
        self.depths = self.L.mgic_mg_depths(h)

    def MGnewOp(self, depth):
        h = C.c_void_p()
        check(self.L.mgic_mg_op(self.h, depth, C.byref(h)))
        if not h:
            return None
        return VariableCoeffPoissonOperator(self.ctx, None, None, handle=h)

    def AMRnewOp(self):
        return self.MGnewOp(0)

    def scratch(self, depth):
        e, r = C.c_void_p(), C.c_void_p()
        check(self.L.mgic_mg_scratch(self.h, depth, C.byref(e), C.byref(r)))
        op = self.MGnewOp(depth)
        return LevelField(op, e, owned=False), LevelField(op, r, owned=False)

    def refresh_coefs(self):
        check(self.L.mgic_mg_refresh_coefs(self.h))

    def set_smoother(self, kind):
        check(self.L.mgic_mg_set_smoother(self.h, kind))

    @property
    def b_is_one(self):
        return bool(self.L.mgic_mg_b_is_one(self.h))

    # [Chombo] MultiGrid::oneCycle(e, r), homogeneous
    def vcycle(self, e, r):
        check(self.L.mgic_mg_vcycle(self.h, e.h, r.h))

    # setToZero(e); oneCycle(e, r) -- the first V-cycle of [Chombo] MultilevelLinearOp::preCond
    def vcycle_from_zero(self, e, r):
        check(self.L.mgic_mg_vcycle_from_zero(self.h, e.h, r.h))

    def bottom_solve(self, e, r):
        it = C.c_int()
        check(self.L.mgic_mg_bottom_solve(self.h, e.h, r.h, C.byref(it)))
        return it.value

    @property
    def last_bottom_iterations(self):
        return self.L.mgic_mg_last_bottom_iterations(self.h)

    # solver.solve(dpsi, rhs): BiCGStab preconditioned by numMGIterations V-cycles (Main_PoissonSolver.cpp:173-184)
    def solve(self, dpsi, rhs, max_norms=512):
        it, st = C.c_int(), C.c_int()
        norms = (C.c_double * max_norms)()
        check(self.L.mgic_mg_outer_solve(self.h, dpsi.h, rhs.h, C.byref(it), C.byref(st), norms, max_norms))
        return it.value, st.value, np.array(norms[: min(it.value + 1, max_norms)])

    def close(self):
        if self.h:
            self.L.mgic_mg_destroy(self.h)
            self.h = None


class AMRHierarchy:
    """[Chombo] AMRMultiGrid::AMRVCycle, MultilevelLinearOp and the outer BiCGStab over a hierarchy of levels: level 0 = the
    factory's MG hierarchy, finer levels = lists of patch operators (VariableCoeffPoissonOperator.patch), each nested with
    ratio 2 in one array of the level below and not touching its siblings.  `patches`: a flat list (a chain, one patch per
    level) or a list of lists (level 1's patches, level 2's patches, ...).  A level vector is a list of fields: the base
    level's, then one per patch in that order."""

    def __init__(self, factory, patches):
        self.L, self.factory = factory.L, factory
        levels = [list(p) if isinstance(p, (list, tuple)) else [p] for p in patches]
        self.patches = [q for lv in levels for q in lv]
        arr = (C.c_void_p * max(len(self.patches), 1))(*[p.h.value for p in self.patches])
        counts = (C.c_int * max(len(levels), 1))(*[len(lv) for lv in levels])
        h = C.c_void_p()
        check(self.L.mgic_amr_create_levels(factory.h, len(levels), counts, arr, C.byref(h)))
        self.h = h
        self.nodes = self.L.mgic_amr_nodes(h)
        self.levels = self.L.mgic_amr_levels(h)

    def _vec(self, fields):
        if len(fields) != self.nodes:
            raise MgicError(f"a level vector of this hierarchy has {self.nodes} fields, got {len(fields)}")
        return (C.c_void_p * self.nodes)(*[f.h.value for f in fields])

    def node_info(self, node):
        lv, par = C.c_int(), C.c_int()
        check(self.L.mgic_amr_node_info(self.h, node, C.byref(lv), C.byref(par)))
        return lv.value, par.value

    def create(self):
        """a new level vector"""
        return [self.factory.MGnewOp(0).create()] + [p.create() for p in self.patches]

    def vcycle(self, corr, res):
        """corr (out) = one AMR V-cycle's correction for the residuals res"""
        check(self.L.mgic_amr_vcycle(self.h, self._vec(corr), self._vec(res)))

    def applyOp(self, lhs, phi, homogeneous=False):
        check(self.L.mgic_amr_apply(self.h, self._vec(lhs), self._vec(phi), int(homogeneous)))

    def residual(self, res, phi, rhs, homogeneous=False):
        check(self.L.mgic_amr_residual(self.h, self._vec(res), self._vec(phi), self._vec(rhs), int(homogeneous)))

    def zeroCovered(self, x):
        check(self.L.mgic_amr_zero_covered(self.h, self._vec(x)))

    def averageDown(self, x):
        check(self.L.mgic_amr_average_down(self.h, self._vec(x)))

    def norm(self, x, ord=0):
        out = C.c_double()
        check(self.L.mgic_amr_norm(self.h, self._vec(x), ord, C.byref(out)))
        return out.value

    def dotProduct(self, x, y):
        out = C.c_double()
        check(self.L.mgic_amr_dot(self.h, self._vec(x), self._vec(y), C.byref(out)))
        return out.value

    def preCond(self, cor, res):
        check(self.L.mgic_amr_precond(self.h, self._vec(cor), self._vec(res)))

    def solve(self, dpsi, rhs, max_norms=256):
        """solver.solve(dpsi, rhs) on the hierarchy: (iterations, exit status, residual max-norm history)"""
        it, st = C.c_int(), C.c_int()
        norms = np.zeros(max_norms)
        check(self.L.mgic_amr_outer_solve(self.h, self._vec(dpsi), self._vec(rhs), C.byref(it), C.byref(st),
                                          norms.ctypes.data_as(C.POINTER(C.c_double)), max_norms))
        return it.value, st.value, norms[:min(it.value + 1, max_norms)].copy()

    def close(self):
        if self.h:
            self.L.mgic_amr_destroy(self.h)
            self.h = None


class MultigridVars:
    """multigrid_vars (8 components, MultigridUserVariables.hpp) + the Set* functions of Source/SetLevelData.cpp."""

    NAMES = ("psi", "A11_0", "A12_0", "A13_0", "A22_0", "A23_0", "A33_0", "phi_0")

    def __init__(self, ctx, params, k0=0, nz_local=None):
        self.ctx, self.L = ctx, ctx.L
        self.params = params if not isinstance(params, dict) else make_params(params)
        nzl = self.params.N[2] - k0 if nz_local is None else nz_local
        h = C.c_void_p()
        check(self.L.mgic_vars_create(ctx.h, C.byref(self.params), k0, nzl, C.byref(h)))
        self.h, self.k0, self.nz_local = h, k0, nzl

    def set_initial_conditions(self, dpsi=None):
        check(self.L.mgic_set_initial_conditions(self.h, dpsi.h if dpsi is not None else None))

    def set_a_coef(self, aCoef, constant_K=0.0):
        check(self.L.mgic_set_a_coef(self.h, aCoef.h, constant_K))

    def set_b_coef(self, bCoef):
        check(self.L.mgic_set_b_coef(self.h, bCoef.h))

    def set_rhs(self, rhs, constant_K=0.0):
        check(self.L.mgic_set_rhs(self.h, rhs.h, constant_K))

    def set_rhs_and_a_coef(self, rhs, aCoef, constant_K=0.0):
        check(self.L.mgic_set_rhs_and_a_coef(self.h, rhs.h, aCoef.h, constant_K))

    def set_update_psi0(self, op0, dpsi):
        out = C.c_double()
        check(self.L.mgic_update_psi0(self.h, op0.h, dpsi.h, C.byref(out)))
        return out.value

    def download(self, comp, ghosted=False):
        N = self.params.N
        if ghosted:
            out = np.zeros((self.nz_local + 2, N[1] + 2, N[0] + 2))
            check(self.L.mgic_vars_download_ghosted(self.h, comp, out))
        else:
            out = np.zeros((N[2], N[1], N[0]))
            check(self.L.mgic_vars_download(self.h, comp, out))
        return out

    def close(self):
        if self.h:
            self.L.mgic_vars_destroy(self.h)
            self.h = None


class Grids:
    """set_grids (Source/SetGrids.cpp:31-148): the hierarchy of boxes from tagging + [Chombo] BRMeshRefine::regrid.
    Grids.generate needs a GPU (the regrid condition is a device kernel); Grids.regrid is the clustering alone on
    caller-supplied tags (host only)."""

    def __init__(self, handle, L):
        self.h, self.L = handle, L

    @classmethod
    def generate(cls, ctx, params, refine_threshold=0.1, fill_ratio=0.5):
        p = params if not isinstance(params, dict) else make_params(params)
        h = C.c_void_p()
        check(ctx.L.mgic_grids_generate(ctx.h, C.byref(p), refine_threshold, fill_ratio, C.byref(h)))
        return cls(h, ctx.L)

    @classmethod
    def regrid(cls, params, boxes_per_level, tags_per_level, fill_ratio=0.5):
        """boxes_per_level[l] = [(lo, hi), ...], tags_per_level[l] = [(i, j, k), ...] for the existing levels 0 .. top"""
        p = params if not isinstance(params, dict) else make_params(params)
        L = lib()
        mk = lambda v: (C.c_int * max(len(v), 1))(*v)
        nb = [len(b) for b in boxes_per_level]
        fb = [v for lv in boxes_per_level for lo, hi in lv for v in (list(lo) + list(hi))]
        nt = [len(t) for t in tags_per_level]
        ft = [int(v) for lv in tags_per_level for t in lv for v in t]
        h = C.c_void_p()
        check(L.mgic_grids_regrid(C.byref(p), fill_ratio, len(boxes_per_level) - 1, mk(nb), mk(fb), mk(nt), mk(ft), C.byref(h)))
        return cls(h, L)

    @property
    def levels(self):
        return self.L.mgic_grids_levels(self.h)

    def boxes(self, level, parts=False):
        n = self.L.mgic_grids_num_boxes(self.h, level)
        b, part, np_ = (C.c_int * max(6 * n, 1))(), (C.c_int * max(n, 1))(), C.c_int()
        check(self.L.mgic_grids_get_boxes(self.h, level, b, part, C.byref(np_)))
        out = [(tuple(b[6 * q:6 * q + 3]), tuple(b[6 * q + 3:6 * q + 6])) for q in range(n)]
        return (out, list(part[:n]), np_.value) if parts else out

    def nodes(self, level):
        """the level as its connected parts: [[(lo, hi), ...], ...] -- what Hierarchy takes"""
        b, part, n = self.boxes(level, parts=True)
        return [[b[q] for q in range(len(b)) if part[q] == c] for c in range(n)]

    def stats(self, level):
        mx, tg, cells = C.c_double(), C.c_longlong(), C.c_longlong()
        check(self.L.mgic_grids_level_stats(self.h, level, C.byref(mx), C.byref(tg), C.byref(cells)))
        return dict(max_condition=mx.value, tagged_cells=tg.value, cells=cells.value)

    def close(self):
        if self.h:
            self.L.mgic_grids_destroy(self.h)
            self.h = None


class Hierarchy:
    """poissonSolve (Main_PoissonSolver.cpp:45-256) on an AMR hierarchy: levels = [[node, ...], ...] for levels 1, 2, ...;
    a node is a list of boxes (lo, hi) -- one connected component of its level, boxes may touch -- or a single (lo, hi)."""

    WHAT = dict(psi=0, A11=1, A12=2, A13=3, A22=4, A23=5, A33=6, phi=7, dpsi=8, rhs=9, aCoef=10)

    def __init__(self, ctx, params, levels):
        self.ctx, self.L = ctx, ctx.L
        self.params = params if not isinstance(params, dict) else make_params(params)
        nnodes, nboxes, flat = [], [], []
        for lv in levels:
            nnodes.append(len(lv))
            for node in lv:
                boxes = node if isinstance(node, list) else [node]
                nboxes.append(len(boxes))
                for lo, hi in boxes:
                    flat += list(lo) + list(hi)
        mk = lambda v: (C.c_int * max(len(v), 1))(*v)
        h = C.c_void_p()
        check(self.L.mgic_hier_create(ctx.h, C.byref(self.params), len(levels), mk(nnodes), mk(nboxes), mk(flat), C.byref(h)))
        self.h = h
        self.nodes = self.L.mgic_hier_nodes(h)

    @classmethod
    def from_grids(cls, ctx, params, grids):
        """the problem on the hierarchy set_grids produced (Main_PoissonSolver.cpp:278-284: set_grids, then poissonSolve)"""
        self = cls.__new__(cls)
        self.ctx, self.L = ctx, ctx.L
        self.params = params if not isinstance(params, dict) else make_params(params)
        h = C.c_void_p()
        check(self.L.mgic_hier_create_from_grids(ctx.h, C.byref(self.params), grids.h, C.byref(h)))
        self.h = h
        self.nodes = self.L.mgic_hier_nodes(h)
        return self

    def node_info(self, q):
        """(level, lo (i, j, k), n (nx, ny, nz), valid cells)"""
        lvl, cells = C.c_int(), C.c_longlong()
        lo, n = (C.c_int * 3)(), (C.c_int * 3)()
        check(self.L.mgic_hier_node_info(self.h, q, C.byref(lvl), lo, n, C.byref(cells)))
        return lvl.value, tuple(lo), tuple(n), cells.value

    def mask(self, q):
        _, _, n, _ = self.node_info(q)
        out = np.zeros((n[2], n[1], n[0]), dtype=np.uint8)
        check(self.L.mgic_hier_get_mask(self.h, q, out))
        return out

    def set_initial_conditions(self):
        check(self.L.mgic_hier_set_initial_conditions(self.h))

    def nl_iteration(self):
        """one pass of Main_PoissonSolver.cpp:131-212: (norm of dpsi, BiCGStab iterations, exit status)"""
        nrm, it, st = C.c_double(), C.c_int(), C.c_int()
        check(self.L.mgic_hier_nl_iteration(self.h, C.byref(nrm), C.byref(it), C.byref(st)))
        return nrm.value, it.value, st.value

    def nl_iteration_steps(self, constant_K=0.0, timer=None):
        """the same pass step by step, as the reference's driver calls it (Main_PoissonSolver.cpp:154-208); timer(name) is called
        before each step and once ("end") after the last"""
        t = timer or (lambda name: None)
        it, st, nrm = C.c_int(), C.c_int(), C.c_double()
        t("set_sources"); check(self.L.mgic_hier_set_sources(self.h, constant_K))
        t("define_solver"); check(self.L.mgic_hier_define_solver(self.h))
        t("solve"); check(self.L.mgic_hier_solve(self.h, C.byref(it), C.byref(st)))
        t("update_psi"); check(self.L.mgic_hier_update_psi(self.h))
        t("dpsi_norm"); check(self.L.mgic_hier_dpsi_norm(self.h, C.byref(nrm)))
        t("release_solver"); check(self.L.mgic_hier_release_solver(self.h))
        t("end")
        return nrm.value, it.value, st.value

    def nl_solve(self):
        norms = (C.c_double * 64)()
        its = C.c_int()
        check(self.L.mgic_hier_nl_solve(self.h, norms, 64, C.byref(its)))
        return np.array(norms[: its.value])

    def write_checkpoint(self, path, constant_K=0.0):
        """output_final_data (Source/WriteOutput.H:127-227): the GRChombo checkpoint as an MGICCHK1 container"""
        check(self.L.mgic_hier_write_checkpoint(self.h, str(path).encode(), constant_K))

    def download(self, q, what="psi"):
        _, _, n, _ = self.node_info(q)
        out = np.zeros((n[2], n[1], n[0]))
        check(self.L.mgic_hier_download(self.h, q, self.WHAT[what], out))
        return out

    def close(self):
        if self.h:
            self.L.mgic_hier_destroy(self.h)
            self.h = None


def level_op_from_params(ctx, params, k0=0, nz_local=None):
    """The AMR-level operator geometry of params (dx = L/N[0], PoissonParameters.cpp:82) without coefficients."""
    p = params if not isinstance(params, dict) else make_params(params)
    per = bool(p.is_periodic)
    lo = [2] * 3 if per else list(p.bc_lo)
    hi = [2] * 3 if per else list(p.bc_hi)
    return VariableCoeffPoissonOperator(ctx, tuple(p.N), p.L / p.N[0], p.alpha, p.beta, lo, hi, p.bc_value, k0, nz_local)


def nl_solve(ctx, params, want_psi=True):
    """The nonlinear loop of Main_PoissonSolver.cpp:131-216 (single level) on the GPU."""
    p = params if not isinstance(params, dict) else make_params(params)
    norms = (C.c_double * 64)()
    its = C.c_int()
    psi = np.zeros((p.N[2], p.N[1], p.N[0])) if want_psi else None
    check(ctx.L.mgic_nl_solve(ctx.h, C.byref(p), norms, 64, C.byref(its), psi.ctypes.data if want_psi else None))
    return np.array(norms[: its.value]), psi
