"""ctypes binding of oracle/build/libmgic_oracle.so (test infrastructure only)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "build", "libmgic_oracle.so")


def build(force=False):
    """Compile the oracle with the committed Makefile (g++ -O3 -fopenmp, no FMA contraction)."""
    if force or not os.path.exists(_SO) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_SO)
        for f in ("mgic_oracle.cpp", "mgic_oracle.h", "Makefile")
    ):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


class OrcParams(C.Structure):
    _fields_ = [
        ("alpha", C.c_double), ("beta", C.c_double),
        ("G_Newton", C.c_double), ("phi_amplitude", C.c_double), ("phi_wavelength", C.c_double),
        ("bh1_bare_mass", C.c_double), ("bh1_spin", C.c_double), ("bh1_momentum", C.c_double), ("bh1_offset", C.c_double),
        ("bh2_bare_mass", C.c_double), ("bh2_spin", C.c_double), ("bh2_momentum", C.c_double), ("bh2_offset", C.c_double),
        ("L", C.c_double), ("bc_value", C.c_double), ("tolerance", C.c_double),
        ("N", C.c_int * 3), ("max_level", C.c_int), ("block_factor", C.c_int), ("max_grid_size", C.c_int),
        ("coefficient_average_type", C.c_int), ("is_periodic", C.c_int),
        ("bc_lo", C.c_int * 3), ("bc_hi", C.c_int * 3),
        ("numMGsmooth", C.c_int), ("numMGIterations", C.c_int), ("preCondSolverDepth", C.c_int),
        ("max_iterations", C.c_int), ("max_NL_iterations", C.c_int), ("verbosity", C.c_int),
    ]


def default_params(**over):
    """The reference's params.txt physics/solver block (params.txt:12-84) with max_level = 0."""
    d = dict(
        alpha=1.0, beta=-1.0, G_Newton=1.0, phi_amplitude=0.1, phi_wavelength=1.0,
        bh1_bare_mass=0.5, bh1_spin=0.1, bh1_momentum=0.05, bh1_offset=10.0,
        bh2_bare_mass=0.5, bh2_spin=0.1, bh2_momentum=-0.05, bh2_offset=-10.0,
        L=100.0, bc_value=0.0, tolerance=1.0e-10, N=(64, 64, 64), max_level=0, block_factor=8,
        max_grid_size=16, coefficient_average_type=1, is_periodic=0, bc_lo=(0, 0, 0), bc_hi=(0, 0, 0),
        numMGsmooth=4, numMGIterations=2, preCondSolverDepth=-1, max_iterations=100, max_NL_iterations=6,
        verbosity=2,
    )
    d.update(over)
    return d


def to_struct(d):
    p = OrcParams()
    for k, v in d.items():
        if k in ("N", "bc_lo", "bc_hi"):
            setattr(p, k, (C.c_int * 3)(*[int(x) for x in v]))
        else:
            setattr(p, k, v)
    return p


FIELD = dict(E=0, R=1, A=2, B=3, LAMBDA=4, TMP=5, DPSI=6, RHS=7, MGVAR0=16)

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(OrcParams)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        for name in ("orc_set_initial_conditions",):
            getattr(L, name).argtypes = [C.c_void_p]
        L.orc_set_coefs_and_rhs.argtypes = [C.c_void_p, C.c_double]
        L.orc_define_solver.argtypes = [C.c_void_p]
        L.orc_define_solver.restype = C.c_int
        L.orc_mg_depths.argtypes = [C.c_void_p]
        L.orc_mg_depths.restype = C.c_int
        L.orc_level_dims.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
        L.orc_get_field.argtypes = [C.c_void_p, C.c_int, C.c_int, dp]
        L.orc_set_field.argtypes = [C.c_void_p, C.c_int, C.c_int, dp]
        L.orc_get_field_ghosted.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, dp]
        L.orc_op_relax.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_op_gsrb_color.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_op_residual.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_op_apply.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_op_restrict.argtypes = [C.c_void_p, C.c_int]
        L.orc_op_prolong.argtypes = [C.c_void_p, C.c_int]
        L.orc_op_precond.argtypes = [C.c_void_p, C.c_int]
        L.orc_op_norm.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_op_norm.restype = C.c_double
        L.orc_op_dot.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_op_dot.restype = C.c_double
        L.orc_vcycle.argtypes = [C.c_void_p]
        L.orc_vcycle.restype = C.c_int
        L.orc_bottom_solve.argtypes = [C.c_void_p]
        L.orc_bottom_solve.restype = C.c_int
        L.orc_load_rhs_zero_e.argtypes = [C.c_void_p]
        L.orc_outer_solve.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double), dp, C.c_int]
        L.orc_outer_solve.restype = C.c_int
        L.orc_update_psi0.argtypes = [C.c_void_p]
        L.orc_set_dpsi_with_bc.argtypes = [C.c_void_p, dp]
        L.orc_output_box.argtypes = [C.POINTER(OrcParams), C.c_double, C.c_int * 3, C.c_int * 3, dp, C.c_double, dp]
        L.orc_condition_box.argtypes = [C.POINTER(OrcParams), C.c_double, C.c_int * 3, C.c_int * 3, C.c_int, dp]
        L.orc_update_psi0.restype = C.c_double
        L.orc_nl_solve.argtypes = [C.c_void_p, dp, C.c_int]
        L.orc_nl_solve.restype = C.c_int
        i3 = C.c_int * 3
        L.orc_patch_create.restype = C.c_void_p
        L.orc_patch_create.argtypes = [i3, i3, i3, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, i3, i3, C.c_double]
        L.orc_patch_create_boxes.restype = C.c_void_p
        L.orc_patch_create_boxes.argtypes = [i3, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, i3, i3, C.c_double]
        L.orc_patch_destroy.argtypes = [C.c_void_p]
        L.orc_patch_set.argtypes = [C.c_void_p, C.c_int, dp]
        L.orc_patch_get.argtypes = [C.c_void_p, C.c_int, dp]
        L.orc_patch_num_boxes.argtypes = [C.c_void_p]
        L.orc_patch_set_initial_conditions.argtypes = [C.c_void_p, C.POINTER(OrcParams)]
        L.orc_patch_set_coefs_and_rhs.argtypes = [C.c_void_p, C.c_double]
        L.orc_patch_update_psi.argtypes = [C.c_void_p, dp]
        L.orc_patch_get_var.argtypes = [C.c_void_p, C.c_int, dp]
        L.orc_patch_relax.argtypes = [C.c_void_p, C.c_int]
        L.orc_patch_gsrb_color.argtypes = [C.c_void_p, C.c_int]
        L.orc_patch_restrict.argtypes = [C.c_void_p]
        L.orc_patch_precond.argtypes = [C.c_void_p]
        L.orc_patch_set_coarse.argtypes = [C.c_void_p, dp]
        L.orc_patch_amr_operator_nf.argtypes = [C.c_void_p, C.c_int]
        L.orc_patch_amr_residual_nf.argtypes = [C.c_void_p, C.c_int]
        L.orc_interp_homo.argtypes = [C.c_double] * 4
        L.orc_interp_homo.restype = C.c_double
        _lib = L
    return _lib


class Oracle:
    """One reference problem (Main_PoissonSolver.cpp poissonSolve state) on the CPU oracle."""

    def __init__(self, **over):
        self.params = default_params(**over)
        self._p = to_struct(self.params)
        self.L = lib()
        self.h = self.L.orc_create(C.byref(self._p))
        if not self.h:
            raise RuntimeError("orc_create failed")
        self.depths = 0

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- setup ---------------------------------------------------------------
    def set_initial_conditions(self):
        self.L.orc_set_initial_conditions(self.h)

    def set_coefs_and_rhs(self, constant_K=0.0):
        self.L.orc_set_coefs_and_rhs(self.h, constant_K)

    def define_solver(self):
        self.depths = self.L.orc_define_solver(self.h)
        return self.depths

    def setup(self):
        self.set_initial_conditions()
        self.set_coefs_and_rhs()
        return self.define_solver()

    def dims(self, depth=0):
        n = (C.c_int * 3)()
        dx = C.c_double()
        self.L.orc_level_dims(self.h, depth, n, C.byref(dx))
        return (n[0], n[1], n[2]), dx.value

    # -- fields (numpy arrays indexed [k, j, i]) -----------------------------
    def get(self, field, depth=0, comp=0):
        (nx, ny, nz), _ = self.dims(depth if field not in ("DPSI", "RHS", "MGVAR0") else 0)
        out = np.empty((nz, ny, nx), dtype=np.float64)
        self.L.orc_get_field(self.h, depth, FIELD[field] + comp, out)
        return out

    def get_ghosted(self, field, ng, depth=0, comp=0):
        (nx, ny, nz), _ = self.dims(depth if field not in ("DPSI", "RHS", "MGVAR0") else 0)
        out = np.zeros((nz + 2 * ng, ny + 2 * ng, nx + 2 * ng), dtype=np.float64)
        self.L.orc_get_field_ghosted(self.h, depth, FIELD[field] + comp, ng, out)
        return out

    def set(self, field, arr, depth=0, comp=0):
        self.L.orc_set_field(self.h, depth, FIELD[field] + comp, np.ascontiguousarray(arr, dtype=np.float64))

    # -- operator -------------------------------------------------------------
    def relax(self, depth, iterations):
        self.L.orc_op_relax(self.h, depth, iterations)

    def gsrb_color(self, depth, which):
        self.L.orc_op_gsrb_color(self.h, depth, which)

    def residual(self, depth, homogeneous=True):
        self.L.orc_op_residual(self.h, depth, int(homogeneous))
        return self.get("TMP", depth)

    def apply(self, depth, homogeneous=True):
        self.L.orc_op_apply(self.h, depth, int(homogeneous))
        return self.get("TMP", depth)

    def restrict(self, depth):
        self.L.orc_op_restrict(self.h, depth)

    def prolong(self, depth):
        self.L.orc_op_prolong(self.h, depth)

    def precond(self, depth):
        self.L.orc_op_precond(self.h, depth)

    def norm(self, depth, field, ord=0):
        return self.L.orc_op_norm(self.h, depth, FIELD[field], ord)

    def dot(self, depth, f1, f2):
        return self.L.orc_op_dot(self.h, depth, FIELD[f1], FIELD[f2])

    def vcycle(self):
        return self.L.orc_vcycle(self.h)

    def bottom_solve(self):
        return self.L.orc_bottom_solve(self.h)

    def load_rhs_zero_e(self):
        self.L.orc_load_rhs_zero_e(self.h)

    def outer_solve(self, max_norms=256):
        st = C.c_int()
        fn = C.c_double()
        norms = np.zeros(max_norms)
        it = self.L.orc_outer_solve(self.h, C.byref(st), C.byref(fn), norms, max_norms)
        return it, st.value, fn.value, norms[: it + 1].copy()

    def update_psi0(self, dpsi=None):
        """psi += dpsi over the ghosted boxes; dpsi = the solver's own (None) or given valid cells whose physical ghosts get
        the inhomogeneous BC fill"""
        if dpsi is not None:
            self.L.orc_set_dpsi_with_bc(self.h, np.ascontiguousarray(dpsi, dtype=np.float64))
        return self.L.orc_update_psi0(self.h)

    def nl_solve(self):
        out = np.zeros(64)
        n = self.L.orc_nl_solve(self.h, out, 64)
        return out[:n].copy()

    @property
    def num_threads(self):
        return self.L.orc_num_threads()


def condition_box(params, dx, lo, hi, mode=0):
    """set_regrid_condition (mode 0) / set_constant_K_integrand (mode 1) on fresh initial data over the index box [lo, hi] of a
    level with spacing dx; [k, j, i]"""
    p = to_struct(params)
    out = np.zeros(tuple(hi[d] - lo[d] + 1 for d in (2, 1, 0)))
    lib().orc_condition_box(C.byref(p), dx, (C.c_int * 3)(*lo), (C.c_int * 3)(*hi), mode, out)
    return out


def output_box(params, dx, lo, hi, psi, constant_K=0.0):
    """set_output_data: the 32 GRChombo variables over the index box [lo, hi] from psi [k, j, i] over the same box"""
    p = to_struct(params)
    shape = tuple(hi[d] - lo[d] + 1 for d in (2, 1, 0))
    psi = np.ascontiguousarray(psi, dtype=np.float64)
    assert psi.shape == shape
    out = np.zeros((32,) + shape)
    lib().orc_output_box(C.byref(p), dx, (C.c_int * 3)(*lo), (C.c_int * 3)(*hi), psi, constant_K, out)
    return out


def use_all_host_cores():
    """OpenMP threads = the cores this process may run on, whatever OMP_NUM_THREADS says (torchrun exports 1)."""
    import os
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().orc_set_num_threads(n)
    return lib().orc_num_threads()


class OraclePatch:
    """One AMR level > 0 on the CPU oracle: a box [lo, hi] of the refined domain, split into max_grid_size boxes, with
    [Chombo] homogeneousCFInterp at its coarse-fine faces (VariableCoeffPoissonOperator.cpp:156,296)."""

    def __init__(self, n_domain, lo, hi, dx, dx_crse=None, max_grid_size=8, alpha=1.0, beta=-1.0, bc_lo=(0, 0, 0),
                 bc_hi=(0, 0, 0), bc_value=0.0, boxes=None):
        """boxes = [(lo, hi), ...]: the level as a list of boxes that may touch (lo / hi are then ignored and become the
        bounding box; arrays are bounding-box shaped, cells outside the boxes are not touched)."""
        self.L = lib()
        i3 = C.c_int * 3
        self.boxes = None if boxes is None else [(tuple(a), tuple(b)) for a, b in boxes]
        if boxes is not None:
            lo = [min(b[0][d] for b in self.boxes) for d in range(3)]
            hi = [max(b[1][d] for b in self.boxes) for d in range(3)]
        self.lo, self.hi = tuple(lo), tuple(hi)
        self.shape = tuple(hi[d] - lo[d] + 1 for d in (2, 1, 0))
        self.cshape = tuple(s // 2 for s in self.shape)
        if boxes is not None:
            flat = [v for b in self.boxes for v in (list(b[0]) + list(b[1]))]
            self.h = self.L.orc_patch_create_boxes(i3(*n_domain), len(self.boxes), (C.c_int * len(flat))(*flat), max_grid_size, dx,
                                                   2 * dx if dx_crse is None else dx_crse, alpha, beta, i3(*bc_lo), i3(*bc_hi), bc_value)
        else:
            self.h = self.L.orc_patch_create(i3(*n_domain), i3(*lo), i3(*hi), max_grid_size, dx, 2 * dx if dx_crse is None else dx_crse,
                                             alpha, beta, i3(*bc_lo), i3(*bc_hi), bc_value)
        if not self.h:
            raise ValueError("patch box must lie in the domain and be coarsenable by 2 (even lo, odd hi), max_grid_size even")

    def close(self):
        if self.h:
            self.L.orc_patch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def num_boxes(self):
        return self.L.orc_patch_num_boxes(self.h)

    def set(self, field, arr):
        self.L.orc_patch_set(self.h, FIELD[field], np.ascontiguousarray(arr, dtype=np.float64))

    def mask(self):
        """1 = cell of the level's boxes, over the bounding box, [k, j, i]"""
        m = np.zeros(self.shape, dtype=np.uint8)
        for a, b in (self.boxes or [(self.lo, self.hi)]):
            m[a[2] - self.lo[2]:b[2] - self.lo[2] + 1, a[1] - self.lo[1]:b[1] - self.lo[1] + 1, a[0] - self.lo[0]:b[0] - self.lo[0] + 1] = 1
        return m

    def get(self, field):
        out = np.zeros(self.cshape if field == "TMP" else self.shape, dtype=np.float64)
        self.L.orc_patch_get(self.h, FIELD[field], out)
        return out

    def relax(self, iterations):
        self.L.orc_patch_relax(self.h, iterations)

    # -- the nonlinear loop's per-level steps (Main_PoissonSolver.cpp:93, 154-160, 189-205)
    def set_initial_conditions(self, params):
        self._params = to_struct(params)
        self.L.orc_patch_set_initial_conditions(self.h, C.byref(self._params))

    def set_coefs_and_rhs(self, constant_K=0.0):
        self.L.orc_patch_set_coefs_and_rhs(self.h, constant_K)

    def update_psi(self, dpsi, coarse_dpsi):
        """psi += dpsi over the ghosted boxes; coarse-fine ghosts of dpsi by QuadCFInterp from the coarser level's dpsi"""
        self.set_coarse(coarse_dpsi)
        self.L.orc_patch_update_psi(self.h, np.ascontiguousarray(dpsi, dtype=np.float64))

    def var(self, comp):
        """multigrid_vars component 0..7 (0 = psi), or 8 = rhs, over the bounding box"""
        out = np.zeros(self.shape, dtype=np.float64)
        self.L.orc_patch_get_var(self.h, comp, out)
        return out

    def gsrb_color(self, which):
        self.L.orc_patch_gsrb_color(self.h, which)

    def restrict(self):
        """restrictResidual(e, r) -> the residual on the patch coarsened by 2"""
        self.L.orc_patch_restrict(self.h)
        return self.get("TMP")

    def precond(self):
        self.L.orc_patch_precond(self.h)

    def set_coarse(self, arr):
        """the coarser AMR level's field over its whole domain, [k, j, i], shape n_domain / 2"""
        self.L.orc_patch_set_coarse(self.h, np.ascontiguousarray(arr, dtype=np.float64))

    def amr_operator_nf(self, homogeneous=True):
        """[Chombo] AMROperatorNF: QuadCFInterp from the coarse field, then applyOp"""
        self.L.orc_patch_amr_operator_nf(self.h, int(homogeneous))
        return self.get("RHS")

    def amr_residual_nf(self, homogeneous=True):
        self.L.orc_patch_amr_residual_nf(self.h, int(homogeneous))
        return self.get("RHS")


def interp_homo(dx, dx_crse, far, near):
    return lib().orc_interp_homo(dx, dx_crse, far, near)
