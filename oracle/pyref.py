"""ctypes loader for oracle/_ref/libmgic_ref.so (TEST INFRASTRUCTURE): the reference's own C++ (SetLevelData.cpp with
SetBinaryBH.H and MyPhiFunction.H, PoissonParameters.cpp, SetBCs.cpp, VariableCoeffPoissonOperator.cpp,
VariableCoeffPoissonOperatorFactory.cpp) compiled unmodified against a stand-in for the Chombo API it touches
(oracle/ref_shim/chombo_standin.H), plus its two .ChF kernel files translated mechanically to C++ (oracle/chf2c.py);
recipe: `make -C oracle ref`.  The library contains nothing of the oracle.  Used by tests/test_reference_pins.py and
tests/golden/make_reference_golden.py to pin the oracle's restatement to the reference."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libmgic_ref.so")
# the same reference C++ with its .ChF symbols resolved by the product's Fortran-ABI drop-ins in libmgic_b200.so
# (include/mgic_chf.h): needs a GPU to run
SO_CUDA = os.path.join(_HERE, "_ref", "libmgic_ref_cuda.so")
B200_SO = os.path.join(os.path.dirname(_HERE), "mg_ic_code_b200", "lib", "libmgic_b200.so")
REFERENCE = os.environ.get("MGIC_REFERENCE", "/root/reference")


class RefParams(C.Structure):
    _fields_ = [("G_Newton", C.c_double), ("phi_amplitude", C.c_double), ("phi_wavelength", C.c_double),
                ("bh1_bare_mass", C.c_double), ("bh1_spin", C.c_double), ("bh1_momentum", C.c_double), ("bh1_offset", C.c_double),
                ("bh2_bare_mass", C.c_double), ("bh2_spin", C.c_double), ("bh2_momentum", C.c_double), ("bh2_offset", C.c_double),
                ("L", C.c_double * 3)]


class RefPoissonParameters(C.Structure):
    _fields_ = [("nCells", C.c_int * 3), ("maxGridSize", C.c_int), ("blockFactor", C.c_int), ("bufferSize", C.c_int),
                ("coefficient_average_type", C.c_int), ("verbosity", C.c_int), ("periodic", C.c_int * 3), ("maxLevel", C.c_int),
                ("numLevels", C.c_int), ("refRatio0", C.c_int), ("refRatioLast", C.c_int), ("domainLo", C.c_int * 3),
                ("domainHi", C.c_int * 3), ("domainPeriodic", C.c_int * 3),
                ("fillRatio", C.c_double), ("refineThresh", C.c_double), ("coarsestDx", C.c_double), ("domainLength", C.c_double * 3),
                ("probLo", C.c_double * 3), ("probHi", C.c_double * 3), ("alpha", C.c_double), ("beta", C.c_double),
                ("G_Newton", C.c_double), ("phi_amplitude", C.c_double), ("phi_wavelength", C.c_double),
                ("bh1_bare_mass", C.c_double), ("bh2_bare_mass", C.c_double), ("bh1_spin", C.c_double), ("bh2_spin", C.c_double),
                ("bh1_momentum", C.c_double), ("bh2_momentum", C.c_double), ("bh1_offset", C.c_double), ("bh2_offset", C.c_double)]


def available():
    return os.path.exists(SO) or os.path.isdir(os.path.join(REFERENCE, "Source"))


def build():
    """compile the reference's files where they lie (only possible where /root/reference exists); returns the .so or None"""
    if os.path.isdir(os.path.join(REFERENCE, "Source")):
        targets = ["ref"] + (["ref_cuda"] if os.path.exists(B200_SO) else [])
        try:
            subprocess.check_call(["make", "-C", _HERE] + targets + [f"REF={REFERENCE}", f"PYTHON={sys.executable}"],
                                  stdout=subprocess.DEVNULL)
        except (subprocess.CalledProcessError, OSError):
            if not os.path.exists(SO):      # a prebuilt library (e.g. on a read-only tree) is still usable
                raise
    return SO if os.path.exists(SO) else None


_libs = {}


def lib(cuda=False):
    """cuda=True: the variant whose kernels are the product's CUDA drop-ins behind the Fortran ABI (compute calls need a GPU)"""
    if cuda not in _libs:
        if build() is None:
            raise RuntimeError("oracle/_ref/libmgic_ref.so is not built and /root/reference is absent")
        if cuda and not os.path.exists(SO_CUDA):
            raise RuntimeError("oracle/_ref/libmgic_ref_cuda.so is not built (needs mg_ic_code_b200/lib/libmgic_b200.so and /root/reference)")
        L = C.CDLL(SO_CUDA if cuda else SO)
        nd = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
        L.ref_set_level_data.argtypes = [C.POINTER(RefParams), C.c_int * 3, C.c_double, C.c_double, C.c_void_p, nd, nd, nd, nd]
        L.ref_set_level_data.restype = C.c_int
        L.ref_condition.argtypes = [C.POINTER(RefParams), C.c_int * 3, C.c_double, C.c_int, nd]
        L.ref_condition.restype = C.c_int
        L.ref_output_data.argtypes = [C.POINTER(RefParams), C.c_int * 3, C.c_double, C.c_double, C.c_void_p, nd]
        L.ref_output_data.restype = C.c_int
        L.ref_point_values.argtypes = [C.POINTER(RefParams), C.c_double * 3, C.c_double * 6, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.ref_point_values.restype = None
        L.ref_m_value.argtypes = [C.POINTER(RefParams), C.c_double, C.c_double]
        L.ref_m_value.restype = C.c_double
        L.ref_get_poisson_parameters.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(RefPoissonParameters),
                                                 C.c_char_p, C.c_int]
        L.ref_get_poisson_parameters.restype = C.c_int
        i3 = C.c_int * 3
        L.ref_op_create.argtypes = [i3, C.c_int, C.c_double, C.c_double, C.c_double, i3, i3, C.c_double]
        L.ref_op_create.restype = C.c_void_p
        L.ref_op_destroy.argtypes = [C.c_void_p]
        L.ref_op_num_boxes.argtypes = [C.c_void_p]
        L.ref_op_set.argtypes = [C.c_void_p, C.c_int, nd]
        L.ref_op_get.argtypes = [C.c_void_p, C.c_int, nd]
        L.ref_op_get_coarse_residual.argtypes = [C.c_void_p, nd]
        L.ref_op_relax.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_op_relax_status.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_op_relax_status_msg.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_int]
        for name in ("ref_op_residual", "ref_op_apply"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_int]
        for name in ("ref_op_restrict", "ref_op_precond"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.ref_factory_create.argtypes = [i3, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, nd, nd]
        L.ref_factory_create.restype = C.c_void_p
        L.ref_factory_destroy.argtypes = [C.c_void_p]
        L.ref_factory_depths.argtypes = [C.c_void_p]
        L.ref_factory_average_type.argtypes = [C.c_void_p]
        L.ref_factory_level.argtypes = [C.c_void_p, C.c_int, i3, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.ref_factory_level.restype = None
        L.ref_factory_get.argtypes = [C.c_void_p, C.c_int, C.c_int, nd]
        _libs[cuda] = L
    return _libs[cuda]


def to_struct(params):
    """params: the oracle's parameter dict (oracle.pyoracle.default_params)"""
    p = RefParams()
    for k, _ in RefParams._fields_:
        if k != "L":
            setattr(p, k, params[k])
    N, L = params["N"], params["L"]
    # PoissonParameters.cpp:70-85: cubic cells, dx = L / N[0], domainLength[d] = N[d] * dx
    p.L = (C.c_double * 3)(*[L * N[d] / N[0] for d in range(3)])
    return p


def set_level_data(params, constant_K=0.0, dpsi_ghosted=None, cuda=False):
    """The reference's set_initial_conditions [+ set_update_psi0(dpsi)] + set_a_coef + set_b_coef + set_rhs on one box:
    (multigrid_vars [8, nz+6, ny+6, nx+6], rhs, aCoef, bCoef [nz, ny, nx])"""
    N = tuple(params["N"])
    dx = params["L"] / N[0]
    g = (N[2] + 6, N[1] + 6, N[0] + 6)
    mg = np.zeros((8,) + g)
    rhs, a, b = (np.zeros((N[2], N[1], N[0])) for _ in range(3))
    d = None
    if dpsi_ghosted is not None:
        d = np.ascontiguousarray(dpsi_ghosted, dtype=np.float64)
        assert d.shape == g
    p = to_struct(params)
    lib(cuda).ref_set_level_data(C.byref(p), (C.c_int * 3)(*N), dx, constant_K, None if d is None else d.ctypes.data, mg, rhs, a, b)
    return mg, rhs, a, b


def condition(params, mode=0, dx=None):
    """The reference's set_regrid_condition (mode 0) / set_constant_K_integrand (mode 1) on freshly initialised data, one
    box of N cells at spacing dx (default L / N[0]): [nz, ny, nx]"""
    N = tuple(params["N"])
    out = np.zeros((N[2], N[1], N[0]))
    p = to_struct(params)
    lib().ref_condition(C.byref(p), (C.c_int * 3)(*N), params["L"] / N[0] if dx is None else dx, mode, out)
    return out


def output_data(params, constant_K=0.0, dpsi_ghosted=None):
    """The reference's set_output_data after set_initial_conditions [+ set_update_psi0(dpsi)]: the 32 GRChombo variables with
    three ghost layers, [32, nz+6, ny+6, nx+6]"""
    N = tuple(params["N"])
    g = (N[2] + 6, N[1] + 6, N[0] + 6)
    out = np.zeros((32,) + g)
    d = None
    if dpsi_ghosted is not None:
        d = np.ascontiguousarray(dpsi_ghosted, dtype=np.float64)
        assert d.shape == g
    p = to_struct(params)
    lib().ref_output_data(C.byref(p), (C.c_int * 3)(*N), params["L"] / N[0], constant_K, None if d is None else d.ctypes.data, out)
    return out


def point_values(params, loc):
    """(Aij[6] in the order A11 A12 A13 A22 A23 A33, psi_bh, phi) from get_Aij / set_binary_bh_psi / my_phi_function"""
    p = to_struct(params)
    A = (C.c_double * 6)()
    psi, phi = C.c_double(), C.c_double()
    lib().ref_point_values(C.byref(p), (C.c_double * 3)(*loc), A, C.byref(psi), C.byref(phi))
    return np.array(A[:]), psi.value, phi.value


def m_value(params, phi_here, constant_K):
    p = to_struct(params)
    return lib().ref_m_value(C.byref(p), phi_here, constant_K)


def get_poisson_parameters(path, overrides=()):
    """The reference's getPoissonParameters (Source/PoissonParameters.cpp:26-131) on an input file + 'key = value' overrides:
    dict of the PoissonParameters members; raises RuntimeError with the MayDay / ParmParse message."""
    out = RefPoissonParameters()
    err = C.create_string_buffer(512)
    ov = (C.c_char_p * max(len(overrides), 1))(*[o.encode() for o in overrides])
    if lib().ref_get_poisson_parameters(str(path).encode(), len(overrides), ov, C.byref(out), err, 512):
        raise RuntimeError(err.value.decode())
    d = {}
    for k, t in RefPoissonParameters._fields_:
        v = getattr(out, k)
        d[k] = list(v) if hasattr(v, "__len__") else v
    return d


class ReferenceOperator:
    """The reference's VariableCoeffPoissonOperator (Source/VariableCoeffPoissonOperator.cpp, compiled unmodified) on one level
    that covers its domain, in boxes of max_grid_size, with the reference's ParseBC (Source/SetBCs.cpp) as boundary
    function.  Inner loops: the C restatement of the .ChF kernels (no Fortran compiler); base class and containers:
    oracle/ref_shim/chombo_standin.H.  Fields E (dpsi, one ghost layer), R (rhs), A, B as arrays [k, j, i]."""
    FIELD = dict(E=0, R=1, A=2, B=3, LAMBDA=4, TMP=5)

    def __init__(self, params, cuda=False):
        """params: the oracle's parameter dict; cuda=True: the .ChF symbols are libmgic_b200.so's CUDA drop-ins"""
        self.L = lib(cuda)
        N = tuple(params["N"])
        i3 = C.c_int * 3
        self.shape = (N[2], N[1], N[0])
        self.h = self.L.ref_op_create(i3(*N), params["max_grid_size"], params["L"] / N[0], params["alpha"], params["beta"],
                                      i3(*params["bc_lo"]), i3(*params["bc_hi"]), params["bc_value"])

    def close(self):
        if self.h:
            self.L.ref_op_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def num_boxes(self):
        return self.L.ref_op_num_boxes(self.h)

    def set(self, field, arr):
        assert self.L.ref_op_set(self.h, self.FIELD[field], np.ascontiguousarray(arr, dtype=np.float64)) == 0

    def get(self, field):
        out = np.empty(self.shape)
        assert self.L.ref_op_get(self.h, self.FIELD[field], out) == 0
        return out

    def relax(self, iterations, mode=1):
        """AMRPoissonOp::relax: mode 1 levelGSRB (Chombo's default), 4 levelJacobi"""
        self.L.ref_op_relax(self.h, iterations, mode)

    def residual(self, homogeneous):
        self.L.ref_op_residual(self.h, int(homogeneous))
        return self.get("TMP")

    def apply(self, homogeneous):
        self.L.ref_op_apply(self.h, int(homogeneous))
        return self.get("TMP")

    def restrict(self):
        self.L.ref_op_restrict(self.h)
        out = np.empty(tuple(s // 2 for s in self.shape))
        self.L.ref_op_get_coarse_residual(self.h, out)
        return out

    def precond(self):
        self.L.ref_op_precond(self.h)


class ReferenceFactory:
    """The reference's VariableCoeffPoissonOperatorFactory (Source/VariableCoeffPoissonOperatorFactory.cpp, compiled
    unmodified), built by defineOperatorFactory (what Main_PoissonSolver.cpp:163-166 calls) on one AMR level, and asked for
    MGnewOp(depth) until it returns NULL (what MultiGrid::define does).  [Chombo] CoarseAverage / coarsenable / s_maxCoarse
    are the stand-in's restatements."""

    def __init__(self, params, aCoef, bCoef, coefficient_average_type=None):
        self.L = lib()
        N = tuple(params["N"])
        t = params["coefficient_average_type"] if coefficient_average_type is None else coefficient_average_type
        self.h = self.L.ref_factory_create((C.c_int * 3)(*N), params["max_grid_size"], params["L"] / N[0], params["alpha"], params["beta"],
                                           t, np.ascontiguousarray(aCoef, dtype=np.float64), np.ascontiguousarray(bCoef, dtype=np.float64))
        self.depths = self.L.ref_factory_depths(self.h)
        self.average_type = self.L.ref_factory_average_type(self.h)

    def level(self, depth):
        """((nx, ny, nz), dx, number of boxes) of MGnewOp(depth)"""
        n, dx, nb = (C.c_int * 3)(), C.c_double(), C.c_int()
        self.L.ref_factory_level(self.h, depth, n, C.byref(dx), C.byref(nb))
        return tuple(n), dx.value, nb.value

    def get(self, field, depth):
        (nx, ny, nz), _, _ = self.level(depth)
        out = np.empty((nz, ny, nx))
        assert self.L.ref_factory_get(self.h, depth, ReferenceOperator.FIELD[field], out) == 0
        return out

    def close(self):
        if self.h:
            self.L.ref_factory_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
