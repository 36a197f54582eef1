"""ctypes binding of libmgic_b200.so (include/mgic.h).  Plumbing only: every compute call runs CUDA kernels."""
import ctypes as C
import os

import numpy as np

from . import build as _build

_lib = None


class MgicError(RuntimeError):
    """Raised where the reference would call MayDay::Error / MayDay::Abort."""


class MgicParams(C.Structure):
    """mgic_params: PoissonParameters (Source/PoissonParameters.H) + solver / BC keys of params.txt."""
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


def library_path():
    return _build.SO


def lib():
    """Load the CUDA library; fail loudly if it is missing or cannot be loaded (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    so = _build.SO
    if not os.path.exists(so):
        raise MgicError(f"{so} is not built: run `python -m mg_ic_code_b200.build` (needs nvcc); there is no CPU fallback")
    L = C.CDLL(so)
    vp, ip, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)
    pvp = C.POINTER(C.c_void_p)
    nd = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    i3 = C.c_int * 3
    L.mgic_last_error.restype = C.c_char_p
    L.mgic_version.restype = C.c_char_p
    sig = {
        "mgic_ctx_create": [C.c_int, pvp], "mgic_ctx_destroy": [vp], "mgic_ctx_sync": [vp],
        "mgic_ctx_set_stream": [vp, vp], "mgic_ctx_profile": [vp, C.c_int],
        "mgic_ctx_set_option": [vp, C.c_char_p, C.c_longlong],
        "mgic_ctx_profile_read": [vp, C.POINTER(C.c_longlong), dp],
        "mgic_ctx_profile_read_tag": [vp, C.c_int, C.POINTER(C.c_longlong), dp], "mgic_ctx_set_rank": [vp, C.c_int, C.c_int],
        "mgic_op_create": [vp, i3, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, i3, i3, C.c_double, pvp],
        "mgic_op_create_patch": [vp, i3, i3, i3, C.c_double, C.c_double, C.c_double, C.c_double, i3, i3, C.c_double, pvp],
        "mgic_op_cf_ghosts": [vp, C.c_int, nd],
        "mgic_op_create_patch_boxes": [vp, i3, C.c_int, ip, C.c_double, C.c_double, C.c_double, C.c_double, i3, i3, C.c_double, pvp],
        "mgic_op_get_mask": [vp, np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")],
        "mgic_amr_create": [vp, C.c_int, pvp, pvp], "mgic_amr_create_levels": [vp, C.c_int, ip, pvp, pvp],
        "mgic_amr_destroy": [vp], "mgic_amr_vcycle": [vp, pvp, pvp], "mgic_amr_node_info": [vp, C.c_int, ip, ip],
        "mgic_amr_apply": [vp, pvp, pvp, C.c_int], "mgic_amr_residual": [vp, pvp, pvp, pvp, C.c_int],
        "mgic_amr_zero_covered": [vp, pvp], "mgic_amr_average_down": [vp, pvp],
        "mgic_amr_norm": [vp, pvp, C.c_int, dp], "mgic_amr_dot": [vp, pvp, pvp, dp], "mgic_amr_precond": [vp, pvp, pvp],
        "mgic_amr_outer_solve": [vp, pvp, pvp, ip, ip, dp, C.c_int],
        "mgic_op_amr_operator_nf": [vp, vp, vp, vp, i3, C.c_int], "mgic_op_amr_residual_nf": [vp, vp, vp, vp, i3, vp, C.c_int],
        "mgic_op_destroy": [vp], "mgic_op_set_coefs": [vp, vp, vp, C.c_double, C.c_double],
        "mgic_op_set_alpha_beta": [vp, C.c_double, C.c_double], "mgic_op_reset_lambda": [vp],
        "mgic_op_compute_lambda": [vp], "mgic_op_get_lambda": [vp, pvp],
        "mgic_op_dims": [vp, i3, ip, ip, dp],
        "mgic_field_create": [vp, pvp], "mgic_field_destroy": [vp], "mgic_field_sync": [vp],
        "mgic_field_upload": [vp, nd], "mgic_field_download": [vp, nd],
        "mgic_field_upload_async": [vp, vp], "mgic_field_download_async": [vp, vp],
        "mgic_field_upload_fab": [vp, nd, i3, i3, i3, i3], "mgic_field_download_fab": [vp, nd, i3, i3, i3, i3],
        "mgic_field_devptr": [vp, pvp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)],
        "mgic_op_relax": [vp, vp, vp, C.c_int], "mgic_op_gsrb_color": [vp, vp, vp, C.c_int],
        "mgic_op_level_jacobi": [vp, vp, vp], "mgic_op_residual": [vp, vp, vp, vp, C.c_int],
        "mgic_op_apply": [vp, vp, vp, C.c_int], "mgic_op_apply_no_boundary": [vp, vp, vp],
        "mgic_op_restrict_residual": [vp, vp, vp, vp], "mgic_op_prolong_increment": [vp, vp, vp],
        "mgic_op_precond": [vp, vp, vp], "mgic_op_norm": [vp, vp, C.c_int, dp], "mgic_op_dot": [vp, vp, vp, dp],
        "mgic_op_incr": [vp, vp, vp, C.c_double], "mgic_op_axby": [vp, vp, vp, vp, C.c_double, C.c_double],
        "mgic_op_scale": [vp, vp, C.c_double], "mgic_op_assign": [vp, vp, vp], "mgic_op_set_to_zero": [vp, vp],
        "mgic_op_set_val": [vp, vp, C.c_double], "mgic_op_set_smoother": [vp, C.c_int],
        "mgic_mg_create": [vp, C.POINTER(MgicParams), vp, vp, pvp],
        "mgic_mg_create_ex": [vp, C.POINTER(MgicParams), vp, vp, C.c_int, pvp],
        "mgic_mg_destroy": [vp], "mgic_mg_op": [vp, C.c_int, pvp], "mgic_mg_scratch": [vp, C.c_int, pvp, pvp],
        "mgic_mg_refresh_coefs": [vp], "mgic_mg_vcycle": [vp, vp, vp], "mgic_mg_vcycle_from_zero": [vp, vp, vp], "mgic_mg_bottom_solve": [vp, vp, vp, ip],
        "mgic_mg_set_smoother": [vp, C.c_int],
        "mgic_mg_outer_solve": [vp, vp, vp, ip, ip, dp, C.c_int],
        "mgic_vars_create": [vp, C.POINTER(MgicParams), C.c_int, C.c_int, pvp], "mgic_vars_destroy": [vp],
        "mgic_vars_download": [vp, C.c_int, nd], "mgic_vars_download_ghosted": [vp, C.c_int, nd],
        "mgic_set_initial_conditions": [vp, vp], "mgic_set_a_coef": [vp, vp, C.c_double], "mgic_set_b_coef": [vp, vp],
        "mgic_set_rhs": [vp, vp, C.c_double], "mgic_set_rhs_and_a_coef": [vp, vp, vp, C.c_double],
        "mgic_update_psi0": [vp, vp, vp, dp],
        "mgic_nl_solve": [vp, C.POINTER(MgicParams), dp, C.c_int, ip, C.c_void_p],
        "mgic_hier_create": [vp, C.POINTER(MgicParams), C.c_int, ip, ip, ip, pvp], "mgic_hier_destroy": [vp],
        "mgic_hier_node_info": [vp, C.c_int, ip, i3, i3, C.POINTER(C.c_longlong)],
        "mgic_hier_get_mask": [vp, C.c_int, np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")],
        "mgic_hier_set_initial_conditions": [vp], "mgic_hier_nl_iteration": [vp, dp, ip, ip],
        "mgic_hier_nl_solve": [vp, dp, C.c_int, ip], "mgic_hier_write_checkpoint": [vp, C.c_char_p, C.c_double], "mgic_hier_download": [vp, C.c_int, C.c_int, nd],
        "mgic_hier_set_solver_params": [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int], "mgic_hier_set_sources": [vp, C.c_double],
        "mgic_hier_define_solver": [vp], "mgic_hier_solve": [vp, ip, ip], "mgic_hier_update_psi": [vp], "mgic_hier_dpsi_norm": [vp, dp],
        "mgic_hier_release_solver": [vp],
        "mgic_vars_create_patch": [vp, C.POINTER(MgicParams), vp, pvp],
        "mgic_grids_generate": [vp, C.POINTER(MgicParams), C.c_double, C.c_double, pvp],
        "mgic_grids_regrid": [C.POINTER(MgicParams), C.c_double, C.c_int, ip, ip, ip, ip, pvp],
        "mgic_grids_destroy": [vp], "mgic_grids_get_boxes": [vp, C.c_int, ip, ip, ip],
        "mgic_grids_level_stats": [vp, C.c_int, dp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)],
        "mgic_hier_create_from_grids": [vp, C.POINTER(MgicParams), vp, pvp],
        "mgic_update_psi0_patch": [vp, vp, vp, vp, i3],
    }
    for name, argtypes in sig.items():
        f = getattr(L, name)
        f.argtypes = argtypes
        f.restype = C.c_int
    L.mgic_ctx_stream.argtypes = [vp]
    L.mgic_ctx_stream.restype = vp
    L.mgic_grids_levels.argtypes = [vp]
    L.mgic_grids_levels.restype = C.c_int
    L.mgic_grids_num_boxes.argtypes = [vp, C.c_int]
    L.mgic_grids_num_boxes.restype = C.c_int
    L.mgic_op_valid_cells.argtypes = [vp]
    L.mgic_op_valid_cells.restype = C.c_longlong
    L.mgic_ctx_launch_count.argtypes = [vp]
    L.mgic_ctx_launch_count.restype = C.c_longlong
    for name in ("mgic_mg_depths", "mgic_mg_last_bottom_iterations", "mgic_mg_b_is_one", "mgic_amr_levels", "mgic_amr_nodes", "mgic_hier_nodes"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = C.c_int
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise MgicError(f"mgic error {rc}: {lib().mgic_last_error().decode()}")
