"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/*.h declares, and
refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import mg_ic_code_b200 as m

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in ("mgic.h", "mgic_chf.h", "mgic_comm.h"):
        p = os.path.join(ROOT, "include", h)
        if not os.path.exists(p):
            continue
        src = open(p).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = "\n".join(l for l in src.splitlines() if not l.strip().startswith("#"))
        src = src.replace("\\\n", " ")
        for mm in re.finditer(r"\b([a-z_0-9]+)\s*\(", src):
            n = mm.group(1)
            if n.startswith("mgic_") or n.endswith("3d_") or n in ("getlaplacianpsif_", "getrhogradphif_", "prolong_"):
                names.add(n)
    return sorted(names)


def test_library_exports_every_declared_symbol():
    L = m.lib()
    syms = declared_symbols()
    assert len(syms) > 60
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_chf_symbol_names_are_fortran_mangled():
    L = m.lib()
    for s in ("gsrbhelmholtzvc3d_", "vccomputeop3d_", "vccomputeres3d_", "restrictresvc3d_", "getlaplacianpsif_",
              "getrhogradphif_", "prolong_"):
        assert hasattr(L, s)


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(m.MgicError, match="no CPU fallback"):
        m.Context(0)


def test_bad_arguments_are_refused_before_any_device_work():
    """every entry point checks its arguments first: NULL handles come back as MGIC_ERR_ARG with a message, with or without a GPU"""
    L = m.lib()
    vp = C.c_void_p
    out = vp()
    d = C.c_double()
    calls = [
        lambda: L.mgic_amr_create_levels(None, 0, None, None, C.byref(out)),
        lambda: L.mgic_amr_create(None, 0, None, C.byref(out)),
        lambda: L.mgic_amr_vcycle(None, None, None),
        lambda: L.mgic_amr_norm(None, None, 0, C.byref(d)),
        lambda: L.mgic_amr_outer_solve(None, None, None, None, None, None, 0),
        lambda: L.mgic_op_norm(None, None, 0, C.byref(d)),
        lambda: L.mgic_op_relax(None, None, None, 1),
        lambda: L.mgic_mg_vcycle(None, None, None),
    ]
    for call in calls:
        rc = call()
        assert rc != 0 and L.mgic_last_error()
        with pytest.raises(m.MgicError):
            m._capi.check(rc)
    assert L.mgic_amr_levels(None) == 0 and L.mgic_amr_nodes(None) == 0 and L.mgic_amr_destroy(None) == 0


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "mg_ic_code_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".H", ".cpp", ".hpp")):
                src = open(os.path.join(dp, f)).read()
                for pat in (r"^\s*(from|import)\s+oracle", r"import_module\(.oracle", r"libmgic_oracle", r"#include\s+.*oracle"):
                    assert not re.search(pat, src, flags=re.M), f"{f} reaches into oracle/ ({pat})"


def test_params_reader(tmp_path):
    txt = """# sample in the reference's params.txt format
alpha = 1.0
beta  = -1.0
L = 100.0
N = 64 32 16
max_level    = 0
block_factor = 8
max_grid_size = 16   # max box size
numMGsmooth = 2 # smooths
numMGIterations = 3
tolerance  = 1.0e-9
coefficient_average_type = harmonic
is_periodic = 0
bc_lo       = 0 1 0
bc_hi       = 1 0 0
bc_value = 0.5
bh1_offset = 12.5
"""
    p = tmp_path / "params.txt"
    p.write_text(txt)
    P = m.read_params(str(p), overrides=["numMGsmooth=4"])
    assert list(P.N) == [64, 32, 16] and P.numMGsmooth == 4 and P.numMGIterations == 3
    assert list(P.bc_lo) == [0, 1, 0] and list(P.bc_hi) == [1, 0, 0] and P.bc_value == 0.5
    assert P.coefficient_average_type == 1 and P.bh1_offset == 12.5 and P.tolerance == 1e-9
    p.write_text(txt.replace("harmonic", "geometric"))
    with pytest.raises(m.MgicError, match="bad coefficient_average_type"):
        m.read_params(str(p))


def test_device_arena_bookkeeping():
    """the sub-allocator that replaces one cudaMalloc per array (csrc/arena.h): random allocations and frees on the host alone --
    no overlap between live ranges, double frees refused, everything coalesces back into one free range"""
    import ctypes as C
    from mg_ic_code_b200._capi import lib
    L = lib()
    L.mgic_arena_selftest.argtypes = [C.c_uint, C.c_int]
    for seed in range(8):
        assert L.mgic_arena_selftest(seed, 20000) == 0
