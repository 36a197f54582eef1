"""The C++ host mirror (mg_ic_code_b200/host): VariableCoeffPoissonOperator / Factory / MultilevelLinearOp / BiCGStabSolver
driven by a reference-format params.txt, against the oracle's nonlinear loop."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PARAMS = os.path.join(ROOT, "tests", "data", "params_32.txt")


def exe():
    from mg_ic_code_b200 import build as b
    from mg_ic_code_b200.host import build_host
    b.build()
    return build_host.build()


def test_driver_builds_and_fails_loudly_without_gpu():
    import torch
    e = exe()
    assert os.path.exists(e)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([e, PARAMS, "--json"], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr and "MayDay::Error" in r.stderr


def test_driver_rejects_bad_input(tmp_path):
    e = exe()
    bad = tmp_path / "bad.txt"
    bad.write_text(open(PARAMS).read().replace("harmonic", "geometric"))
    r = subprocess.run([e, str(bad)], capture_output=True, text=True)
    assert r.returncode != 0 and "bad coefficient_average_type in input" in r.stderr
    r = subprocess.run([e, str(tmp_path / "missing.txt")], capture_output=True, text=True)
    assert r.returncode != 0 and "cannot open" in r.stderr


REFERENCE_MAIN = os.path.join(os.environ.get("MGIC_REFERENCE", "/root/reference"), "Main_PoissonSolver.cpp")
REFMAIN_EXE = os.path.join(ROOT, "mg_ic_code_b200", "lib", "Main_PoissonSolver_b200")


@pytest.mark.skipif(not os.path.exists(REFERENCE_MAIN), reason="needs the reference's Main_PoissonSolver.cpp")
def test_reference_main_compiles_unmodified_against_the_host_layer(tmp_path):
    """The drop-in claim at its strongest: the reference's own Main_PoissonSolver.cpp -- main(), parameter handling, the
    nonlinear loop -- compiles UNMODIFIED, where it lies, against mg_ic_code_b200/host (+ host/dropin: headers with Chombo's and
    the reference's names) and links against the C ABI library.  Without a GPU it behaves like the reference up to the first
    device call and then stops loudly."""
    import torch
    exe()
    assert os.path.exists(REFMAIN_EXE)
    r = subprocess.run([REFMAIN_EXE], capture_output=True, text=True)
    assert r.returncode == 0 and "usage" in r.stderr and "<input_file_name>" in r.stderr        # Main_PoissonSolver.cpp:266-269
    ref_params = os.path.join(os.path.dirname(REFERENCE_MAIN), "params.txt")                     # max_level = 6: set_grids tags on the device
    if not torch.cuda.is_available():
        r = subprocess.run([REFMAIN_EXE, ref_params], capture_output=True, text=True)
        assert r.returncode != 0 and "no CPU fallback" in r.stderr
    bad = tmp_path / "bad.txt"
    bad.write_text(open(PARAMS).read().replace("harmonic", "geometric"))
    r = subprocess.run([REFMAIN_EXE, str(bad)], capture_output=True, text=True)
    assert r.returncode != 0 and "bad coefficient_average_type in input" in r.stderr
    if not torch.cuda.is_available():
        r = subprocess.run([REFMAIN_EXE, PARAMS], capture_output=True, text=True)
        assert r.returncode != 0 and "no CPU fallback" in r.stderr and "MayDay::Error" in r.stderr


@pytest.mark.gpu
def test_reference_main_drives_the_b200_path(tmp_path):
    """the reference's unmodified main() and nonlinear loop on the B200 path: same NL history and psi as the oracle"""
    from oracle import Oracle
    if not os.path.exists(REFMAIN_EXE):
        pytest.skip("Main_PoissonSolver_b200 was not built (the reference's Main_PoissonSolver.cpp is absent)")
    o = Oracle(N=(32, 32, 32), max_grid_size=16, numMGsmooth=4, numMGIterations=2)
    o.set_initial_conditions()
    nl = o.nl_solve()
    psi_o = o.get("MGVAR0", comp=0)
    dump = tmp_path / "psi.bin"
    r = subprocess.run([REFMAIN_EXE, PARAMS], capture_output=True, text=True, timeout=120, env=dict(os.environ, MGIC_DUMP_PSI=str(dump)))
    assert r.returncode == 0, r.stderr[-2000:]
    norms = [float(l.split(" is ")[1]) for l in r.stdout.splitlines() if l.startswith("The norm of dpsi after step")]
    assert len(norms) == len(nl) and np.allclose(norms[:3], nl[:3], rtol=1e-5)      # pout() prints 6 significant digits
    psi = np.fromfile(dump).reshape(32, 32, 32)
    assert np.abs(psi - psi_o).max() / np.abs(psi_o).max() < 1e-10


@pytest.mark.gpu
def test_reference_main_runs_an_amr_hierarchy(tmp_path):
    """the reference's unmodified main() with max_level = 2: its set_grids call builds the hierarchy (tagging + BRMeshRefine in
    the library), its nonlinear loop drives the hierarchy solve call by call (HierarchySession.H), its output_final_data writes
    the GRChombo checkpoint -- same dpsi norms and psi as the library's own loop (mgic_hier_nl_solve), which the GPU tests hold
    to the oracle twin"""
    import mg_ic_code_b200 as m
    from mg_ic_code_b200 import checkpoint
    from oracle import default_params
    if not os.path.exists(REFMAIN_EXE):
        pytest.skip("Main_PoissonSolver_b200 was not built (the reference's Main_PoissonSolver.cpp is absent)")
    dump, chk = tmp_path / "psi.bin", tmp_path / "vcPoissonFinal.3d.mgic"
    r = subprocess.run([REFMAIN_EXE, PARAMS, "max_level=2", "numMGsmooth=2", "max_NL_iterations=3"], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, MGIC_DUMP_PSI=str(dump), MGIC_CHECKPOINT=str(chk)), cwd=str(tmp_path))
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    norms = [float(l.split(" is ")[1]) for l in r.stdout.splitlines() if l.startswith("The norm of dpsi after step")]
    ctx = m.Context(0)
    P = m.make_params(default_params(N=(32, 32, 32), L=100.0, max_grid_size=16, block_factor=8, max_level=2, numMGsmooth=2,
                                     numMGIterations=2, max_iterations=100, max_NL_iterations=3, tolerance=1e-10))
    g = m.Grids.generate(ctx, P, 0.1, 0.5)
    H = m.Hierarchy.from_grids(ctx, P, g)
    want = H.nl_solve()
    assert len(norms) == len(want) == 3 and np.allclose(norms, want, rtol=1e-5)       # pout() prints 6 significant digits
    psi = np.fromfile(dump).reshape(32, 32, 32)
    assert np.array_equal(psi, H.download(0, "psi"))                                   # the same library calls in the same order
    hdr, levels = checkpoint.read(chk)
    assert hdr["root"]["ints"]["num_levels"] == g.levels == 3
    assert [len(lv["boxes"]) for lv in hdr["levels"]] == [len(g.boxes(l)) for l in range(3)]
    assert all("level " + str(l) in r.stdout for l in range(3))
    H.close(); g.close(); ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("route", ["device", "host"])
def test_driver_matches_oracle_nl_loop(tmp_path, route):
    from oracle import Oracle
    o = Oracle(N=(32, 32, 32), max_grid_size=16, numMGsmooth=4, numMGIterations=2)
    o.set_initial_conditions()
    nl = o.nl_solve()
    psi_o = o.get("MGVAR0", comp=0)
    dump = tmp_path / "psi.bin"
    args = [exe(), PARAMS, "--json", "--dump-psi", str(dump)] + (["--host-vcycle"] if route == "host" else [])
    r = subprocess.run(args, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["nl_iterations"] == len(nl)
    assert np.allclose(d["dpsi_norms"][:3], nl[:3], rtol=1e-7)
    assert d["exit_status"] == 0 and d["kernel_launches"] > 0
    psi = np.fromfile(dump).reshape(32, 32, 32)
    assert np.abs(psi - psi_o).max() / np.abs(psi_o).max() < 1e-10
    assert "The norm of dpsi after step 1 is" in r.stdout


@pytest.mark.gpu
def test_driver_runs_the_reference_params_hierarchy(tmp_path):
    """the repo's own driver on a hierarchy: max_level = 3, V(4,4), set_grids + the nonlinear loop + the checkpoint"""
    from mg_ic_code_b200 import checkpoint
    chk = tmp_path / "final.mgic"
    r = subprocess.run([exe(), PARAMS, "--json", "--checkpoint", str(chk), "max_level=3", "max_NL_iterations=4"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert d["levels"] == 4 and d["nodes"] >= 4 and d["exit_status"] == 0
    n = d["dpsi_norms"]
    assert n[0] > 1e-2 and n[1] < 1e-2 * n[0] and min(n) < 1e-5 * n[0]
    hdr, _ = checkpoint.read(chk, load_data=False)
    assert hdr["root"]["ints"]["num_levels"] == 4 and len(hdr["levels"]) == 4


@pytest.mark.gpu
def test_driver_key_value_overrides(tmp_path):
    r = subprocess.run([exe(), PARAMS, "--json", "max_NL_iterations=1", "numMGsmooth=2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert d["nl_iterations"] == 1
