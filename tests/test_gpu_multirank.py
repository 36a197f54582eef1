"""Multi-GPU parity (needs >= 2 GPUs, skipped otherwise), each a tool under torchrun:
  * tools/check_multigpu.py -- z-slab decomposed sweeps / residual bitwise equal to one GPU, V-cycles (halo overlap, level
    agglomeration, graph replay, ncclSend/ncclRecv halos, halo folded into the sweep) equal to 1e-10 / bitwise among themselves;
  * tools/check_multigpu_amr.py -- poissonSolve on an AMR hierarchy with the base level in z-slabs and the refined levels
    replicated: three nonlinear iterations, psi on every node against one GPU;
  * tools/check_multigpu_periodic.py -- is_periodic = 1 with the slabs on a ring."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def torchrun(nproc, tool, arg, port, passed):
    import torch
    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", tool), str(arg)],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert passed in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("nproc", [2, 4])
def test_slab_decomposition_matches_single_gpu(nproc):
    torchrun(nproc, "check_multigpu.py", 128, 29600 + nproc, "MULTIGPU CHECK PASSED")


@pytest.mark.gpu
@pytest.mark.parametrize("nproc", [2, 4])
def test_amr_hierarchy_over_ranks_matches_single_gpu(nproc):
    torchrun(nproc, "check_multigpu_amr.py", 128, 29610 + nproc, "MULTIGPU AMR CHECK PASSED")


@pytest.mark.gpu
@pytest.mark.parametrize("nproc", [2, 4])
def test_periodic_ring_matches_single_gpu(nproc):
    torchrun(nproc, "check_multigpu_periodic.py", 64 if nproc == 2 else 128, 29620 + nproc, "MULTIGPU PERIODIC CHECK PASSED")
