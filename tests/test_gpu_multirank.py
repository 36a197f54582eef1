"""Multi-GPU parity (needs >= 2 GPUs, skipped otherwise): tools/check_multigpu.py under torchrun -- z-slab decomposed
sweeps / residual bitwise equal to one GPU, V-cycles (halo overlap, level agglomeration, graph replay) equal to 1e-10."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("nproc", [2, 4])
def test_slab_decomposition_matches_single_gpu(nproc):
    import torch
    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    port = 29600 + nproc
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "check_multigpu.py"), "128"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert "MULTIGPU CHECK PASSED" in r.stdout
