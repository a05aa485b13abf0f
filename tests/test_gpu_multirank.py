"""Multi-GPU parity: a z-slab run over 2 ranks (NCCL halo exchange) must reproduce
the single-GPU fields bitwise. Needs two visible GPUs; skipped otherwise."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus() -> int:
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("case", ["3d_default"])
def test_two_slabs_reproduce_one_gpu_bitwise(case):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "tools", "multigpu_check.py"), case, "12", "6", "0", "6"]   # + one chunked pdgpu_step_host pass
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "bitwise=no" not in r.stdout
