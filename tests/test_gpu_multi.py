"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): torchrun with 2 ranks, NCCL."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_sharded_nccl_sweep_matches_single_gpu():
    import jwas_b200
    ngpu = jwas_b200.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29711", os.path.join(ROOT, "tools", "multigpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert "MULTIGPU_CHECK PASS" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
