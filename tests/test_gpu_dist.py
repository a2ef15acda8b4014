"""GPU suite: the multi-GPU join (radix partition -> grouped NCCL send/recv -> local join) against the oracle.
Needs >= 2 GPUs on the box; skipped otherwise (the single-GPU driver run)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        return subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.count("GPU ")
    except Exception:
        return 0


@pytest.mark.skipif(_gpu_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_join_matches_oracle():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29517",
           os.path.join(ROOT, "tests", "dist_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "DIST_CHECK_OK world=2" in out.stdout
    assert "DIST_AGG_OK world=2" in out.stdout
