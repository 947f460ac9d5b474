"""Sharded precomp_gpu (tries split across ranks, NCCL list exchange) against the oracle.
Needs >= 2 GPUs on the box; on a single-GPU box it is skipped (the partition logic itself is
covered on CPU by tests/test_dist_plan.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2])
def test_sharded_precomp_matches_oracle(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29617",
           os.path.join(ROOT, "tests", "_dist_gpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    print(out.stdout[-4000:])
    print(out.stderr[-4000:])
    assert out.returncode == 0
    assert "MISMATCH" not in out.stdout
