"""Sharded precomp_gpu (tries split across ranks, NCCL list exchange) against the oracle at
2, 4 and 8 ranks.  Needs that many GPUs on the box; smaller boxes skip the larger worlds (the
partition logic itself is covered on CPU by tests/test_dist_plan.py).  Cases (see the worker):
uneven try ownership, ranks without any try, n not divisible by 32*R, an empty last row slice,
the prefix corner (k*T = 100), save_t in sharded mode, and BASELINE config 3's size checked on
sampled rows."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_precomp_matches_oracle(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29610 + world),
           os.path.join(ROOT, "tests", "_dist_gpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    print(out.stdout[-6000:])
    print(out.stderr[-4000:])
    assert out.returncode == 0
    assert "MISMATCH" not in out.stdout
    assert out.stdout.count(" OK") >= world * 8
