"""ANN_B200_GPUS=N: the whole node behind the reference's own single-process call (csrc/ann_multi.c).
One worker thread per device runs the sharded path; the caller gets ONE malloc()ed result, bit
for bit the single-GPU / oracle result.  Also runs the reference's unmodified time_results
binary that way.  Needs >= 2 GPUs (skipped on the single-GPU boxes)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("gpus", [2, 8])
def test_single_process_multi_gpu_matches_oracle(gpus):
    import torch
    if torch.cuda.device_count() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    env = dict(os.environ, ANN_B200_GPUS=str(gpus))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_single_process_multi_worker.py")],
                         capture_output=True, text=True, timeout=900, env=env)
    print(out.stdout[-4000:])
    print(out.stderr[-3000:])
    assert out.returncode == 0
    assert "MISMATCH" not in out.stdout and out.stdout.count("OK") >= 10


def test_reference_time_results_uses_the_node():
    import torch
    n_gpus = torch.cuda.device_count()
    exe = os.path.join(ROOT, "oracle", "_ref", "bin", "time_results_f32")
    if n_gpus < 2 or not os.path.exists(exe):
        pytest.skip("needs >= 2 GPUs and the reference's programs (oracle/_ref/bin)")
    env = dict(os.environ, ANN_B200_GPUS=str(min(n_gpus, 8)))
    out = subprocess.run([exe, "-n", "1000000", "-d", "64", "-k", "16", "-t", "8", "-o", "3"],
                         capture_output=True, text=True, timeout=600, env=env)
    print(out.stdout, out.stderr[-2000:])
    assert out.returncode == 0 and "Average time for comp" in out.stdout
