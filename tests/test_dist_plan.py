"""world_size-2 CPU (gloo) replay of the multi-GPU plan: shard tries, exchange lists by row
slice, merge, all-gather merged ids, supercharge slices == the single-process reference."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharding_plan_replayed_on_cpu_matches_reference(oracle_mod):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29631",
           os.path.join(ROOT, "tests", "_dist_cpu_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    sys.stdout.write(out.stdout[-3000:])
    sys.stderr.write(out.stderr[-3000:])
    assert out.returncode == 0
    assert out.stdout.count(" OK") == 6 and "MISMATCH" not in out.stdout


def test_partition_functions():
    from approximatenn_b200 import dist as adist
    from approximatenn_b200.api import gpu_backend
    lib = gpu_backend(np.float32).lib
    adist.declare(lib)
    for n, world in [(1_000_000, 8), (5003, 2), (31, 4), (4096, 3)]:
        cuts = [adist.row_slice(lib, n, r, world) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        for (a, b), (c, d) in zip(cuts, cuts[1:]):
            assert b == c and a <= b and (a % 32 == 0 or a == n)
    owners = [adist.try_owner(lib, t, 3) for t in range(8)]
    assert owners == [0, 1, 2, 0, 1, 2, 0, 1]
    # k*T = 100 -> prefix 64: tries 0..5 whole, try 6 contributes 4, the rest nothing
    assert [adist.admitted(lib, 10, 10, t) for t in range(10)] == [10] * 6 + [4, 0, 0, 0]
    assert [adist.admitted(lib, 16, 8, t) for t in range(8)] == [16] * 8
    assert [adist.admitted(lib, 5, 3, t) for t in range(3)] == [5, 5, 5]       # < 16 slots: all
