"""Phase times of query_gpu on a resident index (ANN_B200_HOSTPROF=1 synchronises after each phase)."""
import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
os.environ["ANN_B200_HOSTPROF"] = "1"
from bench import CONFIGS, synth_points
from approximatenn_b200.api import gpu_backend
n, d, k, tries, dtype = CONFIGS["cfg3"]
g = gpu_backend(dtype)
pts = synth_points(n, d, dtype)
y = np.random.default_rng(11).standard_normal((65536, d), dtype=np.float32)
r = g.precomp(pts, k, tries, want_save=True, seed=1001)
for i in range(3):
    print("---- query", i, file=sys.stderr, flush=True)
    g.query(r.save, pts, y)
