import sys, numpy as np
sys.path.insert(0, "/root/repo")
from approximatenn_b200.api import gpu_backend
for dtype, n, d, k, T in [(np.float32, 3000, 64, 16, 4), (np.float32, 2000, 20, 10, 3), (np.float64, 1500, 32, 16, 2), (np.float32, 1200, 16, 3, 2)]:
    rng = np.random.default_rng(1)
    pts = rng.standard_normal((n, d)).astype(dtype)
    g = gpu_backend(dtype)
    r = g.precomp(pts, k, T, want_save=True, seed=5)
    q = g.query(r.save, pts, rng.standard_normal((100, d)).astype(dtype))
    r.save.free()
    print(np.dtype(dtype).name, n, d, k, T, "ok", int(r.ids.sum() % 1000), int(q.ids.sum() % 1000))
for dtype in (np.float32, np.float64):
    gpu_backend(dtype).lib.gpu_cleanup()
