import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import oracle
from approximatenn_b200.api import gpu_backend
from conftest import knn_diff
for dtype,n,d,k,tries,rot,seed in [(np.float32, 5000, 16, 10, 10, (6, 1, 1, 1), 203),(np.float32, 3001, 80, 10, 10, (6, 1, 1, 1), 204),(np.float64, 2000, 128, 20, 4, (2, 8, 2, 2), 205),(np.float32, 6000, 32, 32, 4, (6, 1, 1, 1), 206),(np.float32, 4000, 48, 40, 3, (3, 2, 1, 1), 207),(np.float32, 2500, 9, 10, 10, (1, 4, 1, 2), 208),(np.float32, 20000, 16, 10, 10, (6,1,1,1), 209),(np.float32, 50000, 64, 16, 8, (6,1,1,1), 210)]:
    rng = np.random.default_rng(seed)
    pts = rng.standard_normal((n, d)).astype(dtype)
    want = oracle.restatement(dtype).precomp(pts, k, tries, *rot, want_save=True, seed=seed)
    got = gpu_backend(dtype).precomp(pts, k, tries, *rot, want_save=True, seed=seed)
    dd, hard, soft = knn_diff(got.ids, got.dists, want.ids, want.dists)
    tabs = all(np.array_equal(got.save.which_par(t), want.save.which_par(t)) for t in range(tries))
    print(dtype.__name__, n, d, k, tries, "rows dist-diff", dd, "hard id diff", hard, "tie-only", soft, "tables", tabs,
          "means", np.array_equal(got.save.row_means.view(np.uint8), want.save.row_means.view(np.uint8)),
          "bases", np.array_equal(got.save.bases.view(np.uint8), want.save.bases.view(np.uint8)), flush=True)
    bad = np.flatnonzero((got.ids != want.ids).any(axis=1))[:3]
    for r in bad:
        print("  row", r, "\n   got ", got.ids[r], got.dists[r], "\n   want", want.ids[r], want.dists[r])
