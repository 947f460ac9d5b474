"""Child process of tests/test_gpu_single_process.py: ANN_B200_GPUS is set in its environment,
so the plain precomp()/query() calls below use several devices from ONE process."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from approximatenn_b200.api import gpu_backend  # noqa: E402


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


def main():
    ok = True
    cases = [(np.float32, 8192, 64, 16, 8, 501), (np.float32, 50001, 32, 10, 10, 502), (np.float64, 4099, 16, 10, 5, 503),
             (np.float32, 224, 16, 4, 4, 504)]
    for dtype, n, d, k, tries, seed in cases:
        rng = np.random.default_rng(seed)
        pts = rng.standard_normal((n, d)).astype(dtype)
        gpu, orc = gpu_backend(dtype), oracle.restatement(dtype)
        want = orc.precomp(pts, k, tries, seed=seed)
        for rep in range(2):                                   # second call: cached arenas, same answer
            got = gpu.precomp(pts, k, tries, seed=seed)
            good = np.array_equal(got.ids, want.ids) and same_bits(got.dists, want.dists)
            print(f"{np.dtype(dtype).name} n={n} d={d} k={k} T={tries} call {rep} {'OK' if good else 'MISMATCH'}", flush=True)
            ok &= good
        nod = gpu.precomp(pts, k, tries, want_dists=False, seed=seed)
        ok &= np.array_equal(nod.ids, want.ids)
    # a slice large enough for the staged (multi-threaded) upload inside every worker; sampled rows
    from approximatenn_b200.api import srandom
    n, d, k, tries = 300_000, 64, 16, 2
    rng = np.random.default_rng(505)
    pts = rng.standard_normal((n, d), dtype=np.float32)
    sample = np.sort(rng.choice(n, size=48, replace=False))
    got = gpu_backend(np.float32).precomp(pts, k, tries, seed=505)
    srandom(505)
    want_ids, want_d, _ = oracle.sampled_rows(oracle.restatement(np.float32), pts, k, tries, sample)
    good = np.array_equal(got.ids[sample], want_ids) and same_bits(got.dists[sample], want_d)
    print(f"float32 n={n} staged upload, sampled rows {'OK' if good else 'MISMATCH'}", flush=True)
    ok &= good
    # save_t and query run on device 0
    rng = np.random.default_rng(9)
    pts = rng.standard_normal((6000, 32)).astype(np.float32)
    y = rng.standard_normal((300, 32)).astype(np.float32)
    gpu, orc = gpu_backend(np.float32), oracle.restatement(np.float32)
    want = orc.precomp(pts, 16, 5, want_save=True, seed=77)
    got = gpu.precomp(pts, 16, 5, want_save=True, seed=77)
    good = np.array_equal(got.ids, want.ids) and np.array_equal(got.save.graph, want.save.graph)
    for t in range(5):
        good &= np.array_equal(got.save.which_par(t), want.save.which_par(t))
    qa, qb = gpu.query(got.save, pts, y), orc.query(want.save, pts, y)
    good &= np.array_equal(qa.ids, qb.ids) and same_bits(qa.dists, qb.dists)
    print(f"save + query {'OK' if good else 'MISMATCH'}", flush=True)
    ok &= good
    got.save.free(); want.save.free()
    gpu.lib.gpu_cleanup()
    again = gpu.precomp(pts, 16, 5, seed=77)                  # workers restart after gpu_cleanup
    ok &= np.array_equal(again.ids, want.ids)
    print("after gpu_cleanup", "OK" if ok else "MISMATCH", flush=True)
    gpu.lib.gpu_cleanup()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
