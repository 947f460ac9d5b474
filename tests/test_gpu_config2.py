"""BASELINE config 2 at full size: n=65536 d=32 k=16 double, fixed seed, GPU vs the reference
algorithm on the CPU (compare_results' protocol, compare_results.c:123-171): every graph entry,
bucket-table entry, par_max, and every bit of bases / row_means / distances must agree."""
import time

import numpy as np
import pytest

from conftest import same_bits

pytestmark = pytest.mark.gpu


def test_config2_full_size_parity(oracle_mod):
    from approximatenn_b200.api import gpu_backend
    n, d, k, tries = 65536, 32, 16, 8
    rng = np.random.default_rng(2)
    pts = rng.standard_normal((n, d))
    gpu, orc = gpu_backend(np.float64), oracle_mod.restatement(np.float64)
    t0 = time.time()
    got = gpu.precomp(pts, k, tries, want_save=True, seed=1001)
    t1 = time.time()
    want = orc.precomp(pts, k, tries, want_save=True, seed=1001)
    t2 = time.time()
    print(f"config 2: GPU call {t1 - t0:.3f} s (first call, incl. init), CPU restatement {t2 - t1:.1f} s")
    assert np.array_equal(got.ids, want.ids)
    assert same_bits(got.dists, want.dists)
    assert np.array_equal(got.save.par_maxes, want.save.par_maxes)
    assert same_bits(got.save.bases, want.save.bases) and same_bits(got.save.row_means, want.save.row_means)
    for t in range(tries):
        assert np.array_equal(got.save.which_par(t), want.save.which_par(t))
    y = rng.standard_normal((2048, d))
    qa, qb = gpu.query(got.save, pts, y), orc.query(want.save, pts, y)
    assert np.array_equal(qa.ids, qb.ids) and same_bits(qa.dists, qb.dists)
    got.save.free(); want.save.free()
