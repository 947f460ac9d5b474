"""BASELINE configs 1 and 2 at their exact sizes with the reference programs' default 10 tries
(time_results.c:16-17): fixed seed, GPU vs the reference algorithm on the CPU (compare_results'
protocol, compare_results.c:123-171): every graph entry, bucket-table entry, par_max, and every
bit of bases / row_means / distances must agree.  With 10 tries the merged row has k*T = 160
(config 2) / 100 (config 1) slots, so the power-of-two prefix drops the last tries and the
prefix-corner rule is live at full size.  BASELINE config 3 (n = 10^6) is checked on sampled
rows (the oracle streams rows, a full run would take a quarter of an hour)."""
import time

import numpy as np
import pytest

from conftest import same_bits

pytestmark = pytest.mark.gpu


def test_config1_exact_parity(oracle_mod):
    from approximatenn_b200.api import gpu_backend
    n, d, k, tries = 16384, 16, 10, 10
    rng = np.random.default_rng(1)
    pts = rng.standard_normal((n, d)).astype(np.float32)
    gpu, orc = gpu_backend(np.float32), oracle_mod.restatement(np.float32)
    got = gpu.precomp(pts, k, tries, want_save=True, seed=1001)
    want = orc.precomp(pts, k, tries, want_save=True, seed=1001)
    assert np.array_equal(got.ids, want.ids)
    assert same_bits(got.dists, want.dists)
    assert np.array_equal(got.save.par_maxes, want.save.par_maxes)
    assert same_bits(got.save.bases, want.save.bases) and same_bits(got.save.row_means, want.save.row_means)
    for t in range(tries):
        assert np.array_equal(got.save.which_par(t), want.save.which_par(t))
    y = rng.standard_normal((1024, d)).astype(np.float32)
    qa, qb = gpu.query(got.save, pts, y), orc.query(want.save, pts, y)
    assert np.array_equal(qa.ids, qb.ids) and same_bits(qa.dists, qb.dists)
    got.save.free(); want.save.free()


def test_config3_sampled_rows_parity(oracle_mod):
    """n = 10^6, d = 64, k = 16, 8 tries: the exact final rows of 128 sampled points."""
    import ctypes
    from approximatenn_b200.api import gpu_backend, srandom, _libc, _view
    n, d, k, tries = 1_000_000, 64, 16, 8
    rng = np.random.default_rng(3)
    pts = rng.standard_normal((n, d), dtype=np.float32)
    sample = np.sort(rng.choice(n, size=128, replace=False))
    gpu, orc = gpu_backend(np.float32), oracle_mod.restatement(np.float32)
    dptr = ctypes.c_void_p()
    srandom(1001)
    ids = gpu.precomp_raw(n, k, d, pts.ctypes.data, tries, 6, 1, 1, 1, None, ctypes.byref(dptr))
    got_ids = _view(ids, (n, k), np.uint64)[sample].copy()
    got_d = _view(dptr, (n, k), np.float32)[sample].copy()
    _libc.free(ids); _libc.free(dptr)
    srandom(1001)
    want_ids, want_d, _ = oracle_mod.sampled_rows(orc, pts, k, tries, sample)
    assert np.array_equal(got_ids, want_ids)
    assert same_bits(got_d, want_d)


def test_config2_full_size_parity(oracle_mod):
    from approximatenn_b200.api import gpu_backend
    n, d, k, tries = 65536, 32, 16, 10
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
