"""Pins the restatement against the reference compiled here (oracle/_ref), live, on
seeds that are NOT in the golden set.  Skipped where oracle/_ref does not exist."""
import numpy as np
import pytest

from conftest import same_bits

CASES = [
    (np.float32, 777, 24, 8, 5, (6, 1, 1, 1), 101),
    (np.float64, 1200, 40, 12, 6, (2, 5, 2, 3), 102),
    (np.float32, 2500, 9, 10, 10, (1, 4, 1, 2), 103),
    (np.float64, 513, 33, 3, 2, (6, 1, 1, 1), 104),      # k(k+1) < 16 and k*tries < 16
]


@pytest.mark.parametrize("dtype,n,d,k,tries,rot,seed", CASES)
def test_restatement_equals_reference_live(oracle_mod, dtype, n, d, k, tries, rot, seed):
    if not oracle_mod.reference_available():
        pytest.skip("oracle/_ref not built (no reference checkout on this machine)")
    rng = np.random.default_rng(seed)
    pts = rng.standard_normal((n, d)).astype(dtype)
    y = rng.standard_normal((41, d)).astype(dtype)
    ref, orc = oracle_mod.reference(dtype), oracle_mod.restatement(dtype)
    a = ref.precomp(pts, k, tries, *rot, want_save=True, seed=seed)
    b = orc.precomp(pts, k, tries, *rot, want_save=True, seed=seed)
    assert np.array_equal(a.ids, b.ids) and same_bits(a.dists, b.dists)
    assert np.array_equal(a.save.par_maxes, b.save.par_maxes)
    assert same_bits(a.save.row_means, b.save.row_means) and same_bits(a.save.bases, b.save.bases)
    for t in range(tries):
        assert np.array_equal(a.save.which_par(t), b.save.which_par(t))
    qa, qb = ref.query(a.save, pts, y), orc.query(b.save, pts, y)
    assert np.array_equal(qa.ids, qb.ids) and same_bits(qa.dists, qb.dists)
    # y == points (same pointer) switches self-exclusion on in the reference (compute.cl:145)
    qa, qb = ref.query(a.save, pts, pts), orc.query(b.save, pts, pts)
    assert np.array_equal(qa.ids, qb.ids) and same_bits(qa.dists, qb.dists)
    a.save.free(); b.save.free()
