"""Parity of the CUDA path (through the C-ABI: precomp_gpu) with the reference.

Bar: bit-exact neighbour ids, squared distances and save_t fields against
  * the golden vectors produced by the reference itself (tests/golden), and
  * the CPU restatement on fresh seeded inputs (oracle/ann_oracle.c, pinned to the reference).
"""
import numpy as np
import pytest

from conftest import assert_matches_golden, golden_names, load_golden, same_bits

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from approximatenn_b200.api import gpu_backend
    return {np.dtype(np.float32): gpu_backend(np.float32), np.dtype(np.float64): gpu_backend(np.float64)}


@pytest.mark.parametrize("name", golden_names())
def test_precomp_gpu_reproduces_reference_golden(gpu, name):
    g = load_golden(name)
    res = gpu[g["dtype"]].precomp(g["points"], g["k"], g["tries"], *g["rot"], want_save=True, seed=g["seed"])
    assert_matches_golden(g, res)
    res.save.free()


LIVE = [
    (np.float32, 8192, 64, 16, 8, (6, 1, 1, 1), 201),
    (np.float64, 8192, 32, 16, 8, (6, 1, 1, 1), 202),
    (np.float32, 5000, 16, 10, 10, (6, 1, 1, 1), 203),
    (np.float32, 3001, 80, 10, 10, (6, 1, 1, 1), 204),     # reference defaults' d, odd n
    (np.float64, 2000, 128, 20, 4, (2, 8, 2, 2), 205),
    (np.float32, 6000, 32, 32, 4, (6, 1, 1, 1), 206),
    (np.float32, 4000, 48, 40, 3, (3, 2, 1, 1), 207),      # k > 32: two list registers per lane
    (np.float32, 2500, 9, 10, 10, (1, 4, 1, 2), 208),
    (np.float32, 20000, 16, 10, 10, (6, 1, 1, 1), 209),    # has exact ties; reference keeps an id twice
    (np.float32, 50000, 64, 16, 8, (6, 1, 1, 1), 210),
]


@pytest.mark.parametrize("dtype,n,d,k,tries,rot,seed", LIVE)
def test_precomp_gpu_equals_oracle_live(gpu, oracle_mod, dtype, n, d, k, tries, rot, seed):
    rng = np.random.default_rng(seed)
    pts = rng.standard_normal((n, d)).astype(dtype)
    want = oracle_mod.restatement(dtype).precomp(pts, k, tries, *rot, want_save=True, seed=seed)
    got = gpu[np.dtype(dtype)].precomp(pts, k, tries, *rot, want_save=True, seed=seed)
    # bit-exact, INCLUDING rows with exact distance ties (the literal-network redo)
    assert np.array_equal(got.ids, want.ids)
    assert same_bits(got.dists, want.dists)
    a, b = got.save, want.save
    assert np.array_equal(a.par_maxes, b.par_maxes)
    assert same_bits(a.row_means, b.row_means) and same_bits(a.bases, b.bases)
    assert np.array_equal(a.graph, got.ids)
    for t in range(tries):
        assert np.array_equal(a.which_par(t), b.which_par(t))
    a.free(); b.free()


def test_precomp_gpu_without_save_or_dists(gpu):
    g = load_golden("cfg3shape_f32")
    res = gpu[g["dtype"]].precomp(g["points"], g["k"], g["tries"], *g["rot"], want_save=False,
                                  want_dists=False, seed=g["seed"])
    assert res.dists is None and res.save is None
    assert np.array_equal(res.ids, g["ids"].astype(np.uint64))
