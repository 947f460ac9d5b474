"""S3 screened path (tensor-core distance brackets + exact tree for the survivors) against the
tiled kernel and the oracle.

The screen may only ever drop a candidate that is strictly farther than the k-th best, so the
per-try lists — and with them every output bit — must not depend on whether it runs.  The data
sets below are chosen to stress its error bounds and its fallbacks: a large common offset
(centring), tight clusters (brackets wider than the distances: survivors overflow, buckets go
to the tiled kernel), duplicated points (exact ties), one far outlier (the fp16 scale) and
points almost on the mean (fp16 subnormals).
"""
import ctypes
import zlib

import numpy as np
import pytest

from conftest import same_bits

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu32():
    from approximatenn_b200.api import gpu_backend
    g = gpu_backend(np.float32)
    g.lib.annb_leaf_pairs.restype = ctypes.c_ulonglong
    g.lib.annb_leaf_exact_pairs.restype = ctypes.c_ulonglong
    g.lib.annb_leaf_overflow_buckets.restype = ctypes.c_ulonglong
    yield g
    g.lib.annb_leaf_screen_mode(1)


def _datasets():
    def gauss(rng, n, d):
        return rng.standard_normal((n, d))

    def offset(rng, n, d):
        return rng.standard_normal((n, d)) + 1000.0

    def clusters(rng, n, d):
        centres = rng.standard_normal((32, d)) * 50.0
        return centres[rng.integers(0, 32, n)] + rng.standard_normal((n, d)) * 1e-3

    def duplicates(rng, n, d):
        base = rng.standard_normal((n // 4, d))
        return base[rng.integers(0, n // 4, n)]

    def outlier(rng, n, d):
        x = rng.standard_normal((n, d))
        x[7] *= 3.0e4
        return x

    def near_mean(rng, n, d):
        x = rng.standard_normal((n, d))
        x[: n // 2] *= 1e-6
        return x

    return [("gauss", gauss), ("offset", offset), ("clusters", clusters), ("duplicates", duplicates),
            ("outlier", outlier), ("near_mean", near_mean)]


SHAPES = [(20000, 64, 16, 4), (12000, 32, 10, 5), (9000, 16, 5, 6)]


@pytest.mark.parametrize("name,make", _datasets())
@pytest.mark.parametrize("n,d,k,tries", SHAPES)
def test_screened_lists_equal_tiled_lists(gpu32, name, make, n, d, k, tries):
    rng = np.random.default_rng(zlib.crc32(f"{name}-{n}-{d}".encode()))
    pts = np.ascontiguousarray(make(rng, n, d), dtype=np.float32)
    out = {}
    for mode in (0, 1):
        gpu32.lib.annb_leaf_screen_mode(mode)
        gpu32.lib.annb_leaf_pairs(1)
        gpu32.lib.annb_leaf_exact_pairs(1)
        res = gpu32.precomp(pts, k, tries, 6, 1, 1, 1, want_save=False, seed=77)
        out[mode] = (res.ids, res.dists, int(gpu32.lib.annb_leaf_pairs(0)), int(gpu32.lib.annb_leaf_exact_pairs(0)))
    assert np.array_equal(out[0][0], out[1][0])
    assert same_bits(out[0][1], out[1][1])
    assert out[0][2] == out[1][2]                 # same algorithmic pair count either way
    assert out[0][3] == 0                         # the tiled kernel screens nothing
    if name == "gauss":
        assert 0 < out[1][3] < out[1][2]          # and here the screen does remove work
        if (n, d) == (20000, 64):
            assert out[1][3] < out[1][2] // 3     # most of it where rows are long (k = 16 of ~200)


def test_screened_path_equals_oracle_on_offset_clusters(gpu32, oracle_mod):
    rng = np.random.default_rng(4242)
    n, d, k, tries = 6000, 64, 16, 4
    centres = rng.standard_normal((64, d)) * 20.0 + 300.0
    pts = (centres[rng.integers(0, 64, n)] + rng.standard_normal((n, d))).astype(np.float32)
    want = oracle_mod.restatement(np.float32).precomp(pts, k, tries, 6, 1, 1, 1, want_save=False, seed=5)
    gpu32.lib.annb_leaf_screen_mode(1)
    got = gpu32.precomp(pts, k, tries, 6, 1, 1, 1, want_save=False, seed=5)
    assert np.array_equal(got.ids, want.ids)
    assert same_bits(got.dists, want.dists)
