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


# ---- cutoff carried from try to try (annb_cutoff_update / annb_leaf_topk_cut) ----------------

CUT_SHAPES = [(20000, 64, 16, 8), (12000, 32, 8, 8), (9000, 16, 16, 4), (30000, 64, 16, 2)]


@pytest.mark.parametrize("name,make", _datasets())
@pytest.mark.parametrize("n,d,k,tries", CUT_SHAPES)
def test_cutoff_does_not_change_a_bit(gpu32, name, make, n, d, k, tries):
    """k*tries is a power of two here, so later tries stop their lists at the running cutoff;
    ids and distances of the final rows must be the ones of the full lists."""
    rng = np.random.default_rng(zlib.crc32(f"cut-{name}-{n}-{d}".encode()))
    pts = np.ascontiguousarray(make(rng, n, d), dtype=np.float32)
    gpu32.lib.annb_leaf_screen_mode(1)
    out = {}
    try:
        for mode in (0, 1):
            gpu32.lib.annb_leaf_cutoff_mode(mode)
            gpu32.lib.annb_leaf_exact_pairs(1)
            res = gpu32.precomp(pts, k, tries, 6, 1, 1, 1, want_save=False, seed=31)
            out[mode] = (res.ids, res.dists, int(gpu32.lib.annb_leaf_exact_pairs(0)))
    finally:
        gpu32.lib.annb_leaf_cutoff_mode(1)
    assert np.array_equal(out[0][0], out[1][0])
    assert same_bits(out[0][1], out[1][1])
    if name == "gauss":
        assert out[1][2] < out[0][2]              # fewer pairs reach the exact tree
        if tries == 8:
            assert out[1][2] < 0.6 * out[0][2]


def test_cutoff_equals_oracle(gpu32, oracle_mod):
    rng = np.random.default_rng(808)
    n, d, k, tries = 8000, 64, 16, 8
    pts = np.ascontiguousarray(rng.standard_normal((n, d)), dtype=np.float32)
    pts[: n // 8] = pts[n // 8: n // 4]                        # duplicated points: exact ties
    want = oracle_mod.restatement(np.float32).precomp(pts, k, tries, 6, 1, 1, 1, want_save=False, seed=12)
    gpu32.lib.annb_leaf_screen_mode(1)
    gpu32.lib.annb_leaf_cutoff_mode(1)
    got = gpu32.precomp(pts, k, tries, 6, 1, 1, 1, want_save=False, seed=12)
    assert np.array_equal(got.ids, want.ids)
    assert same_bits(got.dists, want.dists)


@pytest.mark.parametrize("k", [16, 10, 1])
def test_cutoff_update_kernel_matches_its_model(gpu32, k):
    """run = k smallest distinct values of (run, new list); cutoff = the k-th, +inf while unknown."""
    import torch
    rng = np.random.default_rng(k)
    n = 5000
    pool = rng.random((n, 40)).astype(np.float32)               # values shared between the lists
    lib = gpu32.lib
    lib.annb_cutoff_update.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                       ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    lib.annb_cutoff_update.restype = None
    run_d = torch.empty((n, k), dtype=torch.float32, device="cuda")
    cut_d = torch.empty((n,), dtype=torch.float32, device="cuda")
    run = np.full((n, k), np.inf, dtype=np.float32)
    for step in range(4):
        new = np.full((n, k), np.inf, dtype=np.float32)
        for x in range(n):
            m = int(rng.integers(0, k + 1))                      # ragged lists, +inf padded
            new[x, :m] = np.sort(rng.choice(pool[x], size=m, replace=True))   # repeats inside a list too
        if step == 2:
            new[7, 0] = np.nan
        new_d = torch.from_numpy(new).cuda()
        torch.cuda.synchronize()
        lib.annb_cutoff_update(new_d.data_ptr(), run_d.data_ptr(), cut_d.data_ptr(), n, k, int(step == 0), None)
        torch.cuda.synchronize()
        for x in range(n):
            vals = np.unique(np.concatenate([run[x][np.isfinite(run[x])], new[x][np.isfinite(new[x])]]))
            if np.isnan(new[x]).any() or np.isnan(run[x]).any():
                vals = vals[:0]
            m = min(k, len(vals))
            run[x] = np.inf
            run[x, :m] = vals[:m]
        assert same_bits(run_d.cpu().numpy(), run)
        assert same_bits(cut_d.cpu().numpy(), run[:, k - 1].copy())
