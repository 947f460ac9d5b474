"""Edge cases of precomp_gpu / query_gpu against the CPU restatement (bit-exact)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import same_bits

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gpu():
    from approximatenn_b200.api import gpu_backend
    return {np.dtype(np.float32): gpu_backend(np.float32), np.dtype(np.float64): gpu_backend(np.float64)}


def both(gpu, oracle_mod, pts, k, tries, rot, seed, ycnt=40):
    dtype = pts.dtype
    orc, b = oracle_mod.restatement(dtype), gpu[np.dtype(dtype)]
    want = orc.precomp(pts, k, tries, *rot, want_save=True, seed=seed)
    got = b.precomp(pts, k, tries, *rot, want_save=True, seed=seed)
    assert np.array_equal(got.ids, want.ids), "ids"
    assert same_bits(got.dists, want.dists), "dists"
    assert np.array_equal(got.save.par_maxes, want.save.par_maxes)
    for t in range(tries):
        assert np.array_equal(got.save.which_par(t), want.save.which_par(t))
    assert same_bits(got.save.bases, want.save.bases) and same_bits(got.save.row_means, want.save.row_means)
    y = np.random.default_rng(seed + 7).standard_normal((ycnt, pts.shape[1])).astype(dtype)
    qw, qg = orc.query(want.save, pts, y), b.query(got.save, pts, y)
    assert np.array_equal(qg.ids, qw.ids) and same_bits(qg.dists, qw.dists), "query"
    want.save.free(); got.save.free()


CASES = [
    # dtype, n, d, k, tries, rot
    (np.float32, 3000, 128, 16, 4, (6, 1, 1, 1)),      # register hash d_max=128, generic S3, fast S5 (16 coords per lane)
    (np.float32, 1500, 256, 8, 3, (6, 1, 1, 1)),       # d=256: shared-memory hash kernel, warp kernels with 8 coords per lane
    (np.float32, 2000, 100, 12, 5, (4, 10, 2, 2)),     # d not a power of two: generic distance tree everywhere
    (np.float64, 1800, 64, 16, 4, (6, 1, 1, 1)),       # double d=64: no tiled instantiation
    (np.float32, 4000, 32, 100, 2, (6, 1, 1, 1)),      # k=100: four list registers per lane
    (np.float32, 2500, 24, 10, 1, (6, 1, 1, 1)),       # one try (k*tries < 16)
    (np.float32, 2500, 24, 6, 33, (2, 2, 1, 1)),       # many tries, k*tries not a power of two
    (np.float32, 3000, 40, 9, 8, (0, 1, 0, 1)),        # no rotations at all
    (np.float32, 1000, 16, 3, 2, (6, 1, 1, 1)),        # k(k+1) < 16 and k*tries < 16: literal rows everywhere
    (np.float64, 777, 33, 7, 5, (3, 3, 2, 1)),
]


@pytest.mark.parametrize("dtype,n,d,k,tries,rot", CASES)
def test_shapes_and_parameters(gpu, oracle_mod, dtype, n, d, k, tries, rot):
    rng = np.random.default_rng(n + d + k)
    both(gpu, oracle_mod, rng.standard_normal((n, d)).astype(dtype), k, tries, rot, seed=n + k)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_duplicated_points_make_exact_ties_everywhere(gpu, oracle_mod, dtype):
    """Every vector appears four times: zero distances and exact ties in every row, i.e. the
    literal-network redo decides almost every row.  Still bit-identical to the reference."""
    rng = np.random.default_rng(77)
    base = rng.standard_normal((600, 32)).astype(dtype)
    pts = np.ascontiguousarray(np.tile(base, (4, 1))[rng.permutation(2400)])
    both(gpu, oracle_mod, pts, 10, 6, (6, 1, 1, 1), seed=91)
    from approximatenn_b200.api import gpu_backend
    import ctypes
    out = (ctypes.c_ulonglong * 3)()
    gpu[np.dtype(dtype)].lib.annb_literal_rows(out, 1)
    assert sum(out) > 0


def test_quantised_coordinates_tie_heavily(gpu, oracle_mod):
    rng = np.random.default_rng(3)
    pts = np.round(rng.standard_normal((3000, 16)) * 2).astype(np.float32) / 2      # few distinct distances
    both(gpu, oracle_mod, pts, 10, 10, (6, 1, 1, 1), seed=17)


def test_successive_calls_of_different_size_reuse_the_arena(gpu, oracle_mod):
    rng = np.random.default_rng(8)
    for n, d, k in [(5000, 64, 16), (900, 16, 10), (7000, 32, 16), (1200, 64, 16)]:
        both(gpu, oracle_mod, rng.standard_normal((n, d)).astype(np.float32), k, 4, (6, 1, 1, 1), seed=n)


BAD = r"""
import sys, numpy as np
sys.path.insert(0, %r)
from approximatenn_b200.api import gpu_backend
pts = np.random.default_rng(0).standard_normal((256, 16)).astype(np.float32)
gpu_backend(np.float32).precomp(pts, %%d, tries=%%d, rots_before=%%d, rot_len_before=%%d, seed=1)
print("COMPUTED")
""" % ROOT


@pytest.mark.parametrize("k,tries,rb,lb,needle", [
    (300, 2, 6, 1, "k must be smaller than n"),
    (4, 0, 6, 1, "tries >= 1"),
    (4, 2, 6, 9, "rot_len_before"),
])
def test_invalid_arguments_are_fatal_like_the_reference(k, tries, rb, lb, needle):
    out = subprocess.run([sys.executable, "-c", BAD % (k, tries, rb, lb)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 1 and "COMPUTED" not in out.stdout
    assert needle in out.stderr, out.stderr


def test_cleanup_then_reuse(gpu, oracle_mod):
    """gpu_cleanup() releases every device and pinned allocation; the next call re-initialises
    (the reference's programs call gpu_init()/gpu_cleanup() around their loops)."""
    rng = np.random.default_rng(12)
    pts = rng.standard_normal((3000, 32)).astype(np.float32)
    b = gpu[np.dtype(np.float32)]
    want = oracle_mod.restatement(np.float32).precomp(pts, 16, 4, seed=3)
    for _ in range(2):
        got = b.precomp(pts, 16, 4, want_save=True, seed=3)
        q = b.query(got.save, pts, pts[:10])
        assert np.array_equal(got.ids, want.ids) and same_bits(got.dists, want.dists)
        got.save.free()
        b.lib.gpu_cleanup()
    b.lib.gpu_init()


def test_grouped_merge_differs_only_in_tie_rows(oracle_mod, capfd):
    """ANN_B200_MERGE_GROUP forces the running merge used when the per-try lists do not fit the
    device.  It cannot redo exact-tie rows literally (and says so on stderr); every other row
    must still be the reference's, bit for bit.  k*tries is a power of two here, so the
    prefix-corner rule is not involved."""
    import os
    from approximatenn_b200.api import gpu_backend
    n, d, k, tries = 30000, 64, 16, 8
    rng = np.random.default_rng(31)
    pts = rng.standard_normal((n, d)).astype(np.float32)
    want = oracle_mod.restatement(np.float32).precomp(pts, k, tries, seed=31)
    os.environ["ANN_B200_MERGE_GROUP"] = "3"
    try:
        got = gpu_backend(np.float32).precomp(pts, k, tries, seed=31)
    finally:
        del os.environ["ANN_B200_MERGE_GROUP"]
    bad = np.flatnonzero((got.ids != want.ids).any(axis=1) |
                         (got.dists.view(np.uint32) != want.dists.view(np.uint32)).any(axis=1))
    assert len(bad) <= n // 1000, f"{len(bad)} rows differ"
    assert "merging 3 at a time" in capfd.readouterr().err
