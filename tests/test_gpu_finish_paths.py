"""S4 / S5 fast paths against their predecessors and the oracle.

* screened supercharge (fp16 brackets from the original-order copy, exact tree for the survivors)
  vs supercharge_fast_kernel: the screen may only drop candidates that are strictly farther
  than the row's k-th own distance, so every output bit must be the same with it on or off;
* thread-per-point merge vs the warp merge kernels: same rows, same tie reports.
Data sets stress the brackets and the fallbacks: common offset, tight clusters, duplicated
points (exact ties -> literal rows), an outlier (fp16 scale), points almost on the mean, and
data scaled far outside the range in which the screen trusts its brackets.
"""
import ctypes
import zlib

import numpy as np
import pytest

from conftest import same_bits
from test_gpu_screen import _datasets

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu32():
    from approximatenn_b200.api import gpu_backend
    g = gpu_backend(np.float32)
    yield g
    g.lib.annb_supercharge_screen_mode(1)
    g.lib.annb_merge_thread_mode(1)


def _stats(g, reset):
    out = (ctypes.c_ulonglong * 2)()
    g.lib.annb_supercharge_screen_stats(out, reset)
    return int(out[0]), int(out[1])


SHAPES = [(20000, 64, 16, 4), (12000, 32, 10, 5), (9000, 16, 5, 6), (6000, 128, 20, 3), (7000, 32, 32, 2)]
EXTRA = [("tiny_scale", lambda rng, n, d: rng.standard_normal((n, d)) * 1e-15),
         ("huge_scale", lambda rng, n, d: rng.standard_normal((n, d)) * 1e15)]


@pytest.mark.parametrize("name,make", _datasets() + EXTRA)
@pytest.mark.parametrize("n,d,k,tries", SHAPES)
def test_fast_paths_do_not_change_a_bit(gpu32, name, make, n, d, k, tries):
    rng = np.random.default_rng(zlib.crc32(f"s5-{name}-{n}-{d}".encode()))
    pts = np.ascontiguousarray(make(rng, n, d), dtype=np.float32)
    out = {}
    for mode in (0, 1):
        gpu32.lib.annb_supercharge_screen_mode(mode)
        gpu32.lib.annb_merge_thread_mode(mode)
        _stats(gpu32, 1)
        res = gpu32.precomp(pts, k, tries, 6, 1, 1, 1, want_save=False, seed=91)
        out[mode] = (res.ids, res.dists, _stats(gpu32, 0))
    assert np.array_equal(out[0][0], out[1][0])
    assert same_bits(out[0][1], out[1][1])
    assert out[0][2] == (0, 0)
    bracketed, exact = out[1][2]
    assert bracketed > 0 and exact <= bracketed
    if name == "gauss":
        assert exact < bracketed // 2             # the screen removes most of the gathers here
    if name in ("tiny_scale", "huge_scale"):
        assert exact == bracketed                 # outside the trusted range everything is measured


@pytest.mark.parametrize("dtype,n,d,k,tries", [(np.float32, 8000, 64, 16, 8), (np.float32, 5000, 16, 10, 10),
                                                (np.float64, 4000, 32, 16, 10), (np.float32, 3000, 24, 6, 13),
                                                (np.float32, 4000, 32, 12, 20)])
def test_fast_paths_equal_oracle(oracle_mod, dtype, n, d, k, tries):
    """Prefix corner live (k*T not a power of two), more lists than one staging pass holds
    (T = 20 -> warp kernels), double rows, k not a multiple of 4 (scalar staging)."""
    from approximatenn_b200.api import gpu_backend
    rng = np.random.default_rng(n + d)
    pts = rng.standard_normal((n, d)).astype(dtype)
    gpu = gpu_backend(dtype)
    gpu.lib.annb_supercharge_screen_mode(1)
    gpu.lib.annb_merge_thread_mode(1)
    want = oracle_mod.restatement(dtype).precomp(pts, k, tries, seed=17)
    got = gpu.precomp(pts, k, tries, seed=17)
    assert np.array_equal(got.ids, want.ids)
    assert same_bits(got.dists, want.dists)


@pytest.mark.parametrize("n,d,k,tries", [(140000, 64, 16, 4), (20000, 32, 10, 5), (3000, 16, 5, 6)])
def test_locality_order_does_not_change_a_bit(gpu32, n, d, k, tries):
    """With ANN_B200_S5_LOCALITY=1 the supercharge works on its rows in bucket order of one try
    (annb_locality_order); the rows it writes must not depend on that order.  n = 140000 has 8 egress chunks, each with its
    own permutation."""
    import os
    rng = np.random.default_rng(n)
    pts = rng.standard_normal((n, d)).astype(np.float32)
    gpu32.lib.annb_supercharge_screen_mode(1)
    off = gpu32.precomp(pts, k, tries, seed=5)
    os.environ["ANN_B200_S5_LOCALITY"] = "1"
    try:
        on = gpu32.precomp(pts, k, tries, seed=5)
    finally:
        del os.environ["ANN_B200_S5_LOCALITY"]
    assert np.array_equal(on.ids, off.ids)
    assert same_bits(on.dists, off.dists)
