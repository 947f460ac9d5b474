"""Host-side check of the bracket the screened S3 path relies on (annb_leaf_screen.cuh).

The CUDA kernel drops a candidate only if lo = D' - t^2 exceeds an upper bound of the row's
k-th best distance, so everything rests on |D - D'| <= t^2 for every pair, where (all in the
scaled units of the fp16 copy)
    c   = (x - mean) * scale            scale = power of two with max|c| in [4, 8)
    c'  = fp16(c)
    D'  = sum c'_q^2 + sum c'_c^2 - 2 c'_q . c'_c          (fp32 accumulation)
    s   = 0.03226 * (||c|| (1 + 2^-12) + sqrt(d) 2^-14)    t = s_q + s_c
    D   = scale^2 * (the reference's fp32 tree sum of (x_q - x_c)^2)
This test replays those formulas in numpy (no GPU) on the data shapes the GPU tests use to
stress them and asserts the bracket with the margin the derivation in DESIGN.md claims.
"""
import numpy as np
import pytest


def tree_sum_f32(v):
    """compute.cl:160-167 for a power-of-two length: m[z] += m[z + l/2], every step rounded to fp32."""
    v = v.astype(np.float32).copy()
    l = v.shape[-1]
    while l > 1:
        h = l // 2
        v[..., :h] = v[..., :h] + v[..., h:l]
        l = h
    return v[..., 0]


def brackets(x, pairs):
    x = x.astype(np.float32)
    n, d = x.shape
    mean = x.mean(axis=0, dtype=np.float64).astype(np.float32)
    centred = x - mean
    cmax = float(np.abs(centred).max())
    scale = np.float32(2.0 ** (2 - int(np.floor(np.log2(cmax))))) if cmax > 0 else np.float32(1)
    c = centred * scale
    assert 4 <= np.abs(c).max() < 8 or cmax == 0
    c16 = c.astype(np.float16).astype(np.float32)
    n2 = (c16 * c16).sum(axis=1, dtype=np.float32)
    s = np.float32(0.03226) * (np.sqrt((c * c).sum(axis=1, dtype=np.float32)) * np.float32(1 + 2.0 ** -12)
                               + np.float32(np.sqrt(d) * 2.0 ** -14))
    q, cnd = pairs[:, 0], pairs[:, 1]
    dot = np.einsum("ij,ij->i", c16[q], c16[cnd], dtype=np.float32)
    dprime = (n2[q] + n2[cnd]) - np.float32(2) * dot
    t = s[q] + s[cnd]
    diff = x[q] - x[cnd]
    exact = tree_sum_f32(diff * diff).astype(np.float64) * float(scale) ** 2
    return dprime.astype(np.float64), (t * t).astype(np.float64), exact


def datasets(rng, n, d):
    yield "gauss", rng.standard_normal((n, d))
    yield "offset", rng.standard_normal((n, d)) + 1000.0
    centres = rng.standard_normal((32, d)) * 50.0
    yield "clusters", centres[rng.integers(0, 32, n)] + rng.standard_normal((n, d)) * 1e-3
    x = rng.standard_normal((n, d)); x[7] *= 3.0e4
    yield "outlier", x
    x = rng.standard_normal((n, d)); x[: n // 2] *= 1e-6
    yield "near_mean", x
    yield "uniform_positive", rng.random((n, d)) * 255.0


@pytest.mark.parametrize("d", [16, 32, 64])
def test_exact_distance_lies_inside_the_bracket(d):
    rng = np.random.default_rng(1234 + d)
    n = 4000
    for name, x in datasets(rng, n, d):
        pairs = rng.integers(0, n, size=(200000, 2))
        pairs = pairs[pairs[:, 0] != pairs[:, 1]]
        dprime, slack, exact = brackets(x, pairs)
        err = np.abs(exact - dprime)
        worst = float((err / slack).max())
        assert worst < 1.0, (name, worst)
        # the derivation leaves ~2% of the slack as margin for the accumulation order of the
        # tensor cores; here (numpy's fp32 order) the observed error stays well inside
        assert worst < 0.9, (name, worst)


def test_duplicate_points_have_a_positive_slack():
    x = np.tile(np.random.default_rng(5).standard_normal((10, 64)), (20, 1))
    pairs = np.array([[i, i + 10] for i in range(150)])
    dprime, slack, exact = brackets(x, pairs)
    assert np.all(exact == 0) and np.all(slack > 0) and np.all(np.abs(dprime) <= slack)
