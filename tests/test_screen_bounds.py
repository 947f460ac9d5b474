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


def _merge8(m):
    """bitonic_merge8 of annb_leaf_screen.cuh: ascending sort of a bitonic sequence of 8."""
    m = list(m)

    def ce(a, b):
        if m[a] > m[b]:
            m[a], m[b] = m[b], m[a]
    for i in range(4):
        ce(i, i + 4)
    for a, b in ((0, 2), (1, 3), (4, 6), (5, 7)):
        ce(a, b)
    for a, b in ((0, 1), (2, 3), (4, 5), (6, 7)):
        ce(a, b)
    return m


def _quad_sixteenth_smallest(lists):
    """quad_sixteenth_smallest: four lanes, each with two ascending lists of 4 (a, b)."""
    m = [_merge8(list(a) + list(b)[::-1]) for a, b in lists]
    c = []
    for t in range(4):
        other = m[t ^ 1]
        vals = [(max if t & 1 else min)(m[t][i], other[7 - i]) for i in range(8)]
        c.append(_merge8(vals))
    w = []
    for t in range(4):
        other = c[t ^ 3]
        w.append(max(min(c[t][r], other[7 - r]) for r in range(8)))
    out = [max(w[t], w[t ^ 1]) for t in range(4)]
    assert len(set(out)) == 1                       # every lane of the quad gets the same value
    return out[0]


def test_threshold_network_returns_the_sixteenth_smallest_of_32():
    """The selection network behind Theta (an upper bound of the row's 16th smallest distance)."""
    rng = np.random.default_rng(99)
    for trial in range(2000):
        vals = rng.standard_normal(32).astype(np.float32)
        if trial % 5 == 0:
            vals[rng.integers(0, 32, rng.integers(1, 24))] = np.inf     # rows with few candidates
        if trial % 7 == 0:
            vals[:8] = vals[8]                                          # repeated values
        groups = vals.reshape(4, 2, 4)
        lists = [(sorted(g[0]), sorted(g[1])) for g in groups]
        assert _quad_sixteenth_smallest(lists) == np.sort(vals)[15]


def test_keeping_four_per_lane_and_parity_bounds_the_sixteenth_smallest():
    """Pass 1 keeps, per lane and column parity, the 4 smallest upper bounds it has seen; the
    16th smallest of those 32 is never below the 16th smallest of the whole row."""
    rng = np.random.default_rng(7)
    for trial in range(300):
        ncols = int(rng.integers(3, 60)) * 8
        hi = rng.gamma(8.0, 1.0, ncols).astype(np.float32)
        kept = []
        for t in range(4):                          # lane t of the quad: columns j0 + 2t (+1)
            for e in range(2):
                col = hi[(np.arange(ncols) % 8) == 2 * t + e]
                best = np.sort(np.concatenate([col, np.full(4, np.inf, np.float32)]))[:4]
                kept.append(best)
        lists = [(list(kept[2 * t]), list(kept[2 * t + 1])) for t in range(4)]
        theta = _quad_sixteenth_smallest(lists)
        truth = np.sort(np.concatenate([hi, np.full(16, np.inf, np.float32)]))[15]
        assert theta >= truth


# ---- cutoff carried from try to try: the comparison the scan makes -----------------------------

def _half_round_toward_zero(v):
    """cvt.rz.relu.f16.f32 of annb_leaf_screen.cuh (pack_lower_bounds): negatives -> 0, magnitude
    rounded towards zero, +inf stays +inf."""
    v = np.maximum(np.asarray(v, dtype=np.float64), 0.0)
    h = v.astype(np.float16)                               # round to nearest
    up = h.astype(np.float64) > v                          # rounded up: step one value down
    h = np.where(up, np.nextafter(h, np.float16(0)), h).astype(np.float16)
    return h


def _half_round_up(v):
    """__float2half_ru, with the kernel's clamp of +inf to the largest finite half."""
    v = np.asarray(v, dtype=np.float64)
    h = v.astype(np.float16)
    down = h.astype(np.float64) < v
    h = np.where(down, np.nextafter(h, np.float16(np.inf)), h).astype(np.float16)
    return np.where(np.isinf(h), np.float16(65504.0), h)


@pytest.mark.parametrize("d", [16, 64])
def test_a_candidate_at_or_below_the_cutoff_always_passes_the_scan(d):
    """annb_leaf_screen.cuh keeps a candidate when fp16_rz(lo) <= fp16_ru(float_ru(cutoff * scale^2)),
    lo = D' - t^2.  The worst case for the rule is a cutoff EQUAL to the candidate's own exact
    distance (it then is the point's current k-th neighbour and must stay in the list)."""
    rng = np.random.default_rng(4321 + d)
    n = 3000
    for name, x in datasets(rng, n, d):
        pairs = rng.integers(0, n, size=(100000, 2))
        pairs = pairs[pairs[:, 0] != pairs[:, 1]]
        dprime, slack, exact_scaled = brackets(x, pairs)
        lo = _half_round_toward_zero(dprime - slack).astype(np.float64)
        # the kernel scales the fp32 cutoff by scale^2 in double and rounds up twice; `exact_scaled`
        # is that product for a cutoff equal to the pair's distance
        thr = _half_round_up(np.nextafter(exact_scaled.astype(np.float32), np.float32(np.inf))).astype(np.float64)
        assert np.all(lo <= thr), (name, float((lo - thr).max()))
        # and a slightly larger cutoff (any later, looser bound) passes a fortiori
        thr2 = _half_round_up(exact_scaled * 1.5 + 1e-3).astype(np.float64)
        assert np.all(lo <= thr2), name


def _cutoff_update_model(run, new, k, K=16):
    """Line-by-line numpy model of cutoff_update_kernel<16> for one point: duplicate flags, bitonic
    sort of the new list, reverse / min / bitonic merge with the running one."""
    inf = np.float32(np.inf)
    v = np.full(K, inf, np.float32); v[:k] = new
    r = np.full(K, inf, np.float32); r[:k] = run
    bad = bool(np.isnan(v).any() or np.isnan(r).any())
    dup = np.zeros(K, bool)
    for j in range(1, K):
        dup[j] = v[j] == v[j - 1]
    for j in range(K):
        dup[j] |= bool((v[j] == r).any())
    v[dup] = inf

    def ce(a, i, p):
        if a[p] < a[i]:
            a[i], a[p] = a[p], a[i]
    size = 2
    while size <= K:
        stride = size >> 1
        while stride >= 1:
            for i in range(K):
                p = i ^ stride
                if p > i:
                    if (i & size) == 0:
                        ce(v, i, p)
                    else:
                        ce(v, p, i)
            stride >>= 1
        size <<= 1
    r = np.minimum(r, v[::-1])
    stride = K >> 1
    while stride >= 1:
        for i in range(K):
            p = i ^ stride
            if p > i:
                ce(r, i, p)
        stride >>= 1
    if bad:
        r[:] = inf
    return r[:k].copy(), r[k - 1]


@pytest.mark.parametrize("k", [16, 10, 3])
def test_cutoff_update_network_keeps_the_k_smallest_distinct_values(k):
    rng = np.random.default_rng(100 + k)
    for trial in range(400):
        pool = rng.random(24).astype(np.float32)
        run = np.full(k, np.inf, np.float32)
        for step in range(5):
            m = int(rng.integers(0, k + 1))
            new = np.full(k, np.inf, np.float32)
            new[:m] = np.sort(rng.choice(pool, size=m, replace=True))
            want = np.unique(np.concatenate([run[np.isfinite(run)], new[np.isfinite(new)]]))[:k]
            run, cutoff = _cutoff_update_model(run, new, k)
            assert np.array_equal(run[: len(want)], want) and np.all(np.isinf(run[len(want):]))
            assert cutoff == (want[k - 1] if len(want) >= k else np.inf)
