"""Parity of query_gpu (through the C-ABI) with the reference's query_cpu: golden vectors
and live comparisons against the CPU restatement, including the reference's transposed read
of the sign buffer (SURVEY.md "three facts" #3) and the y == points self-exclusion."""
import numpy as np
import pytest

from conftest import golden_names, load_golden, same_bits

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from approximatenn_b200.api import gpu_backend
    return {np.dtype(np.float32): gpu_backend(np.float32), np.dtype(np.float64): gpu_backend(np.float64)}


@pytest.mark.parametrize("name", golden_names())
def test_query_gpu_reproduces_reference_golden(gpu, name):
    g = load_golden(name)
    b = gpu[g["dtype"]]
    res = b.precomp(g["points"], g["k"], g["tries"], *g["rot"], want_save=True, seed=g["seed"])
    q = b.query(res.save, g["points"], g["y"])
    assert np.array_equal(q.ids, g["q_ids"].astype(np.uint64)), "query ids differ"
    assert same_bits(q.dists, g["q_dists"]), "query distances differ"
    # a second call answers from the device-resident index and must not change
    q2 = b.query(res.save, g["points"], g["y"], want_dists=False)
    assert q2.dists is None and np.array_equal(q2.ids, q.ids)
    res.save.free()


LIVE = [
    (np.float32, 8192, 64, 16, 8, (6, 1, 1, 1), 2000, 301),
    (np.float64, 4096, 32, 16, 8, (6, 1, 1, 1), 500, 302),
    (np.float32, 3001, 80, 10, 10, (6, 1, 1, 1), 77, 303),
    (np.float32, 5000, 16, 10, 10, (6, 1, 1, 1), 3000, 304),
    (np.float32, 4000, 48, 40, 3, (3, 2, 1, 1), 100, 305),
]


@pytest.mark.parametrize("dtype,n,d,k,tries,rot,ycnt,seed", LIVE)
def test_query_gpu_equals_oracle_live(gpu, oracle_mod, dtype, n, d, k, tries, rot, ycnt, seed):
    rng = np.random.default_rng(seed)
    pts = rng.standard_normal((n, d)).astype(dtype)
    y = rng.standard_normal((ycnt, d)).astype(dtype)
    orc, b = oracle_mod.restatement(dtype), gpu[np.dtype(dtype)]
    want_pre = orc.precomp(pts, k, tries, *rot, want_save=True, seed=seed)
    got_pre = b.precomp(pts, k, tries, *rot, want_save=True, seed=seed)
    want = orc.query(want_pre.save, pts, y)
    got = b.query(got_pre.save, pts, y)
    assert np.array_equal(got.ids, want.ids) and same_bits(got.dists, want.dists)
    # the index built by the CPU path is interchangeable: same save_t layout
    cross = b.query(want_pre.save, pts, y)
    assert np.array_equal(cross.ids, want.ids) and same_bits(cross.dists, want.dists)
    # y is points (same pointer): the reference switches self-exclusion on (compute.cl:145)
    want_s = orc.query(want_pre.save, pts, pts)
    got_s = b.query(got_pre.save, pts, pts)
    assert np.array_equal(got_s.ids, want_s.ids) and same_bits(got_s.dists, want_s.dists)
    want_pre.save.free(); got_pre.save.free()


def test_query_cache_is_dropped_with_the_save(gpu):
    b = gpu[np.dtype(np.float32)]
    rng = np.random.default_rng(5)
    outs = []
    for rep in range(2):                      # a second index, most likely at recycled addresses
        pts = rng.standard_normal((3000, 32)).astype(np.float32)
        y = rng.standard_normal((64, 32)).astype(np.float32)
        r = b.precomp(pts, 16, 8, want_save=True, seed=40 + rep)
        q = b.query(r.save, pts, y)
        # exact brute force: every returned id must be a real point with the reported distance
        d2 = ((y[:, None, :].astype(np.float64) - pts[q.ids.astype(np.int64)].astype(np.float64)) ** 2).sum(-1)
        assert np.allclose(d2, q.dists, rtol=1e-5)
        outs.append(q.ids)
        r.save.free()
    assert not np.array_equal(outs[0], outs[1])


@pytest.mark.parametrize("n,d,k,tries,ycnt", [(20000, 64, 16, 8, 5000), (9000, 32, 10, 10, 3000), (6000, 16, 32, 3, 1000)])
def test_pipelined_query_rows_equal_the_generic_kernel(gpu, n, d, k, tries, ycnt):
    """query_rows_fast_kernel (buffered ids, whole-line loads), with and without its fp16 screen,
    vs the warp-per-row kernel it replaces: every bit of ids and distances, with and without
    self-exclusion, on Gaussian and on offset/clustered data."""
    import os
    b = gpu[np.dtype(np.float32)]
    rng = np.random.default_rng(n + d)
    pts = rng.standard_normal((n, d)).astype(np.float32)
    y = rng.standard_normal((ycnt, d)).astype(np.float32)
    r = b.precomp(pts, k, tries, want_save=True, seed=9)
    fast, fast_s = b.query(r.save, pts, y), b.query(r.save, pts, pts)          # screened (default)
    os.environ["ANN_B200_QUERY_SCREEN"] = "0"
    try:
        plain, plain_s = b.query(r.save, pts, y), b.query(r.save, pts, pts)    # pipelined, every candidate exact
    finally:
        del os.environ["ANN_B200_QUERY_SCREEN"]
    os.environ["ANN_B200_NO_FAST_QUERY"] = "1"
    try:
        slow, slow_s = b.query(r.save, pts, y), b.query(r.save, pts, pts)      # warp-per-row kernel of round 1
    finally:
        del os.environ["ANN_B200_NO_FAST_QUERY"]
    for a, c in ((fast, slow), (fast_s, slow_s), (plain, slow), (plain_s, slow_s)):
        assert np.array_equal(a.ids, c.ids) and same_bits(a.dists, c.dists)
    # data that stresses the brackets: a common offset and tight clusters
    centres = rng.standard_normal((40, d)) * 30.0 + 500.0
    pts2 = (centres[rng.integers(0, 40, n)] + rng.standard_normal((n, d)) * 0.01).astype(np.float32)
    y2 = (centres[rng.integers(0, 40, ycnt)] + rng.standard_normal((ycnt, d)) * 0.01).astype(np.float32)
    r2 = b.precomp(pts2, k, tries, want_save=True, seed=10)
    a = b.query(r2.save, pts2, y2)
    os.environ["ANN_B200_NO_FAST_QUERY"] = "1"
    try:
        c = b.query(r2.save, pts2, y2)
    finally:
        del os.environ["ANN_B200_NO_FAST_QUERY"]
    assert np.array_equal(a.ids, c.ids) and same_bits(a.dists, c.dists)
    r.save.free(); r2.save.free()


def test_corrected_sign_layout_is_opt_in_and_finds_the_point_itself(gpu):
    """ANN_B200_QUERY_LAYOUT=fixed reads the sign buffer the way it was written ([x][try]).
    Querying copies of indexed points then lands in the point's own bucket in every try, so the
    nearest neighbour is the point itself at distance 0; the reference's transposed read
    (the default, reproduced bit for bit elsewhere in this file) sends most queries to unrelated
    buckets."""
    import os
    b = gpu[np.dtype(np.float32)]
    rng = np.random.default_rng(77)
    n, d, k, tries = 16384, 16, 10, 10
    pts = rng.standard_normal((n, d)).astype(np.float32)
    rows = rng.choice(n, size=2000, replace=False)
    y = pts[rows].copy()                                   # not the same pointer: no self-exclusion
    r = b.precomp(pts, k, tries, want_save=True, seed=13)
    default = b.query(r.save, pts, y)
    os.environ["ANN_B200_QUERY_LAYOUT"] = "fixed"
    try:
        fixed = b.query(r.save, pts, y)
    finally:
        del os.environ["ANN_B200_QUERY_LAYOUT"]
    hit_fixed = float(np.mean((fixed.ids[:, 0] == rows) & (fixed.dists[:, 0] == 0)))
    hit_default = float(np.mean(default.ids[:, 0] == rows))
    print(f"own point found first: fixed layout {hit_fixed:.3f}, reference layout {hit_default:.3f}")
    assert hit_fixed > 0.999
    assert hit_default < 0.9
    r.save.free()
