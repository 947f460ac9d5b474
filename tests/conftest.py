import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    for key in ("n", "d", "k", "tries", "seed", "d_short"):
        g[key] = int(g[key])
    g["rot"] = tuple(int(v) for v in g["rot"])
    g["dtype"] = g["points"].dtype
    return g


def same_bits(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def assert_matches_golden(g, res, qres=None, check_save=True):
    """Bit-exact comparison of a precomp (+query) result with a golden record."""
    assert np.array_equal(res.ids, g["ids"].astype(np.uint64)), "neighbour ids differ"
    if res.dists is not None:
        assert same_bits(res.dists, g["dists"]), "squared distances differ"
    if check_save and res.save is not None:
        s = res.save
        assert (s.tries, s.n, s.k, s.d_short, s.d_long) == (g["tries"], g["n"], g["k"], g["d_short"], g["d"])
        assert np.array_equal(s.par_maxes, g["par_maxes"])
        assert np.array_equal(s.graph, g["graph"].astype(np.uint64))
        assert same_bits(s.row_means, g["row_means"]), "row_means differ"
        assert same_bits(s.bases, g["bases"]), "bases differ"
        for t in range(g["tries"]):
            assert np.array_equal(s.which_par(t), g[f"which_par_{t}"].astype(np.uint64)), f"which_par[{t}] differs"
    if qres is not None:
        assert np.array_equal(qres.ids, g["q_ids"].astype(np.uint64)), "query ids differ"
        if qres.dists is not None:
            assert same_bits(qres.dists, g["q_dists"]), "query distances differ"


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


def knn_diff(got_ids, got_d, want_ids, want_d):
    """Row-level comparison of two kNN results under the north_star rule: squared distances
    must agree exactly; ids must agree except inside a run of EQUAL distances (an exact tie,
    where the reference's order is an artefact of its sorting network).
    Returns (rows whose distances differ, rows whose ids differ outside ties, rows that differ
    only inside ties)."""
    same_d = (np.ascontiguousarray(got_d).view(np.uint8).reshape(got_d.shape[0], -1) ==
              np.ascontiguousarray(want_d).view(np.uint8).reshape(want_d.shape[0], -1)).all(axis=1)
    id_neq = got_ids != want_ids
    tie = np.zeros_like(id_neq)
    tie[:, 1:] |= want_d[:, 1:] == want_d[:, :-1]
    tie[:, :-1] |= want_d[:, :-1] == want_d[:, 1:]
    hard = (id_neq & ~tie).any(axis=1)
    soft = id_neq.any(axis=1) & ~hard
    return int((~same_d).sum()), int(hard.sum()), int(soft.sum())
