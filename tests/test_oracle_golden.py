"""The CPU restatement (oracle/ann_oracle.c) against the reference's golden vectors.

The golden files are outputs of the reference's own precomp_cpu/query_cpu on fixed seeds
(tests/golden/make_golden.py).  The bar is bit-exact: ids, squared distances and every
save_t field, as compare_results.c:152-171 diffs them.
"""
import numpy as np
import pytest

from conftest import assert_matches_golden, golden_names, load_golden


@pytest.mark.parametrize("name", golden_names())
def test_restatement_reproduces_reference_golden(oracle_mod, name):
    g = load_golden(name)
    orc = oracle_mod.restatement(g["dtype"])
    res = orc.precomp(g["points"], g["k"], g["tries"], *g["rot"], want_save=True, seed=g["seed"])
    qres = orc.query(res.save, g["points"], g["y"])
    assert_matches_golden(g, res, qres)
    res.save.free()


def test_restatement_without_save_or_dists(oracle_mod):
    g = load_golden("onetry_f32")
    orc = oracle_mod.restatement(g["dtype"])
    res = orc.precomp(g["points"], g["k"], g["tries"], *g["rot"], want_save=False,
                      want_dists=False, seed=g["seed"])
    assert res.dists is None and res.save is None
    assert np.array_equal(res.ids, g["ids"].astype(np.uint64))


def test_params_follow_reference_rule(oracle_mod):
    orc = oracle_mod.restatement(np.float32)
    # (n, k, d) -> (d_short, d_max)
    assert oracle_mod.params(orc, 16384, 10, 16) == (11, 16)
    assert oracle_mod.params(orc, 1_000_000, 16, 64) == (16, 64)
    assert oracle_mod.params(orc, 10_000_000, 16, 64) == (20, 64)
    assert oracle_mod.params(orc, 100_000_000, 32, 32) == (22, 32)
    assert oracle_mod.params(orc, 1000, 10, 80) == (7, 128)
    orc64 = oracle_mod.restatement(np.float64)
    assert oracle_mod.params(orc64, 65536, 16, 32) == (12, 32)
    # d_short is clipped to d_max (alg.c:356-357)
    assert oracle_mod.params(orc, 1 << 20, 1, 9) == (16, 16)


def test_stage_views_are_consistent(oracle_mod):
    g = load_golden("ragged_f32")
    orc = oracle_mod.restatement(g["dtype"])
    from approximatenn_b200.api import srandom
    srandom(g["seed"])
    st = oracle_mod.Stages(orc, g["points"], g["k"], g["tries"], *g["rot"])
    assert st.d_short == g["d_short"]
    assert st.tmax == g["par_maxes"].tolist()
    for t in range(g["tries"]):
        tab = st.table(t)
        assert np.array_equal(tab, g[f"which_par_{t}"].astype(np.uint64))
        # every point sits in the row named by its hash, rows are descending, padded with n
        for b in np.unique(st.hash[t])[:16]:
            members = np.flatnonzero(st.hash[t] == b)[::-1]
            assert np.array_equal(tab[b, :len(members)], members.astype(np.uint64))
            assert np.all(tab[b, len(members):] == g["n"])
        assert np.array_equal(st.projection(t).view(np.uint8), g["bases"][t].view(np.uint8))
    st.close()
