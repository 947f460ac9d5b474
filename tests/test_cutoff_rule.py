"""The cutoff carried from try to try (include/annb200.h, annb_cutoff_update) on the CPU.

Claim: when k*tries is a power of two (the merged row is sorted as a whole, alg.c:139), the
reference's merge — sort / kill adjacent duplicate ids / sort, then the first k slots
(alg.c:312, compute.cl:181-217) — returns the same k slots if, in every try after the first,
the entries of the per-try list that are farther than the running cutoff are replaced by
(n, +inf), wherever the surviving far entries end up behind the near ones.  The cutoff is the
k-th smallest DISTINCT distance value seen in the earlier lists.

The oracle's per-try lists and its literal merge are used here; the truncation is emulated in
numpy with a random subset of the far entries kept (the fp16 brackets of the screened kernel
let some of them through).  Data with many exact ties (small integer lattices, duplicated
points) exercises the network's tie behaviour, which is where the argument is delicate.
"""
import numpy as np
import pytest

from approximatenn_b200.api import srandom


def running_cutoff(run, new_key, k):
    """numpy model of cutoff_update_kernel: run = k smallest distinct values so far (inf padded)."""
    n = new_key.shape[0]
    out = np.full((n, k), np.inf, dtype=new_key.dtype)
    for x in range(n):
        vals = np.unique(np.concatenate([run[x][np.isfinite(run[x])], new_key[x][np.isfinite(new_key[x])]]))
        m = min(k, len(vals))
        out[x, :m] = vals[:m]
    return out, out[:, k - 1].copy()


def truncate(ids, key, cutoff, n, rng, keep_prob):
    """Entries at or below the cutoff stay where they are; of the others a random subset stays
    (order preserved, packed behind the near ones), the rest becomes (n, +inf)."""
    ids2 = np.full_like(ids, n)
    key2 = np.full_like(key, np.inf)
    for x in range(ids.shape[0]):
        near = key[x] <= cutoff[x]
        far_kept = (~near) & np.isfinite(key[x]) & (rng.random(key.shape[1]) < keep_prob)
        sel = near | far_kept
        m = int(sel.sum())
        ids2[x, :m] = ids[x][sel]
        key2[x, :m] = key[x][sel]
    return ids2, key2


def _datasets():
    def gauss(rng, n, d):
        return rng.standard_normal((n, d))

    def lattice(rng, n, d):                 # squared distances are small integers: ties everywhere
        return rng.integers(0, 7, (n, d)).astype(np.float64)

    def duplicates(rng, n, d):
        base = rng.standard_normal((n // 3, d))
        return base[rng.integers(0, n // 3, n)]

    return [("gauss", gauss), ("lattice", lattice), ("duplicates", duplicates)]


@pytest.mark.parametrize("name,make", _datasets())
@pytest.mark.parametrize("n,d,k,tries", [(1500, 16, 16, 4), (1200, 32, 8, 8), (900, 16, 4, 4)])
@pytest.mark.parametrize("keep_prob", [0.0, 0.4])
def test_truncated_lists_merge_to_the_same_rows(oracle_mod, name, make, n, d, k, tries, keep_prob):
    assert (k * tries) & (k * tries - 1) == 0 and k * tries >= 16
    import zlib
    rng = np.random.default_rng(zlib.crc32(f"{name}-{n}-{d}-{k}".encode()))
    pts = np.ascontiguousarray(make(rng, n, d), dtype=np.float32)
    b = oracle_mod.restatement(np.float32)
    srandom(1234)
    st = oracle_mod.Stages(b, pts, k, tries)
    full_ids, full_key, cut_ids, cut_key = [], [], [], []
    run = np.full((n, k), np.inf, dtype=np.float32)
    cutoff = np.full(n, np.inf, dtype=np.float32)
    dropped = 0
    for t in range(tries):
        ids, key = st.try_lists(t)
        full_ids.append(ids)
        full_key.append(key)
        ti, tk = truncate(ids, key, cutoff, n, rng, keep_prob)
        dropped += int((np.isfinite(key) & ~np.isfinite(tk)).sum())
        cut_ids.append(ti)
        cut_key.append(tk)
        # the kernel folds the list it produced (the truncated one) into the running values
        run, cutoff = running_cutoff(run, tk, k)
    st.close()
    rows_a = np.ascontiguousarray(np.concatenate(full_ids, axis=1))
    keys_a = np.ascontiguousarray(np.concatenate(full_key, axis=1))
    rows_b = np.ascontiguousarray(np.concatenate(cut_ids, axis=1))
    keys_b = np.ascontiguousarray(np.concatenate(cut_key, axis=1))
    oracle_mod.merge_rows(b, rows_a, keys_a)
    oracle_mod.merge_rows(b, rows_b, keys_b)
    if name == "gauss":
        assert dropped > 0                                       # the rule did remove entries
    assert np.array_equal(rows_a[:, :k], rows_b[:, :k])
    assert np.array_equal(keys_a[:, :k].view(np.uint32), keys_b[:, :k].view(np.uint32))


def test_running_cutoff_never_undercuts_the_final_kth_distance(oracle_mod):
    """cutoff after any prefix of the tries >= the k-th distance of the merged row."""
    rng = np.random.default_rng(5)
    n, d, k, tries = 1000, 16, 8, 4
    pts = np.ascontiguousarray(rng.integers(0, 4, (n, d)), dtype=np.float32)
    b = oracle_mod.restatement(np.float32)
    srandom(99)
    st = oracle_mod.Stages(b, pts, k, tries)
    lists = [st.try_lists(t) for t in range(tries)]
    st.close()
    rows = np.ascontiguousarray(np.concatenate([l[0] for l in lists], axis=1))
    keys = np.ascontiguousarray(np.concatenate([l[1] for l in lists], axis=1))
    oracle_mod.merge_rows(b, rows, keys)
    final_kth = keys[:, k - 1]
    run = np.full((n, k), np.inf, dtype=np.float32)
    for t in range(tries):
        run, cutoff = running_cutoff(run, lists[t][1], k)
        assert np.all(cutoff >= final_kth)
