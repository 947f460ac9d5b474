"""CPU prototype of a screened S5 (supercharge): fp16 brackets decide which of the k*k
neighbours-of-neighbours need an exact distance.  Checks, against the oracle's supercharge rows,
that (1) dropping every candidate with lo > tau0 (tau0 = the row's k-th own distance) never
changes a row, and (2) the prefix-corner rule can be decided from brackets, with the undecidable
rows sent to the literal kernel.  Run:  python scratch/next_round/s5_screen_prototype.py
"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import oracle
from test_screen_bounds import tree_sum_f32


def prepare(x):
    x = x.astype(np.float32)
    n, d = x.shape
    mean = x.mean(axis=0, dtype=np.float64).astype(np.float32)
    c = x - mean
    cmax = float(np.abs(c).max())
    scale = np.float32(2.0 ** (2 - int(np.floor(np.log2(cmax)))))
    c = c * scale
    c16 = c.astype(np.float16).astype(np.float32)
    n2 = (c16 * c16).sum(axis=1, dtype=np.float32)
    s = np.float32(0.03226) * (np.sqrt((c * c).sum(axis=1, dtype=np.float32)) * np.float32(1 + 2.0 ** -12)
                               + np.float32(np.sqrt(d) * 2.0 ** -14))
    return c16, n2, s, float(scale) ** 2


def run(n=4096, d=64, k=16, tries=6, seed=3, offset=0.0):
    rng = np.random.default_rng(seed)
    pts = (rng.standard_normal((n, d)) + offset).astype(np.float32)
    b = oracle.restatement(np.float32)
    res = b.precomp(pts, k, tries, 6, 1, 1, 1, want_save=False, seed=seed)
    graph = res.ids.astype(np.uint64)          # a valid merged graph: sorted exact-tree distances
    own_d = res.dists.astype(np.float32)
    want_i, want_d = oracle.supercharge_rows(b, pts, pts, 0, n, graph, own_d, graph, k)
    c16, n2, s, scale2 = prepare(pts)
    wide = k * (k + 1)
    P2 = 1 << (wide.bit_length() - 1)
    ncand = P2 - k
    stats = dict(rows=0, literal_tie=0, literal_corner=0, exact=0, cand=0, mismatch=0, corner_fired=0)
    for x in range(n):
        own = graph[x]
        od = own_d[x]
        cj, cz = np.divmod(np.arange(ncand), k)
        oj = own[cj]
        cid = np.where(oj < n, graph[np.minimum(oj, n - 1), cz], n)
        any_inf = bool(np.isinf(od).any() or (cid >= n).any() or (cid == x).any())
        live = (cid < n) & (cid != x)
        uniq = cid[live].astype(np.int64)
        # brackets (scaled units) and the exact tree distance (original units)
        dot = (c16[uniq] * c16[x]).sum(axis=1, dtype=np.float32)
        dprime = (n2[x] + n2[uniq]) - np.float32(2) * dot
        t = s[x] + s[uniq]
        lo = (dprime - t * t).astype(np.float64) / scale2
        hi = (dprime + t * t).astype(np.float64) / scale2
        diff = pts[uniq] - pts[x]
        exact = tree_sum_f32(diff * diff)
        assert np.all(lo <= exact.astype(np.float64) * (1 + 1e-12)) and np.all(exact.astype(np.float64) <= hi * (1 + 1e-12))
        tau0 = od[k - 1]
        surv = lo <= float(tau0)
        stats["cand"] += len(uniq); stats["exact"] += int(surv.sum()); stats["rows"] += 1
        # the fast kernel's list logic on the survivors only
        best = [(float(od[i]), int(own[i])) for i in range(k)]
        tie = any(best[i][0] == best[i + 1][0] and np.isfinite(best[i][0]) for i in range(k - 1))
        tau = best[k - 1][0]
        for vn, idn in zip(exact[surv], uniq[surv]):
            vn = float(vn); idn = int(idn)
            if vn <= tau and not any(i == idn and np.isfinite(v) for v, i in best):
                if vn < tau:
                    if any(v == vn for v, _ in best):
                        tie = True
                    pos = sum(1 for v, _ in best if v <= vn)
                    best.insert(pos, (vn, idn)); best.pop()
                    tau = best[k - 1][0]
                else:
                    tie = True
        # prefix corner: slot P2 of the row; it kills the prefix's largest entry if the ids agree
        corner_literal = False
        if P2 < wide and not any_inf:
            c = P2 - k
            j, z = divmod(c, k)
            corner = int(graph[own[j]][z]) if own[j] < n else n
            in_own = corner in [int(i) for i in own]
            in_cand = bool((uniq == corner).any())
            if in_own or in_cand:
                hi_c = max([float(od[i]) for i in range(k) if int(own[i]) == corner] +
                           [float(h) for h, u in zip(hi, uniq) if u == corner])
                others_lo = [float(l) for l, u in zip(lo, uniq) if u != corner] + \
                            [float(od[i]) for i in range(k) if int(own[i]) != corner]
                provably_not_max = max(others_lo) > hi_c
                # truth
                allv = [(float(od[i]), int(own[i])) for i in range(k)] + [(float(e), int(u)) for e, u in zip(exact, uniq)]
                mx = max(v for v, _ in allv)
                truth_is_max = any(v == mx and i == corner for v, i in allv)
                if provably_not_max:
                    assert not truth_is_max
                else:
                    corner_literal = True
                    stats["corner_fired"] += int(truth_is_max)
        if tie:
            stats["literal_tie"] += 1
            continue
        if corner_literal:
            stats["literal_corner"] += 1
            continue
        got_i = np.array([i for _, i in best], dtype=np.uint64)
        got_d = np.array([v for v, _ in best], dtype=np.float32)
        if not (np.array_equal(got_i, want_i[x]) and np.array_equal(got_d.view(np.uint32), want_d[x].view(np.uint32))):
            stats["mismatch"] += 1
    return stats


if __name__ == "__main__":
    for kw in (dict(), dict(offset=100.0), dict(n=3000, d=32, k=10, tries=5), dict(n=2000, d=16, k=5, tries=8)):
        st = run(**kw)
        print(kw, st, "exact fraction %.3f" % (st["exact"] / max(1, st["cand"])))
        assert st["mismatch"] == 0
