"""Generates tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref = /root/reference's
precomp_cpu/query_cpu compiled by oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

The reference has no golden vectors of its own (every program seeds from
the clock), so these files are its outputs on fixed seeds.  Each file stores the inputs,
the call parameters and every output, so nothing needs /root/reference at test time.
Protocol (compare_results.c:123-130): srandom(seed) immediately before precomp.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

CASES = [
    # name, dtype, n, d, k, tries, (rots_b, len_b, rots_a, len_a), ycnt, data seed
    ("defaults_f32", np.float32, 1000, 80, 10, 10, (6, 1, 1, 1), 50, 11),
    ("defaults_f64", np.float64, 1000, 80, 10, 10, (6, 1, 1, 1), 50, 12),
    ("cfg2shape_f64", np.float64, 2048, 32, 16, 8, (6, 1, 1, 1), 64, 13),
    ("cfg3shape_f32", np.float32, 2048, 64, 16, 8, (6, 1, 1, 1), 64, 14),
    ("cfg1shape_f32", np.float32, 2048, 16, 10, 10, (6, 1, 1, 1), 64, 15),
    ("ragged_f32", np.float32, 1537, 20, 10, 7, (3, 4, 2, 2), 33, 16),
    ("ragged_f64", np.float64, 999, 17, 5, 3, (2, 3, 1, 2), 17, 17),
    ("onetry_f32", np.float32, 300, 64, 10, 1, (6, 1, 1, 1), 9, 18),
    ("k32_f32", np.float32, 3000, 32, 32, 4, (6, 1, 1, 1), 16, 19),
]


def main():
    if not oracle.reference_available():
        raise SystemExit("oracle/_ref is missing: run `make -C oracle ref` where /root/reference exists")
    here = os.path.dirname(os.path.abspath(__file__))
    for name, dtype, n, d, k, tries, rot, ycnt, seed in CASES:
        rng = np.random.default_rng(seed)
        pts = rng.standard_normal((n, d)).astype(dtype)
        y = rng.standard_normal((ycnt, d)).astype(dtype)
        ref = oracle.reference(dtype)
        r = ref.precomp(pts, k, tries, *rot, want_save=True, seed=seed + 1000)
        q = ref.query(r.save, pts, y)
        s = r.save
        out = dict(points=pts, y=y, n=n, d=d, k=k, tries=tries, rot=np.array(rot), seed=seed + 1000,
                   ids=r.ids.astype(np.uint32), dists=r.dists, d_short=s.d_short,
                   par_maxes=s.par_maxes.copy(), graph=s.graph.astype(np.uint32),
                   row_means=s.row_means.copy(), bases=s.bases.copy(),
                   q_ids=q.ids.astype(np.uint32), q_dists=q.dists)
        for t in range(tries):
            out[f"which_par_{t}"] = s.which_par(t).astype(np.uint32)
        np.savez_compressed(os.path.join(here, name + ".npz"), **out)
        s.free()
        print("wrote", name, "d_short", out["d_short"], "tmax", out["par_maxes"].tolist())


if __name__ == "__main__":
    main()
