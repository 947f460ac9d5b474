"""Test infrastructure: CPU checkers for the approximateNN path.

ONLY tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  It loads
  * oracle/liboracle_{f32,f64}.so  — our CPU restatement (oracle/ann_oracle.c), and
  * oracle/_ref/libannref_{f32,f64}.so — the reference's own pure-C path compiled from
    /root/reference by oracle/Makefile (present wherever `make ref` ran).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from approximatenn_b200.api import Backend, SaveT

HERE = os.path.dirname(os.path.abspath(__file__))
_SUFFIX = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}


def build(quiet: bool = True) -> None:
    """Compile the restatement, and the reference itself when its checkout is present."""
    subprocess.run(["make", "-C", HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _need(path):
    if not os.path.exists(path):
        build()
    return path


def restatement(dtype) -> Backend:
    """Our CPU restatement (`kind: port`)."""
    sfx = _SUFFIX[np.dtype(dtype)]
    b = Backend(_need(os.path.join(HERE, f"liboracle_{sfx}.so")), dtype,
                "precomp_oracle", "query_oracle", "free_save_oracle", mode=ctypes.RTLD_LOCAL)
    _declare_stage_api(b)
    return b


def reference_available() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libannref_f32.so"))


def reference(dtype) -> Backend:
    """The reference's own precomp_cpu/query_cpu (`kind: reference`)."""
    sfx = _SUFFIX[np.dtype(dtype)]
    return Backend(os.path.join(HERE, "_ref", f"libannref_{sfx}.so"), dtype,
                   "precomp_cpu", "query_cpu", None, mode=ctypes.RTLD_LOCAL)


# ---- stage-level access to the restatement (hashes, tables, per-try lists) ---------------

def _declare_stage_api(b: Backend) -> None:
    L, vp, sz = b.lib, ctypes.c_void_p, ctypes.c_size_t
    L.orc_params.argtypes = [sz, sz, sz, ctypes.POINTER(sz), ctypes.POINTER(sz)]
    L.orc_params.restype = None
    L.orc_means.argtypes = [sz, sz, vp, vp]
    L.orc_means.restype = None
    L.orc_prepare.argtypes = [sz, sz, sz, vp, ctypes.c_int, sz, sz, sz, sz]
    L.orc_prepare.restype = vp
    L.orc_release.argtypes = [vp, ctypes.c_int]
    L.orc_release.restype = None
    for name in ("orc_state_d_short", "orc_state_d_max"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = sz
    L.orc_state_tmax.argtypes = [vp, ctypes.c_int]
    L.orc_state_tmax.restype = sz
    for name in ("orc_state_hash", "orc_state_table"):
        getattr(L, name).argtypes = [vp, ctypes.c_int]
        getattr(L, name).restype = vp
    L.orc_state_means.argtypes = [vp]
    L.orc_state_means.restype = vp
    L.orc_state_projection.argtypes = [vp, ctypes.c_int, vp]
    L.orc_state_projection.restype = None
    L.orc_try_lists.argtypes = [vp, vp, ctypes.c_int, vp, sz, vp, vp]
    L.orc_try_lists.restype = None
    L.orc_merge_rows.argtypes = [vp, vp, sz, sz]
    L.orc_merge_rows.restype = None
    L.orc_supercharge_rows.argtypes = [sz, sz, sz, vp, vp, sz, sz, vp, vp, sz, vp, sz, ctypes.c_int, vp, vp]
    L.orc_supercharge_rows.restype = None
    L.orc_sampled_cost.argtypes = [sz, sz, sz, vp, ctypes.c_int, sz, sz, sz, sz, vp, sz,
                                   ctypes.POINTER(ctypes.c_double)]
    L.orc_sampled_cost.restype = sz
    L.orc_sampled_rows.argtypes = [sz, sz, sz, vp, ctypes.c_int, sz, sz, sz, sz, vp, sz,
                                   ctypes.POINTER(ctypes.c_double), vp, vp]
    L.orc_sampled_rows.restype = sz


def params(b: Backend, n, k, d):
    ds, dm = ctypes.c_size_t(), ctypes.c_size_t()
    b.lib.orc_params(n, k, d, ctypes.byref(ds), ctypes.byref(dm))
    return ds.value, dm.value


def means(b: Backend, points):
    pts = np.ascontiguousarray(points, dtype=b.dtype)
    out = np.empty(pts.shape[1], dtype=b.dtype)
    b.lib.orc_means(pts.shape[0], pts.shape[1], pts.ctypes.data, out.ctypes.data)
    return out


class Stages:
    """Everything before the distance work, for stage-by-stage parity tests."""

    def __init__(self, b: Backend, points, k, tries, rots_before=6, rot_len_before=1,
                 rots_after=1, rot_len_after=1):
        from approximatenn_b200.api import _view
        self.b = b
        self.points = np.ascontiguousarray(points, dtype=b.dtype)
        self.n, self.d = self.points.shape
        self.k, self.tries = k, tries
        self.h = b.lib.orc_prepare(self.n, k, self.d, self.points.ctypes.data, tries,
                                   rots_before, rot_len_before, rots_after, rot_len_after)
        self.d_short = b.lib.orc_state_d_short(self.h)
        self.d_max = b.lib.orc_state_d_max(self.h)
        self.means = _view(b.lib.orc_state_means(self.h), (self.d,), b.dtype).copy()
        self.tmax = [b.lib.orc_state_tmax(self.h, t) for t in range(tries)]
        self.hash = [_view(b.lib.orc_state_hash(self.h, t), (self.n,), np.uint64).copy()
                     for t in range(tries)]
        self._view = _view

    def table(self, t):
        return self._view(self.b.lib.orc_state_table(self.h, t),
                          (1 << self.d_short, self.tmax[t]), np.uint64).copy()

    def projection(self, t):
        out = np.empty((self.d_short, self.d), dtype=self.b.dtype)
        self.b.lib.orc_state_projection(self.h, t, out.ctypes.data)
        return out

    def try_lists(self, t, rows=None):
        m = self.n if rows is None else len(rows)
        ids = np.empty((m, self.k), dtype=np.uint64)
        key = np.empty((m, self.k), dtype=self.b.dtype)
        rp = None
        if rows is not None:
            rows = np.ascontiguousarray(rows, dtype=np.uint64)
            rp = rows.ctypes.data
        self.b.lib.orc_try_lists(self.h, self.points.ctypes.data, t, rp, m,
                                 ids.ctypes.data, key.ctypes.data)
        return ids, key

    def close(self):
        if self.h:
            self.b.lib.orc_release(self.h, 0)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sampled_cost(b: Backend, points, k, tries, sample, rots_before=6, rot_len_before=1,
                 rots_after=1, rot_len_after=1):
    """Bounded CPU-baseline sample at full problem size; see orc_sampled_cost()."""
    pts = np.ascontiguousarray(points, dtype=b.dtype)
    sample = np.ascontiguousarray(sample, dtype=np.uint64)
    secs = (ctypes.c_double * 3)()
    rows = b.lib.orc_sampled_cost(pts.shape[0], k, pts.shape[1], pts.ctypes.data, tries,
                                  rots_before, rot_len_before, rots_after, rot_len_after,
                                  sample.ctypes.data, len(sample), secs)
    return {"prepare_s": secs[0], "row_s": secs[1], "supercharge_s": secs[2], "rows": rows}


def sampled_rows(b: Backend, points, k, tries, sample, rots_before=6, rot_len_before=1,
                 rots_after=1, rot_len_after=1):
    """Exact final rows (after supercharging) of the `sample` points of a full-size problem,
    plus the timings of sampled_cost().  Seed libc random() first (api.srandom)."""
    pts = np.ascontiguousarray(points, dtype=b.dtype)
    sample = np.ascontiguousarray(sample, dtype=np.uint64)
    secs = (ctypes.c_double * 3)()
    ids = np.empty((len(sample), k), dtype=np.uint64)
    key = np.empty((len(sample), k), dtype=b.dtype)
    rows = b.lib.orc_sampled_rows(pts.shape[0], k, pts.shape[1], pts.ctypes.data, tries,
                                  rots_before, rot_len_before, rots_after, rot_len_after,
                                  sample.ctypes.data, len(sample), secs, ids.ctypes.data, key.ctypes.data)
    return ids, key, {"prepare_s": secs[0], "row_s": secs[1], "supercharge_s": secs[2], "rows": rows}


def merge_rows(b: Backend, ids, key):
    """In place: the reference's sort / kill duplicates / sort on every row (alg.c:312)."""
    assert ids.dtype == np.uint64 and ids.flags.c_contiguous and key.flags.c_contiguous
    b.lib.orc_merge_rows(ids.ctypes.data, key.ctypes.data, ids.shape[0], ids.shape[1])


def supercharge_rows(b: Backend, points, queries, r0, r1, own_ids, own_key, graph, k, exclude_self=True):
    """alg.c:313-335 for query rows [r0, r1); own_* are indexed by x - r0, graph by point id."""
    pts = np.ascontiguousarray(points, dtype=b.dtype)
    qs = pts if queries is points else np.ascontiguousarray(queries, dtype=b.dtype)
    own_ids = np.ascontiguousarray(own_ids, dtype=np.uint64)
    own_key = np.ascontiguousarray(own_key, dtype=b.dtype)
    graph = np.ascontiguousarray(graph, dtype=np.uint64)
    out_i = np.empty((r1 - r0, k), dtype=np.uint64)
    out_k = np.empty((r1 - r0, k), dtype=b.dtype)
    b.lib.orc_supercharge_rows(pts.shape[0], k, pts.shape[1], qs.ctypes.data, pts.ctypes.data, r0, r1,
                               own_ids.ctypes.data, own_key.ctypes.data, own_ids.shape[1],
                               graph.ctypes.data, graph.shape[1], 1 if exclude_self else 0,
                               out_i.ctypes.data, out_k.ctypes.data)
    return out_i, out_k
