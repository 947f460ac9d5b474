"""ctypes mirror of the approximateNN C API (include/ann.h, include/algg.h).

This is the host-side binding a Python caller uses; it mirrors the reference's public
interface one to one (/root/reference/ann.h:8-65): ``precomp``, ``query``, ``free_save``
and the ``save_t`` record, with the same argument meaning and the same "errors are fatal"
behaviour (the C library prints to stderr and exits, /root/reference/gpu_comp.c:15-19).

`Backend` is deliberately generic over the shared library and the symbol names, so that
the same class binds
  * the product:  libann_b200_f32.so / libann_b200_f64.so  (precomp_gpu, query_gpu)
  * any other library with the same two entry points (the test suite binds its CPU checkers
    through this class; nothing in this package loads or knows them).
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

_libc = ctypes.CDLL(None)
_libc.free.argtypes = [ctypes.c_void_p]
_libc.free.restype = None
_libc.srandom.argtypes = [ctypes.c_uint]
_libc.srandom.restype = None

c_size_p = ctypes.POINTER(ctypes.c_size_t)


def srandom(seed: int) -> None:
    """Seed libc random(): the transforms are drawn from it (compare_results.c:123-130)."""
    _libc.srandom(ctypes.c_uint(seed & 0xFFFFFFFF))


class SaveT(ctypes.Structure):
    """save_t, field for field (include/ann.h; /root/reference/ann.h:8-12)."""

    _fields_ = [
        ("tries", ctypes.c_int),
        ("n", ctypes.c_size_t),
        ("k", ctypes.c_size_t),
        ("d_short", ctypes.c_size_t),
        ("d_long", ctypes.c_size_t),
        ("which_par", ctypes.POINTER(c_size_p)),
        ("par_maxes", c_size_p),
        ("graph", c_size_p),
        ("row_means", ctypes.c_void_p),
        ("bases", ctypes.c_void_p),
    ]


def _view(ptr, shape, dtype):
    """numpy view (no copy) of C memory at `ptr`."""
    count = int(np.prod(shape))
    if count == 0:
        return np.zeros(shape, dtype=dtype)
    addr = ctypes.cast(ptr, ctypes.c_void_p).value
    buf = (ctypes.c_char * (count * np.dtype(dtype).itemsize)).from_address(addr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def _take(ptr, shape, dtype):
    """Copy a malloc()ed result array into numpy and free() it (the caller owns it, ann.h)."""
    out = _view(ptr, shape, dtype).copy()
    _libc.free(ctypes.cast(ptr, ctypes.c_void_p))
    return out


class Save:
    """Owner of one save_t filled by precomp(); arrays are views into the C allocations."""

    def __init__(self, backend: "Backend"):
        self.backend = backend
        self.c = SaveT()
        self._live = False

    # views -----------------------------------------------------------------------------
    @property
    def tries(self): return int(self.c.tries)
    @property
    def n(self): return int(self.c.n)
    @property
    def k(self): return int(self.c.k)
    @property
    def d_short(self): return int(self.c.d_short)
    @property
    def d_long(self): return int(self.c.d_long)
    @property
    def par_maxes(self): return _view(self.c.par_maxes, (self.tries,), np.uint64)
    @property
    def graph(self): return _view(self.c.graph, (self.n, self.k), np.uint64)
    @property
    def row_means(self): return _view(self.c.row_means, (self.d_long,), self.backend.dtype)
    @property
    def bases(self):
        return _view(self.c.bases, (self.tries, self.d_short, self.d_long), self.backend.dtype)

    def which_par(self, t: int):
        width = int(self.c.par_maxes[t])
        return _view(self.c.which_par[t], (1 << self.d_short, width), np.uint64)

    def free(self):
        if self._live:
            self.backend._free_save(ctypes.byref(self.c))
            self._live = False

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


@dataclass
class Result:
    ids: np.ndarray                 # [rows][k] uint64, ascending squared distance
    dists: Optional[np.ndarray]     # [rows][k] squared distances, or None
    save: Optional[Save] = None


class Backend:
    """One shared library implementing precomp/query for one element type."""

    def __init__(self, path: str, dtype, precomp_sym: str, query_sym: str,
                 free_save_sym: Optional[str] = None, mode: int = ctypes.RTLD_LOCAL):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} is missing: build it first (python -c 'import __graft_entry__ as g; g.build()')")
        self.path = path
        self.dtype = np.dtype(dtype)
        self.lib = ctypes.CDLL(path, mode=mode)
        fp = ctypes.c_void_p
        self._precomp = getattr(self.lib, precomp_sym)
        self._precomp.restype = ctypes.c_void_p
        self._precomp.argtypes = [ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, fp,
                                  ctypes.c_int, ctypes.c_size_t, ctypes.c_size_t,
                                  ctypes.c_size_t, ctypes.c_size_t,
                                  ctypes.POINTER(SaveT), ctypes.POINTER(ctypes.c_void_p)]
        self._query = getattr(self.lib, query_sym)
        self._query.restype = ctypes.c_void_p
        self._query.argtypes = [ctypes.POINTER(SaveT), fp, ctypes.c_size_t, fp,
                                ctypes.POINTER(ctypes.c_void_p)]
        if free_save_sym is not None:
            self._free_save = getattr(self.lib, free_save_sym)
            self._free_save.restype = None
            self._free_save.argtypes = [ctypes.POINTER(SaveT)]
        else:
            self._free_save = self._free_save_py

    @staticmethod
    def _free_save_py(save_ref):
        """free_save as the reference's ann.c:25-34 does it (for libraries without one)."""
        s = save_ref._obj
        for t in range(s.tries):
            _libc.free(ctypes.cast(s.which_par[t], ctypes.c_void_p))
        for p in (s.which_par, s.par_maxes, s.graph):
            _libc.free(ctypes.cast(p, ctypes.c_void_p))
        _libc.free(s.row_means)
        _libc.free(s.bases)

    def _as_points(self, a, cols=None):
        a = np.ascontiguousarray(a, dtype=self.dtype)
        if a.ndim != 2 or (cols is not None and a.shape[1] != cols):
            raise ValueError("expected a row-major [rows][d] array")
        return a

    def precomp_raw(self, n, k, d, points_ptr, tries, rots_before, rot_len_before, rots_after,
                    rot_len_after, save_ref, dists_ref):
        """The bare C call (host pointers in, malloc()ed pointer out); bench.py times this."""
        return self._precomp(n, k, d, points_ptr, tries, rots_before, rot_len_before,
                             rots_after, rot_len_after, save_ref, dists_ref)

    def precomp(self, points, k, tries=10, rots_before=6, rot_len_before=1, rots_after=1,
                rot_len_after=1, want_save=False, want_dists=True, seed=None) -> Result:
        """ann.h precomp(); defaults are the reference programs' (time_results.c:16-17)."""
        pts = self._as_points(points)
        n, d = pts.shape
        save = Save(self) if want_save else None
        dptr = ctypes.c_void_p()
        if seed is not None:
            srandom(seed)
        ids = self._precomp(n, k, d, pts.ctypes.data, tries, rots_before, rot_len_before,
                            rots_after, rot_len_after,
                            ctypes.byref(save.c) if save else None,
                            ctypes.byref(dptr) if want_dists else None)
        if save:
            save._live = True
        return Result(_take(ids, (n, k), np.uint64),
                      _take(dptr, (n, k), self.dtype) if want_dists else None, save)

    def query(self, save: Save, points, y, want_dists=True) -> Result:
        """ann.h query(): neighbours of the rows of y among `points` (the indexed set)."""
        pts = self._as_points(points, save.d_long)
        ys = pts if y is points else self._as_points(y, save.d_long)
        dptr = ctypes.c_void_p()
        ids = self._query(ctypes.byref(save.c), pts.ctypes.data, ys.shape[0], ys.ctypes.data,
                          ctypes.byref(dptr) if want_dists else None)
        return Result(_take(ids, (ys.shape[0], save.k), np.uint64),
                      _take(dptr, (ys.shape[0], save.k), self.dtype) if want_dists else None)


# ---- the product libraries ---------------------------------------------------------------

_PKG = os.path.dirname(os.path.abspath(__file__))
_SUFFIX = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}
_loaded = {}


class StageTimes(ctypes.Structure):
    _fields_ = [("ms", ctypes.c_float * 9)]


STAGE_NAMES = ("upload", "means", "hash", "buckets", "leaf", "exchange", "merge", "supercharge",
               "first_to_last_event")


def gpu_backend(dtype) -> Backend:
    """libann_b200_{f32,f64}.so: precomp_gpu / query_gpu on the current CUDA device.

    There is no fallback: a missing library raises here, a missing GPU makes the C library
    print an error and exit (the reference's behaviour, gpu_comp.c:15-19)."""
    dt = np.dtype(dtype)
    if dt not in _loaded:
        path = os.path.join(_PKG, f"libann_b200_{_SUFFIX[dt]}.so")
        # experiments only (tools/variants.py): a side build of the same library
        path = os.environ.get(f"ANN_B200_LIB_{_SUFFIX[dt].upper()}") or path
        b = Backend(path, dt, "precomp_gpu", "query_gpu", "free_save")
        L = b.lib
        L.gpu_init.restype = None
        L.gpu_cleanup.restype = None
        L.annb_launch_count.argtypes = [ctypes.c_int]
        L.annb_launch_count.restype = ctypes.c_ulong
        L.annh_last_times.restype = ctypes.POINTER(StageTimes)
        L.annh_set_timing.argtypes = [ctypes.c_int]
        L.annh_set_timing.restype = None
        _loaded[dt] = b
    return _loaded[dt]


def save_to_file(b: Backend, save: Save, path: str) -> None:
    """include/annb200_io.h ann_save_write()."""
    b.lib.ann_save_write.argtypes = [ctypes.POINTER(SaveT), ctypes.c_char_p]
    b.lib.ann_save_write.restype = ctypes.c_int
    if b.lib.ann_save_write(ctypes.byref(save.c), os.fsencode(path)) != 0:
        raise OSError(f"could not write {path}")


def save_from_file(b: Backend, path: str) -> Save:
    """include/annb200_io.h ann_save_read(); the result is released with free_save as usual."""
    b.lib.ann_save_read.argtypes = [ctypes.POINTER(SaveT), ctypes.c_char_p]
    b.lib.ann_save_read.restype = ctypes.c_int
    s = Save(b)
    if b.lib.ann_save_read(ctypes.byref(s.c), os.fsencode(path)) != 0:
        raise OSError(f"could not read {path} (missing, truncated, or written by the other ftype build)")
    s._live = True
    return s


def stage_times(b: Backend) -> dict:
    t = b.lib.annh_last_times().contents
    return {name: float(t.ms[i]) for i, name in enumerate(STAGE_NAMES)}
