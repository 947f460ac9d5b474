"""The C-ABI shared libraries load (no GPU needed) and export every symbol that include/*.h
declares; the pure host functions behave like the reference's parameter rules."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECL = re.compile(r"^\s*(?:extern\s+)?(?:const\s+)?(?:unsigned\s+long|unsigned\s+long\s+long|size_t|void|int|ftype)\s*\*?\s*"
                  r"(\w+)\s*\(", re.M)
NOT_OURS = {"precomp_cpu", "query_cpu"}          # the reference's CPU path: oracle/_ref only


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names |= set(DECL.findall(text))
    return sorted(names - NOT_OURS)


@pytest.fixture(scope="module", autouse=True)
def built():
    import __graft_entry__
    __graft_entry__.build()


@pytest.mark.parametrize("suffix", ["f32", "f64"])
def test_library_exports_every_declared_symbol(suffix):
    lib = ctypes.CDLL(os.path.join(ROOT, "approximatenn_b200", f"libann_b200_{suffix}.so"))
    names = declared_symbols()
    assert {"precomp_gpu", "query_gpu", "gpu_init", "gpu_cleanup", "register_cleanup", "precomp", "query",
            "free_save", "annb_leaf_topk", "annb_supercharge", "annb200_dist_init"} <= set(names)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/ but not exported: {missing}"


def test_host_parameter_rule_matches_reference():
    from approximatenn_b200.api import gpu_backend
    for dtype in (np.float32, np.float64):
        lib = gpu_backend(dtype).lib
        lib.annh_params.argtypes = [ctypes.c_size_t] * 3 + [ctypes.POINTER(ctypes.c_size_t)] * 2
        lib.annh_params.restype = None

        def params(n, k, d):
            a, b = ctypes.c_size_t(), ctypes.c_size_t()
            lib.annh_params(n, k, d, ctypes.byref(a), ctypes.byref(b))
            return a.value, b.value
        assert params(16384, 10, 16) == (11, 16)
        assert params(65536, 16, 32) == (12, 32)
        assert params(1_000_000, 16, 64) == (16, 64)
        assert params(10_000_000, 16, 64) == (20, 64)
        assert params(100_000_000, 32, 32) == (22, 32)
        assert params(1000, 10, 80) == (7, 128)
        assert params(1 << 20, 1, 9) == (16, 16)


def test_oracle_and_product_agree_on_params(oracle_mod):
    from approximatenn_b200.api import gpu_backend
    rng = np.random.default_rng(0)
    for dtype in (np.float32, np.float64):
        lib = gpu_backend(dtype).lib
        lib.annh_params.argtypes = [ctypes.c_size_t] * 3 + [ctypes.POINTER(ctypes.c_size_t)] * 2
        orc = oracle_mod.restatement(dtype)
        for _ in range(200):
            n, k, d = int(rng.integers(2, 10**7)), int(rng.integers(1, 64)), int(rng.integers(1, 300))
            a, b = ctypes.c_size_t(), ctypes.c_size_t()
            lib.annh_params(n, k, d, ctypes.byref(a), ctypes.byref(b))
            if n > k:
                assert (a.value, b.value) == oracle_mod.params(orc, n, k, d)


def test_save_t_layout_matches_header():
    from approximatenn_b200.api import SaveT
    # int tries; size_t n,k,d_short,d_long; size_t **which_par, *par_maxes, *graph; ftype *row_means, *bases
    assert ctypes.sizeof(SaveT) == 8 + 4 * 8 + 3 * 8 + 2 * 8
    assert SaveT.n.offset == 8 and SaveT.which_par.offset == 40 and SaveT.bases.offset == 72


def test_screened_path_coverage_and_scratch_plan():
    """Host-only entry points of the S3 stage: which shapes take the screened path, and that the
    scratch plan grows by exactly the fp16 copy, the norms and the hand-over list when it does."""
    from approximatenn_b200.api import gpu_backend
    f32, f64 = gpu_backend(np.float32).lib, gpu_backend(np.float64).lib
    for lib in (f32, f64):
        lib.annb_screen_applies.argtypes = [ctypes.c_size_t] * 3
        lib.annb_screen_applies.restype = ctypes.c_int
        lib.annb_leaf_scratch_bytes.argtypes = [ctypes.c_size_t] * 4
        lib.annb_leaf_scratch_bytes.restype = ctypes.c_size_t
        lib.annb_leaf_screen_mode.argtypes = [ctypes.c_int]
    try:
        f32.annb_leaf_screen_mode(1)
        assert [f32.annb_screen_applies(d, 16, 16) for d in (16, 32, 64, 128, 80)] == [1, 1, 1, 0, 0]
        assert f32.annb_screen_applies(64, 16, 17) == 0            # k > 16: the tiled kernel
        assert f32.annb_screen_applies(64, 0, 16) == 0             # a single bucket
        assert f64.annb_screen_applies(64, 16, 16) == 0            # float only
        f32.annb_leaf_screen_mode(0)
        assert f32.annb_screen_applies(64, 16, 16) == 0
    finally:
        f32.annb_leaf_screen_mode(1)
    n, d, ds = 1_000_000, 64, 16
    plain = f32.annb_leaf_scratch_bytes(n, 128, ds, 16)            # shape the screen does not cover
    extra = f32.annb_leaf_scratch_bytes(n, d, ds, 16) - plain
    assert n * d * 2 + n * 8 + 4 * (1 << ds) <= extra <= n * d * 2 + n * 8 + 4 * (1 << ds) + 2048
    assert f64.annb_leaf_scratch_bytes(n, d, ds, 16) == plain
