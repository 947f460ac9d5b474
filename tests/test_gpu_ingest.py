"""The staged upload of pageable caller memory (csrc/ann_ingest.c: >= 8 MB goes through a ring
of pinned slots filled by host threads).  The same points are handed over three ways — plain
malloc()ed memory (staged path), page-locked memory, and pageable memory with the staging
switched off (one cudaMemcpyAsync) — and must give identical results; sampled rows are checked
against the oracle.  Sizes straddle the slot size (8 MB) and the thread count."""
import ctypes
import os

import numpy as np
import pytest

from conftest import same_bits

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,d", [(150_001, 64), (1_050_000, 16)])
def test_pageable_upload_is_bit_identical(oracle_mod, n, d):
    import torch
    from approximatenn_b200.api import gpu_backend, srandom
    k, tries = 16, 2
    rng = np.random.default_rng(n)
    pts = rng.standard_normal((n, d), dtype=np.float32)          # numpy = malloc()ed, pageable
    assert pts.nbytes >= 32 << 20
    gpu = gpu_backend(np.float32)
    staged = gpu.precomp(pts, k, tries, seed=77)
    pinned_t = torch.empty((n, d), dtype=torch.float32, pin_memory=True)
    pinned = pinned_t.numpy()
    pinned[:] = pts
    direct = gpu.precomp(pinned, k, tries, seed=77)
    os.environ["ANN_B200_STAGED_UPLOAD"] = "0"
    try:
        plain = gpu.precomp(pts, k, tries, seed=77)
    finally:
        del os.environ["ANN_B200_STAGED_UPLOAD"]
    for other in (direct, plain):
        assert np.array_equal(staged.ids, other.ids)
        assert same_bits(staged.dists, other.dists)
    sample = np.sort(rng.choice(n, size=48, replace=False))
    srandom(77)
    want_ids, want_d, _ = oracle_mod.sampled_rows(oracle_mod.restatement(np.float32), pts, k, tries, sample)
    assert np.array_equal(staged.ids[sample], want_ids)
    assert same_bits(staged.dists[sample], want_d)
