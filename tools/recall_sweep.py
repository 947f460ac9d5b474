#!/usr/bin/env python
"""Recall-vs-tries sweep (BASELINE config 5 asks for one): for each number of tries, run
precomp_gpu and compare with exact brute force on a sample of points.

    python tools/recall_sweep.py --n 1000000 --d 64 --k 16 --tries 1 2 4 8 16 32

Measurement harness only: the brute force is a torch matmul on the GPU."""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from approximatenn_b200.api import _libc, _view, gpu_backend, srandom, stage_times  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--d", type=int, default=64)
    ap.add_argument("--k", type=int, default=16)
    ap.add_argument("--tries", type=int, nargs="+", default=[1, 2, 4, 8, 16, 32])
    ap.add_argument("--sample", type=int, default=2000)
    args = ap.parse_args()
    import torch
    pts = bench.synth_points(args.n, args.d, np.float32)
    gpu = gpu_backend(np.float32)
    gpu.lib.gpu_init()
    gpu.lib.annh_set_timing(1)
    rng = np.random.default_rng(5)
    rows = np.sort(rng.choice(args.n, size=min(args.sample, args.n), replace=False))
    X = torch.from_numpy(pts).cuda()
    xn = (X * X).sum(1)
    exact = []
    for i in range(0, len(rows), 256):
        q = X[torch.from_numpy(rows[i:i + 256]).cuda()]
        d2 = (q * q).sum(1, keepdim=True) + xn[None, :] - 2.0 * q @ X.T
        d2[torch.arange(len(q)), torch.from_numpy(rows[i:i + 256]).cuda()] = float("inf")
        exact.append(d2.topk(args.k, dim=1, largest=False).indices.cpu().numpy())
    exact = np.concatenate(exact)
    out = []
    for T in args.tries:
        for rep in range(2):                      # second run: warm arena
            dptr = ctypes.c_void_p()
            srandom(1001)
            t0 = time.perf_counter()
            p = gpu.precomp_raw(args.n, args.k, args.d, pts.ctypes.data, T, *bench.ROT, None, ctypes.byref(dptr))
            wall = time.perf_counter() - t0
            ids = _view(p, (args.n, args.k), np.uint64)[rows].astype(np.int64)
            _libc.free(p); _libc.free(dptr)
        st = stage_times(gpu)
        dev = st["first_to_last_event"] - st["upload"]          # spans overlap (S2 beside S3): not their sum
        hits = sum(len(np.intersect1d(a, b)) for a, b in zip(exact, ids))
        row = {"tries": T, "recall_at_k": hits / exact.size, "device_ms": dev, "call_ms": wall * 1e3,
               "points_per_s_device": args.n / dev * 1e3}
        out.append(row)
        print(json.dumps(row), flush=True)
    print(json.dumps({"n": args.n, "d": args.d, "k": args.k, "sample": len(rows), "sweep": out}))


if __name__ == "__main__":
    main()
