"""torchrun worker for tests/test_gpu_multi.py: sharded precomp_gpu vs the single-process
CPU oracle.  Every rank checks the rows it owns (and the full result in gather mode)."""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from approximatenn_b200 import dist as adist  # noqa: E402
from approximatenn_b200.api import _libc, _view, gpu_backend, srandom  # noqa: E402


def run_case(gpu, dtype, n, d, k, tries, seed, gather):
    rank, world = dist.get_rank(), dist.get_world_size()
    rng = np.random.default_rng(seed)
    pts = rng.standard_normal((n, d)).astype(dtype)
    want = oracle.restatement(dtype).precomp(pts, k, tries, seed=seed)
    gpu.lib.annb200_dist_gather(1 if gather else 0)
    lo, hi = (0, n) if gather else adist.row_slice(gpu.lib, n, rank, world)
    dptr = ctypes.c_void_p()
    srandom(seed)
    ids = gpu.precomp_raw(n, k, d, pts.ctypes.data, tries, 6, 1, 1, 1, None, ctypes.byref(dptr))
    got_ids = _view(ids, (hi - lo, k), np.uint64).copy()
    got_d = _view(dptr, (hi - lo, k), dtype).copy()
    _libc.free(ids); _libc.free(dptr)
    ok = np.array_equal(got_ids, want.ids[lo:hi]) and np.array_equal(
        got_d.view(np.uint8), np.ascontiguousarray(want.dists[lo:hi]).view(np.uint8))
    print(f"rank {rank}/{world} {np.dtype(dtype).name} n={n} d={d} k={k} T={tries} gather={gather} "
          f"rows [{lo},{hi}) {'OK' if ok else 'MISMATCH'}", flush=True)
    return ok


def run_sampled_case(gpu, dtype, n, d, k, tries, seed, samples):
    """Full-size problem (the oracle cannot run all of it in test time): every rank checks the
    sampled points that fall into its row slice against the oracle's exact final rows."""
    rank, world = dist.get_rank(), dist.get_world_size()
    rng = np.random.default_rng(seed)
    pts = rng.standard_normal((n, d), dtype=np.float32).astype(dtype)
    gpu.lib.annb200_dist_gather(0)
    lo, hi = adist.row_slice(gpu.lib, n, rank, world)
    sample = np.sort(rng.choice(n, size=samples, replace=False))
    mine = sample[(sample >= lo) & (sample < hi)]
    dptr = ctypes.c_void_p()
    srandom(seed)
    ids = gpu.precomp_raw(n, k, d, pts.ctypes.data, tries, 6, 1, 1, 1, None, ctypes.byref(dptr))
    got_ids = _view(ids, (hi - lo, k), np.uint64)[mine - lo].copy()
    got_d = _view(dptr, (hi - lo, k), dtype)[mine - lo].copy()
    _libc.free(ids); _libc.free(dptr)
    srandom(seed)
    want_ids, want_d, _ = oracle.sampled_rows(oracle.restatement(dtype), pts, k, tries, mine)
    ok = np.array_equal(got_ids, want_ids) and np.array_equal(got_d.view(np.uint8), want_d.view(np.uint8))
    print(f"rank {rank}/{world} {np.dtype(dtype).name} n={n} d={d} k={k} T={tries} sampled rows "
          f"{len(mine)} of [{lo},{hi}) {'OK' if ok else 'MISMATCH'}", flush=True)
    return ok


def run_save_case(gpu, dtype, n, d, k, tries, seed):
    """save != NULL in sharded mode: every rank ends up with the complete index."""
    rank, world = dist.get_rank(), dist.get_world_size()
    rng = np.random.default_rng(seed)
    pts = rng.standard_normal((n, d)).astype(dtype)
    y = rng.standard_normal((50, d)).astype(dtype)
    orc = oracle.restatement(dtype)
    want = orc.precomp(pts, k, tries, want_save=True, seed=seed)
    got = gpu.precomp(pts, k, tries, want_save=True, seed=seed)
    ok = np.array_equal(got.ids, want.ids) and np.array_equal(got.save.graph, want.save.graph)
    ok &= np.array_equal(got.save.par_maxes, want.save.par_maxes)
    ok &= np.array_equal(got.save.bases.view(np.uint8), want.save.bases.view(np.uint8))
    ok &= np.array_equal(got.save.row_means.view(np.uint8), want.save.row_means.view(np.uint8))
    for t in range(tries):
        ok &= np.array_equal(got.save.which_par(t), want.save.which_par(t))
    qa, qb = gpu.query(got.save, pts, y), orc.query(want.save, pts, y)
    ok &= np.array_equal(qa.ids, qb.ids)
    got.save.free(); want.save.free()
    print(f"rank {rank}/{world} {np.dtype(dtype).name} n={n} save+query {'OK' if ok else 'MISMATCH'}", flush=True)
    return ok


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ["ANN_B200_DEVICE"] = str(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    gpus = {}
    for dtype in (np.float32, np.float64):
        gpus[dtype] = gpu_backend(dtype)
        adist.init_from_torch(gpus[dtype].lib)
    cases = [
        (np.float32, 8192, 64, 16, 8, 401, False),
        (np.float32, 8192, 64, 16, 8, 401, True),
        (np.float32, 5003, 16, 10, 10, 402, False),     # k*T = 100: prefix corner, uneven try split
        (np.float64, 4096, 32, 16, 5, 403, False),      # more ranks than some ranks have tries
        (np.float32, 3000, 20, 10, 3, 404, True),
        (np.float32, 8191, 64, 16, 5, 407, False),      # T = 5: ranks 5.. own no try at 8 ranks; odd n
        (np.float32, 2049, 32, 8, 8, 408, False),       # ceil(n/R) rounds up to 32 rows: short or empty last slice
        (np.float64, 6007, 16, 10, 10, 409, True),      # double, prefix corner, n prime
        (np.float32, 224, 16, 4, 4, 411, False),        # 8 ranks: slices of 32 rows, the last one is empty
    ]
    for dtype, n, d, k, tries, seed, gather in cases:
        ok &= run_case(gpus[dtype], dtype, n, d, k, tries, seed, gather)
    ok &= run_sampled_case(gpus[np.float32], np.float32, 1_000_000, 64, 16, 8, 410, 96)   # BASELINE config 3
    ok &= run_save_case(gpus[np.float32], np.float32, 6000, 32, 16, 5, 405)
    ok &= run_save_case(gpus[np.float64], np.float64, 3000, 20, 10, 4, 406)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    for g in gpus.values():
        g.lib.annb200_dist_shutdown()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
