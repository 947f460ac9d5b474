"""gloo worker for tests/test_dist_plan.py: the sharding PLAN of csrc/ann_dist.c replayed on
CPU.  The partition (try owner, row slice, admitted counts) comes from the product library's
own functions; the per-row arithmetic is the oracle's.  Each rank computes the per-try lists
of the tries it owns, the lists are cut by row slice and sent to the slice owners, who merge
all tries for their rows, all-gather the merged ids, and supercharge their rows.  The result
must equal the single-process reference result row for row."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from approximatenn_b200 import dist as adist  # noqa: E402
from approximatenn_b200.api import gpu_backend, srandom  # noqa: E402


def case(dtype, n, d, k, tries, seed):
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = gpu_backend(dtype).lib          # loads without a GPU; only its pure functions are used
    adist.declare(lib)
    orc = oracle.restatement(dtype)
    rng = np.random.default_rng(seed)
    pts = rng.standard_normal((n, d)).astype(dtype)
    want = orc.precomp(pts, k, tries, seed=seed)

    srandom(seed)                          # every rank draws ALL transforms, like the C host
    st = oracle.Stages(orc, pts, k, tries)
    slices = [adist.row_slice(lib, n, p, world) for p in range(world)]
    lo, hi = slices[rank]
    rows_ids = np.empty((hi - lo, tries * k), dtype=np.uint64)
    rows_key = np.empty((hi - lo, tries * k), dtype=dtype)
    for t in range(tries):                 # all-to-all of the per-try lists, by row slice
        owner = adist.try_owner(lib, t, world)
        if owner == rank:
            ids, key = st.try_lists(t)
            for p, (plo, phi) in enumerate(slices):
                if p == rank:
                    rows_ids[:, t * k:(t + 1) * k] = ids[plo:phi]
                    rows_key[:, t * k:(t + 1) * k] = key[plo:phi]
                else:
                    dist.send(torch.from_numpy(ids[plo:phi].astype(np.int64)), dst=p)
                    dist.send(torch.from_numpy(np.ascontiguousarray(key[plo:phi])), dst=p)
        else:
            ti = torch.empty((hi - lo, k), dtype=torch.int64)
            tk = torch.empty((hi - lo, k), dtype=torch.float32 if dtype == np.float32 else torch.float64)
            dist.recv(ti, src=owner)
            dist.recv(tk, src=owner)
            rows_ids[:, t * k:(t + 1) * k] = ti.numpy().astype(np.uint64)
            rows_key[:, t * k:(t + 1) * k] = tk.numpy()
    st.close()
    # admitted counts only describe the prefix rule; the literal merge applies it by itself
    assert sum(adist.admitted(lib, k, tries, t) for t in range(tries)) == (
        tries * k if tries * k < 16 else 1 << ((tries * k).bit_length() - 1))
    oracle.merge_rows(orc, rows_ids, rows_key)
    merged = np.empty((n, k), dtype=np.uint64)
    for p, (plo, phi) in enumerate(slices):            # all-gather(v) of the merged ids
        buf = torch.from_numpy(rows_ids[:, :k].astype(np.int64).copy()) if p == rank else \
            torch.empty((phi - plo, k), dtype=torch.int64)
        dist.broadcast(buf, src=p)
        merged[plo:phi] = buf.numpy().astype(np.uint64)
    got_ids, got_key = oracle.supercharge_rows(orc, pts, pts, lo, hi, rows_ids, rows_key, merged, k)
    ok = np.array_equal(got_ids, want.ids[lo:hi]) and np.array_equal(
        got_key.view(np.uint8), np.ascontiguousarray(want.dists[lo:hi]).view(np.uint8))
    print(f"rank {rank}/{world} {np.dtype(dtype).name} n={n} k={k} T={tries} rows [{lo},{hi}) "
          f"{'OK' if ok else 'MISMATCH'}", flush=True)
    return ok


def main():
    dist.init_process_group("gloo")
    ok = case(np.float32, 1500, 20, 10, 7, 21)
    ok &= case(np.float64, 1111, 17, 5, 3, 22)       # k*T < 16: block-sorted rows
    ok &= case(np.float32, 2048, 32, 16, 8, 23)
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
