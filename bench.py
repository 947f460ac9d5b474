#!/usr/bin/env python
"""bench.py — all-points kNN throughput of the B200 backend on BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg3]

One "step" = one full precomp (centre, hash, bucket tables, per-try lists, merge,
supercharge) over one batch of synthetic Gaussian points.  Workload at N=1 is BASELINE
config 3: n=1,000,000 d=64 k=16 float, 8 tries + supercharge, reference default rotations.

Printed JSON (one line, rank 0):
  value     points/s, device time only: inputs resident in HBM when the timed region starts
            (CUDA events on the library's stream, from after the upload to before the download)
  e2e       points/s through the reference-facing C-ABI call precomp_gpu(host pointers), the
            time_results protocol (time_results.c:90-129): malloc()ed (pageable) host input ->
            H2D -> all stages -> D2H into the malloc()ed result arrays; the clock brackets the
            call, free() of the results happens outside it as in time_results.
            e2e.pinned repeats it with a page-locked input buffer.
  roofline  the dominant kernel against measured peaks (MEASURED_PEAKS.json for HBM, the
            library's own FFMA probe for the FP32 pipe)
  parity_sample  the FINAL rows of sampled points, recomputed by the CPU oracle at full
            problem size, compared bit for bit with the rows of the timed GPU runs
  cpu_baseline  the CPU restatement of the reference (oracle/, pinned bit-exact to the
            reference's C path) timed on this box's host, 1 core, on that bounded sample
`--impl reference` times the reference's own C path (oracle/_ref, compiled from
/root/reference) on the host and prints the same line shape.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (n, d, k, tries, dtype)
    "cfg1": (16384, 16, 10, 10, np.float32),
    "cfg2": (65536, 32, 16, 8, np.float64),
    "cfg3": (1_000_000, 64, 16, 8, np.float32),
    "cfg4": (10_000_000, 64, 16, 8, np.float32),
    "cfg5": (100_000_000, 32, 32, 8, np.float32),        # needs 8 GPUs (sharded)
}
ROT = (6, 1, 1, 1)          # reference defaults (time_results.c:16-17)
METRIC = "all-points kNN points/sec (precomp: hash + per-try lists + merge + supercharge)"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            j = json.load(f)
        return float(j["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def committed_traffic(config, world):
    """DRAM bytes (read + write) per launch of the hot kernels, from committed `ncu --set full`
    captures of this workload (profiles/ncu_traffic.json, written by tools/ncu_summary.py with
    the commit it was captured on).  Nothing is assumed: no matching capture -> no number."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return {}
    with open(p) as f:
        j = json.load(f)
    if j.get("config") != config or j.get("gpus", 1) != world:
        return {}
    return {k_: v for k_, v in j.get("kernels", {}).items()}


def synth_points(n, d, dtype, seed=1):
    rng = np.random.default_rng(seed)
    out = np.empty((n, d), dtype=dtype)
    step = 1 << 18
    for i in range(0, n, step):
        out[i:i + step] = rng.standard_normal((min(step, n - i), d), dtype=np.float32).astype(dtype)
    return out


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons while the timed region runs.  NVML in-process (a
    `nvidia-smi` subprocess every 100 ms costs the timed region ~10 ms per step); falls back to
    one nvidia-smi query per second if NVML is unavailable."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.idx, self.stop_flag = gpu_index, threading.Event()
        self.sm, self.mx, self.reasons, self.how = [], [], set(), "nvml"

    def _nvml(self):
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
        self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)))
        while not self.stop_flag.is_set():
            self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
            mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for name, bit in self.REASONS:
                if mask & bit:
                    self.reasons.add(name)
            self.stop_flag.wait(0.1)

    def _smi(self):
        self.how = "nvidia-smi"
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            out = subprocess.run(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={q}",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                 timeout=10).stdout.strip().split(",")
            if len(out) >= 6:
                self.sm.append(float(out[0])); self.mx.append(float(out[1]))
                for (name, _), v in zip(self.REASONS, out[2:6]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(name)
            self.stop_flag.wait(1.0)

    def run(self):
        try:
            self._nvml()
        except Exception:
            try:
                self._smi()
            except Exception:
                pass

    def summary(self):
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None,
                "sm_max_mhz": max(self.mx) if self.mx else None, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "source": self.how}


SEED = 1001                  # srandom() before every precomp: the transforms of the timed runs


def sample_ids(n, count):
    return np.sort(np.random.default_rng(99).choice(n, size=min(count, n), replace=False))


def cpu_baseline(cfg, pts, sample_points):
    """Reference algorithm on the host (1 core): bounded sample at full problem size.  Returns
    the baseline record and the exact final rows (ids, squared distances) of the sampled points
    for the same seed as the timed GPU runs — the checker of `parity_sample`."""
    import oracle
    n, d, k, tries, dtype = cfg
    orc = oracle.restatement(dtype)
    sample = sample_ids(n, sample_points)
    from approximatenn_b200.api import srandom
    srandom(SEED)
    t0 = time.time()
    ids, key, c = oracle.sampled_rows(orc, pts, k, tries, sample, *ROT)
    per_point = c["prepare_s"] / n + c["row_s"] + c["supercharge_s"]
    return (ids, key), {"value": 1.0 / per_point, "unit": "points/s", "cores": 1, "kind": "port",
            "sample": (f"oracle/ann_oracle.c (bit-exact restatement of the reference's precomp_cpu; the "
                       f"reference itself needs n*L*d*4 B = >100 GB of scratch at this size): hashing+tables "
                       f"for all {n} points, then per-try rows+merge for {c['rows']} points and supercharge "
                       f"for {len(sample)} sampled points; {time.time() - t0:.1f} s of CPU; points/s = "
                       f"1/(prepare/n + row + supercharge)"),
            "detail": {k_: float(v) for k_, v in c.items()}}


def parity_sample(want, got_ids, got_d, sample):
    """Bit-for-bit comparison of the oracle's rows with the GPU's rows of the same points."""
    want_ids, want_d = want
    same = (got_ids == want_ids).all(axis=1) & (
        np.ascontiguousarray(got_d).view(np.uint8).reshape(len(sample), -1) ==
        np.ascontiguousarray(want_d).view(np.uint8).reshape(len(sample), -1)).all(axis=1)
    return {"rows": int(len(sample)), "mismatch": int((~same).sum()),
            "compared": "final neighbour ids and squared distances (all bits) of the sampled points, oracle "
                        "(oracle/ann_oracle.c, pinned to the reference) vs the rows returned by precomp_gpu",
            "first_mismatching_points": [int(v) for v in sample[~same][:5]]}


def measure_recall(gpu, pts, cfg, sample):
    """recall@k of one more (untimed) precomp against exact brute force on `sample` points.
    The brute force is measurement harness only: fp32 torch matmul on the GPU, top-(k+1) minus self."""
    import torch
    from approximatenn_b200.api import srandom, _libc, _view
    n, d, k, tries, dtype = cfg
    dptr = ctypes.c_void_p()
    srandom(1001)
    ids_p = gpu.precomp_raw(n, k, d, pts.ctypes.data, tries, *ROT, None, ctypes.byref(dptr))
    ids = _view(ids_p, (n, k), np.uint64)
    rng = np.random.default_rng(5)
    rows = np.sort(rng.choice(n, size=min(sample, n), replace=False))
    X = torch.from_numpy(np.ascontiguousarray(pts)).cuda().to(torch.float32)
    q = X[torch.from_numpy(rows).cuda()]
    xn = (X * X).sum(1)
    hits = 0
    for i in range(0, len(rows), 256):
        qq = q[i:i + 256]
        d2 = (qq * qq).sum(1, keepdim=True) + xn[None, :] - 2.0 * qq @ X.T
        d2[torch.arange(len(qq)), torch.from_numpy(rows[i:i + 256]).cuda()] = float("inf")
        exact = d2.topk(k, dim=1, largest=False).indices.cpu().numpy()
        got = ids[rows[i:i + 256]].astype(np.int64)
        hits += sum(len(np.intersect1d(a, b)) for a, b in zip(exact, got))
    _libc.free(ids_p)
    _libc.free(dptr)
    return {"recall_at_k": hits / (len(rows) * k), "k": k, "sample_points": int(len(rows)),
            "against": "exact brute force (fp32) over all n points"}


def measure_pair(gpu, pts, cfg, ycnt):
    """The "precomp + query" pair of the metric: precomp WITH the save structure,
    then query() of `ycnt` fresh Gaussian vectors against it — first call (uploads the index to
    the device) and later calls (index resident).  Bare C-ABI calls with host buffers, result
    arrays freed inside the timed spans like the reference's time_results does."""
    from approximatenn_b200.api import SaveT, srandom, _libc
    n, d, k, tries, dtype = cfg
    rng = np.random.default_rng(11)
    y = np.ascontiguousarray(rng.standard_normal((ycnt, d), dtype=np.float32).astype(dtype))
    sv = SaveT()

    def precomp_save():
        dptr = ctypes.c_void_p()
        srandom(1001)
        t0 = time.perf_counter()
        ids = gpu.precomp_raw(n, k, d, pts.ctypes.data, tries, *ROT, ctypes.byref(sv), ctypes.byref(dptr))
        _libc.free(ids); _libc.free(dptr)
        return time.perf_counter() - t0

    def query():
        dptr = ctypes.c_void_p()
        t0 = time.perf_counter()
        ids = gpu._query(ctypes.byref(sv), pts.ctypes.data, ycnt, y.ctypes.data, ctypes.byref(dptr))
        _libc.free(ids); _libc.free(dptr)
        return time.perf_counter() - t0

    precomp_save()                       # warm: table staging buffer, arena
    gpu._free_save(ctypes.byref(sv))
    tp = precomp_save()
    q1 = query()
    q2 = min(query() for _ in range(3))
    gpu._free_save(ctypes.byref(sv))
    return {"precomp_with_save_ms": 1e3 * tp, "precomp_with_save_points_per_s": n / tp,
            "query_count": ycnt, "query_first_call_ms": 1e3 * q1, "query_resident_index_ms": 1e3 * q2,
            "queries_per_s_resident": ycnt / q2,
            "pair_ms": 1e3 * (tp + q1), "pair_points_plus_queries_per_s": (n + ycnt) / (tp + q1)}


def reference_direct(name, n_override=None):
    """The reference's own precomp_cpu (oracle/_ref, compiled unmodified from /root/reference) on a
    BASELINE config, inputs from numpy's generator, time_results' clock placement
    (time_results.c:121-129): wall time of the call including free() of the results."""
    import oracle
    from approximatenn_b200.api import srandom, _libc
    n, d, k, tries, dtype = CONFIGS[name]
    if n_override:
        n = n_override
    ref = oracle.reference(dtype)
    pts = synth_points(n, d, dtype)
    dptr = ctypes.c_void_p()
    srandom(SEED)
    t0 = time.perf_counter()
    ids = ref.precomp_raw(n, k, d, pts.ctypes.data, tries, *ROT, None, ctypes.byref(dptr))
    _libc.free(ids); _libc.free(dptr)
    dt = time.perf_counter() - t0
    ds = int(np.ceil(np.log2(np.float32(n) / np.float32(k)))) if dtype == np.float32 else int(np.ceil(np.log2(n / k)))
    return {"config": name, "n": n, "d": d, "k": k, "tries": tries, "seconds": dt, "points_per_s": n / dt,
            "d_short": min(ds, 1 << (d - 1).bit_length())}


def expected_row_len(n, d_short, trials=3):
    """L = (d_short+1)*tmax of the reference's candidate rows (alg.c:252-277) for uniformly hashed
    points: tmax = largest of 2^d_short bucket counts (Monte Carlo, as SURVEY 8's table)."""
    rng = np.random.default_rng(7)
    tm = [np.bincount(rng.integers(0, 1 << d_short, size=n), minlength=1 << d_short).max() for _ in range(trials)]
    return (d_short + 1) * float(np.median(tm))


def run_reference_arm(args, cfg, name):
    """--impl reference: the reference's OWN C path (oracle/_ref) on this box's host cores.  The
    path is single-threaded by construction (ann.h:37-38), so cores = 1.  Configs the reference
    can hold in memory (cfg1, cfg2) are timed directly, one call per step.  For cfg3-5 its scratch
    buffer n*L*d*w (alg.c:237) exceeds any host, so each step times the reference on a bounded
    problem of the same d, k, tries and the per-point cost is scaled by the ratio of candidate-row
    lengths L (t ~ N*T*L*d) — labelled as extrapolated."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    n, d, k, tries, dtype = cfg
    if not oracle.reference_available():
        line = {"impl": "reference", "unavailable": "oracle/_ref was not built (no /root/reference checkout when build() ran)"}
        print(json.dumps(line), flush=True)
        return
    direct = name in ("cfg1", "cfg2")
    n_run = n if direct else args.ref_sample_n
    vals, runs = [], []
    for i in range(args.warmup + args.steps):
        r = reference_direct(name, None if direct else n_run)
        if i >= args.warmup:
            runs.append(r)
            vals.append(r["points_per_s"])
    v_run = statistics.mean(vals)
    scale, how = 1.0, "measured directly on the named config"
    if not direct:
        ds_full = int(np.ceil(np.log2(np.float32(n) / np.float32(k))))
        L_full, L_run = expected_row_len(n, ds_full), expected_row_len(n_run, runs[-1]["d_short"])
        scale = L_run / L_full
        how = (f"EXTRAPOLATED: reference timed at n={n_run} (same d, k, tries), points/s scaled by the candidate-row "
               f"length ratio L({n_run})/L({n}) = {L_run:.0f}/{L_full:.0f}; the reference itself cannot allocate its "
               f"n*L*d*4 B scratch at n={n}")
    v = v_run * scale
    cb = {"value": v, "unit": "points/s", "cores": 1, "kind": "reference",
          "sample": f"oracle/_ref precomp_cpu (the reference's C path, gcc -O2 -ffp-contract=off), {how}; "
                    f"{len(runs)} timed call(s), mean {statistics.mean(r['seconds'] for r in runs):.2f} s each",
          "measured_points_per_s_at_run_size": v_run, "run_n": n_run}
    if args.ref_cfg1 and name != "cfg1":
        cb["direct_cfg1"] = reference_direct("cfg1")          # BASELINE config 1, measured, never extrapolated
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "points/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * n / v, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32" if dtype == np.float32 else "f64",
            "data": "synthetic", "config": workload_config(cfg, name, args.gpus),
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def extra_config_run(gpu, name, world, rank, steps=3, warmup=1):
    """A second, shorter measurement on another BASELINE config inside the same job (the N=8 line
    carries config 4, n = 10^7, this way while the scaling series itself stays on config 3).
    Device time from the stage events, call time as in `e2e`; instead of the oracle (its
    single-threaded table build alone takes minutes at this size) sampled rows are checked for the
    properties every correct result has: ascending distances, valid ids, no self, and squared
    distances equal to a float64 recomputation within 1e-5 relative (north_star's tolerance)."""
    import torch
    from approximatenn_b200.api import srandom, stage_times, _libc, _view
    n, d, k, tries, dtype = CONFIGS[name]
    pts = synth_points(n, d, dtype)
    lo, hi = 0, n
    if world > 1:
        from approximatenn_b200 import dist as adist
        lo, hi = adist.row_slice(gpu.lib, n, rank, world)
    dev, call, kept = [], [], {}

    def one(i):
        dptr = ctypes.c_void_p()
        srandom(SEED)
        t0 = time.perf_counter()
        ids = gpu.precomp_raw(n, k, d, pts.ctypes.data, tries, *ROT, None, ctypes.byref(dptr))
        dt = time.perf_counter() - t0
        st = stage_times(gpu)
        if i == warmup + steps - 1 and hi > lo:
            rows = np.sort(np.random.default_rng(3 + rank).choice(hi - lo, size=min(64, hi - lo), replace=False))
            kept["rows"] = rows + lo
            kept["ids"] = _view(ids, (hi - lo, k), np.uint64)[rows].copy()
            kept["d"] = _view(dptr, (hi - lo, k), dtype)[rows].copy()
        _libc.free(ids); _libc.free(dptr)
        return dt, st

    for i in range(warmup + steps):
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        dt, st = one(i)
        if i >= warmup:
            call.append(dt)
            dev.append(st["first_to_last_event"] - st["upload"])       # spans overlap (S2 beside S3): not their sum
            last = st
    bad = 0
    if kept:
        ids_, d_ = kept["ids"].astype(np.int64), kept["d"].astype(np.float64)
        exact = ((pts[kept["rows"]][:, None, :].astype(np.float64) - pts[np.minimum(ids_, n - 1)].astype(np.float64)) ** 2).sum(-1)
        ok = (ids_ < n).all(axis=1) & (ids_ != kept["rows"][:, None]).all(axis=1) & (np.diff(d_, axis=1) >= 0).all(axis=1) & \
            (np.abs(exact - d_) <= 1e-5 * np.maximum(exact, 1e-30)).all(axis=1)
        bad = int((~ok).sum())
    t = torch.tensor([sum(dev) / 1e3, sum(call), float(bad), float(len(kept.get("rows", ())))], device="cuda", dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist
        mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        t = torch.stack([mx[0], mx[1], sm[2], sm[3]])
    return {"config": workload_config(CONFIGS[name], name, world), "steps": steps, "warmup": warmup,
            "value": n * steps / float(t[0]), "unit": "points/s", "ms_per_step": 1e3 * float(t[0]) / steps,
            "e2e": {"value": n * steps / float(t[1]), "ms_per_step": 1e3 * float(t[1]) / steps},
            "stage_ms_rank0_last_step": last,
            "sanity_rows": {"rows": int(t[3]), "failing": int(t[2]),
                            "checked": "ascending distances, ids < n and != the point, squared distances vs a "
                                       "float64 recomputation within 1e-5 relative"}}


def workload_config(cfg, name, gpus):
    n, d, k, tries, dtype = cfg
    return {"workload": f"BASELINE {name}: n={n} d={d} k={k} {np.dtype(dtype).name}, {tries} tries + supercharge, "
                        f"rotations {ROT}, iid N(0,1) points",
            "n": n, "d": d, "k": k, "tries": tries, "gpus": gpus,
            "parallelism": "1 GPU" if gpus == 1 else f"tries sharded over {gpus} ranks, rows of merge+supercharge sliced, NCCL all-to-all/all-gather",
            "l2": ("inputs (%.0f MB) plus the per-try lists (%.0f MB) exceed the 126 MB L2 and every step re-uploads "
                   "the points; no flush between steps" % (n * d * np.dtype(dtype).itemsize / 1e6,
                                                            n * k * tries * (4 + np.dtype(dtype).itemsize) / 1e6))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS))
    ap.add_argument("--cpu-sample", type=int, default=256, help="points in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parity-sample", type=int, default=64,
                    help="sampled points checked against the oracle when the CPU baseline does not run (N>1)")
    ap.add_argument("--ref-sample-n", type=int, default=1024,
                    help="--impl reference on cfg3-5: points of the bounded problem the reference is timed on")
    ap.add_argument("--ref-cfg1", action="store_true",
                    help="--impl reference: also time BASELINE config 1 directly (adds ~20-45 s)")
    ap.add_argument("--pair-queries", type=int, default=65536,
                    help="query vectors of the precomp(save)+query pair measurement (0 = skip)")
    ap.add_argument("--extra-config", default="auto", choices=["auto", "none"] + sorted(CONFIGS),
                    help="second, shorter measurement in the same job (auto: cfg4 at 8 GPUs on cfg3)")
    ap.add_argument("--recall-sample", type=int, default=2000,
                    help="points whose exact k nearest neighbours are brute-forced for recall@k (0 = skip)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    n, d, k, tries, dtype = cfg

    if args.impl == "reference":
        run_reference_arm(args, cfg, args.config)
        return

    # keep stdout clean for the one JSON line: native libraries (NCCL's version banner) write
    # to fd 1 directly, so fd 1 points at stderr until the line is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    os.environ.setdefault("ANN_B200_DEVICE", str(local))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from approximatenn_b200.api import gpu_backend, srandom, stage_times, _libc
    gpu = gpu_backend(dtype)
    gpu.lib.gpu_init()
    gpu.lib.annh_set_timing(1)
    if world > 1:
        # tries are sharded across ranks inside the library (csrc/ann_dist.c); torch.distributed
        # only carries NCCL's unique id.  Every rank passes the same points and the same seed and
        # gets back the rows it owns.
        from approximatenn_b200 import dist as adist
        adist.init_from_torch(gpu.lib, gather_full=False)

    # The caller's buffer, as the reference's time_results allocates it (time_results.c:90):
    # plain malloc()ed, i.e. pageable, memory.  A page-locked copy is timed as an extra.
    pts = synth_points(n, d, dtype)
    host = torch.empty((n, d), dtype=torch.float32 if dtype == np.float32 else torch.float64,
                       pin_memory=True)
    pts_pinned = host.numpy()
    pts_pinned[:] = pts

    call_s = [0.0]                     # time inside precomp_gpu, summed (time_results.c:121-129)

    def step(src=pts, keep=None):
        dptr = ctypes.c_void_p()
        srandom(SEED)
        t_in = time.perf_counter()
        ids = gpu.precomp_raw(n, k, d, src.ctypes.data, tries, *ROT, None, ctypes.byref(dptr))
        call_s[0] += time.perf_counter() - t_in
        st = stage_times(gpu)
        if keep is not None:
            keep(ids, dptr)
        _libc.free(ids)
        _libc.free(dptr)
        return st

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    sampler.start()
    gpu.lib.annb_launch_count(1)
    gpu.lib.annb_leaf_pairs.restype = ctypes.c_ulonglong
    gpu.lib.annb_leaf_pairs(1)
    gpu.lib.annb_leaf_exact_pairs.restype = ctypes.c_ulonglong
    gpu.lib.annb_leaf_overflow_buckets.restype = ctypes.c_ulonglong
    gpu.lib.annb_leaf_exact_pairs(1)
    gpu.lib.annb_leaf_overflow_buckets(1)
    # timed region 1: K steps with per-stage CUDA events on the library stream -> `value`
    barrier()
    t0 = time.perf_counter()
    stages = [step() for _ in range(args.steps)]
    barrier()
    wall_instrumented = time.perf_counter() - t0
    launches = int(gpu.lib.annb_launch_count(0))
    leaf_pairs = int(gpu.lib.annb_leaf_pairs(0)) / args.steps        # per step, this rank
    exact_pairs = int(gpu.lib.annb_leaf_exact_pairs(0)) / args.steps  # of those, sent through the exact tree
    overflow_buckets = int(gpu.lib.annb_leaf_overflow_buckets(0)) / args.steps
    # timed region 2: the same K steps through the C-ABI with the event instrumentation off -> `e2e`
    gpu.lib.annh_set_timing(0)
    step()
    barrier()
    t0 = time.perf_counter()
    call_s[0] = 0.0
    for _ in range(args.steps):
        step()
    barrier()
    wall_loop = time.perf_counter() - t0
    wall = call_s[0]
    # timed region 3: page-locked caller buffer (extra)
    step(pts_pinned)
    barrier()
    call_s[0] = 0.0
    for _ in range(args.steps):
        step(pts_pinned)
    barrier()
    wall_pinned = call_s[0]
    gpu.lib.annh_set_timing(1)
    sampler.stop_flag.set()
    sampler.join()

    # one more (untimed) run with the same seed whose rows are kept for the parity sample
    run_baseline = world == 1 and not args.no_cpu_baseline
    sample = sample_ids(n, args.cpu_sample if run_baseline else args.parity_sample)
    lo, hi = 0, n
    if world > 1:
        from approximatenn_b200 import dist as adist
        lo, hi = adist.row_slice(gpu.lib, n, rank, world)
    mine = sample[(sample >= lo) & (sample < hi)]
    kept = {}

    def keep_rows(ids_p, d_p):
        from approximatenn_b200.api import _view
        kept["ids"] = _view(ids_p, (hi - lo, k), np.uint64)[mine - lo].copy()
        kept["d"] = _view(d_p, (hi - lo, k), dtype)[mine - lo].copy()

    step(pts, keep_rows)

    # device time of a step: from the end of the upload to the end of the last S5 chunk, by CUDA
    # events on the library's stream.  (Not the sum of the stage spans: the bucket tables of try
    # j+1 are built on a second stream while try j's lists are computed, so spans overlap.)
    dev_ms = [s["first_to_last_event"] - s["upload"] for s in stages]
    dev_total_s = sum(dev_ms) / 1e3
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([dev_total_s, wall, wall_pinned, wall_loop], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_total_s, wall, wall_pinned, wall_loop = float(t[0]), float(t[1]), float(t[2]), float(t[3])
        parts = [None] * world if rank == 0 else None
        dist.gather_object((mine, kept["ids"], kept["d"]), parts, dst=0)
        stage_all = [None] * world if rank == 0 else None
        dist.gather_object({k_: statistics.mean(s[k_] for s in stages) for k_ in stages[0]}, stage_all, dst=0)
        if rank == 0:
            mine = np.concatenate([p_[0] for p_ in parts])
            kept = {"ids": np.concatenate([p_[1] for p_ in parts]), "d": np.concatenate([p_[2] for p_ in parts])}
            o = np.argsort(mine, kind="stable")
            mine, kept = mine[o], {"ids": kept["ids"][o], "d": kept["d"][o]}
    fp32_probe = None
    if rank == 0:
        gpu.lib.annb_probe_fp32.restype = ctypes.c_double
        gpu.lib.annb_probe_fp32.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        gpu.lib.annh_stream.restype = ctypes.c_void_p
        st_ = gpu.lib.annh_stream()
        fp32_probe = {m_: float(gpu.lib.annb_probe_fp32(i_, 3, st_)) for i_, m_ in
                      enumerate(("ffma", "fmul_fadd", "ffma2", "fmul2_fadd2"))}
    extra_name = args.extra_config
    if extra_name == "auto":
        extra_name = "cfg4" if (world == 8 and args.config == "cfg3") else "none"
    extra = extra_config_run(gpu, extra_name, world, rank) if extra_name != "none" else None
    if world > 1:
        gpu.lib.annb200_dist_shutdown()
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    w = np.dtype(dtype).itemsize
    mean_stage = {k_: statistics.mean(s[k_] for s in stages) for k_ in stages[0]}
    peak, peak_src = measured_peaks()
    dev_mean = statistics.mean(dev_ms)
    # --- dominant kernel: S3 (leaf_screen_kernel, or leaf_topk_tile_kernel where the screen does
    # not apply).  Algorithmic work per (point, candidate) pair = d subtractions + d multiplications
    # + d-1 additions (compute.cl:147-166), counted as 3*d flops.  `peak` is the FP32
    # pipe measured by the library's own probe (annb_probe.cu) as FFMA chains, an FMA counting two
    # flops (the SURVEY 8.D convention); `frac_separately_rounded` divides by the measured rate of
    # packed, separately rounded FMUL2+FADD2 chains instead — the only form the reference's bits
    # allow (an FMA would change them).  The screened kernel reaches its rate by NOT executing most
    # of these flops: fp16 tensor-core brackets leave `exact_pairs_per_step` pairs for the exact
    # tree, so `frac` is algorithmic work over peak, not pipe utilisation (ncu: profiles/).
    traffic = committed_traffic(args.config, world)
    leaf_s = mean_stage["leaf"] / 1e3
    fp32_peak = fp32_probe["ffma"]
    leaf_tflops = 3 * d * leaf_pairs / leaf_s / 1e12 if leaf_s > 0 else 0.0
    screened = exact_pairs > 0
    leaf_kernel = "leaf_screen_kernel" if screened else "leaf_topk_tile_kernel"
    roofline = {"kernel": leaf_kernel + " (S3, summed over the tries of one step; includes prep and literal redo)",
                "bound": "fp32", "achieved": leaf_tflops, "peak": fp32_peak,
                "peak_source": "annb_probe_fp32: FFMA chains on this device, FMA = 2 flops (measured live)",
                "unit": "TFLOP/s", "frac": leaf_tflops / fp32_peak if fp32_peak else None,
                "frac_separately_rounded": leaf_tflops / fp32_probe["fmul2_fadd2"] if fp32_probe["fmul2_fadd2"] else None,
                "fp32_probe_tflops": fp32_probe,
                "traffic": traffic.get(leaf_kernel),
                "pairs_per_step": leaf_pairs, "flops_per_pair": 3 * d,
                "exact_pairs_per_step": exact_pairs, "screen_overflow_buckets_per_step": overflow_buckets,
                "share_of_device_time": mean_stage["leaf"] / dev_mean}
    # --- dominant HBM-bound kernel: supercharge.  Algorithmic bytes per launch set:
    # rows*(P2-k) gathered vectors of (4 + d*w) B, + own lists in, + ids and dists out.
    P2 = 1 << ((k * (k + 1)).bit_length() - 1)
    rows = n / world                      # supercharge rows per rank
    sc_bytes = rows * (P2 - k) * (4 + d * w) + rows * k * (4 + w) + rows * k * (4 + w)
    sc_s = mean_stage["supercharge"] / 1e3
    roofline_hbm = {"kernel": "supercharge kernels (S5, all row chunks of one step)", "bound": "hbm",
                    "achieved": sc_bytes / sc_s / 1e9, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                    "frac": sc_bytes / sc_s / 1e9 / peak, "traffic": traffic.get("supercharge_screen_kernel"),
                    "share_of_device_time": mean_stage["supercharge"] / dev_mean}
    # --- the whole step against HBM (SURVEY 8.D's per-run byte model; w = sizeof ftype, B = buckets)
    ds = int(np.ceil(np.log2((np.float32(n) if dtype == np.float32 else np.float64(n)) / k)))
    ds = min(ds, 1 << (d - 1).bit_length())
    B = 1 << ds
    step_bytes = (n * d * w) + (n * d * w + 4 * n * tries) + tries * (12 * n + 8 * B) + \
        tries * (n * d * w + n * k * (4 + w)) + (n * k * tries * (4 + w) + n * k * (4 + w)) + \
        (n * (P2 - k) * (4 + d * w) + n * k * (4 + w) + n * k * (8 + w))
    step_s = dev_total_s / args.steps
    roofline_step = {"what": "all stages of one step, SURVEY 8.D algorithmic bytes / device time (all ranks)",
                     "bound": "hbm", "algorithmic_bytes": step_bytes, "achieved": step_bytes / step_s / 1e9,
                     "peak": peak * world, "unit": "GB/s", "frac": step_bytes / step_s / 1e9 / (peak * world)}
    line = {"metric": METRIC, "value": n * args.steps / dev_total_s, "unit": "points/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_total_s / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if dtype == np.float32 else "f64", "data": "synthetic",
            "config": workload_config(cfg, args.config, world),
            "e2e": {"value": n * args.steps / wall, "unit": "points/s",
                    "h2d_bytes_per_step": n * d * w, "d2h_bytes_per_step": n * k * (4 + w),
                    "ms_per_step": 1e3 * wall / args.steps,
                    "input": "malloc()ed (pageable) host array",
                    "timed": "time inside the K precomp_gpu calls (host pointers in, malloc()ed arrays out), the "
                             "clock placement of the reference's time_results (time_results.c:121-129): free() of "
                             "the results is outside; max over ranks",
                    "ms_per_step_loop_with_free": 1e3 * wall_loop / args.steps,
                    "ms_per_step_with_stage_events": 1e3 * wall_instrumented / args.steps,
                    "pinned": {"value": n * args.steps / wall_pinned, "ms_per_step": 1e3 * wall_pinned / args.steps,
                               "input": "page-locked host array"}},
            "gpu_launches": launches, "stage_ms": mean_stage,
            "stage_ms_note": "CUDA-event spans per stage, summed over a step. With two or more owned tries the bucket "
                             "tables of try j+1 (buckets) are built on a second stream beside the lists of try j "
                             "(leaf): those spans overlap and stretch, and ms_per_step = first_to_last_event - upload "
                             "is less than their sum (ANN_B200_PIPE=0 serialises them).",
            "roofline": roofline, "roofline_hbm": roofline_hbm,
            "roofline_step": roofline_step, "clocks": sampler.summary(), "host_cores": os.cpu_count()}
    if world > 1:
        line["stage_ms_per_rank"] = stage_all
    if extra is not None:
        line[extra_name] = extra
    if args.recall_sample > 0 and world == 1:
        line["recall"] = measure_recall(gpu, pts, cfg, args.recall_sample)
    if args.pair_queries > 0 and world == 1:
        line["precomp_query_pair"] = pq_ = measure_pair(gpu, pts, cfg, args.pair_queries)
        # Q path against the HBM roof: the reference measures every real candidate of a query's
        # tries*(d_short+1) table rows (alg.c:464-508), about tries*(d_short+1)*n/2^d_short rows of
        # d*w bytes, then the k*k supercharge candidates.  Whole call, host buffers in and out.
        d_short_ = int(np.ceil(np.log2(np.float32(n) / np.float32(k))))
        cand_ = tries * (d_short_ + 1) * n / float(1 << d_short_)
        bytes_ = args.pair_queries * ((cand_ + k * k) * (4 + d * w) + 2 * k * (4 + w))
        ach_ = bytes_ / (pq_["query_resident_index_ms"] * 1e-3) / 1e9
        line["roofline_query"] = {"kernel": "query_gpu on the resident index: query_hash_warp_kernel + query_rows_fast_kernel "
                                            "+ supercharge (whole call, host buffers in and out)",
                                  "bound": "hbm", "achieved": ach_, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                                  "frac": ach_ / peak,
                                  "traffic": committed_traffic(args.config, world).get("query_rows_fast_kernel")
                                  if args.pair_queries == 65536 else None,
                                  "algorithmic_candidates_per_query": cand_ + k * k,
                                  "note": "algorithmic bytes of the reference's query (every candidate row read once); "
                                          "the fp16 screen moves about half of them"}
    if run_baseline:
        want, line["cpu_baseline"] = cpu_baseline(cfg, pts, args.cpu_sample)
    else:
        import oracle
        srandom(SEED)
        ids_, key_, _ = oracle.sampled_rows(oracle.restatement(dtype), pts, k, tries, sample, *ROT)
        want = (ids_, key_)
    assert np.array_equal(mine, sample), "every sampled point must be owned by exactly one rank"
    line["parity_sample"] = parity_sample(want, kept["ids"], kept["d"], sample)
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
