#!/usr/bin/env python
"""bench.py — all-points kNN throughput of the B200 backend on BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg3]

One "step" = one full precomp (centre, hash, bucket tables, per-try lists, merge,
supercharge) over one batch of synthetic Gaussian points.  Workload at N=1 is BASELINE
config 3: n=1,000,000 d=64 k=16 float, 8 tries + supercharge, reference default rotations.

Printed JSON (one line, rank 0):
  value     points/s, device time only: inputs resident in HBM when the timed region starts
            (CUDA events on the library's stream, from after the upload to before the download)
  e2e       points/s through the reference-facing C-ABI call precomp_gpu(host pointers):
            pinned host input -> H2D -> all stages -> D2H into the malloc()ed result arrays
  roofline  the dominant kernel against the measured peak (MEASURED_PEAKS.json)
  cpu_baseline  the CPU restatement of the reference (oracle/, pinned bit-exact to the
            reference's C path) timed on this box's host, 1 core, on a bounded sample
`--impl reference` times that CPU path alone and prints the same line shape.
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


def cpu_baseline(cfg, pts, sample_points):
    """Reference algorithm on the host (1 core): bounded sample at full problem size."""
    import oracle
    n, d, k, tries, dtype = cfg
    orc = oracle.restatement(dtype)
    rng = np.random.default_rng(99)
    sample = rng.choice(n, size=min(sample_points, n), replace=False)
    from approximatenn_b200.api import srandom
    srandom(4242)
    t0 = time.time()
    c = oracle.sampled_cost(orc, pts, k, tries, sample, *ROT)
    per_point = c["prepare_s"] / n + c["row_s"] + c["supercharge_s"]
    return {"value": 1.0 / per_point, "unit": "points/s", "cores": 1, "kind": "port",
            "sample": (f"oracle/ann_oracle.c (bit-exact restatement of the reference's precomp_cpu; the "
                       f"reference itself needs n*L*d*4 B = >100 GB of scratch at this size): hashing+tables "
                       f"for all {n} points, then per-try rows+merge for {c['rows']} points and supercharge "
                       f"for {len(sample)} sampled points; {time.time() - t0:.1f} s of CPU; points/s = "
                       f"1/(prepare/n + row + supercharge)"),
            "detail": {k_: float(v) for k_, v in c.items()}}


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


def run_reference_arm(args, cfg, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, d, k, tries, dtype = cfg
    pts = synth_points(n, d, dtype)
    vals = []
    for i in range(args.warmup + args.steps):
        cb = cpu_baseline(cfg, pts, args.cpu_sample)
        if i >= args.warmup:
            vals.append(cb)
    v = statistics.mean(c["value"] for c in vals)
    cb = vals[-1]
    cb["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "points/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * n / v, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32" if dtype == np.float32 else "f64",
            "data": "synthetic", "config": workload_config(cfg, name, args.gpus),
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(cfg, name, gpus):
    n, d, k, tries, dtype = cfg
    return {"workload": f"BASELINE {name}: n={n} d={d} k={k} {np.dtype(dtype).name}, {tries} tries + supercharge, "
                        f"rotations {ROT}, iid N(0,1) points",
            "n": n, "d": d, "k": k, "tries": tries, "gpus": gpus,
            "parallelism": "1 GPU" if gpus == 1 else f"tries sharded over {gpus} ranks, rows of merge+supercharge sliced, NCCL all-to-all/all-gather",
            "l2": "inputs (n*d*4 B = %.0f MB) exceed the 126 MB L2; no flush needed" % (n * d * np.dtype(dtype).itemsize / 1e6)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS))
    ap.add_argument("--cpu-sample", type=int, default=256, help="points in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pair-queries", type=int, default=65536,
                    help="query vectors of the precomp(save)+query pair measurement (0 = skip)")
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

    # pinned host input: the buffer the caller hands to precomp_gpu
    host = torch.empty((n, d), dtype=torch.float32 if dtype == np.float32 else torch.float64,
                       pin_memory=True)
    pts = host.numpy()
    pts[:] = synth_points(n, d, dtype)

    def step():
        dptr = ctypes.c_void_p()
        srandom(1001)
        ids = gpu.precomp_raw(n, k, d, pts.ctypes.data, tries, *ROT, None, ctypes.byref(dptr))
        st = stage_times(gpu)
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
    for _ in range(args.steps):
        step()
    barrier()
    wall = time.perf_counter() - t0
    gpu.lib.annh_set_timing(1)
    sampler.stop_flag.set()
    sampler.join()

    dev_keys = ("means", "hash", "buckets", "leaf", "exchange", "merge", "supercharge")
    dev_ms = [sum(s[k_] for k_ in dev_keys) for s in stages]
    dev_total_s = sum(dev_ms) / 1e3
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([dev_total_s, wall], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_total_s, wall = float(t[0]), float(t[1])
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
    # + d-1 additions, each rounded separately (the reference's arithmetic, compute.cl:147-166; an
    # FMA would change its bits), counted as 3*d flops.  Peak = one rounded FP32
    # operation per lane per clock: 148 SMs x 128 lanes x max SM clock (half of the usual
    # FFMA-counts-two figure).  The screened kernel reaches its rate by NOT executing most of these
    # flops: fp16 tensor-core brackets leave `exact_pairs_per_step` pairs for the exact tree, so
    # `frac` is algorithmic work over peak, not pipe utilisation (ncu: profiles/).
    leaf_s = mean_stage["leaf"] / 1e3
    fp32_peak = 148 * 128 * 1.965e9 / 1e12
    leaf_tflops = 3 * d * leaf_pairs / leaf_s / 1e12 if leaf_s > 0 else 0.0
    screened = exact_pairs > 0
    roofline = {"kernel": ("leaf_screen_kernel" if screened else "leaf_topk_tile_kernel") +
                          " (S3, summed over the tries of one step; includes prep and literal redo)",
                "bound": "fp32", "achieved": leaf_tflops, "peak": fp32_peak,
                "peak_source": "148 SM x 128 lanes x 1.965 GHz, one separately rounded FP32 op per lane-clock",
                "unit": "TFLOP/s", "frac": leaf_tflops / fp32_peak,
                # DRAM read+write of ONE launch (one try) of leaf_screen_kernel<64> on this workload, from the
                # committed ncu --set full capture (profiles/r1_q_ncu_leaf_screen_raw.csv); not re-measured here
                "traffic": 839.5e6 if (screened and args.config == "cfg3" and world == 1) else None,
                "pairs_per_step": leaf_pairs, "flops_per_pair": 3 * d,
                "exact_pairs_per_step": exact_pairs, "screen_overflow_buckets_per_step": overflow_buckets,
                "share_of_device_time": mean_stage["leaf"] / dev_mean}
    # --- dominant HBM-bound kernel: supercharge.  Algorithmic bytes per launch set:
    # rows*(P2-k) gathered vectors of (4 + d*w) B, + own lists in, + ids and dists out.
    P2 = 1 << ((k * (k + 1)).bit_length() - 1)
    rows = n / world                      # supercharge rows per rank
    sc_bytes = rows * (P2 - k) * (4 + d * w) + rows * k * (4 + w) + rows * k * (4 + w)
    sc_s = mean_stage["supercharge"] / 1e3
    roofline_hbm = {"kernel": "supercharge_fast_kernel (S5, all row chunks of one step)", "bound": "hbm",
                    "achieved": sc_bytes / sc_s / 1e9, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                    "frac": sc_bytes / sc_s / 1e9 / peak, "traffic": None,
                    "share_of_device_time": mean_stage["supercharge"] / dev_mean}
    line = {"metric": METRIC, "value": n * args.steps / dev_total_s, "unit": "points/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_total_s / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if dtype == np.float32 else "f64", "data": "synthetic",
            "config": workload_config(cfg, args.config, world),
            "e2e": {"value": n * args.steps / wall, "unit": "points/s",
                    "h2d_bytes_per_step": n * d * w, "d2h_bytes_per_step": n * k * (4 + w),
                    "ms_per_step": 1e3 * wall / args.steps,
                    "ms_per_step_with_stage_events": 1e3 * wall_instrumented / args.steps},
            "gpu_launches": launches, "stage_ms": mean_stage, "roofline": roofline, "roofline_hbm": roofline_hbm,
            "clocks": sampler.summary()}
    if args.recall_sample > 0 and world == 1:
        line["recall"] = measure_recall(gpu, pts, cfg, args.recall_sample)
    if args.pair_queries > 0 and world == 1:
        line["precomp_query_pair"] = measure_pair(gpu, pts, cfg, args.pair_queries)
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(cfg, pts, args.cpu_sample)
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
