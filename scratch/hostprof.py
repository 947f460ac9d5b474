import sys, os, ctypes, time, numpy as np
sys.path.insert(0, "/root/repo")
import bench, torch
from approximatenn_b200.api import gpu_backend, srandom, stage_times, _libc
cfg = bench.CONFIGS["cfg3"]; n, d, k, tries, dtype = cfg
host = torch.empty((n, d), dtype=torch.float32, pin_memory=True); pts = host.numpy(); pts[:] = bench.synth_points(n, d, dtype)
gpu = gpu_backend(dtype); gpu.lib.gpu_init(); gpu.lib.annh_set_timing(1)
def run():
    dptr = ctypes.c_void_p(); srandom(1001)
    t0 = time.perf_counter()
    ids = gpu.precomp_raw(n, k, d, pts.ctypes.data, tries, *bench.ROT, None, ctypes.byref(dptr))
    t1 = time.perf_counter()
    _libc.free(ids); _libc.free(dptr)
    t2 = time.perf_counter()
    print("call %.2f ms, free %.2f ms" % ((t1-t0)*1e3, (t2-t1)*1e3), file=sys.stderr)
for i in range(3): run()
