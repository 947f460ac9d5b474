import sys, os, ctypes, time, numpy as np
sys.path.insert(0, "/root/repo")
import bench
from approximatenn_b200.api import gpu_backend, srandom, stage_times, _libc
n, d, k, tries, dtype = bench.CONFIGS["cfg1"]
pts = bench.synth_points(n, d, dtype)
gpu = gpu_backend(dtype); gpu.lib.gpu_init()
def run():
    dptr = ctypes.c_void_p(); srandom(1001)
    t0 = time.perf_counter()
    ids = gpu.precomp_raw(n, k, d, pts.ctypes.data, tries, *bench.ROT, None, ctypes.byref(dptr))
    t1 = time.perf_counter()
    _libc.free(ids); _libc.free(dptr)
    print("call %.3f ms" % ((t1-t0)*1e3), file=sys.stderr)
for i in range(4): run()
gpu.lib.annh_set_timing(1); run(); print(stage_times(gpu), file=sys.stderr)
