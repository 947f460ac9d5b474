import sys, os, ctypes, time, numpy as np
sys.path.insert(0, "/root/repo")
import bench
from approximatenn_b200.api import gpu_backend, srandom, SaveT, _libc
cfg = bench.CONFIGS["cfg3"]; n, d, k, tries, dtype = cfg
pts = bench.synth_points(n, d, dtype)
gpu = gpu_backend(dtype); gpu.lib.gpu_init()
def run():
    dptr = ctypes.c_void_p(); srandom(1001); sv = SaveT()
    t0 = time.perf_counter()
    ids = gpu.precomp_raw(n, k, d, pts.ctypes.data, tries, *bench.ROT, ctypes.byref(sv), ctypes.byref(dptr))
    t1 = time.perf_counter()
    gpu._free_save(ctypes.byref(sv)); _libc.free(ids); _libc.free(dptr)
    print("call with save %.2f ms" % ((t1-t0)*1e3), file=sys.stderr)
for i in range(3): run()
