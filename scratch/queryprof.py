import sys, os, ctypes, time, numpy as np
sys.path.insert(0, "/root/repo")
import bench
from approximatenn_b200.api import gpu_backend, srandom, SaveT, _libc
cfg = bench.CONFIGS["cfg3"]; n, d, k, tries, dtype = cfg
pts = bench.synth_points(n, d, dtype)
y = np.random.default_rng(3).standard_normal((65536, d), dtype=np.float32)
gpu = gpu_backend(dtype); gpu.lib.gpu_init()
dptr = ctypes.c_void_p(); srandom(1001); sv = SaveT()
ids = gpu.precomp_raw(n, k, d, pts.ctypes.data, tries, *bench.ROT, ctypes.byref(sv), ctypes.byref(dptr)); _libc.free(ids); _libc.free(dptr)
for i in range(3):
    dptr = ctypes.c_void_p(); t0 = time.perf_counter()
    ids = gpu._query(ctypes.byref(sv), pts.ctypes.data, 65536, y.ctypes.data, ctypes.byref(dptr))
    print("query call %.2f ms" % ((time.perf_counter()-t0)*1e3), file=sys.stderr)
    _libc.free(ids); _libc.free(dptr)
