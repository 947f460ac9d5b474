"""Wall time of bare precomp_gpu calls (pageable input) under different host-thread settings."""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
from bench import CONFIGS, ROT, synth_points
from approximatenn_b200.api import gpu_backend, srandom, _libc
n, d, k, tries, dtype = CONFIGS["cfg3"]
g = gpu_backend(dtype)
pts = synth_points(n, d, dtype)
def call():
    dptr = ctypes.c_void_p(); srandom(1001)
    t0 = time.perf_counter()
    ids = g.precomp_raw(n, k, d, pts.ctypes.data, tries, *ROT, None, ctypes.byref(dptr))
    dt = time.perf_counter() - t0
    _libc.free(ids); _libc.free(dptr)
    return dt
for _ in range(3): call()
for ing in ():
    os.environ["ANN_B200_INGEST_THREADS"] = str(ing)
    call()
    print("ingest threads", ing, "call ms %.2f" % (1e3 * min(call() for _ in range(5))), flush=True)
os.environ["ANN_B200_INGEST_THREADS"] = "8"
for eg in (4, 6, 8):
    os.environ["ANN_B200_HOST_THREADS"] = str(eg)
    call()
    print("egress threads", eg, "call ms %.2f" % (1e3 * min(call() for _ in range(5))), flush=True)
