import sys, os, ctypes, numpy as np
sys.path.insert(0, "/root/repo")
import bench
from approximatenn_b200.api import gpu_backend, srandom, stage_times, _libc
n, d, k, _, dtype = bench.CONFIGS["cfg3"]
pts = bench.synth_points(n, d, dtype)
gpu = gpu_backend(dtype); gpu.lib.gpu_init(); gpu.lib.annh_set_timing(1)
for T in (1, 8):
    for rep in range(2):
        dptr = ctypes.c_void_p(); srandom(1001)
        ids = gpu.precomp_raw(n, k, d, pts.ctypes.data, T, *bench.ROT, None, ctypes.byref(dptr))
        _libc.free(ids); _libc.free(dptr)
    st = stage_times(gpu); print("T", T, {k_: round(v, 3) for k_, v in st.items() if k_ in ("means", "hash", "buckets", "leaf", "merge", "supercharge")})
