import sys, os, ctypes, time, numpy as np
sys.path.insert(0, "/root/repo")
import bench
from approximatenn_b200.api import gpu_backend, srandom, stage_times, _libc
cfg = bench.CONFIGS["cfg3"]; n, d, k, tries, dtype = cfg
pts = bench.synth_points(n, d, dtype)
gpu = gpu_backend(dtype); gpu.lib.gpu_init(); gpu.lib.annh_set_timing(1)
gpu.lib.annb_literal_rows.argtypes = [ctypes.c_void_p, ctypes.c_int]
def run():
    dptr = ctypes.c_void_p(); srandom(1001)
    ids = gpu.precomp_raw(n, k, d, pts.ctypes.data, tries, *bench.ROT, None, ctypes.byref(dptr))
    _libc.free(ids); _libc.free(dptr); return stage_times(gpu)
run()
out = (ctypes.c_ulonglong * 3)()
gpu.lib.annb_literal_rows(out, 1)
st = run()
gpu.lib.annb_literal_rows(out, 1)
print("stage ms", {k_: round(v, 2) for k_, v in st.items()}, "literal rows S3,S4,S5:", list(out))
