import sys, os, ctypes, time, numpy as np
sys.path.insert(0, "/root/repo")
import bench
from approximatenn_b200.api import gpu_backend, srandom, stage_times, _libc
cfg = bench.CONFIGS[os.environ.get("CFG", "cfg3")]; n, d, k, tries, dtype = cfg
pts = bench.synth_points(n, d, dtype)
gpu = gpu_backend(dtype); gpu.lib.gpu_init(); gpu.lib.annh_set_timing(1)
gpu.lib.annb_literal_rows.argtypes = [ctypes.c_void_p, ctypes.c_int]
gpu.lib.annb_leaf_pairs.restype = ctypes.c_ulonglong
gpu.lib.annb_leaf_exact_pairs.restype = ctypes.c_ulonglong
gpu.lib.annb_leaf_overflow_buckets.restype = ctypes.c_ulonglong
def run(keep=False):
    dptr = ctypes.c_void_p(); srandom(1001)
    ids = gpu.precomp_raw(n, k, d, pts.ctypes.data, tries, *bench.ROT, None, ctypes.byref(dptr))
    r = None
    if keep:
        r = (np.ctypeslib.as_array(ctypes.cast(ids, ctypes.POINTER(ctypes.c_size_t)), (n, k)).copy(),
             np.ctypeslib.as_array(ctypes.cast(dptr, ctypes.POINTER(ctypes.c_float)), (n, k)).copy())
    _libc.free(ids); _libc.free(dptr); return stage_times(gpu), r
run()
out = (ctypes.c_ulonglong * 3)()
gpu.lib.annb_literal_rows(out, 1); gpu.lib.annb_leaf_pairs(1); gpu.lib.annb_leaf_exact_pairs(1); gpu.lib.annb_leaf_overflow_buckets(1)
st, res = run(keep=True)
gpu.lib.annb_literal_rows(out, 1)
import hashlib
print("RESULT dbg=%s screen=%s leaf=%.2f total_dev=%.2f lit=%s exact=%d ovf=%d hash=%s/%s" % (
    os.environ.get("ANN_B200_SCREEN_DBG", "0"), os.environ.get("ANN_B200_SCREEN", "0"), st["leaf"],
    sum(v for k_, v in st.items() if k_ not in ("upload", "first_to_last_event")), list(out),
    gpu.lib.annb_leaf_exact_pairs(0), gpu.lib.annb_leaf_overflow_buckets(0),
    hashlib.md5(res[0].tobytes()).hexdigest()[:10], hashlib.md5(res[1].tobytes()).hexdigest()[:10]))
