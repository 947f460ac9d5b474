import sys, os, ctypes, time, numpy as np
sys.path.insert(0, "/root/repo")
import bench, torch, torch.distributed as dist
from approximatenn_b200.api import gpu_backend, srandom, _libc
from approximatenn_b200 import dist as adist
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local); os.environ["ANN_B200_DEVICE"] = str(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = bench.CONFIGS[os.environ.get("CFG", "cfg3")]; n, d, k, tries, dtype = cfg
host = torch.empty((n, d), dtype=torch.float32, pin_memory=True); pts = host.numpy(); pts[:] = bench.synth_points(n, d, dtype)
gpu = gpu_backend(dtype); gpu.lib.gpu_init(); adist.init_from_torch(gpu.lib)
def run(tag):
    dptr = ctypes.c_void_p(); srandom(1001)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    ids = gpu.precomp_raw(n, k, d, pts.ctypes.data, tries, *bench.ROT, None, ctypes.byref(dptr))
    t1 = time.perf_counter()
    _libc.free(ids); _libc.free(dptr)
    if dist.get_rank() == 0: print("%s call %.2f ms" % (tag, (t1-t0)*1e3), file=sys.stderr, flush=True)
for i in range(4): run("warm%d" % i)
os.environ["ANN_B200_HOSTPROF"] = "1"
