"""How long does page-locking the caller's (malloc()ed, touched) array take on this host?  Decides
whether `cudaHostRegister` + a direct DMA could beat the staged upload of ann_ingest.c."""
import ctypes, sys, time
import numpy as np
import torch

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rt = torch.cuda.cudart()
torch.cuda.init()
libc = ctypes.CDLL(None)
libc.malloc.restype = ctypes.c_void_p
libc.malloc.argtypes = [ctypes.c_size_t]
libc.free.argtypes = [ctypes.c_void_p]
n = mb << 20
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
for rep in range(4):
    p = libc.malloc(n)
    a = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(n,))
    a[::4096] = 1                                            # touched, as a caller's array is
    t0 = time.perf_counter()
    err = rt.cudaHostRegister(p, n, 0)
    t1 = time.perf_counter()
    src = torch.from_numpy(a)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    dev.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    err2 = rt.cudaHostUnregister(p)
    t4 = time.perf_counter()
    print(f"{mb} MB: register {1e3 * (t1 - t0):.2f} ms (err {int(err)}), copy {1e3 * (t3 - t2):.2f} ms, "
          f"unregister {1e3 * (t4 - t3):.2f} ms (err {int(err2)})", flush=True)
    libc.free(p)
