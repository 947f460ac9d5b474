"""Multi-GPU bootstrap for the C library (include/annb200_dist.h): one process per GPU.

torch.distributed is only the plumbing that ships NCCL's 128-byte unique id from rank 0 to
the other ranks; the data-path collectives run inside the library (csrc/ann_dist.c).
"""
from __future__ import annotations

import ctypes

import numpy as np


def declare(lib) -> None:
    lib.annb200_dist_unique_id.argtypes = [ctypes.c_char_p]
    lib.annb200_dist_unique_id.restype = None
    lib.annb200_dist_init.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_char_p]
    lib.annb200_dist_init.restype = None
    lib.annb200_dist_shutdown.restype = None
    lib.annb200_dist_gather.argtypes = [ctypes.c_int]
    lib.annb200_dist_gather.restype = None
    lib.annb200_dist_try_owner.argtypes = [ctypes.c_int, ctypes.c_int]
    lib.annb200_dist_try_owner.restype = ctypes.c_int
    lib.annb200_dist_slice.argtypes = [ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                       ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t)]
    lib.annb200_dist_slice.restype = None
    lib.annb200_dist_admit.argtypes = [ctypes.c_size_t, ctypes.c_int, ctypes.c_int]
    lib.annb200_dist_admit.restype = ctypes.c_int


def row_slice(lib, n: int, rank: int, world: int):
    lo, hi = ctypes.c_size_t(), ctypes.c_size_t()
    lib.annb200_dist_slice(n, rank, world, ctypes.byref(lo), ctypes.byref(hi))
    return lo.value, hi.value


def try_owner(lib, t: int, world: int) -> int:
    return lib.annb200_dist_try_owner(t, world)


def admitted(lib, k: int, tries: int, t: int) -> int:
    return lib.annb200_dist_admit(k, tries, t)


def init_from_torch(lib, gather_full: bool = False) -> None:
    """Call on every rank after torch.distributed.init_process_group()."""
    import torch
    import torch.distributed as dist

    declare(lib)
    rank, world = dist.get_rank(), dist.get_world_size()
    buf = ctypes.create_string_buffer(128)
    if rank == 0:
        lib.annb200_dist_unique_id(buf)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.from_numpy(np.frombuffer(buf.raw, dtype=np.uint8).copy()).to(dev)
    dist.broadcast(t, src=0)
    lib.annb200_dist_init(rank, world, t.cpu().numpy().tobytes())
    lib.annb200_dist_gather(1 if gather_full else 0)
