#!/usr/bin/env python
"""Host-to-device bandwidth of ordinary pinned memory against write-combined pinned memory, all ranks at once.

    python -m torch.distributed.run --nproc-per-node 8 tools/h2d_wc_probe.py
"""
import ctypes
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cudart = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else ctypes.CDLL("libcudart.so")
N = 960_000_000                       # bytes: the PCM-16 batch of the bench (10 000 x 48 000 x 2)


def host_alloc(nbytes, flags):
    p = ctypes.c_void_p()
    rc = cudart.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flags))
    assert rc == 0, rc
    arr = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), (nbytes,))
    return torch.from_numpy(arr), p


def bench(t, dst, reps=8):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dst.copy_(t, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return N * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


dst = torch.empty(N, dtype=torch.uint8, device="cuda")
plain = torch.empty(N, dtype=torch.uint8).pin_memory()
plain.fill_(3)
wc, _ = host_alloc(N, 0x04)           # cudaHostAllocWriteCombined
wc.fill_(3)
dflt, _ = host_alloc(N, 0x00)
dflt.fill_(3)
for name, t in (("torch pin_memory", plain), ("cudaHostAlloc default", dflt), ("cudaHostAlloc write-combined", wc)):
    for _ in range(2):
        g = bench(t, dst)
    vals = [None] * world
    if world > 1:
        dist.all_gather_object(vals, g)
    else:
        vals = [g]
    if rank == 0:
        print(f"{name}: per-rank GB/s {[round(v, 1) for v in vals]} total {sum(vals):.1f}", flush=True)
if world > 1:
    dist.destroy_process_group()
