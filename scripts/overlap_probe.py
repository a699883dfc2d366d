#!/usr/bin/env python
"""Does an NCCL all-gather overlap the persistent batch kernel?  (2+ ranks, tuning aid)"""
import os
import sys
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
sys.path.insert(0, ".")
import ctypes as C  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from holo_b200 import capi, synth  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
t = synth.random_topology(10000, 40000, synth.SEED_BASE + 2)
csr = synth.topology_csr(t)
ctx = capi.Context(lr)
g = ctx.upload(csr)
n, V = 1000, csr.n_vertices
roots = torch.arange(n, dtype=torch.int32, device=dev)
buf = torch.empty(n * V * 20 + 4096, dtype=torch.uint8, device=dev)
js = capi.JobsStruct(); js.n_jobs = n; js.roots = C.cast(roots.data_ptr(), C.POINTER(C.c_uint32))
rs = capi.ResultStruct(); p = buf.data_ptr()
rs.dist = C.cast(p, C.POINTER(C.c_uint32)); p += n * V * 4
rs.first_parent = C.cast(p, C.POINTER(C.c_uint32)); p += n * V * 4
rs.nh_mask = C.cast(p, C.POINTER(C.c_uint64)); p += n * V * 8
rs.hops = C.cast(p, C.POINTER(C.c_uint16)); p += n * V * 2
rs.n_parents = C.cast(p, C.POINTER(C.c_uint16)); p += n * V * 2
rs.job_status = C.cast(p, C.POINTER(C.c_uint32)); rs.nh_words = 1
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
comm = torch.cuda.Stream(device=dev, priority=-1)
src = torch.empty(200_000_000, dtype=torch.uint8, device=dev)
dst = torch.empty(world * 200_000_000, dtype=torch.uint8, device=dev)


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def k_only():
    ctx.run_device(g, js, rs, sync=False)


def ag_only():
    with torch.cuda.stream(comm):
        dist.all_gather_into_tensor(dst, src)


def both():
    ctx.run_device(g, js, rs, sync=False)
    with torch.cuda.stream(comm):
        dist.all_gather_into_tensor(dst, src)


for name, fn in (("kernel only", k_only), ("all-gather only", ag_only), ("kernel || all-gather (independent streams)", both)):
    ms = timed(fn)
    if rank == 0:
        print(f"{name}: {ms:.3f} ms / iteration")
for r in (8, 24):
    ctx.reserve_sms(r)
    ms = timed(both)
    if rank == 0:
        print(f"kernel || all-gather with {r} SMs reserved: {ms:.3f} ms")
dist.destroy_process_group()
