#!/usr/bin/env python
"""Self-test of the peer-memory exchange (hspf_xchg_*) on N GPUs of one node.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29533 scripts/xchg_selftest.py [--steps 40] [--slot-mb 8]

Every step each rank fills its own slot with a pattern that depends on (rank, step), pushes
it, and checks on the consumer stream that every slot of the buffer carries that step's
pattern of its owner — data that changes every step, so a stale or torn slot is caught.
Prints one line `xchg_selftest ok ...` on rank 0 and exits 0, or exits 1."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from holo_b200 import capi, shard  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--slot-mb", type=int, default=8)
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    lr = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    ctx = capi.Context(lr)
    nbytes = a.slot_mb << 20
    x = shard.PeerExchange(ctx, lr, rank, world, nbytes, 2)
    compute = torch.cuda.ExternalStream(ctx.stream, device=dev)
    cons = torch.cuda.ExternalStream(x.consumer_stream, device=dev)
    mine = [shard.raw_cuda_tensor(x.slot_ptr(b, rank), nbytes, dev).view(torch.int32) for b in range(2)]
    full = [x.buffer_tensor(b, dev).view(world, -1)[:, :nbytes].view(torch.int32).view(world, -1) for b in range(2)]
    idx = torch.arange(nbytes // 4, dtype=torch.int32, device=dev)
    bad = torch.zeros((), dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    dist.barrier()
    for s in range(a.steps):
        b = s % 2
        x.acquire(b)
        with torch.cuda.stream(compute):
            torch.add(idx, (rank + 1) * 1000003 + s * 7919, out=mine[b])
        x.push(b)
        x.wait(b)
        with torch.cuda.stream(cons):
            for r in range(world):
                bad += (full[b][r] != idx + ((r + 1) * 1000003 + s * 7919)).sum()
        x.release(b)
    x.sync()
    torch.cuda.synchronize()
    tot = bad.clone()
    dist.all_reduce(tot)
    dist.barrier()
    x.close()
    ctx.close()
    ok = int(tot.item()) == 0
    if rank == 0:
        print(f"xchg_selftest {'ok' if ok else 'FAILED'} world={world} steps={a.steps} slot_mb={a.slot_mb} mismatches={int(tot.item())}")
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
