"""Job sharding and result exchange for multi-GPU batches.

Jobs (roots / perturbations) are independent, so ranks take contiguous job ranges over a
replicated graph; the one exchange step is an all-gather of the per-job result planes.
Backend agnostic (NCCL on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def job_range(n_jobs: int, rank: int, world: int):
    """Contiguous, balanced range [lo, hi) of rank `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_jobs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_planes(planes: dict, order=("dist", "hops", "first_parent", "n_parents", "nh_mask", "job_status")):
    """One contiguous uint8 tensor holding every plane of this rank (256-byte aligned
    sections) and the layout needed to unpack it."""
    layout, tot = [], 0
    for k in order:
        t = planes[k].contiguous()
        nbytes = t.numel() * t.element_size()
        layout.append((k, tot, nbytes, t.dtype, tuple(t.shape)))
        tot += (nbytes + 255) // 256 * 256
    buf = torch.empty(tot, dtype=torch.uint8, device=planes[order[0]].device)
    for (k, off, nbytes, _dt, _shape) in layout:
        buf[off: off + nbytes] = planes[k].contiguous().view(torch.uint8).view(-1)
    return buf, layout


def unpack_planes(buf: torch.Tensor, layout):
    return {k: buf[off: off + nbytes].view(dt).view(shape) for (k, off, nbytes, dt, shape) in layout}


def all_gather_planes(planes: dict, world: int):
    """All ranks hold equally sized shards: returns {plane: [world, ...]} on every rank."""
    buf, layout = pack_planes(planes)
    out = torch.empty((world, buf.numel()), dtype=torch.uint8, device=buf.device)
    dist.all_gather_into_tensor(out.view(-1), buf)
    res = {}
    for (k, off, nbytes, dt, shape) in layout:
        res[k] = torch.stack([out[r, off: off + nbytes].view(dt).view(shape) for r in range(world)])
    return res
