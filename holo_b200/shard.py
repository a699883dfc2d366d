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


class PeerExchange:
    """ctypes face of `hspf_xchg_*` (include/holo_spf.h): all-gather of the per-rank result
    planes with the copy engines over NVLink peer memory.  The 64-byte IPC handles are
    exchanged through the process group (any backend that can all-gather a byte tensor)."""

    def __init__(self, ctx, device_index: int, rank: int, world: int, slot_bytes: int, n_buffers: int = 2):
        import ctypes as C
        self.C, self.lib, self.rank, self.world, self.n_buffers = C, ctx.lib, rank, world, n_buffers
        lib = self.lib
        lib.hspf_xchg_create.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_size_t, C.c_uint32,
                                         C.POINTER(C.c_void_p), C.c_char_p]
        lib.hspf_xchg_attach.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p]
        lib.hspf_xchg_slot.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        lib.hspf_xchg_slot.restype = C.c_void_p
        lib.hspf_xchg_slot_bytes.argtypes = [C.c_void_p]
        lib.hspf_xchg_slot_bytes.restype = C.c_size_t
        for f in ("acquire", "push", "wait", "release"):
            getattr(lib, "hspf_xchg_" + f).argtypes = [C.c_void_p, C.c_uint32]
        lib.hspf_xchg_consumer_stream.argtypes = [C.c_void_p]
        lib.hspf_xchg_consumer_stream.restype = C.c_void_p
        lib.hspf_xchg_sync.argtypes = [C.c_void_p]
        lib.hspf_xchg_destroy.argtypes = [C.c_void_p]
        lib.hspf_xchg_last_error.argtypes = [C.c_void_p]
        lib.hspf_xchg_last_error.restype = C.c_char_p
        self.handle = C.c_void_p()
        mine = C.create_string_buffer(64)
        rc = lib.hspf_xchg_create(ctx.handle, device_index, rank, world, slot_bytes, n_buffers,
                                  C.byref(self.handle), mine)
        ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=f"cuda:{device_index}")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            self.close()
            raise RuntimeError(f"hspf_xchg_create failed on some rank (local rc={rc})")
        h = torch.frombuffer(bytearray(mine.raw), dtype=torch.uint8).to(f"cuda:{device_index}")
        allh = torch.empty((world, 64), dtype=torch.uint8, device=h.device)
        dist.all_gather_into_tensor(allh.view(-1), h)
        allh = allh.cpu().numpy()
        rc = 0
        for r in range(world):
            if r != rank and rc == 0:
                rc = lib.hspf_xchg_attach(self.handle, r, allh[r].tobytes())
        ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=h.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            msg = self.last_error()
            dist.barrier()
            self.close()
            raise RuntimeError(f"hspf_xchg_attach failed on some rank (local rc={rc}: {msg})")
        self.slot_bytes = int(lib.hspf_xchg_slot_bytes(self.handle))

    @classmethod
    def local_pair(cls, ctxs, device_index: int, slot_bytes: int, n_buffers: int = 2):
        """Two exchanges of ONE process on one device, attached to each other by pointer
        (hspf_xchg_attach_ptr): exercises the push / wait / release sequencing on a single GPU."""
        import ctypes as C
        xs = []
        for rank, ctx in enumerate(ctxs):
            x = cls.__new__(cls)
            x.C, x.lib, x.rank, x.world, x.n_buffers = C, ctx.lib, rank, len(ctxs), n_buffers
            lib = x.lib
            lib.hspf_xchg_create.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_size_t, C.c_uint32,
                                             C.POINTER(C.c_void_p), C.c_char_p]
            lib.hspf_xchg_attach_ptr.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
            lib.hspf_xchg_base.argtypes = [C.c_void_p]
            lib.hspf_xchg_base.restype = C.c_void_p
            lib.hspf_xchg_slot.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
            lib.hspf_xchg_slot.restype = C.c_void_p
            lib.hspf_xchg_slot_bytes.argtypes = [C.c_void_p]
            lib.hspf_xchg_slot_bytes.restype = C.c_size_t
            lib.hspf_xchg_set_push_bytes.argtypes = [C.c_void_p, C.c_size_t]
            for f in ("acquire", "push", "wait", "release"):
                getattr(lib, "hspf_xchg_" + f).argtypes = [C.c_void_p, C.c_uint32]
            lib.hspf_xchg_consumer_stream.argtypes = [C.c_void_p]
            lib.hspf_xchg_consumer_stream.restype = C.c_void_p
            lib.hspf_xchg_sync.argtypes = [C.c_void_p]
            lib.hspf_xchg_destroy.argtypes = [C.c_void_p]
            lib.hspf_xchg_last_error.argtypes = [C.c_void_p]
            lib.hspf_xchg_last_error.restype = C.c_char_p
            x.handle = C.c_void_p()
            rc = lib.hspf_xchg_create(ctx.handle, device_index, rank, len(ctxs), slot_bytes, n_buffers,
                                      C.byref(x.handle), C.create_string_buffer(64))
            if rc != 0:
                raise RuntimeError(f"hspf_xchg_create rc={rc}")
            x.slot_bytes = int(lib.hspf_xchg_slot_bytes(x.handle))
            xs.append(x)
        for x in xs:
            for y in xs:
                if x is not y:
                    x._ck(x.lib.hspf_xchg_attach_ptr(x.handle, y.rank, x.lib.hspf_xchg_base(y.handle)), "attach_ptr")
        return xs

    def acquire_direct(self, b):
        self.lib.hspf_xchg_acquire_direct.argtypes = [self.C.c_void_p, self.C.c_uint32]
        self._ck(self.lib.hspf_xchg_acquire_direct(self.handle, b), "acquire_direct")

    def publish(self, b):
        self.lib.hspf_xchg_publish.argtypes = [self.C.c_void_p, self.C.c_uint32]
        self._ck(self.lib.hspf_xchg_publish(self.handle, b), "publish")

    def peer_deltas(self, b):
        """Address differences (peer copy of this rank's slot - local slot) for hspf_ctx_set_peer_slots."""
        C = self.C
        self.lib.hspf_xchg_peer_deltas.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_int64), C.POINTER(C.c_uint32)]
        d = (C.c_int64 * 8)()
        n = C.c_uint32()
        self._ck(self.lib.hspf_xchg_peer_deltas(self.handle, b, d, C.byref(n)), "peer_deltas")
        return [int(d[i]) for i in range(n.value)]

    def set_push_bytes(self, nbytes: int):
        self.lib.hspf_xchg_set_push_bytes.argtypes = [self.C.c_void_p, self.C.c_size_t]
        self._ck(self.lib.hspf_xchg_set_push_bytes(self.handle, nbytes), "set_push_bytes")

    def last_error(self) -> str:
        return (self.lib.hspf_xchg_last_error(self.handle) or b"").decode()

    def _ck(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"hspf_xchg_{what} rc={rc}: {self.last_error()}")

    def slot_ptr(self, buffer: int, slot: int) -> int:
        return int(self.lib.hspf_xchg_slot(self.handle, buffer, slot))

    def acquire(self, b): self._ck(self.lib.hspf_xchg_acquire(self.handle, b), "acquire")
    def push(self, b): self._ck(self.lib.hspf_xchg_push(self.handle, b), "push")
    def wait(self, b): self._ck(self.lib.hspf_xchg_wait(self.handle, b), "wait")
    def release(self, b): self._ck(self.lib.hspf_xchg_release(self.handle, b), "release")
    def sync(self): self._ck(self.lib.hspf_xchg_sync(self.handle), "sync")

    @property
    def consumer_stream(self) -> int:
        return int(self.lib.hspf_xchg_consumer_stream(self.handle) or 0)

    def buffer_tensor(self, b: int, device) -> torch.Tensor:
        """uint8 view [world, slot_bytes] of local buffer b (zero-copy, for checks)."""
        return raw_cuda_tensor(self.slot_ptr(b, 0), self.world * self.slot_bytes, device).view(self.world, self.slot_bytes)

    def close(self):
        if self.handle:
            self.lib.hspf_xchg_destroy(self.handle)
            self.handle = None


class _RawCuda:
    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def raw_cuda_tensor(ptr: int, nbytes: int, device) -> torch.Tensor:
    """Zero-copy uint8 tensor over device memory this process did not get from torch."""
    return torch.as_tensor(_RawCuda(ptr, nbytes), device=device)
