"""Peer-memory exchange (hspf_xchg_*) on two GPUs of one node; skipped on a single-GPU box
(the single-GPU suite covers only that the symbols exist, tests/test_abi.py)."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_peer_exchange_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29541", str(ROOT / "scripts" / "xchg_selftest.py"),
           "--steps", "24", "--slot-mb", "4"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "xchg_selftest ok" in out.stdout


def test_peer_exchange_loopback_on_one_gpu():
    """Two exchanges of one process on one device (attached by pointer): the same push / wait /
    release sequencing as between GPUs, with data that changes every step and a pushed prefix
    shorter than the slot; runs past 2 x 70 000 pushes would not fit the old 65 536-entry flag
    table — 300 steps here, the sequence numbers are plain 32-bit values now."""
    import torch
    from holo_b200 import capi, shard
    dev = torch.device("cuda", 0)
    ctxs = [capi.Context(0), capi.Context(0)]
    nbytes = 1 << 20
    xs = shard.PeerExchange.local_pair(ctxs, 0, nbytes, 2)
    prefix = nbytes // 2
    for x in xs:
        x.set_push_bytes(prefix)
    comp = [torch.cuda.ExternalStream(c.stream, device=dev) for c in ctxs]
    cons = [torch.cuda.ExternalStream(x.consumer_stream, device=dev) for x in xs]
    mine = [[shard.raw_cuda_tensor(x.slot_ptr(b, x.rank), nbytes, dev).view(torch.int32) for b in range(2)] for x in xs]
    full = [[x.buffer_tensor(b, dev)[:, :nbytes].view(torch.int32).view(2, -1) for b in range(2)] for x in xs]
    idx = torch.arange(nbytes // 4, dtype=torch.int32, device=dev)
    bad = torch.zeros((), dtype=torch.int64, device=dev)
    npre = prefix // 4
    torch.cuda.synchronize()
    for s in range(300):
        b = s % 2
        for x in xs:
            x.acquire(b)
        for r, x in enumerate(xs):
            with torch.cuda.stream(comp[r]):
                torch.add(idx, (r + 1) * 1000003 + s * 7919, out=mine[r][b])
        for x in xs:
            x.push(b)
        for x in xs:
            x.wait(b)
        for r, x in enumerate(xs):
            with torch.cuda.stream(cons[r]):
                for o in range(2):      # own slot: whole; the peer's: the pushed prefix
                    n = nbytes // 4 if o == r else npre
                    bad += (full[r][b][o][:n] != idx[:n] + ((o + 1) * 1000003 + s * 7919)).sum()
        for x in xs:
            x.release(b)
    for x in xs:
        x.sync()
    torch.cuda.synchronize()
    assert int(bad.item()) == 0
    for x in xs:
        x.close()
    for c in ctxs:
        c.close()


def test_fused_exchange_kernel_stores_into_the_peer_slot_loopback():
    """Fused exchange on one GPU (two contexts, exchanges attached by pointer): the batch kernel of
    rank r writes dist / hops / nh_mask / job status of its jobs into its slot of rank 1-r's buffer
    while it computes; after publish + wait the peer's copy equals the local planes, and the planes
    that do not travel (first_parent, n_parents) stay untouched there."""
    import ctypes as C
    import numpy as np
    import torch
    from holo_b200 import capi, shard, synth
    from oracle import pyoracle
    dev = torch.device("cuda", 0)
    t = synth.random_topology(403, 1800, synth.SEED_BASE + 61)      # odd row length: unaligned heads and tails
    csr = synth.topology_csr(t)
    V, n = csr.n_vertices, 96
    al = lambda x: (x + 255) // 256 * 256
    off, tot = {}, 0
    for k, b in (("dist", 2), ("hops", 2), ("nh", 2), ("status", None), ("fp", 2), ("npar", 2)):
        off[k] = tot
        tot += al(n * 4 if b is None else n * V * b)
    ctxs = [capi.Context(0), capi.Context(0)]
    graphs = [c.upload(csr) for c in ctxs]
    xs = shard.PeerExchange.local_pair(ctxs, 0, tot, 2)
    roots = [np.arange(r * n, (r + 1) * n, dtype=np.uint32) % V for r in range(2)]
    d_roots = [torch.from_numpy(x.astype(np.int64)).to(torch.int32).to(dev) for x in roots]
    for step in range(3):
        b = step % 2
        for r, (c, x) in enumerate(zip(ctxs, xs)):
            base = x.slot_ptr(b, r)
            shard.raw_cuda_tensor(base, tot, dev).zero_()
        torch.cuda.synchronize()
        for r, (c, x, g) in enumerate(zip(ctxs, xs, graphs)):
            x.acquire_direct(b)
            c.set_peer_slots(x.peer_deltas(b))
            js = capi.JobsStruct()
            js.n_jobs = n
            js.roots = C.cast(d_roots[r].data_ptr(), C.POINTER(C.c_uint32))
            base = x.slot_ptr(b, r)
            rs = capi.Result16Struct()
            for k, f in (("dist", "dist"), ("hops", "hops"), ("fp", "first_parent"), ("npar", "n_parents"), ("nh", "nh_mask")):
                setattr(rs, f, C.cast(base + off[k], C.POINTER(C.c_uint16)))
            rs.job_status = C.cast(base + off["status"], C.POINTER(C.c_uint32))
            c.run_device16(g, js, rs, sync=False)
            c.set_peer_slots([])
            x.publish(b)
        for x in xs:
            x.wait(b)
        for x in xs:
            x.sync()
        torch.cuda.synchronize()
        for r, x in enumerate(xs):
            mine = shard.raw_cuda_tensor(x.slot_ptr(b, r), tot, dev).cpu().numpy()
            theirs = shard.raw_cuda_tensor(xs[1 - r].slot_ptr(b, r), tot, dev).cpu().numpy()   # rank r's slot on the peer
            for k in ("dist", "hops", "nh"):
                a0, a1 = off[k], off[k] + n * V * 2
                assert np.array_equal(mine[a0:a1], theirs[a0:a1]), (step, r, k)
            assert np.array_equal(mine[off["status"]: off["status"] + 4 * n], theirs[off["status"]: off["status"] + 4 * n])
            assert not theirs[off["fp"]: off["fp"] + n * V * 2].any()          # does not travel
            d = mine[off["dist"]: off["dist"] + n * V * 2].view(np.uint16).reshape(n, V)
            ref = pyoracle.csr_spf(csr, int(roots[r][5]))
            assert np.array_equal(d[5], ref["dist"].astype(np.uint16))
        for x in xs:
            x.release(b)
    for x in xs:
        x.sync()
    torch.cuda.synchronize()
    for x in xs:
        x.close()
    for g in graphs:
        g.free()
    for c in ctxs:
        c.close()
