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
