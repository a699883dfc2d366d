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
