"""CPU, world_size 2, gloo: the multi-GPU host logic (job sharding + all-gather of the
result planes) without a GPU.  The per-rank planes are produced by the oracle here (test
infrastructure standing in for the device kernel); what is under test is
holo_b200/shard.py, the code bench.py runs over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from holo_b200 import shard, synth
from oracle import pyoracle


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _planes_for(csr, roots):
    n, V = len(roots), csr.n_vertices
    out = dict(dist=np.empty((n, V), np.int32), hops=np.empty((n, V), np.int16), first_parent=np.empty((n, V), np.int32),
               n_parents=np.empty((n, V), np.int16), nh_mask=np.empty((n, V, 1), np.int64), job_status=np.zeros(n, np.int32))
    for j, r in enumerate(roots):
        ref = pyoracle.csr_spf_heap(csr, int(r))
        out["dist"][j] = ref["dist"].view(np.int32)
        out["hops"][j] = ref["hops"].view(np.int16)
        out["first_parent"][j] = ref["first_parent"].view(np.int32)
        out["n_parents"][j] = ref["n_parents"].view(np.int16)
        out["nh_mask"][j] = ref["nh_mask"].view(np.int64)
    return {k: torch.from_numpy(v) for k, v in out.items()}


def _worker(rank, world, port, n_jobs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t = synth.random_topology(200, 900, synth.SEED_BASE + 21, lan_fraction=0.1)
    csr = synth.topology_csr(t)
    lo, hi = shard.job_range(n_jobs, rank, world)
    roots = np.arange(lo, hi) % csr.n_vertices
    mine = _planes_for(csr, roots)
    full = shard.all_gather_planes(mine, world)
    if rank == 0:
        q.put({k: v.numpy() for k, v in full.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_job_range_is_a_balanced_partition():
    for n in (0, 1, 7, 1000, 1001):
        for w in (1, 2, 3, 8):
            r = [shard.job_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(120)
def test_world2_gloo_sharded_batch_equals_single_process():
    world, n_jobs = 2, 24
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_jobs, q)) for r in range(world)]
    for p in procs:
        p.start()
    full = q.get(timeout=100)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    t = synth.random_topology(200, 900, synth.SEED_BASE + 21, lan_fraction=0.1)
    csr = synth.topology_csr(t)
    ref = _planes_for(csr, np.arange(n_jobs) % csr.n_vertices)
    for k, v in ref.items():
        got = full[k].reshape((-1,) + tuple(v.shape[1:]))
        assert np.array_equal(got, v.numpy()), k
