"""Quad-space image (csrc/quad_layout.h) and the CPU model of spf_quad_kernel's SSSP and
parents phases against the oracle — no GPU needed."""
import numpy as np
import pytest

from holo_b200 import synth
from holo_b200.capi import Csr, VF_HOP, quad_image
from oracle import pyoracle
import quad_model


@pytest.mark.parametrize("V,E,seed,kw", [
    (2, 2, 1, {}),
    (5, 12, 2, {}),
    (100, 400, 3, {}),
    (120, 1400, 4, dict(cost_choices=[10, 20])),           # multi-quad chains, ECMP
    (300, 1400, 5, dict(lan_fraction=0.1)),
    (300, 1400, 6, dict(cost_choices=[10, 20], lan_fraction=0.1)),
])
def test_image_and_model_vs_oracle(built, V, E, seed, kw):
    t = synth.random_topology(V, E, synth.SEED_BASE + seed, **kw)
    csr = synth.topology_csr(t)
    q = quad_image(csr)
    assert q.eligible
    quad_model.check_image(q, csr)
    nv = csr.n_vertices
    roots = range(nv) if nv <= 120 else range(0, nv, 23)
    for r in roots:
        ref = pyoracle.csr_spf(csr, int(r))
        dist, rounds = quad_model.sssp(q, int(r))
        od, ofp, onp = quad_model.parents(q, dist, int(r), nv)
        assert np.array_equal(od, ref["dist"]), r
        assert np.array_equal(ofp, ref["first_parent"]), r
        assert np.array_equal(onp, ref["n_parents"]), r


def test_bounded_queue_leaves_work_in_the_bitmap(built):
    t = synth.random_topology(200, 1000, synth.SEED_BASE + 11)
    csr = synth.topology_csr(t)
    q = quad_image(csr)
    ref = pyoracle.csr_spf(csr, 0)
    dist, rounds_small = quad_model.sssp(q, 0, qcap=40)      # the kernel needs qcap >= 32 (one bitmap word)
    od, _, _ = quad_model.parents(q, dist, 0, csr.n_vertices)
    assert np.array_equal(od, ref["dist"])
    _, rounds_big = quad_model.sssp(q, 0)
    assert rounds_small > rounds_big


def test_isolated_vertex_and_high_degree_hub(built):
    # vertex 0: hub with 40 links (10 quads), vertex 41: isolated
    V = 42
    src, dst, cst = [], [], []
    for i in range(1, 41):
        src += [0, i]; dst += [i, 0]; cst += [i, 2 * i]
    order = np.lexsort((np.arange(len(src)), np.asarray(src)))
    src, dst, cst = np.asarray(src)[order], np.asarray(dst)[order], np.asarray(cst)[order]
    row = np.zeros(V + 1, np.uint32)
    np.add.at(row, src + 1, 1)
    row = np.cumsum(row).astype(np.uint32)
    csr = Csr(row, dst.astype(np.uint32), cst.astype(np.uint32), np.full(V, VF_HOP, np.uint8), saturate_at=0xFFFF)
    q = quad_image(csr)
    assert q.eligible and q.max_ichain == 10
    quad_model.check_image(q, csr)
    for r in (0, 7, 41):
        ref = pyoracle.csr_spf(csr, r)
        dist, _ = quad_model.sssp(q, r)
        od, ofp, onp = quad_model.parents(q, dist, r, V)
        assert np.array_equal(od, ref["dist"]) and np.array_equal(ofp, ref["first_parent"])
        assert np.array_equal(onp, ref["n_parents"])


def test_not_eligible_when_costs_do_not_pack(built):
    t = synth.random_topology(50, 200, synth.SEED_BASE + 12, cost_lo=70000, cost_hi=80000)
    csr = synth.topology_csr(t, isis=True)
    assert not quad_image(csr).eligible


def test_bucket_width_covers_the_largest_cost(built):
    t = synth.random_topology(100, 400, synth.SEED_BASE + 13, cost_lo=1, cost_hi=3)
    csr = synth.topology_csr(t)
    csr.cost[5] = 60000                      # one very expensive link
    q = quad_image(csr)
    assert q.eligible and 3 * (1 << q.shift) >= 60000
    ref = pyoracle.csr_spf(csr, 3)
    dist, _ = quad_model.sssp(q, 3)         # asserts every mark stays inside the ring
    od, _, _ = quad_model.parents(q, dist, 3, csr.n_vertices)
    assert np.array_equal(od, ref["dist"])
