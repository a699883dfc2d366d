"""CPU fuzz of the pointer-jumping next-hop algorithm (phase 3J of the kernel, modelled step by
step in tests/jump_model.py) against the reference-faithful oracle: many small graphs with
LANs, ECMP-rich costs, IS-IS pseudonode rules and parallel links, every vertex as root."""
import numpy as np
import pytest

from holo_b200 import synth
from oracle import pyoracle

from jump_model import jump_phase

SHAPES = [
    dict(V=12, E=40, kw=dict(cost_choices=[1])),
    dict(V=30, E=100, kw=dict(cost_choices=[5])),
    dict(V=40, E=160, kw=dict(cost_choices=[10, 20])),
    dict(V=40, E=120, kw=dict(cost_choices=[10, 20], lan_fraction=0.3)),
    dict(V=60, E=260, kw=dict(cost_choices=[3], lan_fraction=0.2)),
    dict(V=80, E=300, kw=dict()),
    dict(V=80, E=400, kw=dict(cost_lo=1, cost_hi=3)),
    dict(V=150, E=700, kw=dict(cost_choices=[7, 14, 21], lan_fraction=0.1)),
]


def nh_int(row):
    x = 0
    for w, word in enumerate(row):
        x |= int(word) << (64 * w)
    return x


@pytest.mark.parametrize("shape", range(len(SHAPES)))
@pytest.mark.parametrize("isis", [False, True])
def test_jump_algorithm_matches_faithful_oracle(shape, isis):
    sh = SHAPES[shape]
    checked = ecmp_total = 0
    for seed in range(16):
        t = synth.random_topology(sh["V"], sh["E"], synth.SEED_BASE + 1000 + 17 * shape + seed, **sh["kw"])
        csr = synth.topology_csr(t, isis=isis)
        for root in range(csr.n_vertices):
            ref = pyoracle.csr_spf(csr, root, vec_mode=int(isis), nh_words=4)
            if ref["status"] != 0:
                continue
            hops, nh, n_atoms, st = jump_phase(csr, root, ref["dist"], ref["first_parent"], ref["n_parents"])
            assert np.array_equal(hops, ref["hops"]), (shape, seed, root)
            exp = [nh_int(r) for r in ref["nh_mask"]]
            assert nh == exp, (shape, seed, root, [v for v in range(csr.n_vertices) if nh[v] != exp[v]][:5])
            checked += 1
            ecmp_total += st["n_ecmp"]
    assert checked > 50 and ecmp_total > 0


def test_jump_algorithm_on_nested_ecmp_ladder():
    """A ladder of equal-cost diamonds: every rung vertex is an ECMP vertex whose parents are
    themselves below ECMP vertices (deep terminal chains, many sweeps)."""
    n = 24
    a = []
    b = []
    for i in range(n - 2):
        a += [i, i]
        b += [i + 1, i + 2]
    a, b = np.asarray(a, np.uint32), np.asarray(b, np.uint32)
    c = np.full(len(a), 4, np.uint32)
    c[1::2] = 8          # i -> i+2 costs as much as i -> i+1 -> i+2
    t = synth.Topology(n, a, b, c, c.copy(), [])
    csr = synth.topology_csr(t)
    for root in (0, n // 2, n - 1):
        ref = pyoracle.csr_spf(csr, root, nh_words=4)
        assert ref["status"] == 0
        hops, nh, _n, st = jump_phase(csr, root, ref["dist"], ref["first_parent"], ref["n_parents"])
        assert np.array_equal(hops, ref["hops"])
        assert nh == [nh_int(r) for r in ref["nh_mask"]]
        assert st["n_ecmp"] >= n // 3


def test_jump_algorithm_with_parallel_links_and_lan_roots():
    """Parallel p2p links (distinct first-hop atoms to the same neighbour) and roots that sit
    on several LANs (atoms behind hops-0 network vertices)."""
    rng = np.random.default_rng(7)
    for trial in range(12):
        R = 14
        a = list(range(1, R)) + [int(x) for x in rng.integers(0, R, 10)]
        b = [int(rng.integers(0, i)) for i in range(1, R)] + [int(x) for x in rng.integers(0, R, 10)]
        keep = [(x, y) for x, y in zip(a, b) if x != y]
        keep += keep[:4]                                   # four parallel links
        a = np.asarray([x for x, _ in keep], np.uint32)
        b = np.asarray([y for _, y in keep], np.uint32)
        c = rng.choice([5, 10], len(a)).astype(np.uint32)
        lans = [([0, 3, 5, 7], [5, 5, 5, 5]), ([0, 2, 4], [10, 5, 5]), ([1, 2, 6, 8], [5, 5, 10, 5])]
        t = synth.Topology(R, a, b, c, c.copy(), lans)
        for isis in (False, True):
            csr = synth.topology_csr(t, isis=isis)
            for root in range(csr.n_vertices):
                ref = pyoracle.csr_spf(csr, root, vec_mode=int(isis), nh_words=4)
                if ref["status"] != 0:
                    continue
                hops, nh, _n, _st = jump_phase(csr, root, ref["dist"], ref["first_parent"], ref["n_parents"])
                assert np.array_equal(hops, ref["hops"]), (trial, isis, root)
                assert nh == [nh_int(r) for r in ref["nh_mask"]], (trial, isis, root)
