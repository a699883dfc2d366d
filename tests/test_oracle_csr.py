"""CPU: the optimised heap baseline must agree with the reference-faithful oracle."""
import numpy as np
import pytest

from holo_b200 import synth
from oracle import pyoracle

PLANES = ["dist", "hops", "first_parent", "n_parents", "nh_mask"]


@pytest.mark.parametrize("V,E,seed,kw,isis", [
    (100, 400, 3, {}, False),
    (100, 400, 4, dict(cost_choices=[10, 20]), False),
    (300, 1400, 5, dict(lan_fraction=0.1), False),
    (300, 1400, 6, dict(cost_choices=[10, 20], lan_fraction=0.1), True),
    (1500, 6000, 7, dict(cost_lo=1, cost_hi=1000), True),
])
def test_heap_matches_faithful(V, E, seed, kw, isis):
    t = synth.random_topology(V, E, synth.SEED_BASE + seed, **kw)
    csr = synth.topology_csr(t, isis=isis)
    for r in range(0, csr.n_vertices, max(1, csr.n_vertices // 40)):
        a = pyoracle.csr_spf(csr, r, vec_mode=int(isis), nh_words=2)
        b = pyoracle.csr_spf_heap(csr, r, nh_words=2)
        for k in PLANES:
            assert np.array_equal(a[k], b[k]), (r, k)
        assert a["status"] == b["status"]


def test_faithful_pop_order_is_dist_then_id():
    t = synth.random_topology(200, 900, synth.SEED_BASE + 11, cost_choices=[10, 20], lan_fraction=0.1)
    csr = synth.topology_csr(t)
    a = pyoracle.csr_spf(csr, csr.n_vertices - 1)
    order = a["pop_order"]
    keys = [(int(a["dist"][v]), int(v)) for v in order]
    assert keys == sorted(keys)


def test_isis_vec_lists():
    t = synth.random_topology(60, 260, synth.SEED_BASE + 12, cost_choices=[10], lan_fraction=0.1)
    csr = synth.topology_csr(t, isis=True)
    a = pyoracle.csr_spf(csr, csr.n_vertices - 1, vec_mode=1, want_lists=True)
    assert a["rc"] == 0
    po = a["parents_off"]
    assert (np.diff(po) == a["n_parents"]).all()
    # de-duplicated Vec == set plane
    no = a["nhvec_off"]
    for v in range(csr.n_vertices):
        s = 0
        for x in a["nhvec"][no[v]:no[v + 1]]:
            s |= 1 << int(x)
        assert s == int(a["nh_mask"][v, 0])


def test_native_batch_pool_matches_single_calls(built):
    """oracle/batch_pool.cc: the thread pool returns the planes of the single-job calls, the
    bounded sample stops early, and the checksum does not depend on the thread count."""
    import numpy as np
    from holo_b200 import synth
    from oracle import pyoracle
    t = synth.random_topology(300, 1400, synth.SEED_BASE + 41, lan_fraction=0.1)
    csr = synth.topology_csr(t)
    roots = np.arange(0, 300, 7, dtype=np.uint32)
    a = pyoracle.csr_batch(csr, roots, mode="heap", threads=3)
    b = pyoracle.csr_batch(csr, roots, mode="faithful", threads=2)
    assert a["jobs_done"] == len(roots) == b["jobs_done"] and a["checksum"] == b["checksum"]
    for j, r in enumerate(roots):
        ref = pyoracle.csr_spf(csr, int(r))
        for k in ("dist", "hops", "first_parent", "n_parents", "nh_mask"):
            assert np.array_equal(a[k][j], ref[k]) and np.array_equal(b[k][j], ref[k]), (j, k)
    c = pyoracle.csr_batch(csr, roots, mode="heap", threads=1, want_planes=False)
    assert c["checksum"] == a["checksum"]
    ov = [[(0, 0xFFFFFFFF)] if j % 2 else [] for j in range(len(roots))]
    d = pyoracle.csr_batch(csr, roots, overrides=ov, mode="heap", threads=2)
    assert np.array_equal(d["dist"][1], pyoracle.csr_spf_heap(csr, int(roots[1]), overrides=ov[1])["dist"])
    assert pyoracle.usable_cores() >= 1 and 1.0 <= pyoracle.effective_cores(2) <= 2.0
