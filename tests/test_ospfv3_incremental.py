"""CPU: hspf_ospfv3_flat_update — an interface cost change of an OSPFv3 router names the CSR edges whose cost
moved (for hspf_graph_update_costs); anything structural rebuilds; the updated flat equals a fresh flatten and
drives the oracle's routes of the new LSDB."""
import copy

import numpy as np
import pytest

from holo_b200 import ospfv3, synth
from oracle import pyoracle


def same_flat(a, b):
    for name in ("row_ptr", "col", "cost", "vflags"):
        assert np.array_equal(getattr(a.csr, name), getattr(b.csr, name)), name
    assert np.array_equal(a.router_ids, b.router_ids) and np.array_equal(a.iface_ids, b.iface_ids)


def links_of(area, rid):
    out = []
    for i in np.nonzero(area.router_lsas["adv_rtr"] == rid)[0]:
        lo, n = int(area.router_lsas["link_off"][i]), int(area.router_lsas["n_links"][i])
        out += list(range(lo, lo + n))
    return out


@pytest.mark.parametrize("seed,frag", [(0, 0), (1, 3), (2, 0), (3, 2)])
def test_cost_change_patches_only_costs(seed, frag):
    rng = np.random.default_rng(seed)
    t = synth.random_topology(100, 450, synth.SEED_BASE + 80 + seed, lan_fraction=0.1)
    area = ospfv3.synth_area(t, root=4, max_links_per_fragment=frag)
    flat = ospfv3.Flat(area)
    new = copy.deepcopy(area)
    for r in rng.choice(100, 3, replace=False):
        for k in links_of(new, int(ospfv3.RID_BASE + int(r))):
            if rng.random() < 0.7:
                new.links["metric"][k] = int(rng.integers(1, 300))
    before = flat.csr.cost.copy()
    kind, edges, costs = flat.update(new)
    fresh = ospfv3.Flat(new)
    same_flat(flat, fresh)
    changed = np.nonzero(before != fresh.csr.cost)[0]
    assert kind == 1 and len(changed) > 0
    assert sorted(edges.tolist()) == changed.tolist() and np.array_equal(fresh.csr.cost[edges], costs)
    # the route stage over the oracle's planes of the updated flat == the faithful oracle on the new LSDB
    res = ospfv3.area_from_planes(new, lambda csr, r, w: tuple(pyoracle.csr_spf(csr, r, nh_words=w)[k] for k in ("dist", "hops", "nh_mask")))
    ref = pyoracle.ospfv3_run_area(new)
    assert np.array_equal(res.routes, ref.routes) and np.array_equal(res.vertices, ref.vertices)


def test_structural_change_rebuilds_and_refresh_is_a_no_op():
    t = synth.random_topology(60, 260, synth.SEED_BASE + 87, lan_fraction=0.1)
    area = ospfv3.synth_area(t, root=0)
    flat = ospfv3.Flat(area)
    assert flat.update(copy.deepcopy(area))[0] == 0                       # nothing moved
    new = copy.deepcopy(area)
    i = int(np.nonzero(new.router_lsas["adv_rtr"] == ospfv3.RID_BASE + 9)[0][0])
    new.router_lsas["n_links"][i] -= 1                                     # an adjacency went away
    kind, edges, _ = flat.update(new)
    assert kind == 2 and len(edges) == 0
    same_flat(flat, ospfv3.Flat(new))
    new2 = copy.deepcopy(new)
    new2.router_lsas["age"][i] = 3600                                      # the fragment ages out
    assert flat.update(new2)[0] == 2
    same_flat(flat, ospfv3.Flat(new2))


ROUTER, NETWORK, INTER_PREFIX, INTER_ROUTER, EXTERNAL, LINK, INTRA_PREFIX, GRACE, ROUTER_INFO = 1, 2, 3, 4, 5, 8, 9, 11, 12


@pytest.mark.parametrize("seed", range(40))
def test_spf_computation_type_matches_restatement(seed):
    rng = np.random.default_rng(900 + seed)
    pool = [("2001:db8:%x::" % i, 64) for i in range(5)] + [("2001:db8::%x" % i, 128) for i in range(3)] + [("10.0.%d.0" % i, 24) for i in range(3)]
    codes = [INTER_PREFIX, INTER_ROUTER, EXTERNAL, INTRA_PREFIX, GRACE, 35, 36, 37, 41] + ([ROUTER, NETWORK, LINK, ROUTER_INFO, 33, 40] if rng.random() < 0.25 else [])
    tr = []
    for _ in range(int(rng.integers(0, 9))):
        c = int(rng.choice(codes))
        k = int(rng.integers(0, 4)) if c in (INTRA_PREFIX, 41) else 1
        pf = [pool[int(j)] for j in rng.integers(0, len(pool), k)]
        tr.append((c, int(rng.integers(1, 4)), int(rng.integers(0, 3)), int(rng.choice([0x01010101, 0x02020202])), pf))
    a = ospfv3.spf_computation_type(tr)
    b = ospfv3.spf_computation_type(tr, fn=pyoracle.lib().oracle_ospfv3_spf_computation_type)
    assert a == b


def test_spf_computation_type_cases():
    p1, p2, v4 = ("2001:db8:1::", 64), ("2001:db8:2::", 64), ("10.0.0.0", 24)
    FULL, PARTIAL = 1, 2
    assert ospfv3.spf_computation_type([(INTER_PREFIX, 1, 1, 0, [p1]), (LINK, 1, 2, 0, [])])[0] == FULL
    assert ospfv3.spf_computation_type([(33, 1, 0, 0, [])])[0] == FULL                       # E-Router-LSA
    assert ospfv3.spf_computation_type([(ROUTER_INFO, 1, 0, 0, [])])[0] == FULL
    kind, intra, inter, rtr, ext = ospfv3.spf_computation_type([
        (INTRA_PREFIX, 1, 0, 0, [p2, p1, v4]),          # prefixes of the new and of the old instance
        (41, 2, 0, 0, [p1]),                            # E-Intra-Area-Prefix-LSA
        (INTER_PREFIX, 1, 5, 0, [p2]), (INTER_ROUTER, 1, 6, 0x09090909, []), (EXTERNAL, 3, 7, 0, [p1]), (GRACE, 1, 0, 0, [])])
    assert kind == PARTIAL
    assert intra == [("10.0.0.0", 24), ("2001:db8:1::", 64), ("2001:db8:2::", 64)]            # IPv4 before IPv6, unique
    assert inter == [p2] and rtr == [0x09090909] and ext == [p1]
    assert ospfv3.spf_computation_type([]) == (PARTIAL, [], [], [], [])
