"""CPU: trigger-keyed recomputation: Ospfv2::spf_computation_type
(holo-ospf/src/ospfv2/spf.rs:98-171) and the incremental flattener hspf_ospfv2_flat_update — an
interface cost change patches cost[] of the flattened area and names the changed CSR edges, nothing else;
anything structural rebuilds.  The patched flat must equal a fresh flatten of the new LSDB image."""
import copy

import numpy as np
import pytest

from holo_b200 import ospfv2, synth
from oracle import pyoracle


def trig(lsa_type, adv, lsa_id, mask=0, opaque=0):
    return (adv, lsa_id, mask, lsa_type, opaque, (0, 0))


def same_flat(a: ospfv2.Flat, b: ospfv2.Flat):
    for name in ("row_ptr", "col", "cost", "vflags"):
        assert np.array_equal(getattr(a.csr, name), getattr(b.csr, name)), name
    assert np.array_equal(a.ids, b.ids) and np.array_equal(a.is_router, b.is_router)
    assert np.array_equal(a.link_index, b.link_index) and np.array_equal(a.link_pos, b.link_pos)


def router_lsa_links(area, rid):
    i = int(np.nonzero(area.router_lsas["adv_rtr"] == rid)[0][0])
    lo, n = int(area.router_lsas["link_off"][i]), int(area.router_lsas["n_links"][i])
    return range(lo, lo + n)


@pytest.mark.parametrize("seed", range(6))
def test_cost_change_patches_only_costs(seed):
    rng = np.random.default_rng(seed)
    t = synth.random_topology(150, 700, synth.SEED_BASE + 30 + seed, lan_fraction=0.1)
    area = ospfv2.synth_area(t, root=0, sr=True)
    flat = ospfv2.Flat(area)
    new = copy.deepcopy(area)
    rids = [int(ospfv2.RID_BASE + int(i)) for i in rng.choice(150, 3, replace=False)]
    touched = 0
    for rid in rids:
        for k in router_lsa_links(new, rid):
            if new.links["link_type"][k] != ospfv2.LINK_STUB and rng.random() < 0.6:
                new.links["metric"][k] = int(rng.integers(1, 200))
                touched += 1
            elif new.links["link_type"][k] == ospfv2.LINK_STUB:
                new.links["metric"][k] = int(rng.integers(1, 50))      # stub metrics are not graph edges
    before = flat.csr.cost.copy()
    kind, edges, costs = ospfv2.flat_update(flat, new, [trig(1, r, r) for r in rids])
    fresh = ospfv2.Flat(new)
    same_flat(flat, fresh)
    changed = np.nonzero(before != fresh.csr.cost)[0]
    assert kind == (ospfv2.FLAT_COSTS if len(changed) else ospfv2.FLAT_UNCHANGED)
    assert sorted(edges.tolist()) == changed.tolist()
    assert np.array_equal(fresh.csr.cost[edges], costs)
    assert 0 < len(changed) <= touched
    # every changed edge leaves one of the trigger routers
    rows = {flat.router_vertex(r) for r in rids}
    for e in edges:
        v = int(np.searchsorted(flat.csr.row_ptr, e, side="right") - 1)
        assert v in rows


def test_structural_changes_rebuild():
    t = synth.random_topology(80, 350, synth.SEED_BASE + 41, lan_fraction=0.1)
    area = ospfv2.synth_area(t, root=0)
    rid = int(ospfv2.RID_BASE + 5)
    # 1. a link turns into a stub (the adjacency went down): same layout, different links
    new = copy.deepcopy(area)
    k = next(k for k in router_lsa_links(new, rid) if new.links["link_type"][k] == ospfv2.LINK_P2P)
    new.links["link_type"][k] = ospfv2.LINK_STUB
    flat = ospfv2.Flat(area)
    kind, edges, _ = ospfv2.flat_update(flat, new, [trig(1, rid, rid)])
    assert kind == ospfv2.FLAT_REBUILT and len(edges) == 0
    same_flat(flat, ospfv2.Flat(new))
    assert flat.csr.n_edges == ospfv2.Flat(area).csr.n_edges - 2          # both directions fall to the mutual check
    # 2. the Router-LSA ages out
    new = copy.deepcopy(area)
    i = int(np.nonzero(new.router_lsas["adv_rtr"] == rid)[0][0])
    new.router_lsas["age"][i] = ospfv2.MAX_AGE
    flat = ospfv2.Flat(area)
    kind, _, _ = ospfv2.flat_update(flat, new, [trig(1, rid, rid)])
    assert kind == ospfv2.FLAT_REBUILT
    same_flat(flat, ospfv2.Flat(new))
    assert flat.csr.n_vertices == ospfv2.Flat(area).csr.n_vertices - 1
    # 3. a Network-LSA loses an attached router
    new = copy.deepcopy(area)
    j = 0
    new.attached[int(new.network_lsas["att_off"][j])] = 0x7F000001
    flat = ospfv2.Flat(area)
    kind, _, _ = ospfv2.flat_update(flat, new, [trig(2, int(new.network_lsas["adv_rtr"][j]), int(new.network_lsas["lsa_id"][j]))])
    assert kind == ospfv2.FLAT_REBUILT
    same_flat(flat, ospfv2.Flat(new))
    # 4. an image with another layout (one more router) always rebuilds
    t2 = synth.random_topology(81, 350, synth.SEED_BASE + 41, lan_fraction=0.1)
    other = ospfv2.synth_area(t2, root=0)
    flat = ospfv2.Flat(area)
    kind, _, _ = ospfv2.flat_update(flat, other, [trig(1, rid, rid)])
    assert kind == ospfv2.FLAT_REBUILT
    same_flat(flat, ospfv2.Flat(other))


def test_triggers_that_do_not_touch_the_graph():
    t = synth.random_topology(60, 250, synth.SEED_BASE + 43)
    area = ospfv2.synth_area(t, root=0, sr=True)
    flat = ospfv2.Flat(area)
    new = copy.deepcopy(area)
    rid = int(ospfv2.RID_BASE + 7)
    # a refreshed Router-LSA (same body), a summary, an external, an SR opaque LSA
    kind, edges, _ = ospfv2.flat_update(flat, new, [trig(1, rid, rid), trig(3, rid, 0x0A0A0000, 0xFFFF0000),
                                                    trig(5, rid, 0x0B000000, 0xFF000000), trig(10, rid, 0x04000000, opaque=4)])
    assert kind == ospfv2.FLAT_UNCHANGED and len(edges) == 0
    same_flat(flat, ospfv2.Flat(area))


def test_updated_flat_gives_the_new_spt():
    """End to end on the CPU: the patched flat drives the same SPT as the oracle computes on the new LSDB."""
    t = synth.random_topology(120, 500, synth.SEED_BASE + 47)
    area = ospfv2.synth_area(t, root=3)
    flat = ospfv2.Flat(area)
    new = copy.deepcopy(area)
    rid = int(ospfv2.RID_BASE + 3)
    for k in router_lsa_links(new, rid):
        if new.links["link_type"][k] == ospfv2.LINK_P2P:
            new.links["metric"][k] = 1 + int(new.links["metric"][k]) * 3
    kind, edges, costs = ospfv2.flat_update(flat, new, [trig(1, rid, rid)])
    assert kind == ospfv2.FLAT_COSTS and len(edges) > 0
    root = flat.router_vertex(new.router_id)
    got = pyoracle.csr_spf(flat.csr, root, nh_words=1)
    res = ospfv2.area_from_planes(new, lambda csr, r, w: (pyoracle.csr_spf(csr, r, nh_words=w)[k] for k in ("dist", "hops", "nh_mask")))
    ref = pyoracle.ospfv2_run_area(new)
    assert np.array_equal(res.routes, ref.routes) and np.array_equal(res.vertices, ref.vertices)
    assert got["status"] == 0


@pytest.mark.parametrize("seed", range(40))
def test_spf_computation_type_matches_restatement(seed):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(0, 12))
    full_ok = rng.random() < 0.3
    types = [3, 4, 5, 9, 10, 11] + ([1, 2] if full_ok else [])
    tr = []
    for _ in range(n):
        ty = int(rng.choice(types))
        op = int(rng.choice([1, 4, 7, 8])) if ty in (10, 11) and full_ok else int(rng.choice([1, 2]))
        mask = int(rng.choice([0xFFFFFF00, 0xFFFF0000, 0xFFFFFFFF, 0]))
        tr.append(trig(ty, int(rng.integers(1, 5)), int(rng.choice([0x0A000001, 0x0A000100, 0x0A010000, 0x0B000000])), mask, op))
    a = ospfv2.spf_computation_type(tr)
    b = pyoracle.ospfv2_spf_computation_type(tr)
    assert a == b
    if a[0] == ospfv2.SPF_PARTIAL:
        assert a[1] == sorted(set(a[1]), key=lambda p: (p[0], bin(p[1]).count("1"))) and a[2] == sorted(set(a[2]))


def test_spf_computation_type_cases():
    r = 0x0A000001
    assert ospfv2.spf_computation_type([trig(3, r, 0x0A000100, 0xFFFFFF00), trig(1, r, r)])[0] == ospfv2.SPF_FULL
    assert ospfv2.spf_computation_type([trig(10, r, 0x07000000, opaque=7)])[0] == ospfv2.SPF_FULL       # Extended-Prefix
    assert ospfv2.spf_computation_type([trig(11, r, 0x08000000, opaque=8)])[0] == ospfv2.SPF_PARTIAL    # AS-scope Ext-Link: not listed
    kind, net, rtr, ext = ospfv2.spf_computation_type([trig(3, r, 0x0A000101, 0xFFFFFF00), trig(3, 2, 0x0A000101, 0xFFFFFF00),
                                                       trig(4, r, 0x0A000005), trig(5, r, 0x0B000000, 0xFF000000)])
    assert kind == ospfv2.SPF_PARTIAL
    assert net == [(0x0A000101, 0xFFFFFF00)]            # one entry, host bits kept (with_netmask, no apply_mask)
    assert rtr == [0x0A000005] and ext == [(0x0B000000, 0xFF000000)]
    assert ospfv2.spf_computation_type([]) == (ospfv2.SPF_PARTIAL, [], [], [])
