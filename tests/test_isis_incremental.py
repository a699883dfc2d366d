"""CPU: trigger-keyed recomputation for IS-IS — the SPF-type decision of lsp_install
(holo-isis/src/lsdb.rs:1450-1465, 1525-1531) against its restatement, and hspf_isis_flat_update: a metric
change names the CSR edges whose cost moved (for hspf_graph_update_costs), anything else rebuilds."""
import copy

import numpy as np
import pytest

from holo_b200 import isis, synth
from oracle import pyoracle


def same_flat(a, b):
    for name in ("row_ptr", "col", "cost", "vflags"):
        assert np.array_equal(getattr(a.csr, name), getattr(b.csr, name)), name
    assert np.array_equal(a.ids, b.ids) and a.csr.reject_above == b.csr.reject_above and a.csr.flags == b.csr.flags


def lsp_index(level, system, fragment=0):
    lan = isis.sysid(system) << 8
    return int(np.nonzero((level.lsps["lan_id"] == lan) & (level.lsps["fragment"] == fragment))[0][0])


@pytest.mark.parametrize("metric_type", [isis.METRIC_WIDE, isis.METRIC_BOTH])
@pytest.mark.parametrize("seed", range(4))
def test_metric_change_patches_only_costs(seed, metric_type):
    rng = np.random.default_rng(seed)
    t = synth.random_topology(120, 520, synth.SEED_BASE + 50 + seed, lan_fraction=0.1)
    lv = isis.synth_level(t, metric_type=metric_type)
    flat = isis.Flat(lv)
    new = copy.deepcopy(lv)
    triggers = []
    for r in rng.choice(120, 3, replace=False):
        i = lsp_index(new, int(r))
        lo, n = int(new.lsps["reach_off"][i]), int(new.lsps["n_reach"][i])
        for k in range(lo, lo + n):
            if rng.random() < 0.7:
                new.reaches["metric"][k] = int(rng.integers(1, 60))
        triggers.append((isis.sysid(int(r)) << 8, 0))
    before = flat.csr.cost.copy()
    assert isis.spf_type(lv, new, triggers) == isis.SPF_FULL              # a metric is part of the IS-reach entry
    kind, edges, costs = isis.flat_update(flat, new)
    fresh = isis.Flat(new)
    same_flat(flat, fresh)
    changed = np.nonzero(before != fresh.csr.cost)[0]
    assert kind == isis.FLAT_COSTS and len(changed) > 0
    assert sorted(edges.tolist()) == changed.tolist() and np.array_equal(fresh.csr.cost[edges], costs)
    # the patched graph drives the oracle's SPT of the new level
    root = isis.sysid(0)
    got = pyoracle.csr_spf(flat.csr, flat.vertex(root << 8), vec_mode=1, nh_words=2)
    spt = flat.spt_from_planes(flat.vertex(root << 8), got["dist"], got["hops"])
    ref = pyoracle.isis_compute_spt(new, root)
    assert np.array_equal(spt.vertices["distance"], ref.vertices["distance"]) and np.array_equal(spt.vertices["hops"], ref.vertices["hops"])


def test_structural_changes_rebuild_and_ip_changes_do_not_touch_the_graph():
    t = synth.random_topology(60, 260, synth.SEED_BASE + 57, lan_fraction=0.1)
    lv = isis.synth_level(t)
    # an adjacency disappears from one LSP
    new = copy.deepcopy(lv)
    i = lsp_index(new, 9)
    new.lsps["n_reach"][i] -= 1
    flat = isis.Flat(lv)
    kind, edges, _ = isis.flat_update(flat, new)
    assert kind == isis.FLAT_REBUILT and len(edges) == 0
    same_flat(flat, isis.Flat(new))
    assert isis.spf_type(lv, new, [(isis.sysid(9) << 8, 0)]) == isis.SPF_FULL
    # overload bit: same edges, another vertex flag
    new = copy.deepcopy(lv)
    new.lsps["flags"][i] |= isis.LSPF_OL
    flat = isis.Flat(lv)
    assert isis.flat_update(flat, new)[0] == isis.FLAT_REBUILT
    assert isis.spf_type(lv, new, [(isis.sysid(9) << 8, 0)]) == isis.SPF_FULL
    # the LSP expires
    new = copy.deepcopy(lv)
    new.lsps["rem_lifetime"][i] = 0
    flat = isis.Flat(lv)
    assert isis.flat_update(flat, new)[0] == isis.FLAT_REBUILT
    assert isis.spf_type(lv, new, [(isis.sysid(9) << 8, 0)]) == isis.SPF_FULL
    # a refresh (new sequence number, same content) and an IP-reachability-only change: the SPTs stand
    new = copy.deepcopy(lv)
    new.lsps["seqno"][i] += 1
    flat = isis.Flat(lv)
    assert isis.flat_update(flat, new) [0] == isis.FLAT_UNCHANGED
    assert isis.spf_type(lv, new, [(isis.sysid(9) << 8, 0)]) == isis.SPF_ROUTE_ONLY
    # Protocols-Supported changes (a TLV, not an LSP flag) are not compared by the reference: RouteOnly,
    # although the flattener's transit gate reads them (the caller of a RouteOnly run keeps its SPT)
    new = copy.deepcopy(lv)
    new.lsps["flags"][i] = int(new.lsps["flags"][i]) & (0xFF ^ isis.LSPF_NLPID_IPV4)
    assert isis.spf_type(lv, new, [(isis.sysid(9) << 8, 0)]) == isis.SPF_ROUTE_ONLY
    # a brand-new LSP is always a topology change
    newer = copy.deepcopy(lv)
    keep = np.ones(len(lv.lsps), bool)
    keep[i] = False
    older = copy.deepcopy(lv)
    older.lsps = lv.lsps[keep]
    assert isis.spf_type(older, newer, [(isis.sysid(9) << 8, 0)]) == isis.SPF_FULL


@pytest.mark.parametrize("seed", range(60))
def test_spf_type_matches_restatement(seed):
    rng = np.random.default_rng(300 + seed)
    t = synth.random_topology(30, 120, synth.SEED_BASE + 58, lan_fraction=0.2)
    mt = int(rng.choice([isis.MT_NONE, isis.MT_STANDARD]))
    lv = isis.synth_level(t, metric_type=int(rng.choice([isis.METRIC_WIDE, isis.METRIC_BOTH, isis.METRIC_STANDARD])), mt_id=mt)
    new = copy.deepcopy(lv)
    triggers = []
    for r in rng.choice(30, int(rng.integers(1, 4)), replace=False):
        i = lsp_index(new, int(r))
        lo, n = int(new.lsps["reach_off"][i]), int(new.lsps["n_reach"][i])
        what = rng.random()
        if what < 0.25 and n:
            new.reaches["metric"][lo + int(rng.integers(0, n))] += 1
        elif what < 0.4:
            new.lsps["flags"][i] ^= int(rng.choice([isis.LSPF_ATT, isis.LSPF_OL, isis.LSPF_HAS_PROTOCOLS]))
        elif what < 0.5:
            new.lsps["rem_lifetime"][i] = 0
        elif what < 0.6 and n:
            new.reaches["kind"][lo] = isis.REACH_MT              # leaves TLV 2 / 22, enters TLV 222
        else:
            new.lsps["seqno"][i] += 1                            # refresh
        triggers.append((isis.sysid(int(r)) << 8, 0))
    a = isis.spf_type(lv, new, triggers)
    b = isis.spf_type(lv, new, triggers, lib=pyoracle.lib(), name="oracle_isis_spf_type")
    assert a == b
    SEEN.add(a)


SEEN = set()


def test_spf_type_fuzz_sees_both_outcomes():
    if len(SEEN) == 0:
        pytest.skip("runs after the fuzz")
    assert SEEN == {isis.SPF_FULL, isis.SPF_ROUTE_ONLY}
