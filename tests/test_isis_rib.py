"""End of holo-isis' update_rib on the CPU: both levels through the product's route stage
(hspf_isis_routes_from_planes, SPT planes from the oracle), hspf_isis_rib_merge (L1 preferred)
and hspf_isis_rib_diff (update_global_rib) from an empty table must give exactly the RouteIpAdd
set the reference sent to the RIB manager (output/ibus.jsonl of its 38 IS-IS conformance
snapshots: prefix, metric, ifindex, next-hop address); product == restatement
(oracle/rib_isis.cc) on these and on perturbed table pairs."""
import numpy as np
import pytest

import golden_util as gu
from holo_b200 import isis, ospfv3
from oracle import pyoracle

SNAPS = [s for s in gu.load_isis() if s.get("ibus_routes") is not None]


def planes(csr, root):
    c = pyoracle.csr_spf(csr, root, vec_mode=1, nh_words=4)
    return c["dist"], c["hops"]


def level_ribs(snap):
    out = {}
    names = None
    for level in snap["levels"]:
        inst = gu.isis_instance_image(snap, level)
        out[level["level"]] = isis.routes_from_planes(inst, planes)
        names = inst["ifnames"]
    return out, names


@pytest.mark.parametrize("snap", SNAPS, ids=[f"{s['topo']}-{s['rt']}" for s in SNAPS])
def test_installs_equal_the_reference_ibus_stream(snap):
    ribs, names = level_ribs(snap)
    merged = isis.rib_merge(ribs.get(2), ribs.get(1))
    ref = isis.rib_merge(ribs.get(2), ribs.get(1), lib=pyoracle.lib(), name="oracle_isis_rib_merge")
    assert merged.routes.tobytes() == ref.routes.tobytes() and merged.nexthops.tobytes() == ref.nexthops.tobytes()
    acts, routes = isis.rib_diff(None, merged)
    acts_o, routes_o = isis.rib_diff(None, merged, lib=pyoracle.lib(), name="oracle_isis_rib_diff")
    assert acts.tobytes() == acts_o.tobytes() and routes.tobytes() == routes_o.tobytes()
    got = {}
    for a in acts:
        assert int(a["kind"]) == 1
        r = merged.routes[int(a["route"])]
        hops = merged.nexthops[int(r["nh_off"]): int(r["nh_off"]) + int(r["n_nh"])]
        got[f"{ospfv3.ip_str(r['prefix'])}/{int(r['len'])}"] = (
            int(r["metric"]), sorted((snap["ifindex"].get(names[int(x["iface"])], 0), ospfv3.ip_str(x["addr"])) for x in hops))
    want = {p: (v["metric"], sorted((n[0], n[1]) for n in v["nexthops"])) for p, v in snap["ibus_routes"].items()}
    assert got == want
    # second run over the unchanged table: nothing to tell the RIB manager, flags carried over
    again, routes2 = isis.rib_diff(isis.IsisRib(routes, merged.nexthops), merged)
    assert len(again) == 0 and routes2.tobytes() == routes.tobytes()


@pytest.mark.parametrize("seed", range(40))
def test_rib_diff_matches_restatement_on_perturbed_tables(seed):
    rng = np.random.default_rng(seed)
    snap = SNAPS[seed % len(SNAPS)]
    ribs, _ = level_ribs(snap)
    merged = isis.rib_merge(ribs.get(2), ribs.get(1))
    _acts, routes = isis.rib_diff(None, merged)
    old = isis.IsisRib(routes, merged.nexthops)
    # perturb: drop some routes, change metrics, change a next-hop label
    keep = rng.random(len(merged.routes)) > 0.2
    new_routes = merged.routes[keep].copy()
    for i in range(len(new_routes)):
        if rng.random() < 0.3:
            new_routes["metric"][i] += 1
    nh = merged.nexthops.copy()
    if len(nh):
        k = int(rng.integers(0, len(nh)))
        nh["has_label"][k] = 1
        nh["sr_label"][k] = 16000 + seed
    new = isis.IsisRib(new_routes, nh)
    a, f = isis.rib_diff(old, new)
    b, g = isis.rib_diff(old, new, lib=pyoracle.lib(), name="oracle_isis_rib_diff")
    assert a.tobytes() == b.tobytes() and f.tobytes() == g.tobytes()
    assert {int(x) for x in a["kind"]} <= {1, 3}


AFTER = [(s, name) for s in gu.load_isis() for name in s.get("after", {})]


@pytest.mark.parametrize("snap,name", AFTER, ids=[f"{n}-{s['topo']}-{s['rt']}" for s, n in AFTER])
def test_step_recomputation_gives_the_reference_ibus_output(snap, name):
    """Step tests of the reference (holo-isis/tests/conformance/mod.rs): the state it reached after
    the step is one more snapshot.  Table of the topology snapshot (installed) vs table of the
    after-state through product host code + hspf_isis_rib_diff == the step's ibus output, message
    for message and in order — 20 cases: configuration changes (att-ignore, max_paths 16 -> 1, an
    address family or an interface disabled / deleted / made passive, interface metric), RPCs
    (clear adjacency / database), interface and adjacency events, and received LSPs (ATT bit,
    overload bit, expiration)."""
    after = dict(snap["after"][name])
    after.setdefault("ifindex", snap["ifindex"])

    def table(s):
        ribs, names = {}, None
        for level in s["levels"]:
            inst = gu.isis_instance_image(s, level)
            inst["att_ignore"] = int(bool(s.get("att_ignore", False)))
            ribs[level["level"]] = isis.routes_from_planes(inst, planes)
            names = inst["ifnames"]
        return isis.rib_merge(ribs.get(2), ribs.get(1)), names

    old, names = table(snap)
    _a, installed = isis.rib_diff(None, old)
    old_inst = isis.IsisRib(installed, old.nexthops)
    new, names2 = table(after)
    names2 = names2 or []
    gone = after.get("deleted_ifaces", [])
    if gone:
        # Deleting an interface is handled by the configuration code before the SPF runs: it removes
        # the next hops over that interface from the local table without telling the RIB manager
        # (holo-isis/src/northbound/configuration.rs:2193-2211), and the arena slot goes away.  The
        # same on the installed table here: prune, and re-index the interfaces by name.
        routes, nhs = old_inst.routes.copy(), []
        for i, r in enumerate(routes):
            hops = [x.copy() for x in old_inst.nexthops[int(r["nh_off"]): int(r["nh_off"]) + int(r["n_nh"])]
                    if names[int(x["iface"])] not in gone]
            for x in hops:
                x["iface"] = names2.index(names[int(x["iface"])])
            routes["nh_off"][i], routes["n_nh"][i] = len(nhs), len(hops)
            nhs += hops
        old_inst = isis.IsisRib(routes, np.asarray(nhs, dtype=isis.NEXTHOP_DT) if nhs else np.zeros(0, isis.NEXTHOP_DT))
    acts, _f = isis.rib_diff(old_inst, new)
    got = []
    for a in acts:
        kind = int(a["kind"])
        if kind == 1:
            r = new.routes[int(a["route"])]
            hops = new.nexthops[int(r["nh_off"]): int(r["nh_off"]) + int(r["n_nh"])]
            got.append(["add", f"{ospfv3.ip_str(r['prefix'])}/{int(r['len'])}", int(r["metric"]),
                        sorted([snap["ifindex"].get(names2[int(x["iface"])], 0), ospfv3.ip_str(x["addr"])] for x in hops)])
        else:
            assert kind == 3
            r = old_inst.routes[int(a["route"])]
            got.append(["del", f"{ospfv3.ip_str(r['prefix'])}/{int(r['len'])}", None, []])
    want = [[k, p, m, sorted(nh)] for (k, p, m, nh) in after["ibus"]]
    assert got == want
    # and the whole table of the after-state is the local-rib the reference reports there
    table = {}
    for r in new.routes:
        hops = new.nexthops[int(r["nh_off"]): int(r["nh_off"]) + int(r["n_nh"])]
        table[f"{ospfv3.ip_str(r['prefix'])}/{int(r['len'])}"] = (
            int(r["metric"]), sorted((names2[int(x["iface"])], ospfv3.ip_str(x["addr"])) for x in hops))
    rib_want = {r["prefix"]: (r["metric"], sorted((a, b) for a, b in r["nexthops"])) for r in after["local_rib"]}
    if rib_want:    # (empty right after the instance was disabled or its database cleared)
        assert table == rib_want
