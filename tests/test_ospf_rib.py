"""Host stage hspf_ospfv2_update_rib_full (holo_b200/csrc/ospf_rib_host.cc) on the CPU:
(1) fed with the per-area results of the oracle's run_area it must reproduce the whole
local-rib of every golden OSPFv2 snapshot of the reference, byte-identical to the oracle's
restatement (oracle/rib_ospfv2.cc); (2) on random multi-area tables with type-3/4/5 LSAs,
ECMP and max_paths truncation it must agree with the restatement record for record."""
import numpy as np
import pytest

import golden_util as gu
from holo_b200 import ospf_rib, ospfv2
from oracle import pyoracle

SNAPS = gu.load_ospfv2()


@pytest.mark.parametrize("snap", SNAPS, ids=[f"{s['topo']}-{s['rt']}" for s in SNAPS])
def test_product_rib_stage_reproduces_reference_local_rib(snap):
    got = gu.ospfv2_full_rib(snap, pyoracle.ospfv2_run_area, ospf_rib.update_rib_full)
    want = gu.golden_rib(snap)
    assert set(got) == set(want)
    for prefix, (metric, rtype, nh) in want.items():
        g = got[prefix]
        assert (g[0], g[1]) == (metric, rtype), (prefix, g)
        assert [(a or "", b or "") for a, b in g[2]] == [(a or "", b or "") for a, b in nh], (prefix, g[2], nh)


def random_instance(seed: int):
    """Fabricated per-area SPF results (router tables with ABR/ASBR flags, intra-area routes,
    ECMP next hops over globally unique interface sort keys) + summary and external LSAs."""
    rng = np.random.default_rng(seed)
    n_areas = int(rng.integers(1, 4))
    area_ids = [0] + sorted(int(x) for x in rng.choice(np.arange(1, 9), n_areas - 1, replace=False))
    rng.shuffle(area_ids)
    router_id = 0x01010101
    pool_prefixes = [(0x0A000000 + (i << 8), 0xFFFFFF00) for i in range(12)] + [(0x0A0A0A00 + i, 0xFFFFFFFF) for i in range(6)] + [(0, 0)]
    rtr_ids = [0x02020200 + i for i in range(10)]
    areas, sk = [], 1
    for aid in area_ids:
        n_if = int(rng.integers(1, 5))
        ifaces = np.zeros(n_if, ospfv2.IFACE_DT)
        for i in range(n_if):
            ifaces[i] = (100 + sk, sk, ospfv2.IF_P2P, (0, 0, 0), 0, 0, 0, 0)
            sk += 1
        nhs = []

        def hops():
            k = int(rng.integers(0, 4))
            off = len(nhs)
            seen = set()
            for _ in range(k):
                i = int(rng.integers(0, n_if))
                ha = int(rng.integers(0, 2))
                addr = int(rng.integers(1, 5)) if ha else 0
                if (i, ha, addr) in seen:
                    continue
                seen.add((i, ha, addr))
                nhs.append((i, addr, int(rng.integers(1, 99)), 0, ha, 1, 0, 0))
            sub = sorted(nhs[off:], key=lambda x: (ifaces[x[0]]["sort_key"], x[4], x[1]))
            nhs[off:] = sub
            return off, len(nhs) - off

        routers = []
        for rid in sorted(rng.choice(rtr_ids, int(rng.integers(1, 7)), replace=False)):
            off, n = hops()
            routers.append((int(rid), int(rng.integers(1, 60)), int(rng.integers(0, 4)), 2, (0, 0), off, n))
        routes = []
        for pi in sorted(rng.choice(len(pool_prefixes) - 1, int(rng.integers(1, 8)), replace=False)):
            p, m = pool_prefixes[int(pi)]
            off, n = hops()
            # LS origin: a router (stub link) or a transit network with a small LSA id, so that the
            # cross-area transit-network rule (route.rs:387-397) meets equal metrics and both id orders
            routes.append((p, m, int(rng.choice([5, 5, 10, 20, 33])), int(rng.integers(0, 2)), int(rng.choice([1, 1, 2])),
                           0, 0, 0, int(rng.integers(1, 6)), 0, 0, 0, (0, 0), 0, off, n))
        res = ospfv2.Ospfv2Result(np.zeros(0, ospfv2.SPT_VERTEX_DT), np.asarray(routers, ospfv2.ROUTE_RTR_DT),
                                  np.asarray(routes, ospfv2.ROUTE_NET_DT), np.asarray(nhs, ospfv2.NEXTHOP_DT)
                                  if nhs else np.zeros(0, ospfv2.NEXTHOP_DT), bool(rng.integers(0, 2)), True)
        sums = []
        for _ in range(int(rng.integers(0, 14))):
            ty = 3 if rng.random() < 0.7 else 4
            adv = int(rng.choice(rtr_ids + [router_id]))
            if ty == 3:
                p, m = pool_prefixes[int(rng.integers(0, len(pool_prefixes)))]
                if rng.random() < 0.1:
                    p |= 1                                     # host bits set: kept unmasked
            else:
                p, m = int(rng.choice(rtr_ids)), 0
            metric = int(rng.choice([1, 5, 10, 10, 20, ospf_rib.LSA_INFINITY]))
            sums.append((adv, p, m, metric, ty, int(rng.random() < 0.1), (0, 0)))
        sums.sort(key=lambda x: (x[4], x[0], x[1]))
        areas.append(ospf_rib.RibArea(aid, res, ifaces, np.asarray(sums, ospf_rib.SUMMARY_LSA_DT)
                                      if sums else np.zeros(0, ospf_rib.SUMMARY_LSA_DT), bool(rng.random() < 0.8)))
    ext = []
    for _ in range(int(rng.integers(0, 10))):
        p, m = pool_prefixes[int(rng.integers(0, len(pool_prefixes)))]
        ext.append((int(rng.choice(rtr_ids)), p, m, int(rng.choice([1, 10, 20, ospf_rib.LSA_INFINITY])), 0,
                    int(rng.integers(0, 5)), int(rng.integers(0, 2)), int(rng.random() < 0.1), (0, 0)))
    ext.sort(key=lambda x: (x[0], x[1]))
    ext = np.asarray(ext, ospf_rib.EXTERNAL_LSA_DT) if ext else np.zeros(0, ospf_rib.EXTERNAL_LSA_DT)
    return router_id, int(rng.choice([1, 2, 16])), areas, ext


@pytest.mark.parametrize("seed", range(200))
def test_product_rib_stage_matches_restatement_on_random_tables(seed):
    router_id, max_paths, areas, ext = random_instance(seed)
    a = ospf_rib.update_rib_full(router_id, max_paths, areas, ext)
    b = pyoracle.ospfv2_update_rib_full(router_id, max_paths, areas, ext)
    assert b.rc == 0
    assert a.routes.tobytes() == b.routes.tobytes(), (a.routes, b.routes)
    assert a.nexthops.tobytes() == b.nexthops.tobytes()


def test_random_tables_exercise_every_stage():
    kinds = set()
    n_multi = 0
    for seed in range(200):
        router_id, max_paths, areas, ext = random_instance(seed)
        rib = ospf_rib.update_rib_full(router_id, max_paths, areas, ext)
        kinds |= set(int(p) for p in rib.routes["path_type"])
        n_multi += int((rib.routes["n_nh"] > 1).sum())
    assert kinds == {0, 1, 2, 3} and n_multi > 20


def test_rib_stage_argument_checks():
    import ctypes as C
    from holo_b200 import capi
    lib = capi.load_library()
    fn = lib.hspf_ospfv2_update_rib_full
    fn.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
    assert fn(1, 16, None, 1, None, 0, None) == capi.HSPF_E_INVAL
    r = ospf_rib.RibStruct()
    assert fn(1, 16, None, 0, None, 0, C.byref(r)) == capi.HSPF_OK and r.n_routes == 0


# ------------------------------------------------------------------------------ OSPFv3
SNAPS_V3 = gu.load_ospfv3()


@pytest.mark.parametrize("snap", SNAPS_V3, ids=[f"{s['topo']}-{s['rt']}" for s in SNAPS_V3])
def test_product_rib_stage_v3_reproduces_reference_local_rib(snap):
    got = gu.ospfv3_full_rib(snap, pyoracle.ospfv3_run_area, ospf_rib.update_rib_full_v3)
    want = gu.golden_rib(snap)
    assert set(got) == set(want)
    for prefix, (metric, rtype, nh) in want.items():
        g = got[prefix]
        assert (g[0], g[1]) == (metric, rtype), (prefix, g)
        assert [(a or "", b or "") for a, b in g[2]] == [(a or "", b or "") for a, b in nh], (prefix, g[2], nh)


def random_instance_v3(seed: int):
    from holo_b200 import ospfv3
    rng = np.random.default_rng(10_000 + seed)
    n_areas = int(rng.integers(1, 4))
    area_ids = [0] + sorted(int(x) for x in rng.choice(np.arange(1, 9), n_areas - 1, replace=False))
    rng.shuffle(area_ids)
    router_id = 0x01010101
    pool = [(f"2001:db8:{i:x}::", 64) for i in range(12)] + [(f"2001:db8:ffff::{i:x}", 128) for i in range(1, 7)] + [("::", 0)]
    rtr_ids = [0x02020200 + i for i in range(10)]
    lla = [f"fe80::{i:x}" for i in range(1, 6)]
    areas, sk = [], 1
    for aid in area_ids:
        n_if = int(rng.integers(1, 5))
        ifaces = np.zeros(n_if, ospfv3.IFACE_DT)
        for i in range(n_if):
            ifaces[i] = (100 + sk, sk, 0, (0, 0, 0))
            sk += 1
        nhs = []

        def hops():
            k = int(rng.integers(0, 4))
            off = len(nhs)
            seen = set()
            for _ in range(k):
                i = int(rng.integers(0, n_if))
                ha = int(rng.integers(0, 2))
                addr = lla[int(rng.integers(0, len(lla)))] if ha else "::"
                if (i, ha, addr) in seen:
                    continue
                seen.add((i, ha, addr))
                nhs.append((i, int(rng.integers(1, 99)), ospfv3.ip_rec(addr), ha, 1, (0, 0)))
            return off, len(nhs) - off

        routers = []
        for rid in sorted(rng.choice(rtr_ids, int(rng.integers(1, 7)), replace=False)):
            off, n = hops()
            routers.append((int(rid), int(rng.integers(1, 60)), int(rng.integers(0, 4)), 2, (0, 0), off, n))
        routes = []
        for pi in sorted(rng.choice(len(pool) - 1, int(rng.integers(1, 8)), replace=False)):
            p, ln = pool[int(pi)]
            off, n = hops()
            routes.append((ospfv3.ip_rec(p), ln, int(rng.integers(0, 2)), int(rng.choice([1, 1, 2])), 0,
                           int(rng.choice([5, 5, 10, 20, 33])), 0, int(rng.integers(1, 6)), off, n))
        mk = lambda rows, dt: np.asarray(rows, dtype=dt) if rows else np.zeros(0, dt)
        res = ospfv3.Ospfv3Result(np.zeros(0, ospfv3.SPT_VERTEX6_DT), mk(routers, ospfv2.ROUTE_RTR_DT),
                                  mk(routes, ospfv3.ROUTE_NET6_DT), mk(nhs, ospfv3.NEXTHOP6_DT),
                                  bool(rng.integers(0, 2)), True)
        sums = []
        for _ in range(int(rng.integers(0, 14))):
            ty = 3 if rng.random() < 0.7 else 4
            adv = int(rng.choice(rtr_ids + [router_id]))
            metric = int(rng.choice([1, 5, 10, 10, 20, ospf_rib.LSA_INFINITY]))
            if ty == 3:
                p, ln = pool[int(rng.integers(0, len(pool)))]
                sums.append((adv, len(sums) + 1, metric, 0, ospfv3.ip_rec(p), ln,
                             ospfv3.PFX_NU if rng.random() < 0.1 else 0, 3, int(rng.random() < 0.1)))
            else:
                sums.append((adv, len(sums) + 1, metric, int(rng.choice(rtr_ids)), ospfv3.ip_rec("::"), 0, 0, 4,
                             int(rng.random() < 0.1)))
        sums.sort(key=lambda x: (x[7], x[0], x[1]))
        areas.append(ospf_rib.RibArea(aid, res, ifaces, mk(sums, ospf_rib.INTER_AREA_LSA_DT), bool(rng.random() < 0.8)))
    ext = []
    for k in range(int(rng.integers(0, 10))):
        p, ln = pool[int(rng.integers(0, len(pool)))]
        ext.append((int(rng.choice(rtr_ids)), k + 1, int(rng.choice([1, 10, 20, ospf_rib.LSA_INFINITY])),
                    int(rng.integers(0, 5)), ospfv3.ip_rec(p), ln, ospfv3.PFX_NU if rng.random() < 0.1 else 0,
                    int(rng.integers(0, 2)), int(rng.random() < 0.1)))
    ext.sort(key=lambda x: (x[0], x[1]))
    ext = np.asarray(ext, ospf_rib.EXTERNAL6_LSA_DT) if ext else np.zeros(0, ospf_rib.EXTERNAL6_LSA_DT)
    return router_id, int(rng.choice([1, 2, 16])), areas, ext


@pytest.mark.parametrize("seed", range(200))
def test_product_rib_stage_v3_matches_restatement_on_random_tables(seed):
    router_id, max_paths, areas, ext = random_instance_v3(seed)
    a = ospf_rib.update_rib_full_v3(router_id, max_paths, areas, ext)
    b = pyoracle.ospfv3_update_rib_full(router_id, max_paths, areas, ext)
    assert b.rc == 0
    assert a.routes.tobytes() == b.routes.tobytes(), (a.routes, b.routes)
    assert a.nexthops.tobytes() == b.nexthops.tobytes()


def test_random_tables_v3_exercise_every_stage():
    kinds = set()
    n_multi = 0
    for seed in range(200):
        router_id, max_paths, areas, ext = random_instance_v3(seed)
        rib = ospf_rib.update_rib_full_v3(router_id, max_paths, areas, ext)
        kinds |= set(int(p) for p in rib.routes["path_type"])
        n_multi += int((rib.routes["n_nh"] > 1).sum())
    assert kinds == {0, 1, 2, 3} and n_multi > 20


# ---- LSDB -> whole routing table through product host code only (SPT planes from the oracle) ----
def _planes(csr, root, nh_words):
    c = pyoracle.csr_spf(csr, root, nh_words=nh_words)
    return c["dist"], c["hops"], c["nh_mask"]


@pytest.mark.parametrize("snap", SNAPS, ids=[f"{s['topo']}-{s['rt']}" for s in SNAPS])
def test_product_host_pipeline_reproduces_reference_local_rib(snap):
    got = gu.ospfv2_full_rib(snap, lambda img: ospfv2.area_from_planes(img, _planes), ospf_rib.update_rib_full)
    assert got == {k: (m, t, [(a, b) for a, b in nh]) for k, (m, t, nh) in gu.golden_rib(snap).items()} or \
        {k: (v[0], v[1], [(a or "", b or "") for a, b in v[2]]) for k, v in got.items()} == \
        {k: (m, t, [(a or "", b or "") for a, b in nh]) for k, (m, t, nh) in gu.golden_rib(snap).items()}


@pytest.mark.parametrize("snap", SNAPS_V3, ids=[f"{s['topo']}-{s['rt']}" for s in SNAPS_V3])
def test_product_host_pipeline_v3_reproduces_reference_local_rib(snap):
    from holo_b200 import ospfv3
    got = gu.ospfv3_full_rib(snap, lambda img: ospfv3.area_from_planes(img, _planes), ospf_rib.update_rib_full_v3)
    assert {k: (v[0], v[1], [(a or "", b or "") for a, b in v[2]]) for k, v in got.items()} == \
        {k: (m, t, [(a or "", b or "") for a, b in nh]) for k, (m, t, nh) in gu.golden_rib(snap).items()}


# ---- update_global_rib: the messages to the RIB manager ------------------------------------------
IBUS_V2 = [s for s in SNAPS if s.get("ibus_routes") is not None]
IBUS_V3 = [s for s in SNAPS_V3 if s.get("ibus_routes") is not None]


@pytest.mark.parametrize("snap", IBUS_V2, ids=[f"{s['topo']}-{s['rt']}" for s in IBUS_V2])
def test_installs_equal_the_reference_ibus_stream(snap):
    """LSDB -> routing table -> update_global_rib from an empty table, all through product host code
    (SPT planes from the oracle): exactly the RouteIpAdd set the reference sent to the RIB manager
    (output/ibus.jsonl), prefix by prefix with metric, ifindex and next-hop address."""
    areas_rib = gu.ospfv2_full_rib  # noqa: F841  (same pipeline, kept as structured arrays below)
    keys = gu.global_sort_keys(snap)
    areas = []
    for area in snap["areas"]:
        img = gu.ospfv2_area_image(snap, area, keys)
        res = ospfv2.area_from_planes(img, _planes)
        if res.root_found:
            active = any((i.get("state") or "down") != "down" for i in area["interfaces"])
            areas.append(ospf_rib.RibArea(gu.ip(area["area_id"]), res, img.ifaces, gu.ospfv2_summaries(area), active))
    rib = ospf_rib.update_rib_full(gu.ip(snap["router_id"]), 16, areas)
    got = gu.installs_from_empty(snap, rib, ospf_rib.rib_diff)
    assert got == gu.golden_ibus(snap)
    # and the restatement says the same
    want = gu.installs_from_empty(snap, rib, lambda o, n: ospf_rib.call_rib_diff(pyoracle.lib().oracle_ospfv2_rib_diff, o, n))
    assert got == want


@pytest.mark.parametrize("snap", IBUS_V3, ids=[f"{s['topo']}-{s['rt']}" for s in IBUS_V3])
def test_installs_v3_equal_the_reference_ibus_stream(snap):
    from holo_b200 import ospfv3
    keys = gu.global_sort_keys(snap)
    areas = []
    for area in snap["areas"]:
        img = gu.ospfv3_area_image(snap, area, keys)
        res = ospfv3.area_from_planes(img, _planes)
        if res.root_found:
            active = any((i.get("state") or "down") != "down" for i in area["interfaces"])
            areas.append(ospf_rib.RibArea(gu.ip(area["area_id"]), res, img.ifaces, gu.ospfv3_inter_area_lsas(area), active))
    rib = ospf_rib.update_rib_full_v3(gu.ip(snap["router_id"]), 16, areas)
    got = gu.installs_from_empty(snap, rib, lambda o, n: ospf_rib.rib_diff(o, n, v3=True), v3=True)
    assert got == gu.golden_ibus(snap)


@pytest.mark.parametrize("seed", range(150))
def test_rib_diff_matches_restatement_on_random_table_pairs(seed):
    """Two random tables of the same instance (the second one perturbed): product vs restatement,
    action by action and flag by flag, then the second diff with the flags of the first."""
    router_id, max_paths, areas, ext = random_instance(seed)
    old = ospf_rib.update_rib_full(router_id, max_paths, areas, ext)
    router_id2, max_paths2, areas2, ext2 = random_instance(seed + 1 if seed % 3 else seed)
    new = ospf_rib.update_rib_full(router_id2, max_paths2, areas2, ext2)
    olib = pyoracle.lib()
    a0, f0 = ospf_rib.rib_diff(None, old)
    b0, g0 = ospf_rib.call_rib_diff(olib.oracle_ospfv2_rib_diff, None, old)
    assert a0.tobytes() == b0.tobytes() and f0.tobytes() == g0.tobytes()
    old_inst = ospf_rib.Rib(f0, old.nexthops)
    a1, f1 = ospf_rib.rib_diff(old_inst, new)
    b1, g1 = ospf_rib.call_rib_diff(olib.oracle_ospfv2_rib_diff, old_inst, new)
    assert a1.tobytes() == b1.tobytes() and f1.tobytes() == g1.tobytes()
    if seed % 3 == 0:          # identical tables: nothing to tell the RIB manager, flags carried over
        assert len(a1) == 0 and f1.tobytes() == f0.tobytes()


def test_step_interface_cost_change_reinstalls_exactly_the_changed_route():
    """The reference's step test nb-config-iface-cost1 (holo-ospf/tests/conformance/ospfv2/mod.rs:
    494-499, topo1-1 rt2): the cost of eth-rt1 goes 10 -> 50, the router re-originates its
    Router-LSA of area 0.0.0.1 with the new metric, and the only message to the RIB manager is the
    reinstall of 1.1.1.1/32 with metric 50 (02-output-ibus.jsonl; 10.0.1.0/24 is connected)."""
    snap = [s for s in SNAPS if s.get("steps", {}).get("nb-config-iface-cost1")][0]
    step = snap["steps"]["nb-config-iface-cost1"]
    keys = gu.global_sort_keys(snap)

    def table(changed):
        areas = []
        for area in snap["areas"]:
            img = gu.ospfv2_area_image(snap, area, keys)
            if changed and area["area_id"] == step["change"]["area"]:
                # own Router-LSA: the p2p link over that interface and its stub link take the new cost
                nbr = [n for i in area["interfaces"] if i["name"] == step["change"]["iface"] for n in i["neighbors"]][0]
                me = [k for k in range(len(img.router_lsas)) if int(img.router_lsas["adv_rtr"][k]) == gu.ip(snap["router_id"])][0]
                off, n = int(img.router_lsas["link_off"][me]), int(img.router_lsas["n_links"][me])
                nbr_net = gu.ip(nbr[1]) & 0xFFFFFF00
                hit = 0
                for k in range(off, off + n):
                    l = img.links[k]
                    if (int(l["link_type"]) == ospfv2.LINK_P2P and int(l["link_id"]) == gu.ip(nbr[0])) or \
                            (int(l["link_type"]) == ospfv2.LINK_STUB and int(l["link_id"]) == nbr_net):
                        img.links["metric"][k] = step["change"]["cost"]
                        hit += 1
                assert hit == 2
            res = ospfv2.area_from_planes(img, _planes)
            if res.root_found:
                active = any((i.get("state") or "down") != "down" for i in area["interfaces"])
                areas.append(ospf_rib.RibArea(gu.ip(area["area_id"]), res, img.ifaces, gu.ospfv2_summaries(area), active))
        return ospf_rib.update_rib_full(gu.ip(snap["router_id"]), 16, areas)

    old = table(False)
    _a, installed = ospf_rib.rib_diff(None, old)
    new = table(True)
    acts, _f = ospf_rib.rib_diff(ospf_rib.Rib(installed, old.nexthops), new)
    key_name = {v: k for k, v in keys.items()}
    got = {}
    for a in acts:
        assert int(a["kind"]) == ospf_rib.RIB_INSTALL
        r = new.routes[int(a["route"])]
        hops = new.nexthops[int(r["nh_off"]): int(r["nh_off"]) + int(r["n_nh"])]
        got[f"{gu.ipstr(r['prefix'])}/{bin(int(r['mask'])).count('1')}"] = (
            int(r["metric"]), sorted((snap["ifindex"].get(key_name[int(x["iface"])], 0), gu.ipstr(x["addr"])) for x in hops))
    want = {p: (v["metric"], sorted((n[0], n[1]) for n in v["nexthops"])) for p, v in step["ibus_routes"].items()}
    assert got == want


AFTER_V2 = [(s, name) for s in SNAPS for name in s.get("after", {})]


def _table_v2(snap, keys0):
    """Whole table of a snapshot through product host code; interfaces keep the sort keys of the
    topology snapshot (arena slots are stable across deletions), new ones are appended."""
    keys = dict(keys0)
    for n in sorted({i["name"] for a in snap["areas"] for i in a["interfaces"]}):
        keys.setdefault(n, max(keys.values(), default=0) + 1)
    key_name = {v: k for k, v in keys.items()}
    if not snap.get("router_id"):            # instance disabled: no table
        return ospf_rib.Rib(np.zeros(0, ospf_rib.RIB_ROUTE_DT), np.zeros(0, ospfv2.NEXTHOP_DT)), key_name
    areas = []
    for area in snap["areas"]:
        img = gu.ospfv2_area_image(snap, area, keys)
        res = ospfv2.area_from_planes(img, _planes)
        if res.root_found:
            active = any((i.get("state") or "down") != "down" for i in area["interfaces"])
            areas.append(ospf_rib.RibArea(gu.ip(area["area_id"]), res, img.ifaces, gu.ospfv2_summaries(area), active))
    return ospf_rib.update_rib_full(gu.ip(snap["router_id"]), 16, areas), key_name


@pytest.mark.parametrize("snap,name", AFTER_V2, ids=[f"{n}-{s['topo']}-{s['rt']}" for s, n in AFTER_V2])
def test_step_recomputation_gives_the_reference_ibus_output(snap, name):
    """19 step tests of the reference (holo-ospf/tests/conformance/ospfv2/mod.rs): LSA expiry,
    area / interface / instance configuration changes, router-id change, neighbour clearing and
    time-outs, interface and address events.  The state the reference reached after the step is one
    more snapshot; table of the topology snapshot (installed) vs table of the after-state, through
    product host code + hspf_ospfv2_rib_diff, == the step's ibus output, message for message."""
    after = dict(snap["after"][name], ifindex=snap["ifindex"])
    keys0 = gu.global_sort_keys(snap)
    old, _kn = _table_v2(snap, keys0)
    _a, installed = ospf_rib.rib_diff(None, old)
    old_inst = ospf_rib.Rib(installed, old.nexthops)
    new, kn = _table_v2(after, keys0)
    acts, _f = ospf_rib.rib_diff(old_inst, new)
    got = []
    for a in acts:
        if int(a["kind"]) == ospf_rib.RIB_INSTALL:
            r = new.routes[int(a["route"])]
            hops = new.nexthops[int(r["nh_off"]): int(r["nh_off"]) + int(r["n_nh"])]
            got.append(["add", f"{gu.ipstr(r['prefix'])}/{bin(int(r['mask'])).count('1')}", int(r["metric"]),
                        sorted([snap["ifindex"].get(kn[int(x["iface"])], 0), gu.ipstr(x["addr"]) if x["has_addr"] else None]
                               for x in hops)])
        else:
            assert int(a["kind"]) == ospf_rib.RIB_UNINSTALL_OLD
            r = old_inst.routes[int(a["route"])]
            got.append(["del", f"{gu.ipstr(r['prefix'])}/{bin(int(r['mask'])).count('1')}", None, []])
    assert got == [[k, p, m, sorted(nh)] for (k, p, m, nh) in after["ibus"]]
    # and the whole table of the after-state is the local-rib the reference reports there
    table = {}
    for r in new.routes:
        hops = new.nexthops[int(r["nh_off"]): int(r["nh_off"]) + int(r["n_nh"])]
        table[f"{gu.ipstr(r['prefix'])}/{bin(int(r['mask'])).count('1')}"] = (
            int(r["metric"]), ospf_rib.PATH_NAMES[int(r["path_type"])],
            sorted(((kn[int(x["iface"])], gu.ipstr(x["addr"]) if x["has_addr"] else None) for x in hops),
                   key=lambda x: (x[0] or "", x[1] or "")))
    want = gu.golden_rib(after)
    if not want:
        return      # the instance has just restarted (router-id change, disable): no SPF has run yet
    assert set(table) == set(want)
    for pfx, (metric, rtype, nh) in want.items():
        assert table[pfx][:2] == (metric, rtype), (pfx, table[pfx])
        assert [(a or "", b or "") for a, b in table[pfx][2]] == [(a or "", b or "") for a, b in nh], pfx


def test_rib_stage_rejects_out_of_range_next_hop_slices():
    from holo_b200 import capi
    router_id, max_paths, areas, ext = random_instance(3)
    bad = areas[0].result.routes.copy()
    if len(bad) == 0:
        pytest.skip("no routes in this instance")
    bad["nh_off"][0] = 10_000_000
    areas[0].result = ospfv2.Ospfv2Result(areas[0].result.vertices, areas[0].result.routers, bad, areas[0].result.nexthops,
                                          areas[0].result.transit_capability, True)
    with pytest.raises(capi.HspfError):
        ospf_rib.update_rib_full(router_id, max_paths, areas, ext)
    rib = ospf_rib.update_rib_full(*random_instance(4))
    if len(rib.routes):
        broken = ospf_rib.Rib(rib.routes.copy(), rib.nexthops)
        broken.routes["n_nh"][0] = 1_000_000
        with pytest.raises(capi.HspfError):
            ospf_rib.rib_diff(None, broken)


def _two_area_tables(origin_b, lsa_id_b, metric_b=10):
    """Area 0 reaches 10.0.0.0/24 through transit network 5 (metric 10, next hop on interface 1);
    area 1 reaches the same prefix with the given origin (next hop on interface 2)."""
    def area(aid, sk, origin_type, lsa_id, metric):
        ifaces = np.zeros(1, ospfv2.IFACE_DT)
        ifaces[0] = (100 + sk, sk, ospfv2.IF_P2P, (0, 0, 0), 0, 0, 0, 0)
        nhs = np.asarray([(0, 7, 0x02020200 + sk, 0, 1, 1, 0, 0)], ospfv2.NEXTHOP_DT)
        routes = np.asarray([(0x0A000000, 0xFFFFFF00, metric, 0, origin_type, 0, 0, 0x02020200 + sk, lsa_id, 0, 0, 0, (0, 0), 0, 0, 1)],
                            ospfv2.ROUTE_NET_DT)
        res = ospfv2.Ospfv2Result(np.zeros(0, ospfv2.SPT_VERTEX_DT), np.zeros(0, ospfv2.ROUTE_RTR_DT), routes, nhs, False, True)
        return ospf_rib.RibArea(aid, res, ifaces, np.zeros(0, ospf_rib.SUMMARY_LSA_DT), True)
    return [area(0, 1, 2, 5, 10), area(1, 2, origin_b, lsa_id_b, metric_b)]


@pytest.mark.parametrize("origin_b,lsa_id_b,metric_b,want", [
    (2, 7, 10, [2]),        # a transit network with a higher LSA id takes the entry over (route.rs:387-397)
    (2, 5, 10, [2]),        # ... or an equal one
    (2, 3, 10, [1]),        # a lower one stays out
    (1, 3, 10, [1, 2]),     # a stub link merges its next hops (route_update, route.rs:916-932)
    (2, 7, 11, [1]),        # a longer path never replaces
    (2, 3, 9, [1]),         # the id rule comes before route_compare: a shorter path with a lower id stays out too
])
def test_transit_network_rule_across_areas(origin_b, lsa_id_b, metric_b, want):
    areas = _two_area_tables(origin_b, lsa_id_b, metric_b)
    empty = np.zeros(0, ospf_rib.EXTERNAL_LSA_DT)
    for fn in (ospf_rib.update_rib_full, pyoracle.ospfv2_update_rib_full):
        rib = fn(0x01010101, 16, areas, empty)
        assert len(rib.routes) == 1
        r = rib.routes[0]
        hops = rib.nexthops[int(r["nh_off"]): int(r["nh_off"]) + int(r["n_nh"])]
        assert [int(h["iface"]) for h in hops] == want, fn
