"""CPU: partial SPF runs of OSPFv2 (update_rib_partial, holo-ospf/src/route.rs:196-340): the product's host
stage (csrc/ospf_rib_partial.cc) against the restatement (oracle/rib_partial.cc) on random multi-area states
and summary / external LSA changes, chained over several runs; hand cases for each branch of the reference."""
import copy

import numpy as np
import pytest

import test_ospf_rib as T
from holo_b200 import ospf_rib, ospfv2
from oracle import pyoracle


def installed_state(router_id, max_paths, areas, ext):
    """Full run + update_global_rib: the table with INSTALLED flags and the router tables."""
    rib = ospf_rib.update_rib_full(router_id, max_paths, areas, ext)
    _, flagged = ospf_rib.rib_diff(None, rib)
    return ospf_rib.Rib(flagged, rib.nexthops), ospf_rib.router_tables(router_id, areas)


def mutate(rng, areas, ext):
    """Change, withdraw or revive some summary / external LSAs; returns the new LSDB parts and the triggers."""
    areas2 = copy.deepcopy(areas)
    ext2 = ext.copy()
    tr = []
    for a in areas2:
        for i in range(len(a.summaries)):
            if rng.random() < 0.3:
                s = a.summaries
                r = rng.random()
                if r < 0.5:
                    s["metric"][i] = int(rng.choice([1, 5, 10, 20, ospf_rib.LSA_INFINITY]))
                elif r < 0.8:
                    s["maxage"][i] = 1 - int(s["maxage"][i])
                tr.append((int(s["adv_rtr"][i]), int(s["lsa_id"][i]), int(s["mask"][i]), int(s["lsa_type"][i]), 0, (0, 0)))
    for i in range(len(ext2)):
        if rng.random() < 0.3:
            if rng.random() < 0.6:
                ext2["metric"][i] = int(rng.choice([1, 10, 20, ospf_rib.LSA_INFINITY]))
            else:
                ext2["maxage"][i] = 1 - int(ext2["maxage"][i])
            tr.append((int(ext2["adv_rtr"][i]), int(ext2["lsa_id"][i]), int(ext2["mask"][i]), 5, 0, (0, 0)))
    return areas2, ext2, tr


def same_tables(a, b):
    assert a.rtrs.tobytes() == b.rtrs.tobytes() and a.nexthops.tobytes() == b.nexthops.tobytes()


@pytest.mark.parametrize("seed", range(150))
def test_partial_runs_match_restatement(seed):
    rng = np.random.default_rng(7000 + seed)
    router_id, max_paths, areas, ext = T.random_instance(seed)
    same_tables(ospf_rib.router_tables(router_id, areas), pyoracle.ospfv2_rib_router_tables(router_id, areas))
    rib, rtrs = installed_state(router_id, max_paths, areas, ext)
    for _ in range(3):                                  # a chain of partial runs over the evolving state
        areas, ext, tr = mutate(rng, areas, ext)
        kind, net, rtr, xt = ospfv2.spf_computation_type(tr)
        assert kind == ospfv2.SPF_PARTIAL
        a = ospf_rib.update_rib_partial(router_id, max_paths, areas, ext, (net, rtr, xt), rib, rtrs)
        b = pyoracle.ospfv2_update_rib_partial(router_id, max_paths, areas, ext, (net, rtr, xt), rib, rtrs)
        assert a[0].routes.tobytes() == b[0].routes.tobytes(), (a[0].routes, b[0].routes)
        assert a[0].nexthops.tobytes() == b[0].nexthops.tobytes()
        same_tables(a[1], b[1])
        assert a[2].tobytes() == b[2].tobytes()
        STATS["runs"] += 1
        STATS["actions"] += len(a[2])
        full = ospf_rib.update_rib_full(router_id, max_paths, areas, ext)
        keys = lambda r: {(int(x["prefix"]), int(x["mask"])): (int(x["path_type"]), int(x["metric"]), int(x["type2_metric"])) for x in r.routes}
        STATS["as_full"] += int(keys(full) == keys(a[0]))
        rib, rtrs = a[0], a[1]


STATS = {"runs": 0, "actions": 0, "as_full": 0}


def test_partial_fuzz_is_not_vacuous():
    if STATS["runs"] < 400:
        pytest.skip("runs after the fuzz")
    assert STATS["actions"] > 100
    # the partial walk usually lands on the table a full recomputation gives (it may not: module docstring of
    # csrc/ospf_rib_partial.cc); both happen in the fuzz
    assert 0.5 * STATS["runs"] < STATS["as_full"] < STATS["runs"]


# ---- hand cases -------------------------------------------------------------------------------------------
RID, ABR, ASBR = 0x01010101, 0x02020201, 0x02020209
NET = (0x0A000000, 0xFFFFFF00)


def one_area(summaries, routers=None):
    """Backbone with one interface (sort key 1); ABR at metric 10 and ASBR at metric 30 as intra-area routers."""
    ifaces = np.zeros(1, ospfv2.IFACE_DT)
    ifaces[0] = (101, 1, ospfv2.IF_P2P, (0, 0, 0), 0, 0, 0, 0)
    nhs = np.asarray([(0, 7, ABR, 0, 1, 1, 0, 0), (0, 9, ASBR, 0, 1, 1, 0, 0)], ospfv2.NEXTHOP_DT)
    rt = routers if routers is not None else [(ABR, 10, 1, 2, (0, 0), 0, 1), (ASBR, 30, 2, 2, (0, 0), 1, 1)]   # flags: B = 1, E = 2
    res = ospfv2.Ospfv2Result(np.zeros(0, ospfv2.SPT_VERTEX_DT), np.asarray(rt, ospfv2.ROUTE_RTR_DT), np.zeros(0, ospfv2.ROUTE_NET_DT),
                              nhs, False, True)
    sm = np.asarray(summaries, ospf_rib.SUMMARY_LSA_DT) if summaries else np.zeros(0, ospf_rib.SUMMARY_LSA_DT)
    return [ospf_rib.RibArea(0, res, ifaces, sm, True)]


def ext_lsas(rows):
    return np.asarray(rows, ospf_rib.EXTERNAL_LSA_DT) if rows else np.zeros(0, ospf_rib.EXTERNAL_LSA_DT)


def run(areas0, ext0, areas1, ext1, sets):
    rib, rtrs = installed_state(RID, 16, areas0, ext0)
    a = ospf_rib.update_rib_partial(RID, 16, areas1, ext1, sets, rib, rtrs)
    b = pyoracle.ospfv2_update_rib_partial(RID, 16, areas1, ext1, sets, rib, rtrs)
    assert a[0].routes.tobytes() == b[0].routes.tobytes() and a[2].tobytes() == b[2].tobytes()
    return rib, a


def test_summary_metric_change_reinstalls_that_route_only():
    other = (0x0A000100, 0xFFFFFF00)
    s0 = [(ABR, NET[0], NET[1], 5, 3, 0, (0, 0)), (ABR, other[0], other[1], 5, 3, 0, (0, 0))]
    s1 = [(ABR, NET[0], NET[1], 8, 3, 0, (0, 0)), (ABR, other[0], other[1], 5, 3, 0, (0, 0))]
    prev, (rib, _, acts) = run(one_area(s0), ext_lsas([]), one_area(s1), ext_lsas([]), ([NET], [], []))
    assert [int(a["kind"]) for a in acts] == [ospf_rib.RIB_INSTALL]
    r = rib.routes[int(acts[0]["route"])]
    assert (int(r["prefix"]), int(r["metric"]), int(r["path_type"])) == (NET[0], 18, ospf_rib.PATH_INTER)
    assert all(int(x["flags"]) & ospf_rib.ROUTE_INSTALLED for x in rib.routes) and len(rib.routes) == 2


def test_withdrawn_summary_falls_back_to_an_external_path():
    """The prefix loses its inter-area route; `partial.external.extend(old_rib.keys())` lets the AS-external LSA
    for the same prefix take over in the same run (route.rs:283-287)."""
    s0 = [(ABR, NET[0], NET[1], 5, 3, 0, (0, 0))]
    s1 = [(ABR, NET[0], NET[1], 5, 3, 1, (0, 0))]                       # MaxAge
    ext = ext_lsas([(ASBR, NET[0], NET[1], 7, 0, 0, 0, 0, (0, 0))])
    prev, (rib, _, acts) = run(one_area(s0), ext, one_area(s1), ext, ([NET], [], []))
    assert int(prev.routes[0]["path_type"]) == ospf_rib.PATH_INTER      # the inter-area route hid the external one
    assert len(rib.routes) == 1 and int(rib.routes[0]["path_type"]) == ospf_rib.PATH_TYPE1 and int(rib.routes[0]["metric"]) == 37
    assert [int(a["kind"]) for a in acts] == [ospf_rib.RIB_INSTALL]
    # without an external LSA the route goes away
    prev, (rib, _, acts) = run(one_area(s0), ext_lsas([]), one_area(s1), ext_lsas([]), ([NET], [], []))
    assert len(rib.routes) == 0 and [int(a["kind"]) for a in acts] == [ospf_rib.RIB_UNINSTALL_OLD] and int(acts[0]["route"]) == 0


def test_type4_change_reevaluates_every_external_route():
    far = 0x02020277                                                     # an ASBR only a type-4 LSA reaches
    s0 = [(ABR, far, 0, 5, 4, 0, (0, 0))]
    s1 = [(ABR, far, 0, 9, 4, 0, (0, 0))]
    ext = ext_lsas([(far, NET[0], NET[1], 7, 0, 0, 0, 0, (0, 0)), (far, 0x0B000000, 0xFF000000, 3, 0, 0, 1, 0, (0, 0))])
    prev, (rib, rtrs, acts) = run(one_area(s0), ext, one_area(s1), ext, ([], [far], []))
    assert sorted(int(a["kind"]) for a in acts) == [ospf_rib.RIB_INSTALL]              # the type-2 route keeps its reported metric
    byp = {int(r["prefix"]): r for r in rib.routes}
    assert int(byp[NET[0]]["metric"]) == 10 + 9 + 7 and int(byp[0x0B000000]["metric"]) == 19 and int(byp[0x0B000000]["type2_metric"]) == 3
    e = [r for r in rtrs.rtrs if int(r["router_id"]) == far][0]
    assert int(e["metric"]) == 19 and int(e["path_type"]) == ospf_rib.PATH_INTER


def test_recomputed_routes_do_not_see_the_routes_that_stayed():
    """A reference quirk kept on purpose: the recomputed inter-area route replaces an intra-area route for the same
    prefix, because it is built in a side table and laid over the previous one (route.rs:206, 337-339)."""
    areas = one_area([(ABR, NET[0], NET[1], 5, 3, 0, (0, 0))])
    intra = np.asarray([(NET[0], NET[1], 3, 0, 1, 0, 0, RID, RID, 0, 0, 0, (0, 0), 0, 0, 1)], ospfv2.ROUTE_NET_DT)
    areas[0].result.routes = intra
    rib, rtrs = installed_state(RID, 16, areas, ext_lsas([]))
    assert int(rib.routes[0]["path_type"]) == ospf_rib.PATH_INTRA
    new, _, acts = ospf_rib.update_rib_partial(RID, 16, areas, ext_lsas([]), ([NET], [], []), rib, rtrs)
    assert int(new.routes[0]["path_type"]) == ospf_rib.PATH_INTER and int(new.routes[0]["metric"]) == 15
    assert [int(a["kind"]) for a in acts] == [ospf_rib.RIB_INSTALL]
    full = ospf_rib.update_rib_full(RID, 16, areas, ext_lsas([]))
    assert int(full.routes[0]["path_type"]) == ospf_rib.PATH_INTRA        # what a full run says


def test_argument_checks():
    areas = one_area([])
    rib, rtrs = installed_state(RID, 16, areas, ext_lsas([]))
    lib = ospf_rib.capi.load_library()
    import ctypes as C
    assert lib.hspf_ospfv2_update_rib_partial(RID, 16, None, None, 0, None, 0, None, None, None, None, None, None, 0, None) == ospf_rib.capi.HSPF_E_INVAL
    assert lib.hspf_ospfv2_rib_router_tables(RID, None, 1, None) == ospf_rib.capi.HSPF_E_INVAL
