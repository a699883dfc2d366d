"""OSPFv2 routing-table stages after the per-area SPFs, from Python: ctypes/numpy twins of
hl_ospfv2_summary_lsa / hl_ospfv2_external_lsa / hl_ospfv2_rib_area / hl_rib_route /
hl_ospfv2_rib (include/holo_lsdb.h) and the `hspf_ospfv2_update_rib_full` call
(update_rib_full, holo-ospf/src/route.rs:146-193: inter-area networks and routers, transit
areas, AS-external routes)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import capi, ospfv2

PATH_INTRA, PATH_INTER, PATH_TYPE1, PATH_TYPE2 = 0, 1, 2, 3
PATH_NAMES = {PATH_INTRA: "intra-area", PATH_INTER: "inter-area", PATH_TYPE1: "external-1", PATH_TYPE2: "external-2"}
LSA_INFINITY = 0x00FFFFFF

SUMMARY_LSA_DT = np.dtype([("adv_rtr", "<u4"), ("lsa_id", "<u4"), ("mask", "<u4"), ("metric", "<u4"),
                           ("lsa_type", "u1"), ("maxage", "u1"), ("_pad", "u1", (2,))], align=True)
EXTERNAL_LSA_DT = np.dtype([("adv_rtr", "<u4"), ("lsa_id", "<u4"), ("mask", "<u4"), ("metric", "<u4"),
                            ("fwd_addr", "<u4"), ("tag", "<u4"), ("e_bit", "u1"), ("maxage", "u1"),
                            ("_pad", "u1", (2,))], align=True)
RIB_ROUTE_DT = np.dtype([("prefix", "<u4"), ("mask", "<u4"), ("metric", "<u4"), ("type2_metric", "<u4"),
                         ("tag", "<u4"), ("area_id", "<u4"), ("path_type", "u1"), ("flags", "u1"), ("has_area", "u1"),
                         ("has_type2", "u1"), ("nh_off", "<u4"), ("n_nh", "<u4"), ("sr_label", "<u4"),
                         ("has_sr_label", "u1"), ("_pad", "u1", (3,))], align=True)
ROUTE_CONNECTED, ROUTE_INSTALLED = 0x01, 0x02
RIB_INSTALL, RIB_UNINSTALL, RIB_UNINSTALL_OLD = 1, 2, 3
ACTION_DT = np.dtype([("route", "<u4"), ("old_sr_label", "<u4"), ("kind", "u1"), ("has_old_sr_label", "u1"),
                      ("_pad", "u1", (2,))], align=True)


# OSPFv3 twins (hl_ospfv3_inter_area_lsa, hl_ospfv3_external_lsa, hl_rib_route6)
def _ip_dt():
    from . import ospfv3
    return ospfv3.IP_DT


INTER_AREA_LSA_DT = np.dtype([("adv_rtr", "<u4"), ("lsa_id", "<u4"), ("metric", "<u4"), ("router_id", "<u4"),
                              ("prefix", _ip_dt()), ("len", "u1"), ("prefix_options", "u1"), ("lsa_type", "u1"),
                              ("maxage", "u1")], align=True)
EXTERNAL6_LSA_DT = np.dtype([("adv_rtr", "<u4"), ("lsa_id", "<u4"), ("metric", "<u4"), ("tag", "<u4"),
                             ("prefix", _ip_dt()), ("len", "u1"), ("prefix_options", "u1"), ("e_bit", "u1"),
                             ("maxage", "u1")], align=True)
RIB_ROUTE6_DT = np.dtype([("prefix", _ip_dt()), ("len", "u1"), ("path_type", "u1"), ("flags", "u1"),
                          ("prefix_options", "u1"), ("has_area", "u1"), ("has_type2", "u1"), ("_pad", "u1", (2,)),
                          ("metric", "<u4"), ("type2_metric", "<u4"), ("tag", "<u4"), ("area_id", "<u4"),
                          ("nh_off", "<u4"), ("n_nh", "<u4")], align=True)


class RibAreaStruct(C.Structure):
    _fields_ = [
        ("area_id", C.c_uint32), ("n_summaries", C.c_uint32),
        ("spf", C.c_void_p), ("ifaces", C.c_void_p), ("summaries", C.c_void_p),
        ("n_ifaces", C.c_uint32), ("active", C.c_uint8), ("_pad", C.c_uint8 * 3),
    ]


class RibStruct(C.Structure):
    _fields_ = [
        ("routes_cap", C.c_uint32), ("n_routes", C.c_uint32), ("routes", C.c_void_p),
        ("nexthops_cap", C.c_uint32), ("n_nexthops", C.c_uint32), ("nexthops", C.c_void_p),
    ]


# appended to hspf_abi_sizes() after the OSPFv3 block
ABI_SIZES = [SUMMARY_LSA_DT.itemsize, EXTERNAL_LSA_DT.itemsize, C.sizeof(RibAreaStruct), RIB_ROUTE_DT.itemsize,
             C.sizeof(RibStruct),
             INTER_AREA_LSA_DT.itemsize, EXTERNAL6_LSA_DT.itemsize, C.sizeof(RibAreaStruct), RIB_ROUTE6_DT.itemsize,
             C.sizeof(RibStruct),       # hl_ospfv3_rib_area / hl_ospfv3_rib have the layouts of the v2 structs
             ACTION_DT.itemsize]


@dataclass
class RibArea:
    """One attached area: the result of run_area, the area's interfaces and Summary-LSAs."""
    area_id: int
    result: ospfv2.Ospfv2Result
    ifaces: np.ndarray
    summaries: np.ndarray
    active: bool = True


@dataclass
class Rib:
    routes: np.ndarray
    nexthops: np.ndarray
    rc: int = 0

    def nh(self, rec):
        return [tuple(int(x[k]) for k in ("iface", "has_addr", "addr", "has_nbr", "nbr_router_id", "has_label", "sr_label"))
                for x in self.nexthops[int(rec["nh_off"]): int(rec["nh_off"]) + int(rec["n_nh"])]]


def _version(v3: bool):
    """(result struct class, result dtypes, iface dtype, summary dtype, external dtype, route dtype, nexthop dtype)"""
    if not v3:
        return (ospfv2.ResultStruct,
                (("vertices", ospfv2.SPT_VERTEX_DT), ("routers", ospfv2.ROUTE_RTR_DT), ("routes", ospfv2.ROUTE_NET_DT),
                 ("nexthops", ospfv2.NEXTHOP_DT)),
                ospfv2.IFACE_DT, SUMMARY_LSA_DT, EXTERNAL_LSA_DT, RIB_ROUTE_DT, ospfv2.NEXTHOP_DT)
    from . import ospfv3
    return (ospfv3.ResultStruct,
            (("vertices", ospfv3.SPT_VERTEX6_DT), ("routers", ospfv2.ROUTE_RTR_DT), ("routes", ospfv3.ROUTE_NET6_DT),
             ("nexthops", ospfv3.NEXTHOP6_DT)),
            ospfv3.IFACE_DT, INTER_AREA_LSA_DT, EXTERNAL6_LSA_DT, RIB_ROUTE6_DT, ospfv3.NEXTHOP6_DT)


def _result_struct(res, keep: list, v3: bool = False):
    cls, fields = _version(v3)[:2]
    r = cls()
    for name, dt in fields:
        a = np.ascontiguousarray(getattr(res, name), dtype=dt)
        keep.append(a)
        setattr(r, name + "_cap", len(a))
        setattr(r, "n_" + name, len(a))
        setattr(r, name, a.ctypes.data if len(a) else None)
    r.transit_capability = int(res.transit_capability)
    r.root_found = int(res.root_found)
    return r


def call_update_rib_full(fn, router_id: int, max_paths: int, areas: list, externals=None, v3: bool = False) -> Rib:
    """`fn` = hspf_ospfv{2,3}_update_rib_full of the product library or the oracle's twin."""
    _cls, _fields, IFACE_DT_, SUM_DT_, EXT_DT_, ROUTE_DT_, NH_DT_ = _version(v3)
    fn.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(RibAreaStruct), C.c_uint32, C.c_void_p, C.c_uint32,
                   C.POINTER(RibStruct)]
    keep = []
    arr = (RibAreaStruct * max(len(areas), 1))()
    for i, a in enumerate(areas):
        rs = _result_struct(a.result, keep, v3)
        keep.append(rs)
        ifs = np.ascontiguousarray(a.ifaces, dtype=IFACE_DT_)
        sm = np.ascontiguousarray(a.summaries, dtype=SUM_DT_)
        keep += [ifs, sm]
        arr[i].area_id, arr[i].n_summaries = a.area_id, len(sm)
        arr[i].spf = C.addressof(rs)
        arr[i].ifaces = ifs.ctypes.data if len(ifs) else None
        arr[i].summaries = sm.ctypes.data if len(sm) else None
        arr[i].n_ifaces, arr[i].active = len(ifs), int(a.active)
    ext = np.ascontiguousarray(externals if externals is not None else np.zeros(0, EXT_DT_), dtype=EXT_DT_)
    caps = [256, 1024]
    for _ in range(3):
        routes = np.zeros(caps[0], ROUTE_DT_)
        nhs = np.zeros(caps[1], NH_DT_)
        r = RibStruct()
        r.routes_cap, r.routes = caps[0], routes.ctypes.data
        r.nexthops_cap, r.nexthops = caps[1], nhs.ctypes.data
        rc = fn(router_id, max_paths, arr, len(areas), ext.ctypes.data if len(ext) else None, len(ext), C.byref(r))
        if rc == capi.HSPF_E_NOMEM:
            caps = [max(caps[0], r.n_routes), max(caps[1], r.n_nexthops)]
            continue
        break
    return Rib(routes[: r.n_routes].copy(), nhs[: r.n_nexthops].copy(), rc)


def update_rib_full(router_id: int, max_paths: int, areas: list, externals=None) -> Rib:
    """The product's host stage (libholo_spf.so); needs no device."""
    lib = capi.load_library()
    rib = call_update_rib_full(lib.hspf_ospfv2_update_rib_full, router_id, max_paths, areas, externals)
    if rib.rc != capi.HSPF_OK:
        raise capi.HspfError(rib.rc, "hspf_ospfv2_update_rib_full failed")
    return rib


def update_rib_full_v3(router_id: int, max_paths: int, areas: list, externals=None) -> Rib:
    """OSPFv3 twin (hspf_ospfv3_update_rib_full); host only."""
    lib = capi.load_library()
    rib = call_update_rib_full(lib.hspf_ospfv3_update_rib_full, router_id, max_paths, areas, externals, v3=True)
    if rib.rc != capi.HSPF_OK:
        raise capi.HspfError(rib.rc, "hspf_ospfv3_update_rib_full failed")
    return rib


def _rib_struct(rib: Rib, keep: list, route_dt, nh_dt) -> RibStruct:
    r = RibStruct()
    routes = np.ascontiguousarray(rib.routes, dtype=route_dt)
    nhs = np.ascontiguousarray(rib.nexthops, dtype=nh_dt)
    keep += [routes, nhs]
    r.routes_cap = r.n_routes = len(routes)
    r.nexthops_cap = r.n_nexthops = len(nhs)
    r.routes = routes.ctypes.data if len(routes) else None
    r.nexthops = nhs.ctypes.data if len(nhs) else None
    return r


def call_rib_diff(fn, old, new: Rib, v3: bool = False):
    """update_global_rib: returns (actions ACTION_DT[], new routes with HL_ROUTE_INSTALLED set as the
    reference would).  `fn` = hspf_ospfv{2,3}_rib_diff or the oracle's twin; `old` may be None."""
    ROUTE_DT_, NH_DT_ = _version(v3)[5:7]
    fn.argtypes = [C.c_void_p, C.POINTER(RibStruct), C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
    keep = []
    ns = _rib_struct(new, keep, ROUTE_DT_, NH_DT_)
    new_routes = keep[0]
    if new_routes is new.routes:                  # never write into the caller's array
        new_routes = new_routes.copy()
        keep[0] = new_routes
        ns.routes = new_routes.ctypes.data if len(new_routes) else None
    os_ = _rib_struct(old, keep, ROUTE_DT_, NH_DT_) if old is not None else None
    cap = len(new.routes) + (len(old.routes) if old is not None else 0) + 1
    acts = np.zeros(cap, ACTION_DT)
    n = C.c_uint32()
    rc = fn(C.addressof(os_) if os_ is not None else None, C.byref(ns), acts.ctypes.data, cap, C.byref(n))
    if rc != capi.HSPF_OK:
        raise capi.HspfError(rc, "rib_diff failed")
    return acts[: n.value].copy(), new_routes


def rib_diff(old, new: Rib, v3: bool = False):
    lib = capi.load_library()
    return call_rib_diff(lib.hspf_ospfv3_rib_diff if v3 else lib.hspf_ospfv2_rib_diff, old, new, v3)


# ---- partial runs (update_rib_partial, holo-ospf/src/route.rs:196-340), OSPFv2 ---------------------------
RIB_RTR_DT = np.dtype([("area_id", "<u4"), ("router_id", "<u4"), ("metric", "<u4"), ("path_type", "u1"), ("flags", "u1"),
                       ("_pad", "u1", (2,)), ("nh_off", "<u4"), ("n_nh", "<u4")], align=True)


class RtrTablesStruct(C.Structure):
    _fields_ = [("rtrs_cap", C.c_uint32), ("n_rtrs", C.c_uint32), ("rtrs", C.c_void_p),
                ("nexthops_cap", C.c_uint32), ("n_nexthops", C.c_uint32), ("nexthops", C.c_void_p)]


@dataclass
class RtrTables:
    rtrs: np.ndarray
    nexthops: np.ndarray


def _area_array(areas: list, keep: list, with_results: bool):
    arr = (RibAreaStruct * max(len(areas), 1))()
    for i, a in enumerate(areas):
        ifs = np.ascontiguousarray(a.ifaces, dtype=ospfv2.IFACE_DT)
        sm = np.ascontiguousarray(a.summaries, dtype=SUMMARY_LSA_DT)
        keep += [ifs, sm]
        arr[i].area_id, arr[i].n_summaries = a.area_id, len(sm)
        if with_results:
            rs = _result_struct(a.result, keep)
            keep.append(rs)
            arr[i].spf = C.addressof(rs)
        arr[i].ifaces = ifs.ctypes.data if len(ifs) else None
        arr[i].summaries = sm.ctypes.data if len(sm) else None
        arr[i].n_ifaces, arr[i].active = len(ifs), int(a.active)
    return arr


def _tables_out(cap_r, cap_h, keep):
    rt, nh = np.zeros(max(cap_r, 1), RIB_RTR_DT), np.zeros(max(cap_h, 1), ospfv2.NEXTHOP_DT)
    keep += [rt, nh]
    s = RtrTablesStruct(len(rt), 0, rt.ctypes.data, len(nh), 0, nh.ctypes.data)
    return s, rt, nh


def router_tables(router_id: int, areas: list, fn=None) -> RtrTables:
    """hspf_ospfv2_rib_router_tables: area.state.routers of every area after a full run."""
    if fn is None:
        fn = capi.load_library().hspf_ospfv2_rib_router_tables
    fn.argtypes = [C.c_uint32, C.POINTER(RibAreaStruct), C.c_uint32, C.POINTER(RtrTablesStruct)]
    keep = []
    arr = _area_array(areas, keep, True)
    caps = [64, 256]
    for _ in range(3):
        s, rt, nh = _tables_out(caps[0], caps[1], keep)
        rc = fn(router_id, arr, len(areas), C.byref(s))
        if rc == capi.HSPF_E_NOMEM:
            caps = [max(caps[0], s.n_rtrs), max(caps[1], s.n_nexthops)]
            continue
        break
    if rc != capi.HSPF_OK:
        raise capi.HspfError(rc, "rib_router_tables failed")
    return RtrTables(rt[: s.n_rtrs].copy(), nh[: s.n_nexthops].copy())


def update_rib_partial(router_id: int, max_paths: int, areas: list, externals, triggers_or_sets, prev_rib: Rib,
                       prev_rtrs: RtrTables, fn=None):
    """hspf_ospfv2_update_rib_partial -> (new Rib, new RtrTables, actions).  `triggers_or_sets`: the
    (inter_network [(addr, mask)], inter_router [id], external [(addr, mask)]) sets of a PARTIAL computation."""
    if fn is None:
        fn = capi.load_library().hspf_ospfv2_update_rib_partial
    fn.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(RibAreaStruct), C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                   C.POINTER(ospfv2.SpfComputationStruct), C.POINTER(RibStruct), C.POINTER(RtrTablesStruct),
                   C.POINTER(RibStruct), C.POINTER(RtrTablesStruct), C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
    keep = []
    arr = _area_array(areas, keep, False)
    transit = np.asarray([int(a.result.transit_capability) for a in areas] or [0], np.uint8)
    ext = np.ascontiguousarray(externals if externals is not None else np.zeros(0, EXTERNAL_LSA_DT), dtype=EXTERNAL_LSA_DT)
    net, rtr, xt = triggers_or_sets
    netA = np.asarray(net or [(0, 0)], ospfv2.IPV4_NET_DT)
    rtrA = np.asarray(rtr or [0], np.uint32)
    xtA = np.asarray(xt or [(0, 0)], ospfv2.IPV4_NET_DT)
    pc = ospfv2.SpfComputationStruct(ospfv2.SPF_PARTIAL, len(net), len(rtr), len(xt), max(len(net), len(rtr), len(xt), 1),
                                     netA.ctypes.data, rtrA.ctypes.data, xtA.ctypes.data)
    prs = _rib_struct(prev_rib, keep, RIB_ROUTE_DT, ospfv2.NEXTHOP_DT)
    prt = np.ascontiguousarray(prev_rtrs.rtrs, RIB_RTR_DT)
    pnh = np.ascontiguousarray(prev_rtrs.nexthops, ospfv2.NEXTHOP_DT)
    pts = RtrTablesStruct(len(prt), len(prt), prt.ctypes.data if len(prt) else None, len(pnh), len(pnh),
                          pnh.ctypes.data if len(pnh) else None)
    caps = [len(prev_rib.routes) + 64, len(prev_rib.nexthops) + 512, len(prt) + 64, len(pnh) + 512, len(prev_rib.routes) + 128]
    for _ in range(3):
        routes, nhs = np.zeros(caps[0], RIB_ROUTE_DT), np.zeros(caps[1], ospfv2.NEXTHOP_DT)
        out = RibStruct(caps[0], 0, routes.ctypes.data, caps[1], 0, nhs.ctypes.data)
        ts, rt, tnh = _tables_out(caps[2], caps[3], keep)
        acts = np.zeros(caps[4], ACTION_DT)
        n = C.c_uint32()
        rc = fn(router_id, max_paths, arr, transit.ctypes.data, len(areas), ext.ctypes.data if len(ext) else None, len(ext),
                C.byref(pc), C.byref(prs), C.byref(pts), C.byref(out), C.byref(ts), acts.ctypes.data, caps[4], C.byref(n))
        if rc == capi.HSPF_E_NOMEM:
            caps = [max(caps[0], out.n_routes), max(caps[1], out.n_nexthops), max(caps[2], ts.n_rtrs), max(caps[3], ts.n_nexthops),
                    max(caps[4], n.value)]
            continue
        break
    if rc != capi.HSPF_OK:
        raise capi.HspfError(rc, "update_rib_partial failed")
    return (Rib(routes[: out.n_routes].copy(), nhs[: out.n_nexthops].copy()), RtrTables(rt[: ts.n_rtrs].copy(), tnh[: ts.n_nexthops].copy()),
            acts[: n.value].copy())
