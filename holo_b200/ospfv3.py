"""OSPFv3 side of the engine from Python: LSDB images (include/holo_lsdb.h), the
run_area call (holo-ospf/src/spf.rs:587-729 with the OSPFv3 hooks of
holo-ospf/src/ospfv3/spf.rs replaced by hspf_ospfv3_run_area) and a synthetic builder."""
from __future__ import annotations

import ctypes as C
import ipaddress
from dataclasses import dataclass, field

import numpy as np

from . import capi
from .ospfv2 import ROUTE_RTR_DT, IF_P2P, IF_BROADCAST, LINK_P2P, LINK_TRANSIT, MAX_AGE
from .synth import Topology

OPT_R, OPT_V6 = 0x01, 0x02
PFX_NU = 0x01
REF_ROUTER, REF_NETWORK = 1, 2

IP_DT = np.dtype([("bytes", "u1", (16,)), ("is_v6", "u1"), ("_pad", "u1", (3,))], align=True)
LINK_DT = np.dtype([("iface_id", "<u4"), ("nbr_iface_id", "<u4"), ("nbr_router_id", "<u4"), ("metric", "<u2"),
                    ("link_type", "u1"), ("_pad", "u1")], align=True)
ROUTER_LSA_DT = np.dtype([("adv_rtr", "<u4"), ("lsa_id", "<u4"), ("age", "<u2"), ("flags", "u1"), ("options", "u1"),
                          ("link_off", "<u4"), ("n_links", "<u4")], align=True)
NETWORK_LSA_DT = np.dtype([("adv_rtr", "<u4"), ("lsa_id", "<u4"), ("age", "<u2"), ("_pad", "<u2"), ("att_off", "<u4"),
                           ("n_att", "<u4")], align=True)
PREFIX_DT = np.dtype([("addr", IP_DT), ("len", "u1"), ("options", "u1"), ("metric", "<u2")], align=True)
IAP_LSA_DT = np.dtype([("adv_rtr", "<u4"), ("lsa_id", "<u4"), ("age", "<u2"), ("ref_type", "u1"), ("_pad", "u1"),
                       ("ref_lsa_id", "<u4"), ("ref_adv_rtr", "<u4"), ("prefix_off", "<u4"), ("n_prefixes", "<u4")],
                      align=True)
LINK_LSA_DT = np.dtype([("iface", "<u4"), ("adv_rtr", "<u4"), ("lsa_id", "<u4"), ("age", "<u2"), ("_pad", "<u2"),
                        ("linklocal", IP_DT)], align=True)
IFACE_DT = np.dtype([("ifindex", "<u4"), ("sort_key", "<u4"), ("if_type", "u1"), ("_pad", "u1", (3,))], align=True)
NEXTHOP6_DT = np.dtype([("iface", "<u4"), ("nbr_router_id", "<u4"), ("addr", IP_DT), ("has_addr", "u1"),
                        ("has_nbr", "u1"), ("_pad", "u1", (2,))], align=True)
SPT_VERTEX6_DT = np.dtype([("router_id", "<u4"), ("iface_id", "<u4"), ("distance", "<u4"), ("hops", "<u2"),
                           ("is_router", "u1"), ("_pad", "u1"), ("nh_off", "<u4"), ("n_nh", "<u4")], align=True)
ROUTE_NET6_DT = np.dtype([("prefix", IP_DT), ("len", "u1"), ("flags", "u1"), ("origin_type", "u1"),
                          ("prefix_options", "u1"), ("metric", "<u4"), ("origin_adv_rtr", "<u4"),
                          ("origin_lsa_id", "<u4"), ("nh_off", "<u4"), ("n_nh", "<u4")], align=True)


class AreaStruct(C.Structure):
    _fields_ = [
        ("router_id", C.c_uint32), ("area_id", C.c_uint32), ("max_paths", C.c_uint16), ("af_ipv6", C.c_uint8),
        ("_pad", C.c_uint8),
        ("n_router_lsas", C.c_uint32), ("router_lsas", C.c_void_p),
        ("n_links", C.c_uint32), ("links", C.c_void_p),
        ("n_network_lsas", C.c_uint32), ("network_lsas", C.c_void_p),
        ("n_attached", C.c_uint32), ("attached", C.c_void_p),
        ("n_iap_lsas", C.c_uint32), ("iap_lsas", C.c_void_p),
        ("n_prefixes", C.c_uint32), ("prefixes", C.c_void_p),
        ("n_ifaces", C.c_uint32), ("ifaces", C.c_void_p),
        ("n_link_lsas", C.c_uint32), ("link_lsas", C.c_void_p),
    ]


class ResultStruct(C.Structure):
    _fields_ = [
        ("vertices_cap", C.c_uint32), ("n_vertices", C.c_uint32), ("vertices", C.c_void_p),
        ("routers_cap", C.c_uint32), ("n_routers", C.c_uint32), ("routers", C.c_void_p),
        ("routes_cap", C.c_uint32), ("n_routes", C.c_uint32), ("routes", C.c_void_p),
        ("nexthops_cap", C.c_uint32), ("n_nexthops", C.c_uint32), ("nexthops", C.c_void_p),
        ("transit_capability", C.c_uint8), ("root_found", C.c_uint8), ("_pad", C.c_uint8 * 2),
    ]


ABI_SIZES = [LINK_DT.itemsize, ROUTER_LSA_DT.itemsize, NETWORK_LSA_DT.itemsize, IP_DT.itemsize, PREFIX_DT.itemsize,
             IAP_LSA_DT.itemsize, LINK_LSA_DT.itemsize, IFACE_DT.itemsize, C.sizeof(AreaStruct), NEXTHOP6_DT.itemsize,
             SPT_VERTEX6_DT.itemsize, ROUTE_NET6_DT.itemsize, C.sizeof(ResultStruct)]

_FIELDS = (("router_lsas", ROUTER_LSA_DT), ("links", LINK_DT), ("network_lsas", NETWORK_LSA_DT),
           ("attached", np.dtype("<u4")), ("iap_lsas", IAP_LSA_DT), ("prefixes", PREFIX_DT), ("ifaces", IFACE_DT),
           ("link_lsas", LINK_LSA_DT))


def ip_rec(addr) -> tuple:
    a = ipaddress.ip_address(addr)
    b = a.packed if a.version == 6 else a.packed + bytes(12)
    return (tuple(b), 1 if a.version == 6 else 0, (0, 0, 0))


def ip_str(rec) -> str:
    b = bytes(int(x) for x in rec["bytes"])
    return str(ipaddress.IPv6Address(b)) if int(rec["is_v6"]) else str(ipaddress.IPv4Address(b[:4]))


@dataclass
class Ospfv3Area:
    router_id: int
    area_id: int = 0
    max_paths: int = 16
    af_ipv6: bool = True
    router_lsas: np.ndarray = field(default_factory=lambda: np.zeros(0, ROUTER_LSA_DT))
    links: np.ndarray = field(default_factory=lambda: np.zeros(0, LINK_DT))
    network_lsas: np.ndarray = field(default_factory=lambda: np.zeros(0, NETWORK_LSA_DT))
    attached: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint32))
    iap_lsas: np.ndarray = field(default_factory=lambda: np.zeros(0, IAP_LSA_DT))
    prefixes: np.ndarray = field(default_factory=lambda: np.zeros(0, PREFIX_DT))
    ifaces: np.ndarray = field(default_factory=lambda: np.zeros(0, IFACE_DT))
    link_lsas: np.ndarray = field(default_factory=lambda: np.zeros(0, LINK_LSA_DT))
    ifnames: list = field(default_factory=list)

    def as_struct(self) -> AreaStruct:
        s = AreaStruct()
        s.router_id, s.area_id, s.max_paths, s.af_ipv6 = self.router_id, self.area_id, self.max_paths, int(self.af_ipv6)
        for name, dt in _FIELDS:
            a = np.ascontiguousarray(getattr(self, name), dtype=dt)
            setattr(self, name, a)
            setattr(s, "n_" + name, len(a))
            setattr(s, name, a.ctypes.data if len(a) else None)
        return s


@dataclass
class Ospfv3Result:
    vertices: np.ndarray
    routers: np.ndarray
    routes: np.ndarray
    nexthops: np.ndarray
    transit_capability: bool
    root_found: bool
    rc: int = 0

    def nh(self, rec):
        return [(int(x["iface"]), ip_str(x["addr"]) if x["has_addr"] else None, int(x["nbr_router_id"]) if x["has_nbr"] else None)
                for x in self.nexthops[int(rec["nh_off"]): int(rec["nh_off"]) + int(rec["n_nh"])]]


def _call_run_area(fn, area: Ospfv3Area, prefix_args=(), tail_args=()):
    s = area.as_struct()
    nv = len(area.router_lsas) + len(area.network_lsas) + 1
    n_routes = len(area.prefixes) + 1
    caps = [nv, nv, n_routes, 64 * (2 * nv + n_routes) + 64]
    for _ in range(2):
        verts = np.zeros(caps[0], SPT_VERTEX6_DT)
        rtrs = np.zeros(caps[1], ROUTE_RTR_DT)
        routes = np.zeros(caps[2], ROUTE_NET6_DT)
        nhs = np.zeros(caps[3], NEXTHOP6_DT)
        r = ResultStruct()
        r.vertices_cap, r.vertices = caps[0], verts.ctypes.data
        r.routers_cap, r.routers = caps[1], rtrs.ctypes.data
        r.routes_cap, r.routes = caps[2], routes.ctypes.data
        r.nexthops_cap, r.nexthops = caps[3], nhs.ctypes.data
        rc = fn(*prefix_args, C.byref(s), *tail_args, C.byref(r))
        if rc == capi.HSPF_E_NOMEM:
            caps = [max(caps[0], r.n_vertices), max(caps[1], r.n_routers), max(caps[2], r.n_routes),
                    max(caps[3], r.n_nexthops)]
            continue
        break
    return Ospfv3Result(verts[: r.n_vertices].copy(), rtrs[: r.n_routers].copy(), routes[: r.n_routes].copy(),
                        nhs[: r.n_nexthops].copy(), bool(r.transit_capability), bool(r.root_found), rc)


def run_area(ctx: capi.Context, area: Ospfv3Area) -> Ospfv3Result:
    lib = ctx.lib
    lib.hspf_ospfv3_run_area.argtypes = [C.c_void_p, C.POINTER(AreaStruct), C.POINTER(ResultStruct)]
    res = _call_run_area(lib.hspf_ospfv3_run_area, area, (ctx.handle,))
    if res.rc != capi.HSPF_OK:
        raise capi.HspfError(res.rc, ctx.last_error())
    return res


def area_from_planes(area: Ospfv3Area, spf) -> Ospfv3Result:
    """hspf_ospfv3_area_from_planes over planes from `spf(csr, root_vertex, nh_words)` (host only;
    see holo_b200.ospfv2.area_from_planes)."""
    lib = capi.load_library()
    lib.hspf_ospfv3_area_from_planes.argtypes = [C.POINTER(AreaStruct), C.POINTER(C.c_uint32), C.POINTER(C.c_uint16),
                                                 C.POINTER(C.c_uint64), C.c_uint32, C.POINTER(ResultStruct)]
    f = Flat(area)
    root = f.router_vertex(area.router_id)
    nhw = 4
    if root == 0xFFFFFFFF:
        V = f.csr.n_vertices
        d, h, m = np.zeros(V, np.uint32), np.zeros(V, np.uint16), np.zeros((V, nhw), np.uint64)
    else:
        d, h, m = spf(f.csr, root, nhw)
    d = np.ascontiguousarray(d, np.uint32)
    h = np.ascontiguousarray(h, np.uint16)
    m = np.ascontiguousarray(m, np.uint64)
    tail = (d.ctypes.data_as(C.POINTER(C.c_uint32)), h.ctypes.data_as(C.POINTER(C.c_uint16)),
            m.ctypes.data_as(C.POINTER(C.c_uint64)), nhw)
    res = _call_run_area(lib.hspf_ospfv3_area_from_planes, area, (), tail)
    if res.rc != capi.HSPF_OK:
        raise capi.HspfError(res.rc, "hspf_ospfv3_area_from_planes failed")
    return res


class RouteTable:
    """hspf_ospfv3_rtable_create: the route table of an OSPFv3 area for the batched route stage (the object
    ospfv2.RouteTable wraps; upload and hspf_ospfv2_routes_batch[16] work on it unchanged)."""

    def __init__(self, flat: "Flat"):
        lib = capi.load_library()
        lib.hspf_ospfv3_rtable_create.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        lib.hspf_ospfv2_rtable_free.argtypes = [C.c_void_p]
        lib.hspf_ospfv2_rtable_free.restype = None
        lib.hspf_ospfv2_rtable_prefixes.argtypes = [C.c_void_p]
        lib.hspf_ospfv2_rtable_prefixes.restype = C.c_uint32
        lib.hspf_ospfv2_rtable_contributors.argtypes = [C.c_void_p]
        lib.hspf_ospfv2_rtable_contributors.restype = C.c_uint32
        lib.hspf_ospfv2_rtable_upload.argtypes = [C.c_void_p, C.c_void_p]
        lib.hspf_ospfv3_rtable_prefixes6.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.POINTER(C.c_uint32))]
        self.lib, self.flat = lib, flat
        h = C.c_void_p()
        rc = lib.hspf_ospfv3_rtable_create(flat.handle, C.byref(h))
        if rc != capi.HSPF_OK:
            raise capi.HspfError(rc, "hspf_ospfv3_rtable_create failed")
        self.handle = h
        self.n_prefixes = int(lib.hspf_ospfv2_rtable_prefixes(h))
        self.n_contributors = int(lib.hspf_ospfv2_rtable_contributors(h))
        pp, pl = C.c_void_p(), C.POINTER(C.c_uint32)()
        lib.hspf_ospfv3_rtable_prefixes6(h, C.byref(pp), C.byref(pl))
        P = self.n_prefixes
        self.prefix = np.frombuffer(C.string_at(pp.value, P * IP_DT.itemsize), IP_DT).copy() if P else np.zeros(0, IP_DT)
        self.plen = np.ctypeslib.as_array(pl, shape=(P,)).copy() if P else np.zeros(0, np.uint32)
        lib.hspf_ospfv2_rtable_arrays.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.POINTER(C.c_uint32)), C.c_void_p]
        po = C.POINTER(C.c_uint32)()
        lib.hspf_ospfv2_rtable_arrays(h, None, None, C.byref(po), None)
        self.off = np.ctypeslib.as_array(po, shape=(P + 1,)).copy()

    def upload(self, ctx: capi.Context):
        rc = self.lib.hspf_ospfv2_rtable_upload(ctx.handle, self.handle)
        if rc != capi.HSPF_OK:
            raise capi.HspfError(rc, ctx.last_error())

    def __del__(self):
        try:
            if self.handle:
                self.lib.hspf_ospfv2_rtable_free(self.handle)
                self.handle = None
        except Exception:
            pass


def routes_from_cells(area: Ospfv3Area, rt: RouteTable, cells: np.ndarray, gather_v, gather_nh) -> Ospfv3Result:
    """hspf_ospfv3_routes_from_cells (host): one job's cells -> the routes / next hops of hspf_ospfv3_run_area."""
    from . import ospfv2
    lib = capi.load_library()
    lib.hspf_ospfv3_routes_from_cells.argtypes = [C.POINTER(AreaStruct), C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32),
                                                  C.POINTER(C.c_uint64), C.c_uint32, C.POINTER(ResultStruct)]
    cells = np.ascontiguousarray(cells, ospfv2.CELL_DT)
    assert cells.shape == (rt.n_prefixes,)
    gv = np.ascontiguousarray(gather_v, np.uint32)
    gn = np.ascontiguousarray(gather_nh, np.uint64)
    res = _call_run_area(lib.hspf_ospfv3_routes_from_cells, area, (),
                         (rt.handle, cells.ctypes.data, gv.ctypes.data_as(C.POINTER(C.c_uint32)),
                          gn.ctypes.data_as(C.POINTER(C.c_uint64)), len(gv)))
    if res.rc not in (capi.HSPF_OK, capi.HSPF_E_UNSUPPORTED):
        raise capi.HspfError(res.rc, "hspf_ospfv3_routes_from_cells failed")
    return res


IP_PREFIX_DT = np.dtype([("addr", IP_DT), ("len", "u1"), ("_pad", "u1", (3,))], align=True)
TRIGGER6_DT = np.dtype([("adv_rtr", "<u4"), ("lsa_id", "<u4"), ("router_id", "<u4"), ("prefix_off", "<u4"), ("n_prefixes", "<u4"),
                        ("function_code", "<u2"), ("_pad", "u1", (2,))], align=True)


class SpfComputation6Struct(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("n_intra", C.c_uint32), ("n_inter_network", C.c_uint32), ("n_inter_router", C.c_uint32),
                ("n_external", C.c_uint32), ("cap", C.c_uint32), ("intra", C.c_void_p), ("inter_network", C.c_void_p),
                ("inter_router", C.c_void_p), ("external", C.c_void_p)]


def spf_computation_type(triggers, fn=None):
    """hspf_ospfv3_spf_computation_type.  triggers: [(function_code, adv_rtr, lsa_id, router_id, [(addr, len), ...])]
    -> (kind, intra, inter_network, inter_router, external) with prefixes as (addr string, len)."""
    if fn is None:
        fn = capi.load_library().hspf_ospfv3_spf_computation_type
    fn.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(SpfComputation6Struct)]
    tr = np.zeros(len(triggers), TRIGGER6_DT)
    pf = []
    for i, (code, adv, lsa_id, router_id, prefixes) in enumerate(triggers):
        tr[i] = (adv, lsa_id, router_id, len(pf), len(prefixes), code, (0, 0))
        pf += [(ip_rec(a), ln, (0, 0, 0)) for a, ln in prefixes]
    pfa = np.asarray(pf, IP_PREFIX_DT) if pf else np.zeros(0, IP_PREFIX_DT)
    cap = max(len(pfa), len(tr), 1)
    a, b, c, d = np.zeros(cap, IP_PREFIX_DT), np.zeros(cap, IP_PREFIX_DT), np.zeros(cap, np.uint32), np.zeros(cap, IP_PREFIX_DT)
    s = SpfComputation6Struct(0, 0, 0, 0, 0, cap, a.ctypes.data, b.ctypes.data, c.ctypes.data, d.ctypes.data)
    rc = fn(tr.ctypes.data if len(tr) else None, len(tr), pfa.ctypes.data if len(pfa) else None, len(pfa), C.byref(s))
    if rc != capi.HSPF_OK:
        raise capi.HspfError(rc, "ospfv3 spf_computation_type failed")
    out = lambda arr, k: [(ip_str(x["addr"]), int(x["len"])) for x in arr[:k]]
    return s.kind, out(a, s.n_intra), out(b, s.n_inter_network), [int(x) for x in c[: s.n_inter_router]], out(d, s.n_external)


class Flat:
    def __init__(self, area: Ospfv3Area):
        lib = capi.load_library()
        lib.hspf_ospfv3_flatten.argtypes = [C.POINTER(AreaStruct), C.POINTER(C.c_void_p)]
        lib.hspf_ospfv3_flat_free.argtypes = [C.c_void_p]
        lib.hspf_ospfv3_flat_free.restype = None
        lib.hspf_ospfv3_flat_csr.argtypes = [C.c_void_p, C.POINTER(capi.CsrStruct)]
        lib.hspf_ospfv3_flat_vertices.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_uint32)),
                                                  C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.POINTER(C.c_uint8)),
                                                  C.POINTER(C.c_uint32)]
        lib.hspf_ospfv3_flat_router_vertex.argtypes = [C.c_void_p, C.c_uint32]
        lib.hspf_ospfv3_flat_router_vertex.restype = C.c_uint32
        self.lib, self.area = lib, area
        self._s = area.as_struct()
        h = C.c_void_p()
        rc = lib.hspf_ospfv3_flatten(C.byref(self._s), C.byref(h))
        if rc != capi.HSPF_OK:
            raise capi.HspfError(rc, "hspf_ospfv3_flatten failed")
        self.handle = h
        self._load()

    def _load(self):
        lib, h = self.lib, self.handle
        cs = capi.CsrStruct()
        lib.hspf_ospfv3_flat_csr(h, C.byref(cs))
        V, E = cs.n_vertices, cs.n_edges
        as_np = lambda p, n, dt: np.ctypeslib.as_array(p, shape=(n,)).astype(dt).copy() if n else np.zeros(0, dt)
        self.csr = capi.Csr(as_np(cs.row_ptr, V + 1, np.uint32), as_np(cs.col, E, np.uint32),
                            as_np(cs.cost, E, np.uint32), as_np(cs.vflags, V, np.uint8),
                            reject_above=cs.reject_above, saturate_at=cs.saturate_at, flags=cs.flags, delta=cs.delta)
        rid, ifid, isr, n = C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint8)(), C.c_uint32()
        lib.hspf_ospfv3_flat_vertices(h, C.byref(rid), C.byref(ifid), C.byref(isr), C.byref(n))
        self.router_ids = as_np(rid, n.value, np.uint32)
        self.iface_ids = as_np(ifid, n.value, np.uint32)
        self.is_router = as_np(isr, n.value, np.uint8)

    def router_vertex(self, router_id: int) -> int:
        return int(self.lib.hspf_ospfv3_flat_router_vertex(self.handle, router_id))

    def update(self, new_area: "Ospfv3Area"):
        """hspf_ospfv3_flat_update -> (kind, edges, costs) with kind 0 unchanged / 1 costs / 2 rebuilt."""
        self.lib.hspf_ospfv3_flat_update.argtypes = [C.c_void_p, C.POINTER(AreaStruct), C.POINTER(C.c_uint32), C.c_void_p,
                                                     C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        cap = max(int(self.csr.n_edges), 1)
        edges, costs = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32)
        kind, n = C.c_uint32(), C.c_uint32()
        s = new_area.as_struct()
        rc = self.lib.hspf_ospfv3_flat_update(self.handle, C.byref(s), C.byref(kind), edges.ctypes.data, costs.ctypes.data, cap,
                                              C.byref(n))
        if rc != capi.HSPF_OK:
            raise capi.HspfError(rc, "hspf_ospfv3_flat_update failed")
        self.area, self._s = new_area, s
        self._load()
        return kind.value, edges[: n.value].copy(), costs[: n.value].copy()

    def __del__(self):
        try:
            if self.handle:
                self.lib.hspf_ospfv3_flat_free(self.handle)
                self.handle = None
        except Exception:
            pass


# ------------------------------------------------------------------------------ synthetic
RID_BASE = 0x0A000001


def synth_area(t: Topology, root: int = 0, max_links_per_fragment: int = 0, max_paths: int = 16,
               rids=None, area_id: int = 0) -> Ospfv3Area:
    """OSPFv3 area LSDB for topology `t` seen by router `root`.  Interface ids are
    per-router link ordinals (1-based); Router-LSAs are split into fragments of
    `max_links_per_fragment` links when > 0 (RFC 5340 4.8.1 aggregate); one
    Intra-Area-Prefix-LSA per router (a /128 loopback, metric 0, plus one /64 per
    p2p link) and one per LAN (referencing the Network-LSA)."""
    R = t.n_routers
    rid = (lambda i: RID_BASE + int(i)) if rids is None else (lambda i: int(rids[int(i)]))
    per = [[] for _ in range(R)]          # (iface_id, nbr_iface_id, nbr_rid, metric, type)
    nif = [0] * R
    p2p_if = []
    for k in range(t.n_p2p):
        a, b = int(t.p2p_a[k]), int(t.p2p_b[k])
        nif[a] += 1; ia = nif[a]
        nif[b] += 1; ib = nif[b]
        p2p_if.append((ia, ib))
        per[a].append((ia, ib, rid(b), int(t.p2p_cost_ab[k]), LINK_P2P))
        per[b].append((ib, ia, rid(a), int(t.p2p_cost_ba[k]), LINK_P2P))
    lan_if = []
    net_lsas, attached = [], []
    for members, costs in t.lans:
        ids = []
        for m in members:
            nif[m] += 1
            ids.append(nif[m])
        dr, dr_if = members[0], ids[0]
        for m, c, i in zip(members, costs, ids):
            per[m].append((i, dr_if, rid(dr), int(c), LINK_TRANSIT))
        lan_if.append(ids)
        net_lsas.append((rid(dr), dr_if, len(attached), len(members)))
        attached += sorted(rid(m) for m in members)
    rl, links = [], []
    for i in range(R):
        chunks = [per[i]]
        if max_links_per_fragment > 0:
            chunks = [per[i][j:j + max_links_per_fragment] for j in range(0, len(per[i]), max_links_per_fragment)] or [[]]
        for frag, ch in enumerate(chunks):
            rl.append((rid(i), frag, 1, 0, OPT_R | OPT_V6, len(links), len(ch)))
            links += [(a, b, c, d, e, 0) for (a, b, c, d, e) in ch]
    area = Ospfv3Area(router_id=rid(root), max_paths=max_paths, area_id=area_id)
    rl.sort(key=lambda x: (x[0], x[1]))          # LsaKey order (adv_rtr, lsa_id)
    area.router_lsas = np.asarray(rl, dtype=ROUTER_LSA_DT)
    area.links = np.asarray(links, dtype=LINK_DT) if links else np.zeros(0, LINK_DT)
    nl = np.zeros(len(net_lsas), NETWORK_LSA_DT)
    for i, (adv, lsid, ao, na) in enumerate(net_lsas):
        nl[i] = (adv, lsid, 1, 0, ao, na)
    area.network_lsas = nl[np.lexsort((nl["lsa_id"], nl["adv_rtr"]))] if len(nl) else nl
    area.attached = np.asarray(attached, dtype=np.uint32)
    # prefixes
    iaps, prefixes = [], []
    for i in range(R):
        off = len(prefixes)
        prefixes.append((ip_rec(ipaddress.IPv6Address((0x20010DB8 << 96) | (0x1000 << 80) | (i + 1))), 128, 0, 0))
        iaps.append((rid(i), 0, 1, REF_ROUTER, 0, 0, rid(i), off, len(prefixes) - off))
    for k in range(t.n_p2p):
        a = int(t.p2p_a[k])
        off = len(prefixes)
        prefixes.append((ip_rec(ipaddress.IPv6Address((0x20010DB8 << 96) | (0x2000 << 80) | (k << 64))), 64, 0,
                         int(t.p2p_cost_ab[k])))
        iaps.append((rid(a), 1 + k, 1, REF_ROUTER, 0, 0, rid(a), off, 1))
    for li, (members, costs) in enumerate(t.lans):
        off = len(prefixes)
        prefixes.append((ip_rec(ipaddress.IPv6Address((0x20010DB8 << 96) | (0x3000 << 80) | (li << 64))), 64, 0, 0))
        iaps.append((rid(members[0]), 0x10000 + li, 1, REF_NETWORK, 0, lan_if[li][0], rid(members[0]), off, 1))
    ia = np.zeros(len(iaps), IAP_LSA_DT)
    for i, x in enumerate(iaps):
        ia[i] = x
    area.iap_lsas = ia[np.lexsort((ia["lsa_id"], ia["adv_rtr"]))]
    pa = np.zeros(len(prefixes), PREFIX_DT)
    for i, x in enumerate(prefixes):
        pa[i] = x
    area.prefixes = pa
    # local interfaces of the root + the neighbours' Link-LSAs
    ifaces, llsas, names = [], [], []
    ll = lambda r, i: ip_rec(ipaddress.IPv6Address((0xFE80 << 112) | (rid(r) << 32) | i))
    for k in range(t.n_p2p):
        a, b = int(t.p2p_a[k]), int(t.p2p_b[k])
        ia_, ib_ = p2p_if[k]
        if a == root:
            llsas.append((len(ifaces), rid(b), ib_, 1, 0, ll(b, ib_)))
            ifaces.append((ia_, 1000 - len(ifaces), IF_P2P, (0, 0, 0)))
        if b == root:
            llsas.append((len(ifaces), rid(a), ia_, 1, 0, ll(a, ia_)))
            ifaces.append((ib_, 1000 - len(ifaces), IF_P2P, (0, 0, 0)))
    for li, (members, costs) in enumerate(t.lans):
        if root in members:
            me = lan_if[li][members.index(root)]
            for m, i in zip(members, lan_if[li]):
                if m != root:
                    llsas.append((len(ifaces), rid(m), i, 1, 0, ll(m, i)))
            ifaces.append((me, 1000 - len(ifaces), IF_BROADCAST, (0, 0, 0)))
    area.ifaces = np.asarray(ifaces, dtype=IFACE_DT) if ifaces else np.zeros(0, IFACE_DT)
    la = np.zeros(len(llsas), LINK_LSA_DT)
    for i, x in enumerate(llsas):
        la[i] = x
    area.link_lsas = la
    area.ifnames = [f"if{int(x[0])}" for x in ifaces]
    return area
