#!/usr/bin/env python
"""Host-side cost of the LSDB-level drop-in call at BASELINE C5 size (10 000 routers, LANs, ECMP,
SR): hspf_ospfv2_flatten and hspf_ospfv2_area_from_planes (flatten + everything run_area does
after the SPT) on the CPU, with SPT planes from the heap oracle.  No GPU needed."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from holo_b200 import capi, ospfv2, synth  # noqa: E402
from oracle import pyoracle  # noqa: E402

t = synth.random_topology(10000, 40000, synth.SEED_BASE + 5, cost_choices=[10, 20], lan_fraction=0.05)
lib = capi.load_library()
for sr in (False, True):
    a = ospfv2.synth_area(t, 0, sr=sr)
    f = ospfv2.Flat(a)
    root = f.router_vertex(a.router_id)
    c = pyoracle.csr_spf_heap(f.csr, root, nh_words=4)
    d = np.ascontiguousarray(c["dist"], np.uint32)
    h = np.ascontiguousarray(c["hops"], np.uint16)
    m = np.ascontiguousarray(c["nh_mask"], np.uint64)
    s = a.as_struct()
    nv = len(a.router_lsas) + len(a.network_lsas) + 1
    nr = len(a.links) + len(a.network_lsas) + 1
    verts, rt = np.zeros(nv, ospfv2.SPT_VERTEX_DT), np.zeros(nv, ospfv2.ROUTE_RTR_DT)
    ro, nh = np.zeros(nr, ospfv2.ROUTE_NET_DT), np.zeros(64 * (2 * nv + nr), ospfv2.NEXTHOP_DT)
    r = ospfv2.ResultStruct()
    r.vertices_cap, r.vertices = nv, verts.ctypes.data
    r.routers_cap, r.routers = nv, rt.ctypes.data
    r.routes_cap, r.routes = nr, ro.ctypes.data
    r.nexthops_cap, r.nexthops = len(nh), nh.ctypes.data
    lib.hspf_ospfv2_area_from_planes.argtypes = [C.POINTER(ospfv2.AreaStruct), C.POINTER(C.c_uint32), C.POINTER(C.c_uint16),
                                                 C.POINTER(C.c_uint64), C.c_uint32, C.POINTER(ospfv2.ResultStruct)]
    lib.hspf_ospfv2_flatten.argtypes = [C.POINTER(ospfv2.AreaStruct), C.POINTER(C.c_void_p)]
    lib.hspf_ospfv2_flat_free.argtypes = [C.c_void_p]
    tf, ts = [], []
    for _ in range(5):
        hd = C.c_void_p()
        t0 = time.perf_counter()
        lib.hspf_ospfv2_flatten(C.byref(s), C.byref(hd))
        tf.append(time.perf_counter() - t0)
        lib.hspf_ospfv2_flat_free(hd)
        t0 = time.perf_counter()
        rc = lib.hspf_ospfv2_area_from_planes(C.byref(s), d.ctypes.data_as(C.POINTER(C.c_uint32)),
                                              h.ctypes.data_as(C.POINTER(C.c_uint16)),
                                              m.ctypes.data_as(C.POINTER(C.c_uint64)), 4, C.byref(r))
        ts.append(time.perf_counter() - t0)
        assert rc == 0
    print(f"sr={sr}: flatten {min(tf) * 1e3:.1f} ms, flatten + post-SPT {min(ts) * 1e3:.1f} ms "
          f"({r.n_vertices} vertices, {r.n_routes} routes, {r.n_nexthops} next hops)")
