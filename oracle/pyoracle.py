"""ctypes access to oracle/_build/liboracle.so (TEST INFRASTRUCTURE)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_u32p = C.POINTER(C.c_uint32)
_u16p = C.POINTER(C.c_uint16)
_u64p = C.POINTER(C.c_uint64)

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        from holo_b200.build import ORACLE_LIB, build_oracle
        if not ORACLE_LIB.exists():
            build_oracle()
        _lib = C.CDLL(str(ORACLE_LIB))
    return _lib


def _p(a, ty):
    return C.cast(None, ty) if a is None else a.ctypes.data_as(ty)


def csr_spf(csr, root: int, overrides=(), vec_mode: int = 0, nh_words: int = 1, want_lists: bool = False):
    """Reference-faithful SPF over a flattened CSR.  Returns a dict of planes."""
    L = lib()
    V = csr.n_vertices
    s = csr.as_struct()
    ove = np.asarray([e for e, _ in overrides] or [0], dtype=np.uint32)
    ovc = np.asarray([c for _, c in overrides] or [0], dtype=np.uint32)
    dist = np.empty(V, np.uint32)
    hops = np.empty(V, np.uint16)
    fp = np.empty(V, np.uint32)
    npar = np.empty(V, np.uint16)
    nh = np.empty((V, nh_words), np.uint64)
    order = np.empty(V, np.uint32)
    n_popped = C.c_uint32()
    status = C.c_uint32()
    poff = parents = nvoff = nhvec = None
    pcap = ncap = 0
    if want_lists:
        poff = np.empty(V + 1, np.uint32)
        pcap = max(16, 4 * csr.n_edges)
        parents = np.empty(pcap, np.uint32)
        nvoff = np.empty(V + 1, np.uint32)
        ncap = 1 << 22
        nhvec = np.empty(ncap, np.uint32)
    rc = L.oracle_csr_spf(C.byref(s), C.c_uint32(root), C.c_uint32(len(overrides)), _p(ove, _u32p), _p(ovc, _u32p),
                          C.c_int(vec_mode), _p(dist, _u32p), _p(hops, _u16p), _p(fp, _u32p), _p(npar, _u16p),
                          _p(nh, _u64p), C.c_uint32(nh_words), _p(poff, _u32p), _p(parents, _u32p), C.c_uint32(pcap),
                          _p(nvoff, _u32p), _p(nhvec, _u32p), C.c_uint32(ncap), _p(order, _u32p),
                          C.byref(n_popped), C.byref(status))
    out = dict(dist=dist, hops=hops, first_parent=fp, n_parents=npar, nh_mask=nh, status=status.value,
               pop_order=order[: n_popped.value].copy(), rc=rc)
    if want_lists:
        out["parents_off"] = poff
        out["parents"] = parents[: poff[V]].copy()
        out["nhvec_off"] = nvoff
        out["nhvec"] = nhvec[: nvoff[V]].copy()
    return out


def csr_spf_heap(csr, root: int, overrides=(), nh_words: int = 1):
    """Optimised CPU baseline (binary heap), same planes, static-order rules."""
    L = lib()
    V = csr.n_vertices
    s = csr.as_struct()
    ove = np.asarray([e for e, _ in overrides] or [0], dtype=np.uint32)
    ovc = np.asarray([c for _, c in overrides] or [0], dtype=np.uint32)
    dist = np.empty(V, np.uint32)
    hops = np.empty(V, np.uint16)
    fp = np.empty(V, np.uint32)
    npar = np.empty(V, np.uint16)
    nh = np.empty((V, nh_words), np.uint64)
    status = C.c_uint32()
    L.oracle_csr_spf_heap(C.byref(s), C.c_uint32(root), C.c_uint32(len(overrides)), _p(ove, _u32p), _p(ovc, _u32p),
                          _p(dist, _u32p), _p(hops, _u16p), _p(fp, _u32p), _p(npar, _u16p), _p(nh, _u64p),
                          C.c_uint32(nh_words), C.byref(status))
    return dict(dist=dist, hops=hops, first_parent=fp, n_parents=npar, nh_mask=nh, status=status.value)


def ospfv2_run_area(area):
    """Reference-faithful run_area + update_rib_intra_area over an LSDB image."""
    from holo_b200 import ospfv2
    L = lib()
    L.oracle_ospfv2_run_area.argtypes = [C.POINTER(ospfv2.AreaStruct), C.POINTER(ospfv2.ResultStruct)]
    return ospfv2._call_run_area(L.oracle_ospfv2_run_area, area)


def ospfv2_spf_computation_type(triggers):
    """Restatement of Ospfv2::spf_computation_type (oracle/spf_ospfv2.cc)."""
    from holo_b200 import ospfv2
    return ospfv2.spf_computation_type(triggers, fn=lib().oracle_ospfv2_spf_computation_type)


def ospfv2_update_rib_full(router_id: int, max_paths: int, areas: list, externals=None):
    """Reference-faithful update_rib_full stages after the per-area SPFs (oracle/rib_ospfv2.cc)."""
    from holo_b200 import ospf_rib
    return ospf_rib.call_update_rib_full(lib().oracle_ospfv2_update_rib_full, router_id, max_paths, areas, externals)


def ospfv2_rib_router_tables(router_id: int, areas: list):
    from holo_b200 import ospf_rib
    return ospf_rib.router_tables(router_id, areas, fn=lib().oracle_ospfv2_rib_router_tables)


def ospfv2_update_rib_partial(router_id, max_paths, areas, externals, sets, prev_rib, prev_rtrs):
    """Restatement of update_rib_partial (oracle/rib_partial.cc)."""
    from holo_b200 import ospf_rib
    return ospf_rib.update_rib_partial(router_id, max_paths, areas, externals, sets, prev_rib, prev_rtrs,
                                       fn=lib().oracle_ospfv2_update_rib_partial)


def ospfv3_update_rib_full(router_id: int, max_paths: int, areas: list, externals=None):
    from holo_b200 import ospf_rib
    return ospf_rib.call_update_rib_full(lib().oracle_ospfv3_update_rib_full, router_id, max_paths, areas, externals,
                                         v3=True)


def isis_compute_spt(level, root_system_id: int):
    """Reference-faithful compute_spt (local = false) over an IS-IS level image."""
    from holo_b200 import isis
    L = lib()
    L.oracle_isis_compute_spt.argtypes = [C.POINTER(isis.LevelStruct), C.c_uint64, C.POINTER(isis.SptStruct)]
    s = level.as_struct()
    return isis._call_spt(L.oracle_isis_compute_spt, len(level.lsps), len(level.reaches),
                          (C.byref(s), C.c_uint64(root_system_id)))


def ospfv3_run_area(area):
    """Reference-faithful OSPFv3 run_area + update_rib_intra_area over an LSDB image."""
    from holo_b200 import ospfv3
    L = lib()
    L.oracle_ospfv3_run_area.argtypes = [C.POINTER(ospfv3.AreaStruct), C.POINTER(ospfv3.ResultStruct)]
    return ospfv3._call_run_area(L.oracle_ospfv3_run_area, area)


def isis_compute_routes(inst):
    """Reference-faithful compute_spt(local = true) + compute_routes for one level."""
    from holo_b200 import isis
    L = lib()
    L.oracle_isis_compute_routes.argtypes = [C.POINTER(isis.InstanceStruct), C.POINTER(isis.RibStruct)]
    return isis._call_rib(L.oracle_isis_compute_routes, inst)


def usable_cores() -> int:
    """Cores this process may use (affinity mask capped by the cgroup CPU quota)."""
    return int(lib().oracle_usable_cores())


def effective_cores(threads: int = 0) -> float:
    """Measured parallel throughput of `threads` threads in units of one core (batch_pool.cc)."""
    L = lib()
    L.oracle_effective_cores.restype = C.c_double
    L.oracle_effective_cores.argtypes = [C.c_int]
    return float(L.oracle_effective_cores(threads if threads > 0 else usable_cores()))


def csr_batch(csr, roots, overrides=None, mode: str = "heap", vec_mode: int = 0, threads: int = 0,
              stop_after_s: float = 0.0, nh_words: int = 1, want_planes: bool = True):
    """Run a batch on a native thread pool (oracle/batch_pool.cc).  mode: "faithful" | "heap".
    Returns a dict: planes [n, V] (when want_planes), status [n], jobs_done, seconds, checksum."""
    L = lib()
    V = csr.n_vertices
    roots = np.ascontiguousarray(roots, dtype=np.uint32)
    n = len(roots)
    s = csr.as_struct()
    off = ove = ovc = None
    if overrides is not None:
        off = np.zeros(n + 1, np.uint32)
        ed, co = [], []
        for j, ov in enumerate(overrides):
            for e, c in ov:
                ed.append(e); co.append(c)
            off[j + 1] = len(ed)
        ove = np.asarray(ed or [0], np.uint32)
        ovc = np.asarray(co or [0], np.uint32)
    out = {}
    if want_planes:
        out = dict(dist=np.empty((n, V), np.uint32), hops=np.empty((n, V), np.uint16),
                   first_parent=np.empty((n, V), np.uint32), n_parents=np.empty((n, V), np.uint16),
                   nh_mask=np.empty((n, V, nh_words), np.uint64))
    status = np.zeros(n, np.uint32)
    done, secs, chk = C.c_uint32(), C.c_double(), C.c_uint64()
    if threads <= 0:
        threads = usable_cores()
    L.oracle_csr_batch.argtypes = [C.c_void_p, C.c_uint32, _u32p, _u32p, _u32p, _u32p, C.c_int, C.c_int, C.c_int,
                                   C.c_double, C.c_uint32, _u32p, _u16p, _u32p, _u16p, _u64p, _u32p,
                                   C.POINTER(C.c_uint32), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    rc = L.oracle_csr_batch(C.cast(C.byref(s), C.c_void_p), n, _p(roots, _u32p), _p(off, _u32p), _p(ove, _u32p),
                            _p(ovc, _u32p), 0 if mode == "faithful" else 1, vec_mode, threads, float(stop_after_s),
                            nh_words, _p(out.get("dist"), _u32p), _p(out.get("hops"), _u16p),
                            _p(out.get("first_parent"), _u32p), _p(out.get("n_parents"), _u16p),
                            _p(out.get("nh_mask"), _u64p), _p(status, _u32p), C.byref(done), C.byref(secs), C.byref(chk))
    if rc != 0:
        raise RuntimeError(f"oracle_csr_batch rc={rc}")
    out.update(status=status, jobs_done=done.value, seconds=secs.value, checksum=chk.value, threads=threads)
    return out
