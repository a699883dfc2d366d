"""CPU: the batched intra-area route stage (holo-ospf/src/route.rs:343-446 for every job of a batch).

The device kernel's body (route_cell_eval, holo_b200/csrc/route_cells.h) is compiled into a test
harness and run on the CPU over the oracle's SPT planes; the cells, decoded by the product's host
function hspf_ospfv2_routes_from_cells, must equal the routes and next hops (SR labels included) of
the reference-faithful LSDB-level oracle for every root — the same comparison tests/test_ospfv2_gpu.py
makes with the cells computed on the device."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from holo_b200 import capi, ospfv2, synth
from oracle import pyoracle

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def harness(built):
    out = ROOT / "tests" / "_build" / "libroute_cells_harness.so"
    src = ROOT / "tests" / "native" / "route_cells_harness.cc"
    hdr = ROOT / "holo_b200" / "csrc" / "route_cells.h"
    if not out.exists() or out.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        out.parent.mkdir(parents=True, exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-I", str(ROOT / "include"),
                        "-o", str(out), str(src)], check=True)
    lib = C.CDLL(str(out))
    lib.harness_route_cells.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.harness_route_cells16.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def oracle_planes(csr, root):
    c = pyoracle.csr_spf(csr, root, nh_words=1)
    assert c["status"] == 0
    return (np.ascontiguousarray(c["dist"], np.uint32), np.ascontiguousarray(c["hops"], np.uint16),
            np.ascontiguousarray(c["nh_mask"], np.uint64).reshape(-1))


def cells_on_cpu(harness, rt, planes, narrow=False):
    d, h, m = planes
    cells = np.zeros(rt.n_prefixes, ospfv2.CELL_DT)
    if narrow:
        assert int(m.max()) < 65536 and int(d[d != 0xFFFFFFFF].max()) < 0xFFFF
        d16 = np.where(d == 0xFFFFFFFF, 0xFFFF, d).astype(np.uint16)
        m16 = m.astype(np.uint16)
        harness.harness_route_cells16(rt.handle, 1, d16.ctypes.data, h.ctypes.data, m16.ctypes.data, cells.ctypes.data)
    else:
        harness.harness_route_cells(rt.handle, 1, d.ctypes.data, h.ctypes.data, m.ctypes.data, cells.ctypes.data)
    return cells


def gather_for(flat, root, planes):
    """nh_mask of the transit networks next to the root (what run_area_batch brings back per job)."""
    csr = flat.csr
    nets = sorted({int(v) for v in csr.col[csr.row_ptr[root]: csr.row_ptr[root + 1]] if not flat.is_router[v]})
    return np.array(nets, np.uint32), planes[2][nets] if nets else np.zeros(0, np.uint64)


def check_root(harness, t, root, sr=True, narrow=False, max_paths=16, mutate=None):
    area = ospfv2.synth_area(t, root=root, sr=sr, max_paths=max_paths)
    if mutate:
        mutate(area)
    flat = ospfv2.Flat(area)
    rt = ospfv2.RouteTable(flat)
    rv = flat.router_vertex(area.router_id)
    planes = oracle_planes(flat.csr, rv)
    cells = cells_on_cpu(harness, rt, planes, narrow)
    gv, gn = gather_for(flat, rv, planes)
    res = ospfv2.routes_from_cells(area, rt, cells, gv, gn)
    ref = pyoracle.ospfv2_run_area(area)
    return area, rt, cells, res, ref


def same_routes(res, ref):
    assert res.rc == capi.HSPF_OK
    assert len(res.routes) == len(ref.routes)
    drop = [n for n in res.routes.dtype.names if n != "nh_off"]     # offsets differ: ref also lists vertex next hops
    assert np.array_equal(res.routes[drop], ref.routes[drop])
    for a, b in zip(res.routes, ref.routes):
        assert res.nh(a) == ref.nh(b), (hex(int(a["prefix"])), res.nh(a), ref.nh(b))


@pytest.mark.parametrize("V,E,seed,kw,sr", [
    (100, 400, 1, {}, False),
    (100, 400, 1, {}, True),
    (300, 1400, 6, dict(cost_choices=[10, 20], lan_fraction=0.1), True),       # C5 shape: ECMP, LANs, SR
    (200, 900, 11, dict(cost_choices=[10, 20], lan_fraction=0.15), True),
])
def test_cells_decode_to_the_oracle_routes_for_every_root(harness, V, E, seed, kw, sr):
    t = synth.random_topology(V, E, synth.SEED_BASE + seed, **kw)
    n_ecmp = n_conn = 0
    for root in range(V):
        _, rt, cells, res, ref = check_root(harness, t, root, sr=sr)
        same_routes(res, ref)
        n_ecmp += int((res.routes["n_nh"] > 1).sum())
        n_conn += int((cells["flags"] & ospfv2.CELL_CONNECTED != 0).sum())
        assert rt.n_prefixes >= len(res.routes) > V // 2
    assert n_conn > 0
    if kw:
        assert n_ecmp > 0


def test_lan_members_as_roots_and_narrow_planes(harness):
    t = synth.random_topology(200, 900, synth.SEED_BASE + 11, cost_choices=[10, 20], lan_fraction=0.15)
    for members, _ in t.lans[:6]:
        for m in (members[0], members[-1]):
            for narrow in (False, True):
                try:
                    _, _, _, res, ref = check_root(harness, t, m, narrow=narrow)
                except AssertionError:
                    if narrow:
                        continue            # more than 16 atoms: no narrow planes for this root
                    raise
                same_routes(res, ref)


def test_max_paths_truncation(harness):
    t = synth.random_topology(60, 600, synth.SEED_BASE + 13, cost_choices=[10])
    seen = 0
    for root in (0, 7, 31):
        _, _, _, res, ref = check_root(harness, t, root, max_paths=2)
        same_routes(res, ref)
        seen += int((res.routes["n_nh"] == 2).sum())
    assert seen > 0


def test_equal_cost_advertisers_with_different_sids_are_flagged(harness):
    """Two routers advertise the same stub prefix with different Prefix-SIDs at equal cost: the
    merged route's labels depend on the order of the advertisers per next hop — the cell is flagged
    and the decode refuses it (caller: hspf_ospfv2_area_from_planes for that job)."""
    t = synth.random_topology(40, 200, synth.SEED_BASE + 17, cost_choices=[10])
    area, rt, cells, res, ref = check_root(harness, t, 0)
    same_routes(res, ref)
    # find two routers at the same distance from the root and give both the same extra stub prefix
    flat = ospfv2.Flat(area)
    rv = flat.router_vertex(area.router_id)
    d = oracle_planes(flat.csr, rv)[0]
    rtr = [v for v in range(len(d)) if flat.is_router[v] and v != rv]
    by_d = {}
    for v in rtr:
        by_d.setdefault(int(d[v]), []).append(v)
    a, b = next(vs for vs in by_d.values() if len(vs) >= 2)[:2]
    ida, idb = int(flat.ids[a]), int(flat.ids[b])

    def mutate(ar, same_sid):
        links = ar.links
        ext = ar.ext_prefixes
        # re-point each router's loopback stub (the prefix that carries its Prefix-SID) at one shared prefix
        for rid, sid in ((ida, 700), (idb, 700 if same_sid else 701)):
            li = int(np.nonzero(ar.router_lsas["adv_rtr"] == rid)[0][0])
            lo, n = int(ar.router_lsas["link_off"][li]), int(ar.router_lsas["n_links"][li])
            k = next(i for i in range(lo, lo + n)
                     if links["link_type"][i] == ospfv2.LINK_STUB and int(links["link_data"][i]) == 0xFFFFFFFF)
            old = (int(links["link_id"][k]), int(links["link_data"][k]))
            links["link_id"][k], links["link_data"][k], links["metric"][k] = 0xC6336400, 0xFFFFFF00, 5
            e = next(i for i in range(len(ext)) if int(ext["adv_rtr"][i]) == rid and int(ext["prefix"][i]) == old[0])
            ext["prefix"][e], ext["mask"][e], ext["sid_value"][e] = 0xC6336400, 0xFFFFFF00, sid

    _, _, cells, res, ref = check_root(harness, t, 0, mutate=lambda ar: mutate(ar, True))
    same_routes(res, ref)                                 # same SID on both: an ordinary ECMP merge
    _, rt2, cells, res, ref = check_root(harness, t, 0, mutate=lambda ar: mutate(ar, False))
    p = int(np.nonzero(rt2.prefix == 0xC6336400)[0][0])
    assert cells["flags"][p] & ospfv2.CELL_MIXED_SID
    assert res.rc == capi.HSPF_E_UNSUPPORTED


def test_table_shape_and_order():
    t = synth.random_topology(120, 500, synth.SEED_BASE + 19, lan_fraction=0.1)
    area = ospfv2.synth_area(t, root=3, sr=True)
    rt = ospfv2.RouteTable(ospfv2.Flat(area))
    key = rt.prefix.astype(np.uint64) << 8 | rt.plen
    assert np.all(np.diff(key.astype(np.int64)) > 0)                    # Ipv4Network order, unique
    assert rt.off[0] == 0 and rt.off[-1] == rt.n_contributors and np.all(np.diff(rt.off.astype(np.int64)) >= 1)
    # every p2p /30 is advertised by both ends, every LAN prefix by its transit network
    assert rt.n_contributors > rt.n_prefixes
    assert int(rt.contribs["is_network"].sum()) == len(area.network_lsas)


FUZZ_STATS = {}


def collide(area, rng):
    """Make prefixes meet: stub links of several routers re-pointed at a small pool of prefixes (with few
    distinct metrics, so ties are common), some of them a transit network's own prefix; transit networks
    widened to /16 so that several map to one prefix; Prefix-SIDs for some of the new (router, prefix) pairs,
    equal or different; per-router SRGBs of different bases; a few PHP / explicit-null flag variants."""
    links, rl = area.links, area.router_lsas
    nets = area.network_lsas
    pool = [(0xC6336400 + (i << 8), 0xFFFFFF00) for i in range(4)]
    if len(nets):
        for j in rng.choice(len(nets), min(len(nets), 3), replace=False):
            if rng.random() < 0.5:
                nets["mask"][j] = 0xFFFF0000                       # several LANs -> one /16
            pool.append((int(nets["lsa_id"][j]) & int(nets["mask"][j]), int(nets["mask"][j])))
    ext = list(area.ext_prefixes) if area.sr_enabled else []
    for i in rng.choice(len(rl), min(len(rl), 14), replace=False):
        lo, n = int(rl["link_off"][i]), int(rl["n_links"][i])
        stubs = [k for k in range(lo, lo + n) if links["link_type"][k] == ospfv2.LINK_STUB and int(links["link_data"][k]) != 0xFFFFFFFF]
        if not stubs:
            continue
        k = int(rng.choice(stubs))
        p, m = pool[int(rng.integers(0, len(pool)))]
        links["link_id"][k], links["link_data"][k], links["metric"][k] = p, m, int(rng.choice([0, 5, 5, 10]))
        if area.sr_enabled and rng.random() < 0.7:
            flags = int(rng.choice([ospfv2.PSID_NP, ospfv2.PSID_NP, 0, ospfv2.PSID_NP | ospfv2.PSID_E]))
            ext.append((int(rl["adv_rtr"][i]), p, m, 1, 1, 1, flags, 0, (0, 0), int(rng.choice([900, 900, 901]))))
    if area.sr_enabled:
        area.ext_prefixes = np.asarray(ext, ospfv2.EXT_PREFIX_DT)
        area.ext_prefixes = area.ext_prefixes[np.lexsort((area.ext_prefixes["prefix"], area.ext_prefixes["adv_rtr"]))]
        for i in range(len(area.srgbs)):
            area.srgbs["first"][i] = 16000 + 1000 * (i % 5)


@pytest.mark.parametrize("seed", range(12))
def test_colliding_prefixes_fuzz(harness, seed):
    """Every way two advertisers can meet on one prefix (stub / stub, stub / transit network, network /
    network, equal and unequal metrics, equal and different Prefix-SIDs): the decoded cells equal the
    faithful oracle's routes, or the decode refuses the job (different SIDs merged) — never a wrong route."""
    rng = np.random.default_rng(500 + seed)
    V = int(rng.integers(30, 90))
    t = synth.random_topology(V, int(V * rng.uniform(2.5, 5)), synth.SEED_BASE + 70 + seed,
                              cost_choices=[int(x) for x in rng.choice([5, 10, 10, 20], 2)], lan_fraction=float(rng.uniform(0.1, 0.35)))
    sr = bool(rng.random() < 0.75)
    mp = int(rng.choice([1, 2, 16]))
    mut_seed = int(rng.integers(0, 1 << 30))
    n_ok = n_refused = n_multi = 0
    for root in rng.choice(V, 8, replace=False):
        area, rt, cells, res, ref = check_root(harness, t, int(root), sr=sr, max_paths=mp,
                                               mutate=lambda ar: collide(ar, np.random.default_rng(mut_seed)))
        n_multi += int((np.diff(rt.off.astype(np.int64)) > 2).sum())
        if res.rc == capi.HSPF_E_UNSUPPORTED:
            assert sr and (cells["flags"] & ospfv2.CELL_MIXED_SID).any()
            n_refused += 1
            continue
        same_routes(res, ref)
        n_ok += 1
    FUZZ_STATS[seed] = (n_ok, n_refused)
    assert n_ok + n_refused == 8 and n_multi > 0


def test_colliding_prefixes_fuzz_covers_both_outcomes():
    if len(FUZZ_STATS) < 12:
        pytest.skip("runs after the whole fuzz")
    assert sum(a for a, _ in FUZZ_STATS.values()) >= 40 and sum(b for _, b in FUZZ_STATS.values()) >= 8


def test_c5_size_lsdb(harness):
    """BASELINE C5 shape at full size (10 000 routers, 5 % of the adjacencies on LANs, costs {10, 20}, Prefix-SIDs):
    29 213 prefixes, 48 212 advertisers; the cells of three roots — the bench's local router, a LAN's designated
    router, a router in the middle — decode to the faithful oracle's route tables."""
    t = synth.random_topology(10000, 40000, synth.SEED_BASE + 5, cost_choices=[10, 20], lan_fraction=0.05)
    for root in (0, int(t.lans[0][0][0]), 5000):
        area, rt, cells, res, ref = check_root(harness, t, root)
        same_routes(res, ref)
        assert rt.n_prefixes == 29213 and rt.n_contributors == 48212
        assert len(res.routes) == rt.n_prefixes and int((res.routes["n_nh"] > 1).sum()) > 1000
        assert int(res.routes["has_sr_label"].sum()) >= 9999
