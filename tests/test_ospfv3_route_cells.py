"""CPU: the batched route stage for OSPFv3 areas.  The device kernel's body (route_cell_eval, the same one the
OSPFv2 tests run) over the oracle's SPT planes and the OSPFv3 route table (advertisers in Intra-Area-Prefix-LSA
order, ospfv3/spf.rs:420-477), decoded by hspf_ospfv3_routes_from_cells, must equal the faithful oracle's
run_area routes for every root."""
import numpy as np
import pytest

from holo_b200 import capi, ospfv2, ospfv3, synth
from oracle import pyoracle
from test_ospfv2_route_cells import harness  # noqa: F401  (the CPU harness fixture)


def check_root(harness, t, root, frag=0, max_paths=16, mutate=None):
    area = ospfv3.synth_area(t, root=root, max_links_per_fragment=frag, max_paths=max_paths)
    if mutate:
        mutate(area)
    flat = ospfv3.Flat(area)
    rt = ospfv3.RouteTable(flat)
    rv = flat.router_vertex(area.router_id)
    c = pyoracle.csr_spf(flat.csr, rv, nh_words=1)
    assert c["status"] == 0
    d, h, m = (np.ascontiguousarray(c["dist"], np.uint32), np.ascontiguousarray(c["hops"], np.uint16),
               np.ascontiguousarray(c["nh_mask"], np.uint64).reshape(-1))
    cells = np.zeros(rt.n_prefixes, ospfv2.CELL_DT)
    harness.harness_route_cells(rt.handle, 1, d.ctypes.data, h.ctypes.data, m.ctypes.data, cells.ctypes.data)
    csr = flat.csr
    nets = sorted({int(v) for v in csr.col[csr.row_ptr[rv]: csr.row_ptr[rv + 1]] if not flat.is_router[v]})
    res = ospfv3.routes_from_cells(area, rt, cells, np.array(nets, np.uint32), m[nets] if nets else np.zeros(0, np.uint64))
    return area, rt, cells, res, pyoracle.ospfv3_run_area(area)


def same_routes(res, ref):
    assert res.rc == capi.HSPF_OK and len(res.routes) == len(ref.routes)
    keep = [n for n in res.routes.dtype.names if n != "nh_off"]
    assert np.array_equal(res.routes[keep], ref.routes[keep])
    for a, b in zip(res.routes, ref.routes):
        assert res.nh(a) == ref.nh(b)


@pytest.mark.parametrize("V,E,seed,kw,frag", [
    (80, 320, 1, {}, 0),
    (150, 700, 6, dict(cost_choices=[10, 20], lan_fraction=0.1), 0),
    (150, 700, 6, dict(cost_choices=[10, 20], lan_fraction=0.1), 3),          # Router-LSAs in several fragments
])
def test_cells_decode_to_the_oracle_routes_for_every_root(harness, V, E, seed, kw, frag):
    t = synth.random_topology(V, E, synth.SEED_BASE + 90 + seed, **kw)
    n_ecmp = 0
    for root in range(V):
        _, rt, cells, res, ref = check_root(harness, t, root, frag=frag)
        same_routes(res, ref)
        n_ecmp += int((res.routes["n_nh"] > 1).sum())
        assert rt.n_prefixes >= len(res.routes) > V
    if kw:
        assert n_ecmp > 0


def test_colliding_prefixes_and_max_paths(harness):
    """Several routers advertise one prefix with few distinct metrics; LAN prefixes are also advertised by attached
    routers: stub / stub, network / stub and network / network meetings, with a small max_paths."""
    t = synth.random_topology(90, 420, synth.SEED_BASE + 97, cost_choices=[10], lan_fraction=0.2)

    def mutate(area, rng=np.random.default_rng(3)):
        px = area.prefixes
        shared = [i for i in range(len(px)) if int(px["len"][i]) == 64][:40]
        pool = [px["addr"][shared[0]].copy(), px["addr"][shared[1]].copy(), px["addr"][shared[2]].copy()]
        for i in shared[3:]:
            if rng.random() < 0.6:
                px["addr"][i] = pool[int(rng.integers(0, 3))]
                px["metric"][i] = int(rng.choice([0, 5, 5, 10]))
    multi = 0
    for root in range(0, 90, 4):
        for mp in (1, 2, 16):
            _, rt, cells, res, ref = check_root(harness, t, root, max_paths=mp, mutate=mutate)
            same_routes(res, ref)
            multi += int((np.diff(rt.off.astype(np.int64)) > 2).sum())
    assert multi > 0


def test_table_is_a_v3_table():
    t = synth.random_topology(40, 160, synth.SEED_BASE + 98, lan_fraction=0.1)
    area = ospfv3.synth_area(t, root=0)
    rt = ospfv3.RouteTable(ospfv3.Flat(area))
    assert rt.n_prefixes == len(rt.prefix) == len(rt.plen) and rt.n_contributors >= rt.n_prefixes
    key = [(int(p["is_v6"]), bytes(int(b) for b in p["bytes"]), int(l)) for p, l in zip(rt.prefix, rt.plen)]
    assert key == sorted(set(key))                                       # IpNetwork order, unique
    # an OSPFv2 decode of a v3 table is refused and vice versa
    lib = capi.load_library()
    import ctypes as C
    assert lib.hspf_ospfv3_rtable_prefixes6(None, None, None) == capi.HSPF_E_INVAL
