"""GPU parity of the OSPFv2 LSDB-level path (hspf_ospfv2_run_area through the C ABI)
against the line-faithful oracle and against the reference's golden local-ribs."""
import numpy as np
import pytest

import golden_util as gu
from holo_b200 import ospfv2, synth
from oracle import pyoracle

pytestmark = pytest.mark.gpu


def assert_same(res, ref):
    assert res.root_found == ref.root_found
    assert res.transit_capability == ref.transit_capability
    for name in ("vertices", "routers", "routes", "nexthops"):
        a, b = getattr(res, name), getattr(ref, name)
        assert len(a) == len(b), (name, len(a), len(b))
        if not np.array_equal(a, b):
            bad = [i for i in range(len(a)) if a[i] != b[i]][:3]
            raise AssertionError(f"{name} differ at {bad}: got {[a[i] for i in bad]} want {[b[i] for i in bad]}")


@pytest.mark.parametrize("V,E,seed,kw,root,sr", [
    (100, 400, 1, {}, 0, False),                                    # BASELINE C1 shape
    (100, 400, 1, {}, 37, True),
    (300, 1400, 6, dict(cost_choices=[10, 20], lan_fraction=0.1), 5, True),
    (300, 1400, 6, dict(cost_choices=[10, 20], lan_fraction=0.1), 17, False),
    (2000, 8000, 7, dict(lan_fraction=0.05), 1234, True),
])
def test_run_area_matches_oracle(ctx, V, E, seed, kw, root, sr):
    t = synth.random_topology(V, E, synth.SEED_BASE + seed, **kw)
    area = ospfv2.synth_area(t, root=root, sr=sr)
    res = ospfv2.run_area(ctx, area)
    ref = pyoracle.ospfv2_run_area(area)
    assert_same(res, ref)
    assert len(res.routes) > V


def test_lan_members_as_root(ctx):
    t = synth.random_topology(200, 900, synth.SEED_BASE + 11, cost_choices=[10, 20], lan_fraction=0.15)
    for members, _ in t.lans[:6]:
        for m in (members[0], members[-1]):        # the DR and another attached router
            area = ospfv2.synth_area(t, root=m, sr=True)
            assert_same(ospfv2.run_area(ctx, area), pyoracle.ospfv2_run_area(area))


def test_c5_shape_full_route_table(ctx):
    # BASELINE config 5: 10k-node ECMP-rich LSDB with LANs and SR prefix-SIDs
    t = synth.random_topology(10000, 40000, synth.SEED_BASE + 5, cost_choices=[10, 20], lan_fraction=0.05)
    area = ospfv2.synth_area(t, root=0, sr=True)
    res = ospfv2.run_area(ctx, area)
    ref = pyoracle.ospfv2_run_area(area)
    assert_same(res, ref)
    assert res.routes["has_sr_label"].sum() >= 9999
    assert (res.routes["n_nh"] > 1).sum() > 100          # ECMP present


def test_max_paths_truncation(ctx):
    t = synth.random_topology(60, 600, synth.SEED_BASE + 13, cost_choices=[10])
    area = ospfv2.synth_area(t, root=0, max_paths=2)
    res = ospfv2.run_area(ctx, area)
    assert_same(res, pyoracle.ospfv2_run_area(area))
    assert res.routes["n_nh"].max() == 2


def test_adversarial_lsdb_features(ctx):
    t = synth.random_topology(80, 360, synth.SEED_BASE + 14, lan_fraction=0.1)
    area = ospfv2.synth_area(t, root=2, sr=True)
    # MaxAge Router-LSA, one-way link (drop a router's links), missing LSA target
    area.router_lsas["age"][10] = ospfv2.MAX_AGE
    area.router_lsas["n_links"][20] = 1
    area.links["link_id"][int(area.router_lsas["link_off"][30])] = 0x01020304
    if len(area.network_lsas):
        area.network_lsas["age"][0] = ospfv2.MAX_AGE
    assert_same(ospfv2.run_area(ctx, area), pyoracle.ospfv2_run_area(area))


def test_root_not_found(ctx):
    t = synth.random_topology(10, 30, synth.SEED_BASE + 15)
    area = ospfv2.synth_area(t, root=0)
    area.router_id = 0x7F000001
    res = ospfv2.run_area(ctx, area)
    assert not res.root_found and len(res.vertices) == 0


SNAPS = gu.load_ospfv2()


@pytest.mark.parametrize("snap", SNAPS, ids=[f"{s['topo']}-{s['rt']}" for s in SNAPS])
def test_reference_golden_local_rib(ctx, snap):
    """The reference's own conformance snapshots, through the GPU path."""
    want = gu.golden_intra(snap)
    per_area = []
    for area in snap["areas"]:
        img = gu.ospfv2_area_image(snap, area)
        res = ospfv2.run_area(ctx, img)
        if res.root_found:
            per_area.append(gu.routes_as_dict(res, img.ifnames))
        assert_same(res, pyoracle.ospfv2_run_area(img))
    got = gu.merge_area_routes(per_area)
    has_vlink = any(i["cfg_type"] == "virtual-link" for a in snap["areas"] for i in a["interfaces"])
    norm = lambda nh: sorted(((a or ""), (b or "")) for a, b in nh)
    for prefix, (metric, nh) in want.items():
        assert got[prefix][0] == metric
        if has_vlink and not got[prefix][1]:
            continue   # virtual-link next hops: SURVEY §8f f1 (see test_oracle_golden.py)
        assert norm(got[prefix][1]) == norm(nh)


@pytest.mark.parametrize("snap", SNAPS, ids=[f"{s['topo']}-{s['rt']}" for s in SNAPS])
def test_reference_golden_whole_local_rib(ctx, snap):
    """LSDB -> full routing table through the product only: hspf_ospfv2_run_area on the GPU per
    attached area, then hspf_ospfv2_update_rib_full (inter-area, transit areas, externals):
    every route of the reference's local-rib, and nothing else (the same helper runs on the
    CPU with the oracle in tests/test_ospf_rib.py and tests/test_oracle_golden.py)."""
    from holo_b200 import ospf_rib
    got = gu.ospfv2_full_rib(snap, lambda img: ospfv2.run_area(ctx, img), ospf_rib.update_rib_full)
    want = gu.golden_rib(snap)
    assert set(got) == set(want)
    for prefix, (metric, rtype, nh) in want.items():
        g = got[prefix]
        assert (g[0], g[1]) == (metric, rtype), (prefix, g)
        assert [(a or "", b or "") for a, b in g[2]] == [(a or "", b or "") for a, b in nh], (prefix, g[2], nh)


# ---- batched route stage on the device (hspf_ospfv2_run_area_batch, csrc/ospfv2_routes.cu) ----------
def same_routes(res, ref):
    from holo_b200 import capi
    assert res.rc == capi.HSPF_OK
    assert len(res.routes) == len(ref.routes)
    keep = [n for n in res.routes.dtype.names if n != "nh_off"]
    assert np.array_equal(res.routes[keep], ref.routes[keep])
    for a, b in zip(res.routes, ref.routes):
        assert res.nh(a) == ref.nh(b), (hex(int(a["prefix"])), res.nh(a), ref.nh(b))


@pytest.mark.parametrize("V,E,seed,kw", [
    (100, 400, 1, {}),
    (300, 1400, 6, dict(cost_choices=[10, 20], lan_fraction=0.1)),          # C5 shape: ECMP, LANs, SR
])
def test_run_area_batch_every_root_matches_oracle(ctx, V, E, seed, kw):
    """One device batch over EVERY router of the area as root: SPTs and the intra-area route cells; each
    root's cells, decoded with that root's interface state, equal the faithful oracle's run_area routes
    (prefixes, metrics, origins, flags, Prefix-SIDs, next hops, SR labels)."""
    t = synth.random_topology(V, E, synth.SEED_BASE + seed, **kw)
    area0 = ospfv2.synth_area(t, root=0, sr=True)
    rids = [int(ospfv2.RID_BASE + i) for i in range(V)]
    b = ospfv2.run_area_batch(ctx, area0, rids)
    assert b.rc == 0 and not b.status.any()
    rt = ospfv2.RouteTable(ospfv2.Flat(area0))
    assert b.cells.shape == (V, rt.n_prefixes)
    n_ecmp = 0
    for root in range(V):
        area = ospfv2.synth_area(t, root=root, sr=True)
        gv, gn = b.gather(root)
        res = ospfv2.routes_from_cells(area, rt, b.cells[root], gv, gn)
        same_routes(res, pyoracle.ospfv2_run_area(area))
        n_ecmp += int((res.routes["n_nh"] > 1).sum())
    if kw:
        assert n_ecmp > 0


def test_routes_batch_on_device_planes_wide_and_narrow(ctx):
    """hspf_ospfv2_routes_batch / _batch16 over planes that never leave the device: both plane widths
    give the same cells as the one-call batch."""
    import ctypes as C
    import torch
    from holo_b200 import capi
    V = 200
    t = synth.random_topology(V, 900, synth.SEED_BASE + 23, cost_choices=[10, 20])
    area = ospfv2.synth_area(t, root=0, sr=True)
    flat = ospfv2.Flat(area)
    rt = ospfv2.RouteTable(flat)
    rt.upload(ctx)
    rids = [int(ospfv2.RID_BASE + i) for i in range(V)]
    want = ospfv2.run_area_batch(ctx, area, rids)
    roots_v = np.array([flat.router_vertex(r) for r in rids], np.uint32)
    g = ctx.upload(flat.csr)
    nv = flat.csr.n_vertices
    roots = torch.from_numpy(roots_v.astype(np.int32)).cuda()
    js = capi.JobsStruct()
    js.n_jobs = V
    js.roots = C.cast(roots.data_ptr(), C.POINTER(C.c_uint32))
    P = rt.n_prefixes
    for narrow in (False, True):
        dt = torch.int16 if narrow else torch.int32
        d = torch.zeros((V, nv), dtype=dt, device="cuda")
        h = torch.zeros((V, nv), dtype=torch.int16, device="cuda")
        m = torch.zeros((V, nv), dtype=torch.int16 if narrow else torch.int64, device="cuda")
        st = torch.zeros((V,), dtype=torch.int32, device="cuda")
        cells = torch.zeros((V * P * ospfv2.CELL_DT.itemsize,), dtype=torch.uint8, device="cuda")
        if narrow:
            rs = capi.Result16Struct()
            rs.dist = C.cast(d.data_ptr(), C.POINTER(C.c_uint16))
            rs.nh_mask = C.cast(m.data_ptr(), C.POINTER(C.c_uint16))
        else:
            rs = capi.ResultStruct()
            rs.dist = C.cast(d.data_ptr(), C.POINTER(C.c_uint32))
            rs.nh_mask = C.cast(m.data_ptr(), C.POINTER(C.c_uint64))
            rs.nh_words = 1
        rs.hops = C.cast(h.data_ptr(), C.POINTER(C.c_uint16))
        rs.job_status = C.cast(st.data_ptr(), C.POINTER(C.c_uint32))
        (ctx.run_device16 if narrow else ctx.run_device)(g, js, rs, sync=False)
        ospfv2.routes_batch_device(ctx, rt, V, rs, cells.data_ptr())
        ctx.sync()
        status = st.cpu().numpy()
        got = np.frombuffer(cells.cpu().numpy().tobytes(), ospfv2.CELL_DT).reshape(V, P)
        ok = status == 0                         # a root with more than 16 atoms has no narrow planes
        assert ok.sum() > V // 2 and (narrow or ok.all())
        assert got[ok].tobytes() == want.cells[ok].tobytes()
        assert not (got[~ok]["flags"]).any()     # refused jobs: empty cells


def test_interface_cost_change_reuses_the_uploaded_graph(ctx):
    """SURVEY 8f f2 end to end: a Router-LSA comes back with other link metrics; hspf_ospfv2_flat_update names
    the changed CSR edges, hspf_graph_update_costs patches the device image in place, and the SPT + routes of
    the patched graph equal the oracle's on the new LSDB."""
    import copy
    t = synth.random_topology(300, 1400, synth.SEED_BASE + 6, cost_choices=[10, 20], lan_fraction=0.1)
    area = ospfv2.synth_area(t, root=5, sr=True)
    flat = ospfv2.Flat(area)
    g = ctx.upload(flat.csr)
    new = copy.deepcopy(area)
    rids = [int(ospfv2.RID_BASE + i) for i in (5, 6, 40)]
    for rid in rids:
        i = int(np.nonzero(new.router_lsas["adv_rtr"] == rid)[0][0])
        lo, n = int(new.router_lsas["link_off"][i]), int(new.router_lsas["n_links"][i])
        for k in range(lo, lo + n):
            if new.links["link_type"][k] != ospfv2.LINK_STUB:
                new.links["metric"][k] = 7 + 3 * (k % 5)
    kind, edges, costs = ospfv2.flat_update(flat, new, [(r, r, 0, 1, 0, (0, 0)) for r in rids])
    assert kind == ospfv2.FLAT_COSTS and len(edges) > 0
    ctx.update_costs(g, edges, costs)
    root = flat.router_vertex(new.router_id)

    def planes(csr, r, nhw):
        res = ctx.run(g, [r], nh_words=nhw)
        return res.dist[0], res.hops[0], res.nh_mask[0]
    res = ospfv2.area_from_planes(new, planes)
    assert_same(res, pyoracle.ospfv2_run_area(new))
    assert_same(ospfv2.run_area(ctx, new), res)          # and the full re-flatten + upload path agrees
    g.free()
