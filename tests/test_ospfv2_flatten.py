"""CPU: the product's LSDB->CSR flattener (host code, no GPU) agrees with the
reference-faithful LSDB-level oracle: running the abstract CSR oracle over the
flattened graph reproduces the LSDB oracle's SPT."""
import numpy as np
import pytest

import golden_util as gu
from holo_b200 import ospfv2, synth
from oracle import pyoracle


def check_flat_against_lsdb_oracle(area):
    ref = pyoracle.ospfv2_run_area(area)
    f = ospfv2.Flat(area)
    rv = f.router_vertex(area.router_id)
    if not ref.root_found:
        assert rv == 0xFFFFFFFF
        return
    # VertexId order: networks first, each ascending
    key = f.is_router.astype(np.uint64) << np.uint64(32) | f.ids.astype(np.uint64)
    assert (np.diff(key.astype(np.int64)) > 0).all()
    spt = pyoracle.csr_spf(f.csr, rv, nh_words=4)
    want = {(int(v["is_router"]), int(v["id"])): (int(v["distance"]), int(v["hops"]), int(v["n_nh"])) for v in ref.vertices}
    got = {}
    for v in range(f.csr.n_vertices):
        if spt["dist"][v] != 0xFFFFFFFF:
            got[(int(f.is_router[v]), int(f.ids[v]))] = (int(spt["dist"][v]), int(spt["hops"][v]))
    assert set(got) == set(want)
    for k, (d, h) in got.items():
        assert (d, h) == want[k][:2], k


@pytest.mark.parametrize("seed,kw,root", [
    (1, {}, 0), (6, dict(cost_choices=[10, 20], lan_fraction=0.1), 5), (8, dict(lan_fraction=0.3), 11)])
def test_flatten_synthetic(seed, kw, root):
    t = synth.random_topology(150, 700, synth.SEED_BASE + seed, **kw)
    check_flat_against_lsdb_oracle(ospfv2.synth_area(t, root=root))


def test_flatten_structural_filters():
    t = synth.random_topology(80, 360, synth.SEED_BASE + 14, lan_fraction=0.1)
    area = ospfv2.synth_area(t, root=2)
    area.router_lsas["age"][10] = ospfv2.MAX_AGE          # MaxAge LSA hides the vertex
    area.router_lsas["n_links"][20] = 1                   # one-way links fail the mutual check
    area.links["link_id"][int(area.router_lsas["link_off"][30])] = 0x01020304   # dangling link
    area.network_lsas["age"][0] = ospfv2.MAX_AGE
    check_flat_against_lsdb_oracle(area)
    f = ospfv2.Flat(area)
    assert f.router_vertex(int(area.router_lsas["adv_rtr"][10])) == 0xFFFFFFFF


def test_flatten_golden_snapshots():
    for snap in gu.load_ospfv2():
        for area in snap["areas"]:
            check_flat_against_lsdb_oracle(gu.ospfv2_area_image(snap, area))


def test_ospfv3_flatten_agrees_with_lsdb_oracle():
    from holo_b200 import ospfv3
    for seed, kw, root, frag in [(1, {}, 0, 0), (6, dict(cost_choices=[10, 20], lan_fraction=0.1), 5, 3)]:
        t = synth.random_topology(150, 700, synth.SEED_BASE + seed, **kw)
        area = ospfv3.synth_area(t, root=root, max_links_per_fragment=frag)
        ref = pyoracle.ospfv3_run_area(area)
        f = ospfv3.Flat(area)
        spt = pyoracle.csr_spf(f.csr, f.router_vertex(area.router_id), nh_words=4)
        want = {(int(v["is_router"]), int(v["router_id"]), int(v["iface_id"])): (int(v["distance"]), int(v["hops"]))
                for v in ref.vertices}
        got = {(int(f.is_router[v]), int(f.router_ids[v]), int(f.iface_ids[v])): (int(spt["dist"][v]), int(spt["hops"][v]))
               for v in range(f.csr.n_vertices) if spt["dist"][v] != 0xFFFFFFFF}
        assert got == want
    for snap in gu.load_ospfv3():
        for area in snap["areas"]:
            img = gu.ospfv3_area_image(snap, area)
            ref = pyoracle.ospfv3_run_area(img)
            f = ospfv3.Flat(img)
            if not ref.root_found:
                continue
            spt = pyoracle.csr_spf(f.csr, f.router_vertex(img.router_id), nh_words=4)
            assert (spt["dist"] != 0xFFFFFFFF).sum() == len(ref.vertices)
