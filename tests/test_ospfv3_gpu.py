"""GPU parity of the OSPFv3 LSDB-level path (hspf_ospfv3_run_area through the C ABI)
against the line-faithful oracle and the reference's golden OSPFv3 local-ribs."""
import numpy as np
import pytest

import golden_util as gu
from holo_b200 import ospfv3, synth
from oracle import pyoracle

pytestmark = pytest.mark.gpu


def assert_same(res, ref):
    assert res.root_found == ref.root_found and res.transit_capability == ref.transit_capability
    for name in ("vertices", "routers", "routes", "nexthops"):
        a, b = getattr(res, name), getattr(ref, name)
        assert len(a) == len(b), (name, len(a), len(b))
        assert a.tobytes() == b.tobytes(), name


@pytest.mark.parametrize("V,E,seed,kw,root,frag", [
    (100, 400, 1, {}, 0, 0),
    (300, 1400, 6, dict(cost_choices=[10, 20], lan_fraction=0.1), 5, 3),
    (2000, 8000, 7, dict(lan_fraction=0.05), 1234, 2),        # one BASELINE-C4-sized area
])
def test_run_area_matches_oracle(ctx, V, E, seed, kw, root, frag):
    t = synth.random_topology(V, E, synth.SEED_BASE + seed, **kw)
    area = ospfv3.synth_area(t, root=root, max_links_per_fragment=frag)
    assert_same(ospfv3.run_area(ctx, area), pyoracle.ospfv3_run_area(area))


def test_lan_members_and_filters(ctx):
    t = synth.random_topology(200, 900, synth.SEED_BASE + 11, cost_choices=[10, 20], lan_fraction=0.15)
    for members, _ in t.lans[:4]:
        for m in (members[0], members[-1]):
            area = ospfv3.synth_area(t, root=m, max_links_per_fragment=2)
            assert_same(ospfv3.run_area(ctx, area), pyoracle.ospfv3_run_area(area))
    area = ospfv3.synth_area(t, root=3, max_links_per_fragment=2)
    area.router_lsas["age"][7] = ospfv3.MAX_AGE                 # one fragment aged out
    area.router_lsas["options"][20] = 0                        # R-bit clear: fragment ignored
    area.network_lsas["age"][0] = ospfv3.MAX_AGE
    area.prefixes["options"][5] = ospfv3.PFX_NU
    assert_same(ospfv3.run_area(ctx, area), pyoracle.ospfv3_run_area(area))


SNAPS = gu.load_ospfv3()


@pytest.mark.parametrize("snap", SNAPS, ids=[f"{s['topo']}-{s['rt']}" for s in SNAPS])
def test_reference_golden_local_rib(ctx, snap):
    want = gu.golden_intra(snap)
    per_area = []
    for area in snap["areas"]:
        img = gu.ospfv3_area_image(snap, area)
        res = ospfv3.run_area(ctx, img)
        assert_same(res, pyoracle.ospfv3_run_area(img))
        if res.root_found:
            per_area.append(gu.routes6_as_dict(res, img.ifnames))
    got = gu.merge_area_routes(per_area)
    has_vlink = any(i["state"] == "virtual-link" for a in snap["areas"] for i in a["interfaces"])
    norm = lambda nh: sorted(((a or ""), (b or "")) for a, b in nh)
    for prefix, (metric, nh) in want.items():
        assert got[prefix][0] == metric
        if has_vlink and not got[prefix][1]:
            continue
        assert norm(got[prefix][1]) == norm(nh)


@pytest.mark.parametrize("snap", SNAPS, ids=[f"{s['topo']}-{s['rt']}" for s in SNAPS])
def test_reference_golden_whole_local_rib(ctx, snap):
    """LSDB -> full OSPFv3 routing table through the product only: hspf_ospfv3_run_area on the GPU
    per attached area, then hspf_ospfv3_update_rib_full (the same helper runs on the CPU with the
    oracle in tests/test_ospf_rib.py and tests/test_oracle_golden.py)."""
    from holo_b200 import ospf_rib
    got = gu.ospfv3_full_rib(snap, lambda img: ospfv3.run_area(ctx, img), ospf_rib.update_rib_full_v3)
    want = gu.golden_rib(snap)
    assert set(got) == set(want)
    for prefix, (metric, rtype, nh) in want.items():
        g = got[prefix]
        assert (g[0], g[1]) == (metric, rtype), (prefix, g)
        assert [(a or "", b or "") for a, b in g[2]] == [(a or "", b or "") for a, b in nh], (prefix, g[2], nh)
