"""CPU twins of tests/test_ospfv3_gpu.py: hspf_ospfv3_area_from_planes (csrc/ospfv3_host.cc:
fragment aggregation, Link-LSA next hops, Intra-Area-Prefix routes) fed with the oracle's SPT
planes equals the reference-faithful LSDB-level oracle byte for byte."""
import pytest

import golden_util as gu
from holo_b200 import ospfv3, synth
from oracle import pyoracle


def oracle_planes(csr, root, nh_words):
    c = pyoracle.csr_spf(csr, root, nh_words=nh_words)
    assert c["status"] == 0
    return c["dist"], c["hops"], c["nh_mask"]


def assert_same(res, ref):
    assert res.root_found == ref.root_found and res.transit_capability == ref.transit_capability
    for name in ("vertices", "routers", "routes", "nexthops"):
        a, b = getattr(res, name), getattr(ref, name)
        assert len(a) == len(b), (name, len(a), len(b))
        assert a.tobytes() == b.tobytes(), name


def twin(area):
    assert_same(ospfv3.area_from_planes(area, oracle_planes), pyoracle.ospfv3_run_area(area))


@pytest.mark.parametrize("V,E,seed,kw,root,frag", [
    (100, 400, 1, {}, 0, 0),
    (300, 1400, 6, dict(cost_choices=[10, 20], lan_fraction=0.1), 5, 3),
    (1000, 4000, 7, dict(lan_fraction=0.05), 234, 2),
])
def test_area_from_planes_matches_oracle(V, E, seed, kw, root, frag):
    t = synth.random_topology(V, E, synth.SEED_BASE + seed, **kw)
    twin(ospfv3.synth_area(t, root=root, max_links_per_fragment=frag))


def test_lan_members_and_filters():
    t = synth.random_topology(200, 900, synth.SEED_BASE + 11, cost_choices=[10, 20], lan_fraction=0.15)
    for members, _ in t.lans[:4]:
        for m in (members[0], members[-1]):
            twin(ospfv3.synth_area(t, root=m, max_links_per_fragment=2))
    area = ospfv3.synth_area(t, root=3, max_links_per_fragment=2)
    area.router_lsas["age"][7] = ospfv3.MAX_AGE
    area.router_lsas["options"][20] = 0
    area.network_lsas["age"][0] = ospfv3.MAX_AGE
    area.prefixes["options"][5] = ospfv3.PFX_NU
    twin(area)


SNAPS = gu.load_ospfv3()


@pytest.mark.parametrize("snap", SNAPS, ids=[f"{s['topo']}-{s['rt']}" for s in SNAPS])
def test_reference_golden_snapshots(snap):
    for area in snap["areas"]:
        twin(gu.ospfv3_area_image(snap, area))
