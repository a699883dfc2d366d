"""CPU twins of tests/test_ospfv2_gpu.py: the product's post-SPT half of run_area
(hspf_ospfv2_area_from_planes: Vertex.nexthops from the atom sets, router table,
transit_capability, intra-area routes, SR labels — csrc/ospfv2_host.cc) fed with the oracle's
SPT planes must equal the reference-faithful LSDB-level oracle record for record, on synthetic
LSDBs and on every golden snapshot of the reference."""
import numpy as np
import pytest

import golden_util as gu
from holo_b200 import ospfv2, synth
from oracle import pyoracle


def oracle_planes(csr, root, nh_words):
    c = pyoracle.csr_spf(csr, root, nh_words=nh_words)
    assert c["status"] == 0
    return c["dist"], c["hops"], c["nh_mask"]


def assert_same(res, ref):
    assert res.root_found == ref.root_found
    assert res.transit_capability == ref.transit_capability
    for name in ("vertices", "routers", "routes", "nexthops"):
        a, b = getattr(res, name), getattr(ref, name)
        assert len(a) == len(b), (name, len(a), len(b))
        if not np.array_equal(a, b):
            bad = [i for i in range(len(a)) if a[i] != b[i]][:3]
            raise AssertionError(f"{name} differ at {bad}: got {[a[i] for i in bad]} want {[b[i] for i in bad]}")


def twin(area):
    res = ospfv2.area_from_planes(area, oracle_planes)
    assert_same(res, pyoracle.ospfv2_run_area(area))
    return res


@pytest.mark.parametrize("V,E,seed,kw,root,sr", [
    (100, 400, 1, {}, 0, False),
    (100, 400, 1, {}, 37, True),
    (300, 1400, 6, dict(cost_choices=[10, 20], lan_fraction=0.1), 5, True),
    (300, 1400, 6, dict(cost_choices=[10, 20], lan_fraction=0.1), 17, False),
    (1000, 4000, 7, dict(lan_fraction=0.05), 234, True),
])
def test_area_from_planes_matches_oracle(V, E, seed, kw, root, sr):
    t = synth.random_topology(V, E, synth.SEED_BASE + seed, **kw)
    res = twin(ospfv2.synth_area(t, root=root, sr=sr))
    assert len(res.routes) > V


def test_lan_members_as_root():
    t = synth.random_topology(200, 900, synth.SEED_BASE + 11, cost_choices=[10, 20], lan_fraction=0.15)
    for members, _ in t.lans[:6]:
        for m in (members[0], members[-1]):
            twin(ospfv2.synth_area(t, root=m, sr=True))


def test_max_paths_truncation():
    t = synth.random_topology(60, 600, synth.SEED_BASE + 13, cost_choices=[10])
    res = twin(ospfv2.synth_area(t, root=0, max_paths=2))
    assert res.routes["n_nh"].max() == 2


def test_adversarial_lsdb_features():
    t = synth.random_topology(80, 360, synth.SEED_BASE + 14, lan_fraction=0.1)
    area = ospfv2.synth_area(t, root=2, sr=True)
    area.router_lsas["age"][10] = ospfv2.MAX_AGE
    area.router_lsas["n_links"][20] = 1
    area.links["link_id"][int(area.router_lsas["link_off"][30])] = 0x01020304
    if len(area.network_lsas):
        area.network_lsas["age"][0] = ospfv2.MAX_AGE
    twin(area)


def test_root_not_found():
    t = synth.random_topology(10, 30, synth.SEED_BASE + 15)
    area = ospfv2.synth_area(t, root=0)
    area.router_id = 0x7F000001
    res = ospfv2.area_from_planes(area, oracle_planes)
    assert not res.root_found and len(res.vertices) == 0


SNAPS = gu.load_ospfv2()


@pytest.mark.parametrize("snap", SNAPS, ids=[f"{s['topo']}-{s['rt']}" for s in SNAPS])
def test_reference_golden_snapshots(snap):
    for area in snap["areas"]:
        twin(gu.ospfv2_area_image(snap, area))


AFTER = [(s, n) for s in SNAPS for n in s.get("after", {})]


@pytest.mark.parametrize("snap,name", AFTER, ids=[f"{n}-{s['topo']}-{s['rt']}" for s, n in AFTER])
def test_step_after_state_snapshots(snap, name):
    """The LSDBs the reference reached after its step tests (expired LSAs, removed areas and
    interfaces, changed costs ...): product host stage == LSDB-level oracle on each."""
    after = dict(snap["after"][name], ifindex=snap["ifindex"])
    if not after.get("router_id"):
        pytest.skip("instance disabled in this state")
    for area in after["areas"]:
        twin(gu.ospfv2_area_image(after, area))
