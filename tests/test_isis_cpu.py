"""CPU: IS-IS host logic (flatten, Spt reconstruction) against the faithful oracle.
The device planes are stood in for by the CSR-level oracle here; the GPU parity
proper is tests/test_isis_gpu.py."""
import numpy as np
import pytest

from holo_b200 import isis, synth
from oracle import pyoracle

MODES = [
    (isis.MT_STANDARD, isis.METRIC_WIDE, isis.MODE_NORMAL),
    (isis.MT_STANDARD, isis.METRIC_STANDARD, isis.MODE_NORMAL),
    (isis.MT_STANDARD, isis.METRIC_BOTH, isis.MODE_NORMAL),
    (isis.MT_IPV6, isis.METRIC_WIDE, isis.MODE_NORMAL),
    (isis.MT_NONE, isis.METRIC_WIDE, isis.MODE_HOPCOUNT),
]


def same_spt(a, b):
    for name in ("vertices", "parents", "nexthops", "first_hops", "second_hops"):
        x, y = getattr(a, name), getattr(b, name)
        assert len(x) == len(y), (name, len(x), len(y))
        assert np.array_equal(x, y), name


@pytest.mark.parametrize("mt,mtype,mode", MODES)
def test_flatten_and_reconstruct(mt, mtype, mode):
    t = synth.random_topology(80, 340, synth.SEED_BASE + 3, cost_lo=1, cost_hi=30, lan_fraction=0.15)
    lv = isis.synth_level(t, metric_type=mtype, mt_id=mt, metric_mode=mode, max_reach_per_fragment=3,
                          overload=(5, 11), no_protocols=(9,))
    f = isis.Flat(lv)
    for r in (0, 5, 9, 40):
        ref = pyoracle.isis_compute_spt(lv, isis.sysid(r))
        assert ref.rc == 0
        root = f.vertex(isis.sysid(r) << 8)
        c = pyoracle.csr_spf(f.csr, root, vec_mode=1)
        same_spt(f.spt_from_planes(root, c["dist"], c["hops"]), ref)


def test_stale_fragments_and_missing_zeroth():
    t = synth.random_topology(40, 170, synth.SEED_BASE + 4, cost_lo=1, cost_hi=20)
    lv = isis.synth_level(t, max_reach_per_fragment=2)
    lsps = lv.lsps.copy()
    lsps["seqno"][3] = 0                 # purged fragment
    lsps["rem_lifetime"][7] = 0          # expired fragment
    # a router with several fragments loses its zeroth LSP -> still a vertex, but a leaf
    multi = [int(x >> 8) - isis.SYSID_BASE for x in lsps["lan_id"][lsps["fragment"] == 2]]
    victim = multi[0]
    z = np.nonzero((lsps["lan_id"] == (isis.sysid(victim) << 8)) & (lsps["fragment"] == 0))[0][0]
    lsps["rem_lifetime"][z] = 0
    lv.lsps = lsps
    f = isis.Flat(lv)
    assert f.vertex(isis.sysid(victim) << 8) != 0xFFFFFFFF
    for r in (0 if victim != 0 else 1, victim):
        ref = pyoracle.isis_compute_spt(lv, isis.sysid(r))
        root = f.vertex(isis.sysid(r) << 8)
        c = pyoracle.csr_spf(f.csr, root, vec_mode=1)
        same_spt(f.spt_from_planes(root, c["dist"], c["hops"]), ref)


def test_narrow_metric_path_limit():
    # MAX_PATH_METRIC_STANDARD = 1023 rejects long narrow-metric paths (spf.rs:636-645)
    t = synth.random_topology(120, 260, synth.SEED_BASE + 6, cost_lo=50, cost_hi=63)
    lv = isis.synth_level(t, metric_type=isis.METRIC_STANDARD)
    f = isis.Flat(lv)
    assert f.csr.reject_above == 1023
    ref = pyoracle.isis_compute_spt(lv, isis.sysid(0))
    root = f.vertex(isis.sysid(0) << 8)
    c = pyoracle.csr_spf(f.csr, root, vec_mode=1)
    same_spt(f.spt_from_planes(root, c["dist"], c["hops"]), ref)


# ---- route stage on the CPU: hspf_isis_routes_from_planes with the oracle's planes -----------
import golden_util as gu  # noqa: E402
from isis_synth import synth_instance  # noqa: E402

SNAPS = gu.load_isis()


def oracle_planes(csr, root):
    c = pyoracle.csr_spf(csr, root, vec_mode=1, nh_words=4)
    return c["dist"], c["hops"]


def same_rib(a, b):
    assert len(a.routes) == len(b.routes) and len(a.nexthops) == len(b.nexthops)
    assert a.routes.tobytes() == b.routes.tobytes()
    assert a.nexthops.tobytes() == b.nexthops.tobytes()


@pytest.mark.parametrize("snap", SNAPS, ids=[f"{s['topo']}-{s['rt']}" for s in SNAPS])
def test_route_stage_from_planes_on_reference_goldens(snap):
    """The product's route stage (local next hops + compute_routes, isis_host.cc) fed with the
    oracle's SPT planes: bit-identical to the oracle's route path and equal to the golden
    local-rib of every IS-IS conformance snapshot (CPU twin of the GPU test)."""
    from holo_b200 import ospfv3
    for level in snap["levels"]:
        inst = gu.isis_instance_image(snap, level)
        rib = isis.routes_from_planes(inst, oracle_planes)
        same_rib(rib, pyoracle.isis_compute_routes(inst))
        got = {f"{ospfv3.ip_str(r['prefix'])}/{int(r['len'])}":
               (int(r["metric"]), sorted((inst["ifnames"][i], a) for (i, a, _s) in rib.nh(r))) for r in rib.routes}
        for r in snap["local_rib"]:
            if r["level"] == level["level"]:
                assert got[r["prefix"]] == (r["metric"], sorted((a, b) for a, b in r["nexthops"]))


@pytest.mark.parametrize("seed,kw,root,mtype,frag", [
    (3, dict(cost_lo=1, cost_hi=30, lan_fraction=0.15), 0, isis.METRIC_WIDE, 3),
    (4, dict(cost_choices=[10], lan_fraction=0.2), 7, isis.METRIC_WIDE, 0),
    (5, dict(cost_lo=1, cost_hi=20), 11, isis.METRIC_BOTH, 2),
])
def test_route_stage_from_planes_synthetic(seed, kw, root, mtype, frag):
    t = synth.random_topology(300, 1300, synth.SEED_BASE + seed, **kw)
    inst = synth_instance(t, root, mtype, frag)
    same_rib(isis.routes_from_planes(inst, oracle_planes), pyoracle.isis_compute_routes(inst))


@pytest.mark.parametrize("seed,kw,root,mtype,frag", [
    (3, dict(cost_lo=1, cost_hi=30, lan_fraction=0.15), 0, isis.METRIC_WIDE, 3),
    (4, dict(cost_choices=[10], lan_fraction=0.2), 7, isis.METRIC_WIDE, 0),          # ECMP: merged next hops
    (5, dict(cost_lo=1, cost_hi=20), 11, isis.METRIC_BOTH, 2),
    (6, dict(cost_choices=[5, 10]), 4, isis.METRIC_WIDE, 2),
    (7, dict(cost_lo=1, cost_hi=9, lan_fraction=0.3), 12, isis.METRIC_WIDE, 1),
])
def test_route_stage_sr_prefix_sid_labels(seed, kw, root, mtype, frag):
    """IS-IS SR Prefix-SID labels (holo-isis/src/sr.rs:33-99): product route stage vs the restatement,
    byte for byte, plus the rules that can be read off the result."""
    t = synth.random_topology(200, 800, synth.SEED_BASE + seed, **kw)
    inst = synth_instance(t, root, mtype, frag, sr=True)
    got = isis.routes_from_planes(inst, oracle_planes)
    same_rib(got, pyoracle.isis_compute_routes(inst))
    assert int(got.routes["has_sr_label"].sum()) > 20 and int(got.nexthops["has_label"].sum()) > 20
    labels = got.nexthops["sr_label"][got.nexthops["has_label"] == 1]
    assert (labels == 3).any() and (labels >= 16000).any()            # implicit null and SRGB labels
    # sr disabled: no label anywhere, everything else unchanged
    inst_off = dict(inst, sr_enabled=0)
    off = isis.routes_from_planes(inst_off, oracle_planes)
    assert int(off.routes["has_sr_label"].sum()) == 0 and int(off.nexthops["has_label"].sum()) == 0
    assert np.array_equal(off.routes[["metric", "len", "route_type", "flags", "nh_off", "n_nh"]],
                          got.routes[["metric", "len", "route_type", "flags", "nh_off", "n_nh"]])
    # the local prefixes (hops 0) carry an input label only with P set and E clear
    from holo_b200 import ospfv3
    me = root
    own = {f"10.{(me >> 16) & 255}.{(me >> 8) & 255}.{me & 255}"}
    for r in got.routes:
        if ospfv3.ip_str(r["prefix"]) in own:
            kind = me % 7
            assert bool(r["has_sr_label"]) == (kind == 1)


AFTER = [(s, n) for s in SNAPS for n in s.get("after", {})]


@pytest.mark.parametrize("snap,name", AFTER, ids=[f"{n}-{s['topo']}-{s['rt']}" for s, n in AFTER])
def test_route_stage_on_step_after_states(snap, name):
    """The LSDBs the reference reached after its step tests (overload and ATT bits, expired LSPs,
    removed adjacencies ...): product route stage == oracle route path on each."""
    after = snap["after"][name]
    for level in after["levels"]:
        inst = gu.isis_instance_image(after, level)
        inst["att_ignore"] = int(bool(after.get("att_ignore", False)))
        same_rib(isis.routes_from_planes(inst, oracle_planes), pyoracle.isis_compute_routes(inst))
