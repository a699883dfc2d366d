"""GPU parity of the IS-IS path (hspf_isis_compute_spt and the batched
flatten -> hspf_run_batch -> hspf_isis_spt_from_planes route) against the
line-faithful oracle of holo-isis compute_spt."""
import numpy as np
import pytest

from holo_b200 import isis, synth
from holo_b200.capi import COST_DISABLED
from oracle import pyoracle
from test_isis_cpu import MODES, same_spt

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mt,mtype,mode", MODES)
def test_compute_spt_matches_oracle(ctx, mt, mtype, mode):
    t = synth.random_topology(150, 640, synth.SEED_BASE + 3, cost_lo=1, cost_hi=40, lan_fraction=0.12)
    lv = isis.synth_level(t, metric_type=mtype, mt_id=mt, metric_mode=mode, max_reach_per_fragment=4,
                          overload=(5, 11), no_protocols=(9,))
    for r in (0, 5, 9, 77):
        same_spt(isis.compute_spt(ctx, lv, isis.sysid(r)), pyoracle.isis_compute_spt(lv, isis.sysid(r)))


def test_manet_style_multi_root_hopcount(ctx):
    """flooding::manet::init_cache: one hop-count SPT per adjacency of a router
    (flooding/manet.rs:47-69), batched as one launch."""
    t = synth.random_topology(400, 1700, synth.SEED_BASE + 8, lan_fraction=0.1)
    lv = isis.synth_level(t, mt_id=isis.MT_NONE, metric_mode=isis.MODE_HOPCOUNT)
    f = isis.Flat(lv)
    g = ctx.upload(f.csr)
    local = 7
    nbrs = sorted({int(b) for a, b in zip(t.p2p_a, t.p2p_b) if a == local} | {int(a) for a, b in zip(t.p2p_a, t.p2p_b) if b == local})
    roots = np.asarray([f.vertex(isis.sysid(n) << 8) for n in nbrs], dtype=np.uint32)
    res = ctx.run(g, roots, nh_words=2)
    for j, n in enumerate(nbrs):
        ref = pyoracle.isis_compute_spt(lv, isis.sysid(n))
        same_spt(f.spt_from_planes(int(roots[j]), res.dist[j], res.hops[j]), ref)
        # the device planes themselves vs the abstract oracle (first_parent, n_parents, nh sets)
        c = pyoracle.csr_spf(f.csr, int(roots[j]), vec_mode=1, nh_words=2)
        for k in ("dist", "hops", "first_parent", "n_parents", "nh_mask"):
            assert np.array_equal(getattr(res, k)[j], c[k]), (n, k)
    g.free()


def test_c3_shape_perturbation_batch(ctx):
    """BASELINE config 3 (scaled for oracle time): same root, job j removes adjacency j."""
    t = synth.random_topology(2000, 8000, synth.SEED_BASE + 3, cost_lo=1, cost_hi=1000)
    lv = isis.synth_level(t)
    f = isis.Flat(lv)
    g = ctx.upload(f.csr)
    row, col = f.csr.row_ptr, f.csr.col
    root = f.vertex(isis.sysid(0) << 8)
    n_jobs = 64
    overrides, seen = [], {}
    for k in range(n_jobs):
        a = f.vertex(isis.sysid(int(t.p2p_a[k])) << 8)
        b = f.vertex(isis.sysid(int(t.p2p_b[k])) << 8)
        key = (min(a, b), max(a, b))
        nth = seen.get(key, 0)
        seen[key] = nth + 1
        e1 = [e for e in range(row[a], row[a + 1]) if col[e] == b][nth]
        e2 = [e for e in range(row[b], row[b + 1]) if col[e] == a][nth]
        overrides.append([(e1, COST_DISABLED), (e2, COST_DISABLED)])
    res = ctx.run(g, np.full(n_jobs, root, np.uint32), overrides=overrides)
    for j in range(0, n_jobs, 7):
        # reference semantics: the LSDB without that adjacency
        lv2 = isis.synth_level(synth.Topology(t.n_routers, np.delete(t.p2p_a, j), np.delete(t.p2p_b, j),
                                              np.delete(t.p2p_cost_ab, j), np.delete(t.p2p_cost_ba, j), t.lans))
        ref = pyoracle.isis_compute_spt(lv2, isis.sysid(0))
        got = f.spt_from_planes(root, res.dist[j], res.hops[j], overrides=overrides[j])
        same_spt(got, ref)
    g.free()


# ---- route path: compute_spt(local = true) + compute_routes ------------------------------
import golden_util as gu  # noqa: E402

SNAPS = gu.load_isis()


def same_rib(a, b):
    assert len(a.routes) == len(b.routes) and len(a.nexthops) == len(b.nexthops)
    assert a.routes.tobytes() == b.routes.tobytes()
    assert a.nexthops.tobytes() == b.nexthops.tobytes()


@pytest.mark.parametrize("snap", SNAPS, ids=[f"{s['topo']}-{s['rt']}" for s in SNAPS])
def test_route_stage_on_reference_goldens(ctx, snap):
    """Every IS-IS conformance snapshot of the reference through the GPU route path:
    bit-identical to the oracle and equal to the golden local-rib."""
    from holo_b200 import ospfv3
    for level in snap["levels"]:
        inst = gu.isis_instance_image(snap, level)
        rib = isis.compute_routes(ctx, inst)
        same_rib(rib, pyoracle.isis_compute_routes(inst))
        got = {f"{ospfv3.ip_str(r['prefix'])}/{int(r['len'])}":
               (int(r["metric"]), sorted((inst["ifnames"][i], a) for (i, a, _s) in rib.nh(r))) for r in rib.routes}
        for r in snap["local_rib"]:
            if r["level"] == level["level"]:
                assert got[r["prefix"]] == (r["metric"], sorted((a, b) for a, b in r["nexthops"]))


from isis_synth import synth_instance  # noqa: E402


@pytest.mark.parametrize("seed,kw,root,mtype,frag", [
    (3, dict(cost_lo=1, cost_hi=30, lan_fraction=0.15), 0, isis.METRIC_WIDE, 3),
    (4, dict(cost_choices=[10], lan_fraction=0.2), 7, isis.METRIC_WIDE, 0),          # ECMP + parallel adjacencies
    (5, dict(cost_lo=1, cost_hi=20), 11, isis.METRIC_BOTH, 2),
])
def test_route_stage_synthetic(ctx, seed, kw, root, mtype, frag):
    t = synth.random_topology(300, 1300, synth.SEED_BASE + seed, **kw)
    inst = synth_instance(t, root, mtype, frag)
    same_rib(isis.compute_routes(ctx, inst), pyoracle.isis_compute_routes(inst))


@pytest.mark.parametrize("seed,kw,root,frag", [
    (3, dict(cost_lo=1, cost_hi=30, lan_fraction=0.15), 0, 3),
    (4, dict(cost_choices=[10], lan_fraction=0.2), 7, 0),
])
def test_route_stage_sr_prefix_sid_labels(ctx, seed, kw, root, frag):
    """SR Prefix-SID labels through the GPU route path (CPU twin: tests/test_isis_cpu.py)."""
    t = synth.random_topology(200, 800, synth.SEED_BASE + seed, **kw)
    inst = synth_instance(t, root, isis.METRIC_WIDE, frag, sr=True)
    rib = isis.compute_routes(ctx, inst)
    same_rib(rib, pyoracle.isis_compute_routes(inst))
    assert int(rib.nexthops["has_label"].sum()) > 20
