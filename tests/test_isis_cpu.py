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
