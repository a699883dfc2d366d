"""IS-IS flooding reduction over the hop-count SPTs of the neighbour batch (SURVEY.md §8f f4;
holo-isis/src/flooding/manet.rs): the reference's own known-answer vectors for
flood_reduction_hash, and the product's Remote Neighbor List / reflood_list against the
restatement (oracle/isis_manet.cc) on synthetic levels.  CPU only: the SPTs are built by the
product's hspf_isis_spt_from_planes from the oracle's hop-count planes (on the GPU they come from
one hspf_run_batch with HSPF_GF_HOPCOUNT, tests/test_isis_gpu.py)."""
import numpy as np
import pytest

from holo_b200 import isis, synth
from oracle import pyoracle


def sysid_bytes(b):
    return int.from_bytes(bytes(b[:6]), "big")


# holo-isis/src/flooding/manet.rs:201-231 (draft-ietf-lsr-distoptflood-12 section 1.2.3)
KAT = [
    ([0x01, 0x02, 0x03, 0x04, 0x05, 0x06, 0x00, 0x00], 0x6215),
    ([0x01, 0x02, 0x03, 0x04, 0x05, 0x06, 0x00, 0x07], 0x6215),
    ([0x01, 0x02, 0x03, 0x04, 0x05, 0x06, 0x00, 0x0F], 0x6316),
    ([0x00, 0x01, 0x02, 0x03, 0x04, 0x05, 0x00, 0x01], 0x410F),
]


@pytest.mark.parametrize("lsp_id,expected", KAT)
def test_flood_reduction_hash_known_answers(lsp_id, expected):
    args = (sysid_bytes(lsp_id), lsp_id[6], lsp_id[7])
    assert isis.flood_reduction_hash(*args) == expected
    assert isis.flood_reduction_hash(*args, lib=pyoracle.lib(), name="oracle_isis_flood_reduction_hash") == expected


def test_flood_reduction_hash_product_equals_restatement():
    rng = np.random.default_rng(5)
    for _ in range(2000):
        args = (int(rng.integers(0, 1 << 48)), int(rng.integers(0, 256)), int(rng.integers(0, 256)))
        assert isis.flood_reduction_hash(*args) == \
            isis.flood_reduction_hash(*args, lib=pyoracle.lib(), name="oracle_isis_flood_reduction_hash")


@pytest.mark.parametrize("seed,kw", [
    (21, dict(cost_lo=1, cost_hi=30)),
    (22, dict(cost_choices=[10], lan_fraction=0.2)),
    (23, dict(cost_lo=1, cost_hi=5, lan_fraction=0.1)),
])
def test_remote_neighbors_and_reflood_list_match_restatement(seed, kw):
    R = 60
    t = synth.random_topology(R, 260, synth.SEED_BASE + seed, **kw)
    lv = isis.synth_level(t, mt_id=isis.MT_NONE, metric_mode=isis.MODE_HOPCOUNT, max_reach_per_fragment=3)
    rng = np.random.default_rng(seed)
    lsps = lv.lsps.copy()
    for i in range(len(lsps)):                        # flooding algorithm sub-TLVs on some LSPs
        if (int(lsps["lan_id"][i]) & 0xFF) == 0 and rng.random() < 0.8:
            lsps["flood_algo"][i] = int(rng.choice([1, 2, 2, 2, 9]))
    lv.lsps = lsps
    f = isis.Flat(lv)
    olib = pyoracle.lib()
    local = 0
    nbrs = sorted({int(t.p2p_b[k]) for k in range(t.n_p2p) if int(t.p2p_a[k]) == local} |
                  {int(t.p2p_a[k]) for k in range(t.n_p2p) if int(t.p2p_b[k]) == local} |
                  {m for members, _ in t.lans if local in members for m in members if m != local})
    assert nbrs
    n_nonempty = 0
    for tn in nbrs:                                   # one hop-count SPT per transmitting neighbour
        root = f.vertex(isis.sysid(tn) << 8)
        c = pyoracle.csr_spf(f.csr, root, vec_mode=1)
        spt = f.spt_from_planes(root, c["dist"], c["hops"])
        rnl = isis.remote_neighbors(lv, spt)
        ref = isis.remote_neighbors(lv, spt, lib=olib, name="oracle_isis_remote_neighbors")
        assert rnl.tobytes() == ref.tobytes() and len(rnl) > 0
        assert isis.sysid(local) in set(int(x) for x in rnl["system_id"])     # we are a neighbour of our neighbour
        for _ in range(40):
            org = int(rng.integers(0, R))
            lsp = (isis.sysid(org), int(rng.integers(0, 3)), int(rng.integers(0, 64)))
            got = isis.reflood_list(spt, rnl, isis.sysid(local), *lsp)
            want = isis.reflood_list(spt, rnl, isis.sysid(local), *lsp, lib=olib, name="oracle_isis_reflood_list")
            assert got == want
            n_nonempty += bool(got)
    assert n_nonempty > 0
