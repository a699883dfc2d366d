"""BASELINE.json configs at their named sizes (GPU): C3 (IS-IS 10k-node, 10k perturbation
SPFs) and C4 (OSPFv3 multi-area, all-routers batch).  EVERY job of both batches is compared
plane by plane with the binary-heap oracle (native thread pool, oracle/batch_pool.cc); samples
are also checked against the reference-faithful oracle and through size-independent properties."""
import numpy as np
import pytest

from holo_b200 import isis, ospfv3, synth
from holo_b200.capi import COST_DISABLED, DIST_INF
from oracle import pyoracle

pytestmark = pytest.mark.gpu
PLANES = ["dist", "hops", "first_parent", "n_parents", "nh_mask"]


def edge_sources(csr):
    return np.repeat(np.arange(csr.n_vertices), np.diff(csr.row_ptr.astype(np.int64)))


def check_properties(csr, res, j, root, disabled=()):
    """SSSP fixpoint + tree consistency of one job (no oracle needed)."""
    d = res.dist[j].astype(np.int64)
    src = edge_sources(csr)
    cost = csr.cost.astype(np.int64).copy()
    big = np.int64(1) << 40
    cost[list(disabled)] = big
    reach = d != DIST_INF
    assert d[root] == 0
    ok = reach[src]
    assert (d[csr.col][ok] <= d[src][ok] + cost[ok]).all()          # no edge can improve a distance
    fp = res.first_parent[j]
    has = reach.copy(); has[root] = False
    assert (fp[has] != 0xFFFFFFFF).all() and (fp[~has] == 0xFFFFFFFF).all()
    hop = (csr.vflags & 1).astype(np.int64)
    assert ((d[fp[has]] < d[has]) | (hop[fp[has]] == 0)).all()      # parents are strictly closer (networks: <=)
    assert (d[fp[has]] <= d[has]).all()
    assert (res.hops[j][has].astype(np.int64) == res.hops[j][fp[has]].astype(np.int64) + hop[has]).all()
    assert (res.n_parents[j][has] >= 1).all()


def test_c3_isis_10k_nodes_10k_perturbation_jobs(ctx):
    t = synth.random_topology(10000, 40000, synth.SEED_BASE + 3, cost_lo=1, cost_hi=1000)
    lv = isis.synth_level(t)
    f = isis.Flat(lv)
    csr = f.csr
    assert csr.n_vertices == 10000 and csr.n_edges == 40000
    g = ctx.upload(csr)
    root = f.vertex(isis.sysid(0) << 8)
    row, col = csr.row_ptr, csr.col
    # adjacency k <-> its two directed CSR edges
    first = {}
    for u in range(csr.n_vertices):
        for e in range(row[u], row[u + 1]):
            first.setdefault((u, int(col[e])), []).append(e)
    seen, pair = {}, []
    for k in range(t.n_p2p):
        a = f.vertex(isis.sysid(int(t.p2p_a[k])) << 8)
        b = f.vertex(isis.sysid(int(t.p2p_b[k])) << 8)
        key = (min(a, b), max(a, b))
        nth = seen.get(key, 0)
        seen[key] = nth + 1
        pair.append((first[(a, b)][nth], first[(b, a)][nth]))
    n_jobs = 10000
    overrides = [[(pair[j % 20000][0], COST_DISABLED), (pair[j % 20000][1], COST_DISABLED)] for j in range(n_jobs)]
    base = ctx.run(g, np.asarray([root], np.uint32))
    chunk = 1000
    for c0 in range(0, n_jobs, chunk):
        ov = overrides[c0:c0 + chunk]
        roots = np.full(len(ov), root, np.uint32)
        res = ctx.run(g, roots, overrides=ov)
        assert (res.job_status == 0).all()
        ref = pyoracle.csr_batch(csr, roots, overrides=ov, mode="heap", vec_mode=1)
        assert ref["jobs_done"] == len(ov)
        for k in PLANES:                                  # every job, every plane, bit for bit
            assert np.array_equal(getattr(res, k), ref[k]), (c0, k)
        for j in (0, len(ov) // 2):
            check_properties(csr, res, j, root, disabled=pair[(c0 + j) % 20000])
        # a removed adjacency never shortens anything
        assert (res.dist.astype(np.int64) >= base.dist[0].astype(np.int64)).all()
        if c0 == 0:
            rf = pyoracle.csr_spf(csr, root, overrides=ov[3], vec_mode=1)      # the reference-faithful restatement
            for k in PLANES:
                assert np.array_equal(getattr(res, k)[3], rf[k]), k
    g.free()


def test_c4_ospfv3_multi_area_all_routers_batch(ctx):
    """25 areas x 2000 routers / 8000 directed links (50k routers, 200k links): every router
    of an area is an SPF root over that area's graph; all 25 areas, all 50 000 jobs."""
    n_areas, per = 25, 2000
    total_jobs = 0
    for k in range(0, n_areas):
        t = synth.random_topology(per, 8000, synth.SEED_BASE + 4 + 100 * k, cost_lo=1, cost_hi=100,
                                  lan_fraction=0.05 if k % 8 == 0 else 0.0)
        rids = ospfv3.RID_BASE + k * per + np.arange(per)
        area = ospfv3.synth_area(t, root=0, max_links_per_fragment=6, rids=rids, area_id=k)
        f = ospfv3.Flat(area)
        csr = f.csr
        g = ctx.upload(csr)
        roots = np.nonzero(f.is_router)[0].astype(np.uint32)
        assert len(roots) == per
        res = ctx.run(g, roots, nh_words=2)
        assert (res.job_status == 0).all()
        total_jobs += len(roots)
        ref = pyoracle.csr_batch(csr, roots, mode="heap", nh_words=2)
        for name in PLANES:                               # every job, every plane, bit for bit
            assert np.array_equal(getattr(res, name), ref[name]), (k, name)
        # the fast path (one next-hop word) on the same batch
        res1 = ctx.run(g, roots, nh_words=1)
        ok = res1.job_status == 0
        assert ok.sum() > per // 2
        assert np.array_equal(res1.dist, ref["dist"]) and np.array_equal(res1.hops, ref["hops"])
        assert np.array_equal(res1.first_parent, ref["first_parent"])
        assert np.array_equal(res1.nh_mask[ok][:, :, 0], ref["nh_mask"][ok][:, :, 0])
        for j in (0, per - 1):
            check_properties(csr, res, j, int(roots[j]))
        g.free()
        if k % 6 == 0:
            # and the LSDB-level call for the area's first router against the faithful oracle
            r1 = ospfv3.run_area(ctx, area)
            r2 = pyoracle.ospfv3_run_area(area)
            assert r1.vertices.tobytes() == r2.vertices.tobytes() and r1.routes.tobytes() == r2.routes.tobytes()
            assert r1.nexthops.tobytes() == r2.nexthops.tobytes()
    assert total_jobs == n_areas * per
