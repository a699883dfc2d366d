"""CPU model of spf_quad_kernel's phases 1 and 2 over the quad-space image
(holo_b200/csrc/quad_layout.h, built by the product's host code).

TEST INFRASTRUCTURE.  It walks the same arrays the kernel reads, in the same roles
(slot-indexed distances, ring of four bucket bitmaps, chain spreading with the
continuation bits, stale-mark filter, four-record quads with pad records, in-quad
parents pass with the chain combine), sequentially.  It pins the layout builder and
the algorithm against the oracle without a GPU; the CUDA kernel itself is checked by
the `-m gpu` suites.
"""
from __future__ import annotations

import numpy as np

INF = 0xFFFFFFFF


def sssp(q, root: int, reject_above: int = 0xFFFFFFFE, qcap: int = 1 << 30):
    """Phase 1.  Returns the slot-indexed distance array (continuation slots hold copies)."""
    NQ, NBW, sh = q.NQ, q.NQ // 32, q.shift
    dist = np.full(NQ, INF, dtype=np.uint64)
    ring = np.zeros((4, NBW), dtype=np.uint64)
    rs = int(q.slot_of[root])
    dist[rs] = 0
    ring[0, rs >> 5] = 1 << (rs & 31)
    cur = empties = rounds = 0
    while True:
        bm = ring[cur & 3]
        queue = []
        pos = 0                               # the kernel's S.cnt: positions are handed out to every word
        for w in range(NBW):
            bits = int(bm[w])
            if not bits:
                continue
            C = int(q.fcont[w])
            allb, m = bits, bits
            while True:
                m = (m << 1) & C & 0xFFFFFFFF
                if not m:
                    break
                allb |= m
            n = bin(allb).count("1")
            pos += n
            if pos > qcap:
                continue                      # stays in the bitmap for the next round
            bm[w] = 0
            first = None
            for bit in range(32):
                if (allb >> bit) & 1:
                    qq = w * 32 + bit
                    if not (C >> bit) & 1:
                        first = qq
                    queue.append((qq, first))
        if not queue:
            empties += 1
            if empties == 4:
                break
            cur += 1
            continue
        empties = 0
        rounds += 1
        for qq, first in queue:
            du = int(dist[first])
            if du == INF or (du >> sh) != cur:
                continue
            for r in q.fq[qq]:
                hs, c = int(r) & 0xFFFF, int(r) >> 16
                nd = du + c
                if nd < int(dist[hs]) and nd <= reject_above:
                    dist[hs] = nd
                    b = nd >> sh
                    assert cur <= b <= cur + 3, (cur, b)
                    ring[b & 3, hs >> 5] = int(ring[b & 3, hs >> 5]) | (1 << (hs & 31))
    return dist, rounds


def parents(q, dist, root: int, V: int):
    """Phase 2.  Returns (dist[V], first_parent[V], n_parents[V])."""
    o_dist = np.full(V, INF, dtype=np.uint32)
    o_fp = np.full(V, INF, dtype=np.uint32)
    o_np = np.zeros(V, dtype=np.uint16)
    NIQ = q.NIQ
    part = []
    for i in range(NIQ):
        mx, my = int(q.imeta[i, 0]), int(q.imeta[i, 1])
        valid = mx != 0xFFFFFFFF
        sv = mx & 0xFFFF if valid else 0
        dv = int(dist[sv])
        cnt, bd, bs = 0, INF, INF
        for r in q.iq[i]:
            su, c = int(r) & 0xFFFF, int(r) >> 16
            du = int(dist[su])
            ok = du != INF and du + c == dv
            if ok:
                cnt += 1
                if du < bd or (du == bd and su < bs):
                    bd, bs = du, su
        part.append([cnt, bd, bs, valid, mx >> 16, my & 0xFF, (my >> 8) & 0xFF, dv])
    steps = 0
    while (1 << steps) < q.max_ichain:
        steps += 1
    d = 1
    for _ in range(steps):
        new = [p[:] for p in part]
        for i in range(NIQ):
            lane = i & 31
            if lane + d < 32 and d <= part[i][5]:
                o = part[i + d]
                new[i][0] = part[i][0] + o[0]
                if o[1] < part[i][1] or (o[1] == part[i][1] and o[2] < part[i][2]):
                    new[i][1], new[i][2] = o[1], o[2]
            elif d <= part[i][5]:
                raise AssertionError("in-quad chain straddles a warp")
        part = new
        d <<= 1
    for i in range(NIQ):
        cnt, bd, bs, valid, v, rem, pos, dv = part[i]
        if valid and pos == 0:
            if v == root or dv == INF:
                cnt, bs = 0, INF
            o_dist[v] = dv
            o_fp[v] = int(q.vert_of[bs]) if cnt else INF
            o_np[v] = min(cnt, 0xFFFF)
    return o_dist, o_fp, o_np


def check_image(q, csr):
    """Structural invariants of the image."""
    V, E = csr.n_vertices, csr.n_edges
    row, col, cost = csr.row_ptr, csr.col, csr.cost
    assert q.NQ % 32 == 0 and q.NIQ % 32 == 0 and q.NQ < 0xFFFF
    slot = q.slot_of.astype(np.int64)
    assert (np.diff(slot) > 0).all() if V > 1 else True            # slot order == vertex order
    for v in range(V):
        deg = int(row[v + 1] - row[v])
        nq = max(1, (deg + 3) // 4)
        s = int(slot[v])
        assert (s & 31) + nq <= 32                                  # chain inside one bitmap word
        for j in range(nq):
            assert q.vert_of[s + j] == v
            assert ((int(q.fcont[(s + j) >> 5]) >> ((s + j) & 31)) & 1) == (1 if j else 0)
        recs = q.fq[s:s + nq].reshape(-1)
        for i in range(deg):
            e = int(row[v]) + i
            assert recs[i] == (int(slot[col[e]]) | (int(cost[e]) << 16)), (v, i)
            assert q.fpos[e] == s * 4 + i
        for i in range(deg, nq * 4):
            assert recs[i] == (s | 0xFFFF0000)                      # pad: owner's slot, cost 65535
    # every forward edge appears exactly once among the in-quad records, at ipos
    seen = 0
    for e in range(E):
        ip = int(q.ipos[e])
        u = int(np.searchsorted(row, e, side="right") - 1)
        assert q.iq[ip >> 2, ip & 3] == (int(slot[u]) | (int(cost[e]) << 16))
        owner = int(q.imeta[ip >> 2, 0])
        assert owner >> 16 == col[e] and (owner & 0xFFFF) == slot[col[e]]
        seen += 1
    assert seen == E
