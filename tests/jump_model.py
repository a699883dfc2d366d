"""Executable model of phase 3J of spf_batch_kernel (holo_b200/csrc/spf_kernel.cuh): hops and
next-hop sets by pointer jumping over the first-parent tree, ECMP vertices as jump terminals,
monotone sweeps over the ECMP vertices, final OR with the top's set.

The model follows the kernel step by step (same cut rules, same seeds, same terminal test,
same sweep formula) but runs the rounds synchronously in numpy, so the ALGORITHM can be
fuzzed on the CPU over thousands of graphs against the reference-faithful oracle
(tests/test_jump_model.py).  It consumes what the kernel has at that point: the distance
plane and the parents pass outputs (first_parent, n_parents)."""
from __future__ import annotations

import numpy as np

VF_HOP = 1
GF_NOHOP_TARGET_NO_NEXTHOP = 1
INF = 0xFFFFFFFF


def jump_phase(csr, root: int, dist: np.ndarray, first_parent: np.ndarray, n_parents: np.ndarray):
    """Returns (hops u16[V], nh python-int bitsets [V], n_atoms, stats)."""
    V = csr.n_vertices
    row, col, cost = csr.row_ptr.astype(np.int64), csr.col.astype(np.int64), csr.cost.astype(np.int64)
    is_hop = (csr.vflags & VF_HOP) != 0
    nohop_rule = bool(csr.flags & GF_NOHOP_TARGET_NO_NEXTHOP)
    d = dist.astype(np.int64)
    reached = dist != INF
    NONE = -1
    fp = np.where(first_parent == INF, NONE, first_parent.astype(np.int64))

    # ---- root edge table: first-hop atom bases behind the root's non-HOP neighbours
    rb, re_ = int(row[root]), int(row[root + 1])
    table = []                       # (target, base, cost) in root-edge order
    nextbase = re_ - rb
    for e in range(rb, re_):
        h = int(col[e])
        if not is_hop[h]:
            table.append((h, nextbase, int(cost[e])))
            nextbase += int(row[h + 1] - row[h])
    n_atoms = nextbase
    h0 = np.zeros(V, bool)           # hops-0 vertices besides the root
    for (h, _b, c) in table:
        if d[h] == c:
            h0[h] = True
    hops0 = h0.copy()
    hops0[root] = True

    # ---- seeds: one "thread" per atom
    seed = [0] * V
    for atom in range(n_atoms):
        u, e = root, None
        if atom < re_ - rb:
            e = rb + atom
        else:
            for k, (N, nb, _c) in enumerate(table):
                if not (nb <= atom < nb + int(row[N + 1] - row[N])):
                    continue
                first = all(t[0] != N for t in table[:k])
                if first and hops0[N]:
                    u, e = N, int(row[N]) + (atom - nb)
                break
        if e is None:
            continue
        v, c = int(col[e]), int(cost[e])
        if reached[u] and reached[v] and c != INF and d[u] + c == d[v] and not (nohop_rule and not is_hop[v]):
            seed[v] |= 1 << atom

    # ---- J1: hops = sum of HOP flags over (root, v], pointer doubling
    anc = np.where((fp == NONE), np.arange(V), fp)
    agg = np.where(fp == NONE, 0, is_hop.astype(np.int64))
    rounds1 = 0
    while True:
        act = (anc != root) & (anc != np.arange(V))
        if not act.any():
            break
        a = anc[act]
        agg_new = agg.copy()
        anc_new = anc.copy()
        agg_new[act] = agg[act] + agg[a]
        anc_new[act] = anc[a]
        anc, agg = anc_new, agg_new
        rounds1 += 1
        assert rounds1 <= 64
    hops = np.where(anc == np.arange(V), 0, agg).astype(np.uint16)
    hops[root] = 0

    # ---- J2: segments up to the nearest terminal (root-ish cut or ECMP vertex)
    ecmp = n_parents >= 2
    cut = (fp == NONE) | hops0[np.where(fp == NONE, 0, fp)]
    top = np.where(cut, root, fp)
    acc = list(seed)
    rounds2 = 0
    while True:
        act = np.nonzero((top != root) & ~ecmp[top])[0]
        if len(act) == 0:
            break
        new_top = top.copy()
        new_acc = list(acc)
        for v in act:
            a = int(top[v])
            new_acc[v] = acc[v] | acc[a]
            new_top[v] = top[a]
        top, acc = new_top, new_acc
        rounds2 += 1
        assert rounds2 <= 64

    # ---- ECMP vertices among themselves
    elist = np.nonzero(ecmp)[0]
    # in-edges
    src = np.repeat(np.arange(V), np.diff(row))
    order = np.argsort(col, kind="stable")
    icol_sorted = col[order]
    istart = np.searchsorted(icol_sorted, np.arange(V + 1))
    sweeps = 0
    while True:
        changed = False
        for x in elist:
            need = acc[int(top[x])] if top[x] != root else 0
            for j in range(istart[x], istart[x + 1]):
                e = order[j]
                u = int(src[e])
                if not reached[u] or cost[e] == INF or d[u] + cost[e] != d[x] or hops0[u]:
                    continue
                need |= acc[u]
                if top[u] != root:
                    need |= acc[int(top[u])]
            if need & ~acc[x]:
                acc[x] |= need
                changed = True
        sweeps += 1
        if not changed:
            break
        assert sweeps <= 64
    nh = [acc[v] | (acc[int(top[v])] if top[v] != root else 0) for v in range(V)]
    return hops, nh, n_atoms, dict(rounds1=rounds1, rounds2=rounds2, sweeps=sweeps, n_ecmp=int(ecmp.sum()))
