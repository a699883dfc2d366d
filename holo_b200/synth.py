"""Seeded synthetic link-state topologies of the BASELINE.json shapes.

PRNG: SplitMix64 (counter based, vectorised), seed 0x484F4C4F + config index
(SURVEY.md §8d).  `random_topology` returns an abstract topology (routers,
point-to-point adjacencies with per-direction costs, broadcast LANs); the
protocol modules turn it into an OSPFv2 / IS-IS LSDB image, and `topology_csr`
flattens it directly for engine-level tests.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .capi import Csr, VF_HOP, GF_NOHOP_TARGET_NO_NEXTHOP

SEED_BASE = 0x484F4C4F
_M64 = (1 << 64) - 1


def splitmix64(seed: int, n: int, stream: int = 0) -> np.ndarray:
    """n 64-bit outputs of SplitMix64 started at `seed` (+ stream offset)."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = (np.uint64((seed + stream * 0x632BE59BD9B4E019) & _M64)
             + idx * np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


@dataclass
class Topology:
    n_routers: int
    # point-to-point adjacencies: endpoints a,b and per-direction costs
    p2p_a: np.ndarray
    p2p_b: np.ndarray
    p2p_cost_ab: np.ndarray
    p2p_cost_ba: np.ndarray
    # broadcast LANs: list of (member routers (first = DR), per-member cost)
    lans: list = field(default_factory=list)

    @property
    def n_p2p(self) -> int:
        return len(self.p2p_a)


def random_topology(n_routers: int, n_directed_edges: int, seed: int, cost_lo: int = 1,
                    cost_hi: int = 100, cost_choices=None, lan_fraction: float = 0.0,
                    lan_min: int = 3, lan_max: int = 8) -> Topology:
    """Random spanning tree + uniform extra links; E/2 bidirectional adjacencies.

    `lan_fraction` of the adjacencies are replaced by broadcast LANs of
    lan_min..lan_max routers (a LAN with m routers stands for m-1 adjacencies).
    """
    R = n_routers
    n_adj = n_directed_edges // 2
    if R > 1 and n_adj < R - 1:
        raise ValueError("not enough edges for a connected topology")
    r = splitmix64(seed, 4 * max(n_adj, 1) + 16, stream=0)
    # spanning tree: router i attaches to a uniformly chosen earlier router
    a = np.arange(1, R, dtype=np.int64)
    b = (r[: R - 1] % np.maximum(a.astype(np.uint64), np.uint64(1))).astype(np.int64)
    n_extra = n_adj - (R - 1)
    xa = (r[R: R + n_extra] % np.uint64(R)).astype(np.int64)
    xb = (r[R + n_extra: R + 2 * n_extra] % np.uint64(max(R - 1, 1))).astype(np.int64)
    xb = np.where(xb >= xa, xb + 1, xb)          # no self adjacency
    pa = np.concatenate([a, xa])
    pb = np.concatenate([b, xb])
    rc = splitmix64(seed, 2 * n_adj + 8, stream=1)
    if cost_choices is not None:
        ch = np.asarray(cost_choices, dtype=np.uint32)
        cab = ch[(rc[:n_adj] % np.uint64(len(ch))).astype(np.int64)]
        cba = ch[(rc[n_adj: 2 * n_adj] % np.uint64(len(ch))).astype(np.int64)]
    else:
        span = np.uint64(cost_hi - cost_lo + 1)
        cab = (rc[:n_adj] % span).astype(np.uint32) + np.uint32(cost_lo)
        cba = (rc[n_adj: 2 * n_adj] % span).astype(np.uint32) + np.uint32(cost_lo)
    lans = []
    if lan_fraction > 0 and n_extra > 0:
        # replace extra (non-tree) adjacencies by LANs so connectivity is kept
        rl = splitmix64(seed, 8 * n_adj + 64, stream=2)
        budget = int(lan_fraction * n_adj)
        keep = np.ones(n_adj, dtype=bool)
        k = 0
        cursor = n_adj - 1          # consume extra adjacencies from the end
        while budget > 0 and cursor >= R - 1:
            m = lan_min + int(rl[k] % np.uint64(lan_max - lan_min + 1)); k += 1
            members = []
            while len(members) < m:
                cand = int(rl[k] % np.uint64(R)); k += 1
                if cand not in members:
                    members.append(cand)
            if cost_choices is not None:
                costs = [int(ch[int(rl[k + i] % np.uint64(len(ch)))]) for i in range(m)]
            else:
                costs = [cost_lo + int(rl[k + i] % np.uint64(cost_hi - cost_lo + 1)) for i in range(m)]
            k += m
            lans.append((members, costs))
            take = min(m - 1, cursor - (R - 2))
            keep[cursor - take + 1: cursor + 1] = False
            cursor -= take
            budget -= (m - 1)
        pa, pb, cab, cba = pa[keep], pb[keep], cab[keep], cba[keep]
    return Topology(R, pa.astype(np.uint32), pb.astype(np.uint32), cab.astype(np.uint32),
                    cba.astype(np.uint32), lans)


def topology_csr(t: Topology, isis: bool = False, saturate_at: int = 0xFFFF,
                 reject_above: int = 0xFFFFFFFE, delta: int = 0) -> Csr:
    """Flatten a Topology straight to the engine CSR (engine-level tests only;
    protocol-level flatteners live in csrc/ and see real LSDB images).

    Vertex numbering: LAN (network / pseudonode) vertices 0..L-1, then routers
    L..L+R-1 — the reference's VertexId order (non-HOP vertices first).  A
    router's edges: its p2p links in adjacency order, then its LAN attachments.
    """
    L = len(t.lans)
    R = t.n_routers
    V = L + R
    src = [t.p2p_a.astype(np.int64) + L, t.p2p_b.astype(np.int64) + L]
    dst = [t.p2p_b.astype(np.int64) + L, t.p2p_a.astype(np.int64) + L]
    cst = [t.p2p_cost_ab.astype(np.int64), t.p2p_cost_ba.astype(np.int64)]
    order = [np.arange(t.n_p2p, dtype=np.int64) * 2, np.arange(t.n_p2p, dtype=np.int64) * 2 + 1]
    base = 2 * t.n_p2p
    for li, (members, costs) in enumerate(t.lans):
        m = np.asarray(members, dtype=np.int64) + L
        c = np.asarray(costs, dtype=np.int64)
        # router -> LAN (cost), LAN -> router (0, attached routers ascending)
        src += [m, np.full(len(m), li, dtype=np.int64)]
        dst += [np.full(len(m), li, dtype=np.int64), np.sort(m)]
        cst += [c, np.zeros(len(m), dtype=np.int64)]
        order += [base + np.arange(len(m)), base + len(m) + np.arange(len(m))]
        base += 2 * len(m)
    src = np.concatenate(src); dst = np.concatenate(dst)
    cst = np.concatenate(cst); order = np.concatenate(order)
    k = np.lexsort((order, src))
    src, dst, cst = src[k], dst[k], cst[k]
    row = np.zeros(V + 1, dtype=np.uint32)
    np.add.at(row, src + 1, 1)
    row = np.cumsum(row, dtype=np.uint64).astype(np.uint32)
    vflags = np.zeros(V, dtype=np.uint8)
    vflags[L:] = VF_HOP
    return Csr(row, dst.astype(np.uint32), cst.astype(np.uint32), vflags,
               reject_above=reject_above, saturate_at=0 if isis else saturate_at,
               flags=GF_NOHOP_TARGET_NO_NEXTHOP if isis else 0, delta=delta)
