"""Synthetic IS-IS instance (level LSDB + local interfaces / adjacencies of one router) shared by
the GPU route-stage tests and their CPU twins."""
import numpy as np

from holo_b200 import isis


def synth_instance(t, root, metric_type=isis.METRIC_WIDE, frag=0, sr=False):
    """Synthetic level-2 instance with local interfaces/adjacencies of router `root`.  With sr=True
    the routers advertise SR capabilities (a few without the sub-TLV, without SPF in SR-Algorithm,
    without the IPv4 flag, or with two label blocks) and Prefix-SIDs of every flavour on their
    prefixes (index / absolute label, P and E flags, none)."""
    from holo_b200 import ospfv3
    lv = isis.synth_level(t, metric_type=metric_type, max_reach_per_fragment=frag)
    # IP reachability: one /32 per router on its zeroth fragment
    lsps = lv.lsps.copy()
    ipr = []
    for i in range(len(lsps)):
        lid = int(lsps["lan_id"][i])
        if lsps["fragment"][i] == 0 and (lid & 0xFF) == 0:
            r = (lid >> 8) - isis.SYSID_BASE
            lsps["ipreach_off"][i] = len(ipr)
            lsps["n_ipreach"][i] = 2
            p1 = p2 = None
            if sr:
                kind = r % 7
                p1 = [(0, 0, r), (isis.PSID_P, 0, r), (isis.PSID_P | isis.PSID_E, 0, r), (isis.PSID_E, 0, r),
                      (isis.PSID_V | isis.PSID_L, 1, 30000 + r), (isis.PSID_N, 0, 9000 + r), None][kind]
                p2 = (isis.PSID_P, 0, 4000 + r) if r % 3 == 0 else None      # index beyond the first block of most
            ipr.append(isis.ipreach_rec(ospfv3.ip_rec(f"10.{(r >> 16) & 255}.{(r >> 8) & 255}.{r & 255}"), 0, 0, 32, isis.IP_V4_EXT, 0, p1))
            ipr.append(isis.ipreach_rec(ospfv3.ip_rec(f"10.255.{(r >> 8) & 255}.{r & 255}"), 5, 0, 32, isis.IP_V4_EXT, 0, p2))
    srgbs = []
    if sr:
        for i in range(len(lsps)):
            lid = int(lsps["lan_id"][i])
            if (lid & 0xFF) != 0:
                continue
            r = (lid >> 8) - isis.SYSID_BASE
            # capabilities on the zeroth fragment, or (every 5th router) only on the next one
            carrier = 0 if r % 5 else 1
            if int(lsps["fragment"][i]) != carrier and not (carrier == 1 and int(lsps["fragment"][i]) == 0 and
                                                             not ((lsps["lan_id"] == lid) & (lsps["fragment"] == 1)).any()):
                continue
            if r % 11 == 10:
                continue                                       # no SR sub-TLVs at all
            fl = isis.LSP_SR_HAS_CAP | isis.LSP_SR_CAP_V
            if r % 13 != 12:
                fl |= isis.LSP_SR_CAP_I                         # a few do not do MPLS IPv4
            if r % 9 != 8:
                fl |= isis.LSP_SR_ALGO_SPF
            lsps["srgb_off"][i] = len(srgbs)
            if r % 4 == 0:
                srgbs += [(16000 + 10 * r, 4000, 0, (0, 0, 0)), (500000, 8000, 0, (0, 0, 0))]
                lsps["n_srgb"][i] = 2
            elif r % 4 == 1:
                srgbs += [(7, 100, 1, (0, 0, 0)), (20000, 8000, 0, (0, 0, 0))]     # an index-typed block is skipped
                lsps["n_srgb"][i] = 2
            else:
                srgbs += [(16000, 8000, 0, (0, 0, 0))]
                lsps["n_srgb"][i] = 1
            lsps["sr_flags"][i] = fl
    lv.srgbs = np.asarray(srgbs, dtype=isis.SRGB_DT) if srgbs else np.zeros(0, isis.SRGB_DT)
    lv.lsps = lsps
    arr = np.zeros(len(ipr), isis.IPREACH_DT)
    for k, x in enumerate(ipr):
        arr[k] = x
    lv.ipreaches = arr
    ifaces, adjs = [], []
    for k in range(t.n_p2p):
        a, b = int(t.p2p_a[k]), int(t.p2p_b[k])
        for me, other, cost in ((a, b, int(t.p2p_cost_ab[k])), (b, a, int(t.p2p_cost_ba[k]))):
            if me == root:
                n = len(adjs) + 1
                adjs.append((isis.sysid(other), (2, 0, 0, 1, n >> 8, n & 255), 1, 2, 1, 0, 1, 0, 0, (0, 0, 0),
                             0xAC100000 + 4 * k + (2 if me == a else 1), ospfv3.ip_rec("::")))
                ifaces.append((len(ifaces) + 1, cost, 0, (0, 0, 0), len(adjs) - 1, 1))
    for members, costs in t.lans:
        if root in members:
            off = len(adjs)
            for m in sorted(members):
                if m != root:
                    n = len(adjs) + 1
                    adjs.append((isis.sysid(m), (2, 0, 0, 2, n >> 8, n & 255), 1, 2, 1, 0, 1, 0, 0, (0, 0, 0),
                                 0xC0A80000 + 256 * len(ifaces) + m % 250, ospfv3.ip_rec("::")))
            ifaces.append((len(ifaces) + 1, costs[members.index(root)], 1, (0, 0, 0), off, len(adjs) - off))
    return dict(level=lv, system_id=isis.sysid(root), max_paths=4, level_no=2, level_type=2, att_ignore=0, mt_ipv6=0,
                sr_enabled=int(sr),
                ifaces=np.asarray(ifaces, dtype=isis.IFACE_DT) if ifaces else np.zeros(0, isis.IFACE_DT),
                adjs=np.asarray(adjs, dtype=isis.ADJ_DT) if adjs else np.zeros(0, isis.ADJ_DT))
