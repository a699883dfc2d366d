"""Synthetic IS-IS instance (level LSDB + local interfaces / adjacencies of one router) shared by
the GPU route-stage tests and their CPU twins."""
import numpy as np

from holo_b200 import isis


def synth_instance(t, root, metric_type=isis.METRIC_WIDE, frag=0):
    """Synthetic level-2 instance with local interfaces/adjacencies of router `root`."""
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
            ipr.append((ospfv3.ip_rec(f"10.{(r >> 16) & 255}.{(r >> 8) & 255}.{r & 255}"), 0, 0, 32, isis.IP_V4_EXT, 0, (0, 0, 0)))
            ipr.append((ospfv3.ip_rec(f"10.255.{(r >> 8) & 255}.{r & 255}"), 5, 0, 32, isis.IP_V4_EXT, 0, (0, 0, 0)))
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
                ifaces=np.asarray(ifaces, dtype=isis.IFACE_DT) if ifaces else np.zeros(0, isis.IFACE_DT),
                adjs=np.asarray(adjs, dtype=isis.ADJ_DT) if adjs else np.zeros(0, isis.ADJ_DT))
