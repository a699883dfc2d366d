"""Helpers turning tests/golden/*.json snapshots (extracted from the reference's
conformance topologies by tests/golden/make_golden.py) into LSDB images."""
from __future__ import annotations

import ipaddress
import json
from pathlib import Path

import numpy as np

from holo_b200 import ospfv2

GOLDEN = Path(__file__).resolve().parent / "golden"


def ip(s: str) -> int:
    return int(ipaddress.IPv4Address(s))


def ipstr(v: int) -> str:
    return str(ipaddress.IPv4Address(int(v)))


def load_ospfv2():
    return json.loads((GOLDEN / "ospfv2.json").read_text())


def global_sort_keys(snap):
    """One sort key per interface of the instance (hl_ospf_iface.sort_key must be unique per
    instance for the multi-area table): rank in byte-wise name order over all areas."""
    names = sorted({i["name"] for a in snap["areas"] for i in a["interfaces"]}, key=lambda n: n.encode())
    return {n: k + 1 for k, n in enumerate(names)}


def ospfv2_area_image(snap, area, sort_keys=None):
    """hl_ospfv2_area for one area of a golden snapshot."""
    rl = sorted(area["router_lsas"], key=lambda l: (ip(l["adv"]), ip(l["id"])))
    nlinks = sum(len(l["links"]) for l in rl)
    links = np.zeros(nlinks, ospfv2.LINK_DT)
    rlsa = np.zeros(len(rl), ospfv2.ROUTER_LSA_DT)
    off = 0
    for i, l in enumerate(rl):
        rlsa[i] = (ip(l["adv"]), ip(l["id"]), ospfv2.MAX_AGE if l.get("maxage") else 1, l["flags"], 0x02, off, len(l["links"]))
        for (ty, lid, ld, m) in l["links"]:
            links[off] = (ip(lid), ip(ld), m, ty, 0)
            off += 1
    nl = sorted(area["network_lsas"], key=lambda l: (ip(l["adv"]), ip(l["id"])))
    nlsa = np.zeros(len(nl), ospfv2.NETWORK_LSA_DT)
    att = []
    for i, l in enumerate(nl):
        a = sorted(ip(x) for x in l["attached"])
        nlsa[i] = (ip(l["adv"]), ip(l["id"]), ip(l["mask"]), ospfv2.MAX_AGE if l.get("maxage") else 1, 0, len(att), len(a))
        att += a
    # interfaces in name order (BTreeMap<String, _>: byte-wise string order)
    ifs = sorted(area["interfaces"], key=lambda i: i["name"].encode())
    ifaces = np.zeros(len(ifs), ospfv2.IFACE_DT)
    nbrs, names = [], []
    for i, f in enumerate(ifs):
        if f["cfg_type"] == "virtual-link":
            ty = ospfv2.IF_VLINK
        elif f["state"] == "loopback":
            ty = ospfv2.IF_LOOPBACK
        elif f["cfg_type"] == "point-to-point" or f["state"] == "point-to-point":
            ty = ospfv2.IF_P2P
        elif f["cfg_type"] == "point-to-multipoint":
            ty = ospfv2.IF_P2MP
        else:
            ty = ospfv2.IF_BROADCAST
        nb = sorted((ip(r), ip(a)) for r, a in f["neighbors"])
        sk = sort_keys[f["name"]] if sort_keys else i + 1
        ifaces[i] = (snap["ifindex"].get(f["name"], 0), sk, ty, (0, 0, 0), 0, 0, len(nbrs), len(nb))
        nbrs += nb
        names.append(f["name"])
    img = ospfv2.Ospfv2Area(router_id=ip(snap["router_id"]), area_id=ip(area["area_id"]))
    img.router_lsas, img.links, img.network_lsas = rlsa, links, nlsa
    img.attached = np.asarray(att, dtype=np.uint32)
    img.ifaces = ifaces
    img.nbrs = np.asarray(nbrs, dtype=ospfv2.NBR_DT) if nbrs else np.zeros(0, ospfv2.NBR_DT)
    img.ifnames = names
    return img


def routes_as_dict(res, names):
    """{prefix_str: (metric, sorted [(ifname, addr_str|None)])} of an Ospfv2Result."""
    out = {}
    for r in res.routes:
        plen = bin(int(r["mask"])).count("1")
        key = f"{ipstr(r['prefix'])}/{plen}"
        nh = sorted((names[i], ipstr(a) if ha else None) for (i, ha, a, _hn, _n, _hl, _l) in res.nh(r))
        out[key] = (int(r["metric"]), nh)
    return out


def merge_area_routes(dicts):
    """RIB merge across areas for intra-area routes (route_update, route.rs:895-942)."""
    rib = {}
    for d in dicts:
        for k, (m, nh) in d.items():
            if k not in rib or m < rib[k][0]:
                rib[k] = (m, list(nh))
            elif m == rib[k][0]:
                rib[k] = (m, sorted(set(rib[k][1]) | set(nh), key=lambda x: (x[0], x[1] or "")))
    return rib


def ospfv2_summaries(area):
    """hl_ospfv2_summary_lsa[] of one area of a golden snapshot, LsaKey order (type, adv_rtr, lsa_id)."""
    from holo_b200 import ospf_rib
    ls = sorted(area.get("summary_lsas", []), key=lambda l: (l["type"], ip(l["adv"]), ip(l["id"])))
    out = np.zeros(len(ls), ospf_rib.SUMMARY_LSA_DT)
    for i, l in enumerate(ls):
        out[i] = (ip(l["adv"]), ip(l["id"]), ip(l["mask"]), l["metric"], l["type"], int(bool(l.get("maxage"))), (0, 0))
    return out


def ospfv2_full_rib(snap, run_area, update_rib_full):
    """The whole OSPFv2 routing table of a golden snapshot: `run_area(img)` per attached area, then
    `update_rib_full(router_id, max_paths, [RibArea...])`.  Returns {prefix: (metric, type, [(ifname, addr)])}
    so that it compares directly with `golden_rib`."""
    from holo_b200 import ospf_rib
    keys = global_sort_keys(snap)
    key_name = {v: k for k, v in keys.items()}
    areas = []
    for area in snap["areas"]:
        img = ospfv2_area_image(snap, area, keys)
        res = run_area(img)
        assert res.rc == 0
        if not res.root_found:
            continue
        active = any((i.get("state") or "down") != "down" for i in area["interfaces"])
        areas.append(ospf_rib.RibArea(ip(area["area_id"]), res, img.ifaces, ospfv2_summaries(area), active))
    rib = update_rib_full(ip(snap["router_id"]), 16, areas)
    assert rib.rc == 0
    out = {}
    for r in rib.routes:
        plen = bin(int(r["mask"])).count("1")
        nh = sorted(((key_name.get(i, "?"), ipstr(a) if ha else None) for (i, ha, a, _hn, _n, _hl, _l) in rib.nh(r)),
                    key=lambda x: (x[0] or "", x[1] or ""))
        out[f"{ipstr(r['prefix'])}/{plen}"] = (int(r["metric"]), ospf_rib.PATH_NAMES[int(r["path_type"])], nh)
    return out


def golden_rib(snap):
    """Every route of the reference's local-rib: {prefix: (metric, route-type, [(ifname, addr)])}."""
    out = {}
    for r in snap["local_rib"]:
        nh = sorted(((n[0], n[1]) for n in r["nexthops"]), key=lambda x: (x[0] or "", x[1] or ""))
        out[r["prefix"]] = (r["metric"], r["type"], nh)
    return out


def golden_intra(snap):
    out = {}
    for r in snap["local_rib"]:
        if r["type"] != "intra-area":
            continue
        nh = sorted(((n[0], n[1]) for n in r["nexthops"]), key=lambda x: (x[0] or "", x[1] or ""))
        out[r["prefix"]] = (r["metric"], nh)
    return out


# ------------------------------------------------------------------------------ OSPFv3
def load_ospfv3():
    return json.loads((GOLDEN / "ospfv3.json").read_text())


def ospfv3_area_image(snap, area, sort_keys=None):
    """hl_ospfv3_area for one area of a golden OSPFv3 snapshot."""
    from holo_b200 import ospfv3
    rl = sorted(area["router_lsas"], key=lambda l: (ip(l["adv"]), l["id"]))
    links, rlsa = [], []
    for l in rl:
        rlsa.append((ip(l["adv"]), l["id"], 1, l["flags"], (ospfv3.OPT_R if l["r"] else 0) | (ospfv3.OPT_V6 if l["v6"] else 0),
                     len(links), len(l["links"])))
        links += [(ifid, nifid, ip(nrid), metric, ty, 0) for (ty, ifid, nifid, nrid, metric) in l["links"]]
    nl = sorted(area["network_lsas"], key=lambda l: (ip(l["adv"]), l["id"]))
    nlsa, att = [], []
    for l in nl:
        a = sorted(ip(x) for x in l["attached"])
        nlsa.append((ip(l["adv"]), l["id"], 1, 0, len(att), len(a)))
        att += a
    il = sorted(area["iap_lsas"], key=lambda l: (ip(l["adv"]), l["id"]))
    iaps, prefixes = [], []
    for l in il:
        ref = {"ospfv3-router-lsa": ospfv3.REF_ROUTER, "ospfv3-network-lsa": ospfv3.REF_NETWORK}.get(l["ref_type"], 0)
        off = len(prefixes)
        for (p, metric, opts) in l["prefixes"]:
            net = ipaddress.ip_network(p, strict=False)
            prefixes.append((ospfv3.ip_rec(net.network_address), net.prefixlen,
                             ospfv3.PFX_NU if "nu-bit" in opts else 0, metric))
        iaps.append((ip(l["adv"]), l["id"], 1, ref, 0, l["ref_id"], ip(l["ref_adv"]), off, len(prefixes) - off))
    ifs = sorted(area["interfaces"], key=lambda i: i["name"].encode())
    ifaces, llsas, names = [], [], []
    for i, f in enumerate(ifs):
        if f["state"] == "virtual-link":
            ty = 4
        elif f["state"] == "loopback":
            ty = 5
        elif f["state"] == "point-to-point":
            ty = 0
        else:
            ty = 1
        ifaces.append((f["interface_id"] or 0, sort_keys[f["name"]] if sort_keys else i + 1, ty, (0, 0, 0)))
        names.append(f["name"])
        for (adv, lsid, ll) in f["link_lsas"]:
            llsas.append((i, ip(adv), lsid, 1, 0, ospfv3.ip_rec(ll)))
    img = ospfv3.Ospfv3Area(router_id=ip(snap["router_id"]), area_id=ip(area["area_id"]))
    mk = lambda rows, dt: np.asarray(rows, dtype=dt) if rows else np.zeros(0, dt)
    img.router_lsas, img.links = mk(rlsa, ospfv3.ROUTER_LSA_DT), mk(links, ospfv3.LINK_DT)
    img.network_lsas, img.attached = mk(nlsa, ospfv3.NETWORK_LSA_DT), np.asarray(att, dtype=np.uint32)
    img.ifaces = mk(ifaces, ospfv3.IFACE_DT)
    for name, rows, dt in (("iap_lsas", iaps, ospfv3.IAP_LSA_DT), ("prefixes", prefixes, ospfv3.PREFIX_DT),
                           ("link_lsas", llsas, ospfv3.LINK_LSA_DT)):
        arr = np.zeros(len(rows), dt)
        for k, r in enumerate(rows):
            arr[k] = r
        setattr(img, name, arr)
    img.ifnames = names
    return img


def ospfv3_inter_area_lsas(area):
    """hl_ospfv3_inter_area_lsa[] of one area of a golden OSPFv3 snapshot, LsaKey order."""
    from holo_b200 import ospf_rib, ospfv3
    ls = sorted(area.get("inter_area_lsas", []), key=lambda l: (l["type"], ip(l["adv"]), l["id"]))
    out = np.zeros(len(ls), ospf_rib.INTER_AREA_LSA_DT)
    for i, l in enumerate(ls):
        if l["type"] == 3:
            net = ipaddress.ip_network(l["prefix"], strict=False)
            out[i] = (ip(l["adv"]), l["id"], l["metric"], 0, ospfv3.ip_rec(net.network_address), net.prefixlen,
                      ospfv3.PFX_NU if "nu-bit" in l.get("options", []) else 0, 3, 0)
        else:
            out[i] = (ip(l["adv"]), l["id"], l["metric"], ip(l["router_id"]), ospfv3.ip_rec("::"), 0, 0, 4, 0)
    return out


def ospfv3_full_rib(snap, run_area, update_rib_full):
    """The whole OSPFv3 routing table of a golden snapshot (see ospfv2_full_rib)."""
    from holo_b200 import ospf_rib, ospfv3
    keys = global_sort_keys(snap)
    key_name = {v: k for k, v in keys.items()}
    areas = []
    for area in snap["areas"]:
        img = ospfv3_area_image(snap, area, keys)
        res = run_area(img)
        assert res.rc == 0
        if not res.root_found:
            continue
        active = any((i.get("state") or "down") != "down" for i in area["interfaces"])
        areas.append(ospf_rib.RibArea(ip(area["area_id"]), res, img.ifaces, ospfv3_inter_area_lsas(area), active))
    rib = update_rib_full(ip(snap["router_id"]), 16, areas)
    assert rib.rc == 0
    out = {}
    for r in rib.routes:
        hops = rib.nexthops[int(r["nh_off"]): int(r["nh_off"]) + int(r["n_nh"])]
        nh = sorted(((key_name.get(int(x["iface"]), "?"), ospfv3.ip_str(x["addr"]) if x["has_addr"] else None) for x in hops),
                    key=lambda x: (x[0] or "", x[1] or ""))
        out[f"{ospfv3.ip_str(r['prefix'])}/{int(r['len'])}"] = (int(r["metric"]), ospf_rib.PATH_NAMES[int(r["path_type"])], nh)
    return out


def routes6_as_dict(res, names):
    from holo_b200 import ospfv3
    out = {}
    for r in res.routes:
        key = f"{ospfv3.ip_str(r['prefix'])}/{int(r['len'])}"
        nh = sorted((names[i], a) for (i, a, _n) in res.nh(r))
        out[key] = (int(r["metric"]), nh)
    return out


# ------------------------------------------------------------------------------ IS-IS
def load_isis():
    return json.loads((GOLDEN / "isis.json").read_text())


def lan_id(s: str) -> int:
    """'0000.0000.0003.01' -> (sysid << 8) | pseudonode"""
    parts = s.split(".")
    sysid = int("".join(parts[:3]), 16)
    return (sysid << 8) | int(parts[3], 16)


def isis_level_image(snap, level, mt_id=None):
    from holo_b200 import isis
    mtype = {"old-only": isis.METRIC_STANDARD, "wide-only": isis.METRIC_WIDE, "both": isis.METRIC_BOTH}[snap["metric_type"]]
    lsps, reaches = [], []
    for l in level["lsps"]:
        lid, frag = l["id"].split("-")
        flags = 0
        if "lsp-overload-flag" in l["flags"]:
            flags |= isis.LSPF_OL
        if l["protocols"] is not None:
            flags |= isis.LSPF_HAS_PROTOCOLS
            if 204 in l["protocols"]:
                flags |= isis.LSPF_NLPID_IPV4
            if 142 in l["protocols"]:
                flags |= isis.LSPF_NLPID_IPV6
        rr = [(lan_id(n), m, 0, isis.REACH_LEGACY, 0) for (n, m, _mt) in l["is"]]
        rr += [(lan_id(n), m, 0, isis.REACH_EXT, 0) for (n, m, _mt) in l["ext_is"]]
        rr += [(lan_id(n), m, 2 if mt is None else mt, isis.REACH_MT, 0) for (n, m, mt) in l["mt_is"]]
        lsps.append(isis.lsp_rec(lan_id(lid), 1, 1200, int(frag, 16), flags, len(reaches), len(rr)))
        reaches += rr
    lv = isis.IsisLevel(metric_type=mtype, mt_id=isis.MT_STANDARD if mt_id is None else mt_id,
                        ipv4_enabled="ipv4" in snap["afs"], ipv6_enabled="ipv6" in snap["afs"])
    la = np.zeros(len(lsps), isis.LSP_DT)
    for i, x in enumerate(lsps):
        la[i] = x
    lv.lsps = la[np.lexsort((la["fragment"], la["lan_id"]))]
    ra = np.zeros(len(reaches), isis.REACH_DT)
    for i, x in enumerate(reaches):
        ra[i] = x
    # reach_off refers to the unsorted reach array, which is fine (offsets travel with the LSP)
    lv.reaches = ra
    return lv


def isis_expected_ipv4_routes(snap, level, spt):
    """IPv4 routes implied by an SPT: metric = dist + prefix metric (route.rs:97), next hops
    = adjacencies of the next-hop systems of the best vertices (route.rs:105-139)."""
    mtype = snap["metric_type"]
    by_lan = {}
    for l in level["lsps"]:
        by_lan.setdefault(lan_id(l["id"].split("-")[0]), []).append(l)
    adj = {}
    for a in snap["adjacencies"]:
        if a["state"] == "up" and a["ipv4"]:
            adj.setdefault(int(a["sysid"].replace(".", ""), 16), []).append((a["iface"], a["ipv4"][0]))
    best = {}
    for v in spt.vertices:
        nh = set()
        for sid in spt.nexthops[int(v["nh_off"]): int(v["nh_off"]) + int(v["n_nh"])]:
            for x in adj.get(int(sid), []):
                nh.add(x)
        for l in by_lan.get(int(v["lan_id"]), []):
            pf = []
            if mtype in ("old-only", "both"):
                pf += l["ipv4_int"] + l["ipv4_ext"]
            if mtype in ("wide-only", "both"):
                pf += l["ext_ipv4"]
            for (p, m, _mt) in pf:
                tot = int(v["distance"]) + m
                cur = best.get(p)
                if cur is None or tot < cur[0]:
                    best[p] = (tot, set(nh))
                elif tot == cur[0]:
                    cur[1].update(nh)
    return best


def isis_instance_image(snap, level):
    """hl_isis_instance (LSDB with IP reachability + local interfaces/adjacencies) for one
    level of a golden IS-IS snapshot.  SNPAs are synthesised (the fixtures do not export
    them): one distinct MAC per adjacency, as on a real network."""
    import ipaddress
    from holo_b200 import isis, ospfv3
    lv = isis_level_image(snap, level)
    # IP reachability per fragment, aligned with lv.lsps order
    by_id = {}
    for l in level["lsps"]:
        lid, frag = l["id"].split("-")
        by_id[(lan_id(lid), int(frag, 16))] = l
    ipr = []
    lsps = lv.lsps.copy()
    for i in range(len(lsps)):
        l = by_id[(int(lsps["lan_id"][i]), int(lsps["fragment"][i]))]
        off = len(ipr)
        def add(items, kind, ext=0):
            for (p, m, mt) in items:
                net = ipaddress.ip_network(p, strict=False)
                ipr.append(isis.ipreach_rec(ospfv3.ip_rec(net.network_address), m,
                                            2 if (mt is None and kind == isis.IP_MT_V6) else (mt or 0),
                                            net.prefixlen, kind, ext))
        add(l["ipv4_int"], isis.IP_V4_INTERNAL)
        add(l["ipv4_ext"], isis.IP_V4_EXTERNAL, 1)
        add(l["ext_ipv4"], isis.IP_V4_EXT)
        add(l["ipv6"], isis.IP_V6)
        add(l["mt_ipv6"], isis.IP_MT_V6)
        lsps["ipreach_off"][i] = off
        lsps["n_ipreach"][i] = len(ipr) - off
        if "lsp-attached-default-metric-flag" in l["flags"]:
            lsps["flags"][i] |= isis.LSPF_ATT
        mtf = l.get("mt_flags", {}).get("2", [])
        if "tlv229-attached-flag" in mtf:
            lsps["flags"][i] |= isis.LSPF_MT_IPV6_ATT
        if "tlv229-overload-flag" in mtf:
            lsps["flags"][i] |= isis.LSPF_MT_IPV6_OL
    lv.lsps = lsps
    arr = np.zeros(len(ipr), isis.IPREACH_DT)
    for k, r in enumerate(ipr):
        arr[k] = r
    lv.ipreaches = arr
    lvl_no = level["level"]
    ifs = sorted(snap["interfaces"], key=lambda i: i["name"].encode())
    ifaces, adjs, names = [], [], []
    n_adj = 0
    for idx, f in enumerate(ifs):
        mine = [a for a in snap["adjacencies"] if a["iface"] == f["name"]]
        if f["type"] != "point-to-point":
            want = "level-1" if lvl_no == 1 else "level-2"
            mine = sorted([a for a in mine if a["usage"] in (want, "level-all")], key=lambda a: a["sysid"])
        off = len(adjs)
        for a in mine:
            n_adj += 1
            usage = {"level-1": 1, "level-2": 2, "level-all": 3}.get(a["usage"], 3)
            v6 = ospfv3.ip_rec(a["ipv6"][0]) if a["ipv6"] else ospfv3.ip_rec("::")
            adjs.append((int(a["sysid"].replace(".", ""), 16), (2, 0, 0, 0, n_adj >> 8, n_adj & 0xFF), int(a["state"] == "up"),
                         usage, int(0 in a["topologies"]), int(2 in a["topologies"]), int(bool(a["ipv4"])),
                         int(bool(a["ipv6"])), int(not (set(a["areas"]) & set(snap["areas"]))), (0, 0, 0),
                         ip(a["ipv4"][0]) if a["ipv4"] else 0, v6))
        ifaces.append((idx + 1, f["metric"], int(f["type"] != "point-to-point"), (0, 0, 0), off, len(adjs) - off))
        names.append(f["name"])
    inst = dict(level=lv, system_id=int(snap["system_id"].replace(".", ""), 16), max_paths=snap.get("max_paths", 16),
                level_no=lvl_no, level_type={"level-1": 1, "level-2": 2, "level-all": 3}[snap["level_type"]],
                att_ignore=0, mt_ipv6=int(snap["mt_ipv6"]),
                ifaces=np.asarray(ifaces, dtype=isis.IFACE_DT) if ifaces else np.zeros(0, isis.IFACE_DT),
                adjs=np.asarray(adjs, dtype=isis.ADJ_DT) if adjs else np.zeros(0, isis.ADJ_DT), ifnames=names)
    return inst


# ------------------------------------------------------------------------------ RIB manager stream
def installs_from_empty(snap, rib, diff, v3=False):
    """{prefix: (metric, sorted [(ifindex, addr)])} of the routes update_global_rib installs when the
    previous table is empty, with interfaces named by system ifindex like the RouteIpAdd messages
    of the reference (output/ibus.jsonl)."""
    from holo_b200 import ospf_rib, ospfv3
    keys = global_sort_keys(snap)
    key_name = {v: k for k, v in keys.items()}
    acts, routes = diff(None, rib)
    assert all(int(a["kind"]) == ospf_rib.RIB_INSTALL for a in acts)
    out = {}
    for a in acts:
        r = rib.routes[int(a["route"])]
        assert routes[int(a["route"])]["flags"] & ospf_rib.ROUTE_INSTALLED
        hops = rib.nexthops[int(r["nh_off"]): int(r["nh_off"]) + int(r["n_nh"])]
        if v3:
            pfx = f"{ospfv3.ip_str(r['prefix'])}/{int(r['len'])}"
            nh = sorted((snap["ifindex"].get(key_name[int(x["iface"])], 0), ospfv3.ip_str(x["addr"]) if x["has_addr"] else None)
                        for x in hops)
        else:
            pfx = f"{ipstr(r['prefix'])}/{bin(int(r['mask'])).count('1')}"
            nh = sorted((snap["ifindex"].get(key_name[int(x["iface"])], 0), ipstr(x["addr"]) if x["has_addr"] else None)
                        for x in hops)
        out[pfx] = (int(r["type2_metric"] if int(r["path_type"]) == 3 else r["metric"]), nh)
    # routes that are not installed: connected, or without next hops
    for i, r in enumerate(rib.routes):
        if not (routes[i]["flags"] & ospf_rib.ROUTE_INSTALLED):
            assert (int(r["flags"]) & ospf_rib.ROUTE_CONNECTED) or int(r["n_nh"]) == 0
    return out


def golden_ibus(snap):
    return {p: (v["metric"], sorted((n[0], n[1]) for n in v["nexthops"])) for p, v in snap["ibus_routes"].items()}
