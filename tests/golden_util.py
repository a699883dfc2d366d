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


def ospfv2_area_image(snap, area):
    """hl_ospfv2_area for one area of a golden snapshot."""
    rl = sorted(area["router_lsas"], key=lambda l: (ip(l["adv"]), ip(l["id"])))
    nlinks = sum(len(l["links"]) for l in rl)
    links = np.zeros(nlinks, ospfv2.LINK_DT)
    rlsa = np.zeros(len(rl), ospfv2.ROUTER_LSA_DT)
    off = 0
    for i, l in enumerate(rl):
        rlsa[i] = (ip(l["adv"]), ip(l["id"]), 1, l["flags"], 0x02, off, len(l["links"]))
        for (ty, lid, ld, m) in l["links"]:
            links[off] = (ip(lid), ip(ld), m, ty, 0)
            off += 1
    nl = sorted(area["network_lsas"], key=lambda l: (ip(l["adv"]), ip(l["id"])))
    nlsa = np.zeros(len(nl), ospfv2.NETWORK_LSA_DT)
    att = []
    for i, l in enumerate(nl):
        a = sorted(ip(x) for x in l["attached"])
        nlsa[i] = (ip(l["adv"]), ip(l["id"]), ip(l["mask"]), 1, 0, len(att), len(a))
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
        ifaces[i] = (snap["ifindex"].get(f["name"], 0), i + 1, ty, (0, 0, 0), 0, 0, len(nbrs), len(nb))
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


def golden_intra(snap):
    out = {}
    for r in snap["local_rib"]:
        if r["type"] != "intra-area":
            continue
        nh = sorted(((n[0], n[1]) for n in r["nexthops"]), key=lambda x: (x[0] or "", x[1] or ""))
        out[r["prefix"]] = (r["metric"], nh)
    return out
