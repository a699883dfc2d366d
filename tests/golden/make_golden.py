#!/usr/bin/env python
"""Extract golden vectors for the SPF path from the reference's own conformance
fixtures (run in the build container; /root/reference does not exist on the GPU
box).  Source (read only):

  holo-ospf/tests/conformance/ospfv2/topologies/<topo>/<rt>/output/northbound-state.json
  holo-ospf/tests/conformance/ospfv2/topologies/<topo>/<rt>/events.jsonl (ifindex map)
  holo-isis/tests/conformance/topologies/<topo>/<rt>/{config.json,output/northbound-state.json}

Each snapshot holds BOTH the instance's converged LSDB (every Router/Network LSA,
or LSP, with links and metrics), the local interface/neighbour state, and the
`local-rib` the reference computed from it; so (LSDB, local state) -> local-rib is
a known-answer test of run_area/compute_spt + the intra-area route stage
(SURVEY.md §8c, Appendix A).  Output: tests/golden/ospfv2.json, tests/golden/isis.json.

Usage: python tests/golden/make_golden.py [/root/reference]
"""
from __future__ import annotations

import json
import re
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent

LINK_TYPES = {"point-to-point-link": 1, "transit-network-link": 2, "stub-network-link": 3, "virtual-link": 4}
RTR_BITS = {"abr-bit": 0x01, "asbr-bit": 0x02, "vlink-end-bit": 0x04}


def ospf_root(d):
    return d["ietf-routing:routing"]["control-plane-protocols"]["control-plane-protocol"][0]["ietf-ospf:ospf"]


def ifindex_map(events_path: Path):
    m = {}
    if not events_path.exists():
        return m
    for line in events_path.read_text().splitlines():
        for mm in re.finditer(r'"InterfaceUpd":\{"ifname":"([^"]+)","ifindex":(\d+)', line):
            m[mm.group(1)] = int(mm.group(2))
    return m


def ibus_routes(path: Path):
    """Final state of the RouteIpAdd / RouteIpDel stream the instance sent to the RIB manager
    (output/ibus.jsonl; produced by update_global_rib, holo-ospf/src/route.rs:833-893):
    {prefix: {"metric", "distance", "nexthops": [[ifindex, addr, [labels]]]}}, or None without the file."""
    if not path.exists():
        return None
    out = {}
    for line in path.read_text().splitlines():
        try:
            msg = json.loads(line)
        except json.JSONDecodeError:
            continue
        if "RouteIpAdd" in msg:
            r = msg["RouteIpAdd"]
            nh = [[n["Address"]["ifindex"], n["Address"].get("addr"), n["Address"].get("labels", [])]
                  for n in r.get("nexthops", []) if "Address" in n]
            out[r["prefix"]] = {"metric": r["metric"], "distance": r["distance"], "tag": r.get("tag"), "nexthops": nh}
        elif "RouteIpDel" in msg:
            out.pop(msg["RouteIpDel"]["prefix"], None)
    return out


# Step tests whose input is a pure recomputation parameter, so that their ibus output pins the
# re-computation + update_global_rib without emulating the protocol machines
# (holo-ospf/tests/conformance/ospfv2/mod.rs:494-499, holo-isis/tests/conformance/mod.rs:994-999).
STEP_TESTS = {
    "ospfv2": [("nb-config-iface-cost1", "topo1-1", "rt2", "02-output-ibus.jsonl",
                {"area": "0.0.0.1", "iface": "eth-rt1", "cost": 50})],
    "isis": [("nb-config-spf-paths1", "topo2-1", "rt1", "01-output-ibus.jsonl", {"max_paths": 1})],
}


def step_outputs(base: Path, proto: str, topo: str, rt: str):
    out = {}
    for (name, t, r, fn, change) in STEP_TESTS.get(proto, []):
        if (t, r) == (topo, rt):
            routes = ibus_routes(base / name / fn)
            if routes is not None:
                out[name] = {"change": change, "ibus_routes": routes}
    return out


# Step tests of the reference whose ibus output carries route messages and whose after-state dump
# is a complete snapshot (holo-ospf/tests/conformance/ospfv2/mod.rs: the cases named here): the
# table computed from the after-state, diffed against the table of the topology snapshot, must
# give the step's ibus output message for message.
OSPFV2_STEPS = [
    ("lsa-expiry1", "topo2-1", "rt2", "02"), ("lsa-expiry2", "topo2-1", "rt2", "02"),
    ("nb-config-area1", "topo1-1", "rt2", "01"), ("nb-config-enable1", "topo1-1", "rt3", "01"),
    ("nb-config-enable2", "topo1-1", "rt3", "01"), ("nb-config-iface1", "topo1-1", "rt2", "02"),
    ("nb-config-iface-cost1", "topo1-1", "rt2", "02"), ("nb-config-router-id1", "topo1-1", "rt3", "01"),
    ("nb-rpc-clear-neighbor1", "topo1-1", "rt6", "02"), ("nb-rpc-clear-neighbor2", "topo1-1", "rt6", "02"),
    ("ibus-addr-add3", "topo2-1", "rt6", "02"), ("ibus-addr-del1", "topo2-1", "rt2", "02"),
    ("ibus-iface-update1", "topo2-1", "rt2", "02"), ("ibus-iface-update2", "topo2-1", "rt2", "02"),
    ("ibus-iface-update3", "topo2-1", "rt3", "02"), ("ibus-iface-update4", "topo2-1", "rt2", "02"),
    ("ibus-iface-update6", "topo2-1", "rt2", "02"), ("timeout-nbr1", "topo1-2", "rt3", "02"),
    ("timeout-nbr2", "topo1-2", "rt3", "02"),
]


def ospfv2_snapshot(rt: Path, state_path: Path):
    """One OSPFv2 snapshot dict from a northbound-state dump, the router's config.json and events.jsonl."""
    o = ospf_root(json.loads(state_path.read_text()))
    cfg = json.loads((rt / "config.json").read_text())
    cfg_ospf = ospf_root(cfg)
    iftype_cfg = {}
    for a in cfg_ospf.get("areas", {}).get("area", []):
        for i in a.get("interfaces", {}).get("interface", []):
            iftype_cfg[(a["area-id"], i["name"])] = i.get("interface-type", "broadcast")
    snap = {"topo": rt.parent.name, "rt": rt.name, "router_id": o.get("router-id"),
            "ifindex": ifindex_map(rt / "events.jsonl"), "areas": [], "local_rib": []}
    for a in o.get("areas", {}).get("area", []):
        area = {"area_id": a["area-id"], "router_lsas": [], "network_lsas": [], "summary_lsas": [],
                "interfaces": []}
        for t in a.get("database", {}).get("area-scope-lsa-type", []):
            for l in t.get("area-scope-lsas", {}).get("area-scope-lsa", []):
                h = l["ospfv2"]["header"]
                b = l["ospfv2"].get("body", {})
                if t["lsa-type"] == 1 and "router" in b:
                    r = b["router"]
                    flags = 0
                    for bit in r.get("router-bits", {}).get("rtr-lsa-bits", []):
                        flags |= RTR_BITS.get(bit.split(":")[-1], 0)    # some dumps prefix "ietf-ospf:"
                    links = [[LINK_TYPES[x["type"]], x["link-id"], x["link-data"],
                              x["topologies"]["topology"][0]["metric"]]
                             for x in r.get("links", {}).get("link", [])]
                    area["router_lsas"].append({"adv": h["adv-router"], "id": h["lsa-id"], "flags": flags,
                                                "links": links, "maxage": "holo-ospf-dev:maxage" in h})
                elif t["lsa-type"] in (3, 4) and "summary" in b:
                    # Summary-LSAs (type 3 network / type 4 ASBR), inputs of the inter-area stage
                    # (holo-ospf/src/ospfv2/spf.rs:539-589)
                    sm = b["summary"]
                    area["summary_lsas"].append({
                        "type": t["lsa-type"], "adv": h["adv-router"], "id": h["lsa-id"],
                        "mask": sm["network-mask"], "metric": sm["topologies"]["topology"][0]["metric"],
                        "maxage": "holo-ospf-dev:maxage" in h})
                elif t["lsa-type"] == 2 and "network" in b:
                    n = b["network"]
                    area["network_lsas"].append({
                        "adv": h["adv-router"], "id": h["lsa-id"], "mask": n["network-mask"],
                        "attached": n.get("attached-routers", {}).get("attached-router", []),
                        "maxage": "holo-ospf-dev:maxage" in h})
        for i in a.get("interfaces", {}).get("interface", []):
            nb = [[x["neighbor-router-id"], x["address"]]
                  for x in i.get("neighbors", {}).get("neighbor", [])]
            area["interfaces"].append({"name": i["name"], "state": i.get("state"),
                                       "cfg_type": iftype_cfg.get((a["area-id"], i["name"]), "broadcast"),
                                       "neighbors": nb})
        for v in a.get("virtual-links", {}).get("virtual-link", []) if a.get("virtual-links") else []:
            nb = [[x["neighbor-router-id"], x["address"]]
                  for x in v.get("neighbors", {}).get("neighbor", [])]
            # holo names it vlink-<transit-area>-<router-id>
            # (holo-ospf/src/northbound/configuration.rs:606)
            area["interfaces"].append({"name": f"vlink-{v['transit-area-id']}-{v['router-id']}",
                                       "state": v.get("state"), "cfg_type": "virtual-link", "neighbors": nb})
        snap["areas"].append(area)
    # AS-external LSAs (type 5), inputs of update_rib_external (ospfv2/spf.rs:591-615)
    snap["external_lsas"] = []
    for t in o.get("database", {}).get("as-scope-lsa-type", []):
        for l in t.get("as-scope-lsas", {}).get("as-scope-lsa", []):
            h = l["ospfv2"]["header"]
            ex = l["ospfv2"].get("body", {}).get("external")
            if t["lsa-type"] == 5 and ex:
                tp = ex["topologies"]["topology"][0]
                snap["external_lsas"].append({
                    "adv": h["adv-router"], "id": h["lsa-id"], "mask": ex["network-mask"],
                    "e_bit": "flags" in tp and "E" in str(tp.get("flags")), "metric": tp["metric"],
                    "fwd": tp.get("forwarding-address", "0.0.0.0"), "tag": tp.get("external-route-tag", 0)})
    for r in o.get("local-rib", {}).get("route", []):
        nhs = [[n.get("outgoing-interface"), n.get("next-hop")]
               for n in r.get("next-hops", {}).get("next-hop", [])]
        snap["local_rib"].append({"prefix": r["prefix"], "metric": r.get("metric"),
                                  "type": r.get("route-type"), "nexthops": nhs})
    return snap


def extract_ospfv2(ref: Path):
    base = ref / "holo-ospf/tests/conformance/ospfv2/topologies"
    out = []
    for topo in sorted(p for p in base.iterdir() if p.is_dir()):
        for rt in sorted(p for p in topo.iterdir() if p.is_dir()):
            st = rt / "output" / "northbound-state.json"
            if not st.exists():
                continue
            snap = ospfv2_snapshot(rt, st)
            snap["ibus_routes"] = ibus_routes(rt / "output" / "ibus.jsonl")
            snap["steps"] = step_outputs(ref / "holo-ospf/tests/conformance/ospfv2", "ospfv2", topo.name, rt.name)
            snap["after"] = {}
            for (name, t, r, nn) in OSPFV2_STEPS:
                if (t, r) != (topo.name, rt.name):
                    continue
                from make_golden_isis import ibus_stream
                sd = ref / "holo-ospf/tests/conformance/ospfv2" / name
                after = ospfv2_snapshot(rt, sd / f"{nn}-output-northbound-state.json")
                after.pop("ifindex", None)          # the events of the topology replay name the interfaces
                after["ibus"] = ibus_stream(sd / f"{nn}-output-ibus.jsonl")
                snap["after"][name] = after
            out.append(snap)
    return out


V3_LINK_TYPES = {"point-to-point-link": 1, "transit-network-link": 2, "virtual-link": 4}
V3_BITS = {"abr-bit": 0x01, "asbr-bit": 0x02, "vlink-end-bit": 0x04}


def ospfv3_snapshot(rt: Path, state_path: Path):
    """One OSPFv3 snapshot dict from a northbound-state dump."""
    o = ospf_root(json.loads(state_path.read_text()))
    snap = {"topo": rt.parent.name, "rt": rt.name, "router_id": o.get("router-id"), "areas": [], "local_rib": []}
    for a in o.get("areas", {}).get("area", []):
        area = {"area_id": a["area-id"], "router_lsas": [], "network_lsas": [], "iap_lsas": [],
                "inter_area_lsas": [], "interfaces": []}
        for t in a.get("database", {}).get("area-scope-lsa-type", []):
            for l in t.get("area-scope-lsas", {}).get("area-scope-lsa", []):
                h = l["ospfv3"]["header"]
                b = l["ospfv3"].get("body", {})
                if "router" in b:
                    r = b["router"]
                    flags = 0
                    for bit in r.get("router-bits", {}).get("rtr-lsa-bits", []):
                        flags |= V3_BITS.get(bit.split(":")[-1], 0)
                    opts = r.get("lsa-options", {}).get("lsa-options", [])
                    links = [[V3_LINK_TYPES[x["type"]], x["interface-id"], x["neighbor-interface-id"],
                              x["neighbor-router-id"], x["metric"]] for x in r.get("links", {}).get("link", [])]
                    area["router_lsas"].append({"adv": h["adv-router"], "id": h["lsa-id"], "flags": flags,
                                                "r": "r-bit" in opts, "v6": "v6-bit" in opts, "links": links})
                elif "network" in b:
                    area["network_lsas"].append({"adv": h["adv-router"], "id": h["lsa-id"],
                                                 "attached": b["network"].get("attached-routers", {}).get("attached-router", [])})
                elif "inter-area-prefix" in b:
                    # inputs of the inter-area stage (holo-ospf/src/ospfv3/spf.rs:479-503)
                    p = b["inter-area-prefix"]
                    area["inter_area_lsas"].append({
                        "type": 3, "adv": h["adv-router"], "id": h["lsa-id"], "prefix": p["prefix"],
                        "metric": p.get("metric", 0),
                        "options": p.get("prefix-options", {}).get("prefix-options", [])})
                elif "inter-area-router" in b:
                    p = b["inter-area-router"]
                    area["inter_area_lsas"].append({
                        "type": 4, "adv": h["adv-router"], "id": h["lsa-id"],
                        "router_id": p.get("destination-router-id"), "metric": p.get("metric", 0)})
                elif "intra-area-prefix" in b:
                    p = b["intra-area-prefix"]
                    pf = [[x["prefix"], x.get("metric", 0), x.get("prefix-options", {}).get("prefix-options", [])]
                          for x in p.get("prefixes", {}).get("prefix", [])]
                    area["iap_lsas"].append({"adv": h["adv-router"], "id": h["lsa-id"],
                                             "ref_type": p["referenced-ls-type"],
                                             "ref_id": p["referenced-link-state-id"],
                                             "ref_adv": p["referenced-adv-router"], "prefixes": pf})
        for i in a.get("interfaces", {}).get("interface", []):
            llsas = []
            for t in i.get("database", {}).get("link-scope-lsa-type", []):
                for l in t.get("link-scope-lsas", {}).get("link-scope-lsa", []):
                    b = l["ospfv3"].get("body", {})
                    if "link" in b:
                        llsas.append([l["adv-router"], l["ospfv3"]["header"]["lsa-id"],
                                      b["link"]["link-local-interface-address"]])
            area["interfaces"].append({"name": i["name"], "state": i.get("state"),
                                       "interface_id": i.get("interface-id"),
                                       "neighbors": [[x["neighbor-router-id"], x["address"]]
                                                     for x in i.get("neighbors", {}).get("neighbor", [])],
                                       "link_lsas": llsas})
        for v in a.get("virtual-links", {}).get("virtual-link", []) if a.get("virtual-links") else []:
            area["interfaces"].append({"name": f"vlink-{v['transit-area-id']}-{v['router-id']}",
                                       "state": "virtual-link", "interface_id": v.get("interface-id"),
                                       "neighbors": [], "link_lsas": []})
        snap["areas"].append(area)
    for r in o.get("local-rib", {}).get("route", []):
        nhs = [[n.get("outgoing-interface"), n.get("next-hop")]
               for n in r.get("next-hops", {}).get("next-hop", [])]
        snap["local_rib"].append({"prefix": r["prefix"], "metric": r.get("metric"),
                                  "type": r.get("route-type"), "nexthops": nhs})
    return snap


def extract_ospfv3(ref: Path):
    """OSPFv3 snapshots: Router/Network/Intra-Area-Prefix LSAs, the interfaces with their
    interface ids, neighbours and link-scope Link-LSAs, and the golden local-rib."""
    base = ref / "holo-ospf/tests/conformance/ospfv3/topologies"
    out = []
    for topo in sorted(p for p in base.iterdir() if p.is_dir()):
        for rt in sorted(p for p in topo.iterdir() if p.is_dir()):
            st = rt / "output" / "northbound-state.json"
            if not st.exists():
                continue
            snap = ospfv3_snapshot(rt, st)
            snap["ibus_routes"] = ibus_routes(rt / "output" / "ibus.jsonl")
            snap["ifindex"] = ifindex_map(rt / "events.jsonl")
            out.append(snap)
    return out


def main():
    ref = Path(sys.argv[1]) if len(sys.argv) > 1 else Path("/root/reference")
    v2 = extract_ospfv2(ref)
    (HERE / "ospfv2.json").write_text(json.dumps(v2, separators=(",", ":"), sort_keys=True))
    print(f"ospfv2: {len(v2)} router snapshots -> {HERE / 'ospfv2.json'}")
    v3 = extract_ospfv3(ref)
    (HERE / "ospfv3.json").write_text(json.dumps(v3, separators=(",", ":"), sort_keys=True))
    print(f"ospfv3: {len(v3)} router snapshots -> {HERE / 'ospfv3.json'}")
    try:
        from make_golden_isis import extract_isis   # optional second extractor
        isis = extract_isis(ref)
        (HERE / "isis.json").write_text(json.dumps(isis, separators=(",", ":"), sort_keys=True))
        print(f"isis: {len(isis)} router snapshots -> {HERE / 'isis.json'}")
    except ImportError:
        pass


if __name__ == "__main__":
    sys.path.insert(0, str(HERE))
    main()
