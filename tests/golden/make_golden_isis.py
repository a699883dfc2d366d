"""IS-IS part of tests/golden/make_golden.py: LSDB (per level: LSPs with IS and IP
reachability), configured system id / metric type / address families / MT topologies, the
local adjacencies, and the golden local-rib, from
holo-isis/tests/conformance/topologies/<topo>/<rt>/{config.json,output/northbound-state.json}."""
from __future__ import annotations

import json
from pathlib import Path


def isis_root(d):
    return d["ietf-routing:routing"]["control-plane-protocols"]["control-plane-protocol"][0]["ietf-isis:isis"]


def _reach(lsp, key, metric_of):
    out = []
    for n in (lsp.get(key) or {}).get("neighbor", []):
        for inst in n.get("instances", {}).get("instance", []):
            out.append([n["neighbor-id"], metric_of(inst), n.get("mt-id", inst.get("mt-id"))])
    return out


def _pfx(lsp, key, metric_of):
    return [[f"{p['ip-prefix']}/{p['prefix-len']}", metric_of(p), p.get("mt-id")]
            for p in (lsp.get(key) or {}).get("prefixes", [])]


def extract_isis(ref: Path):
    base = ref / "holo-isis/tests/conformance/topologies"
    out = []
    for topo in sorted(p for p in base.iterdir() if p.is_dir()):
        for rt in sorted(p for p in topo.iterdir() if p.is_dir()):
            st = rt / "output" / "northbound-state.json"
            if not st.exists():
                continue
            o = isis_root(json.loads(st.read_text()))
            cfg = isis_root(json.loads((rt / "config.json").read_text()))
            # instance.config.is_af_enabled(af): enabled unless configured `enabled: false`
            afs = [af for af in ("ipv4", "ipv6")
                   if all(a.get("enabled", True) for a in cfg.get("address-families", {}).get("address-family-list", [])
                          if a["address-family"] == af)]
            snap = {"topo": topo.name, "rt": rt.name, "system_id": cfg["system-id"],
                    "metric_type": (cfg.get("metric-type") or {}).get("value", "wide-only"),
                    "afs": afs, "mt_ipv6": bool(cfg.get("topologies")), "levels": [], "adjacencies": [],
                    "level_type": cfg.get("level-type", "level-all"), "areas": cfg.get("area-address", []),
                    "max_paths": (cfg.get("spf-control") or {}).get("paths", 16),
                    "interfaces": [{"name": i["name"], "type": i.get("interface-type", "broadcast"),
                                    "metric": (i.get("metric") or {}).get("value", 10)}
                                   for i in cfg.get("interfaces", {}).get("interface", [])],
                    "local_rib": []}
            for lv in o.get("database", {}).get("levels", []):
                lsps = []
                for l in lv.get("lsp", []):
                    dm = lambda x: x["default-metric"]["metric"]
                    lsps.append({
                        "id": l["lsp-id"], "flags": l.get("attributes", {}).get("lsp-flags", []),
                        "protocols": l.get("protocol-supported"),
                        "mt_flags": {str(t["mt-id"]): t.get("attributes", {}).get("flags", [])
                                     for t in (l.get("mt-entries") or {}).get("topology", [])},
                        "is": _reach(l, "is-neighbor", dm),
                        "ext_is": _reach(l, "extended-is-neighbor", lambda x: x["metric"]),
                        "mt_is": _reach(l, "mt-is-neighbor", lambda x: x["metric"]),
                        "ipv4_int": _pfx(l, "ipv4-internal-reachability", dm),
                        "ipv4_ext": _pfx(l, "ipv4-external-reachability", dm),
                        "ext_ipv4": _pfx(l, "extended-ipv4-reachability", lambda x: x["metric"]),
                        "ipv6": _pfx(l, "ipv6-reachability", lambda x: x["metric"]),
                        "mt_ipv6": _pfx(l, "mt-ipv6-reachability", lambda x: x["metric"]),
                    })
                snap["levels"].append({"level": lv["level"], "lsps": lsps})
            for i in o.get("interfaces", {}).get("interface", []):
                for a in (i.get("adjacencies") or {}).get("adjacency", []):
                    snap["adjacencies"].append({"iface": i["name"], "sysid": a["neighbor-sysid"], "state": a.get("state"),
                                                "usage": a.get("usage"), "ipv4": a.get("holo-isis:ipv4-addresses", []),
                                                "ipv6": a.get("holo-isis:ipv6-addresses", []),
                                                "areas": a.get("holo-isis:area-addresses", []),
                                                "topologies": a.get("holo-isis:topologies", [])})
            for r in o.get("local-rib", {}).get("route", []):
                nhs = [[n.get("outgoing-interface"), n.get("next-hop")]
                       for n in r.get("next-hops", {}).get("next-hop", [])]
                snap["local_rib"].append({"prefix": r["prefix"], "metric": r.get("metric"), "level": r.get("level"),
                                          "nexthops": nhs})
            # what the instance sent to the RIB manager (update_global_rib, holo-isis/src/route.rs:255-314)
            from make_golden import ibus_routes, ifindex_map, step_outputs
            snap["steps"] = step_outputs(ref / "holo-isis/tests/conformance", "isis", topo.name, rt.name)
            snap["ibus_routes"] = ibus_routes(rt / "output" / "ibus.jsonl")
            snap["ifindex"] = ifindex_map(rt / "events.jsonl")
            out.append(snap)
    return out
