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


def _updown(lsp):
    """prefixes of this LSP whose up/down bit is set (never propagated L1 -> L2, lsdb.rs:1331)"""
    out = []
    for key in ("ipv4-internal-reachability", "ipv4-external-reachability", "extended-ipv4-reachability",
                "ipv6-reachability", "mt-ipv6-reachability"):
        out += [f"{p['ip-prefix']}/{p['prefix-len']}" for p in (lsp.get(key) or {}).get("prefixes", []) if p.get("up-down")]
    return out


# Step tests: the state the reference reached after the step (full database, interfaces and
# adjacencies in <step>/<NN>-output-northbound-state.json) is one more snapshot; the table computed
# from it, diffed against the table of the topology snapshot, must give the step's ibus output
# (holo-isis/tests/conformance/mod.rs: the cases named here).  `iface_metric` are the interface
# metrics the step configures (they are configuration, not state).
ISIS_STEPS = [
    # (test, topology, router, step whose ibus output carries the route messages, configuration the
    #  step changed — configuration is not part of the state dump)
    ("nb-config-att-ignore1", "topo1-2", "rt7", "01", {"att_ignore": True}),
    ("nb-config-spf-paths1", "topo2-1", "rt1", "01", {"max_paths": 1}),
    ("nb-config-af1", "topo2-1", "rt1", "02", {"afs": ["ipv6"]}),
    ("nb-config-af2", "topo2-1", "rt1", "02", {"afs": ["ipv4"]}),
    ("nb-config-iface-delete1", "topo2-1", "rt6", "02", {"delete_iface": "eth-rt5"}),
    ("nb-config-iface-metric1", "topo2-1", "rt6", "02", {"iface_metric": {"eth-rt4": 50}}),
    ("nb-config-enabled1", "topo2-1", "rt6", "01", {}),
    ("nb-config-enabled2", "topo2-1", "rt6", "01", {}),
    ("nb-config-iface-enabled1", "topo2-1", "rt6", "02", {}),
    ("nb-config-iface-enabled2", "topo2-1", "rt6", "02", {}),
    ("nb-config-iface-passive1", "topo2-1", "rt6", "02", {}),
    ("nb-rpc-clear-adjacency1", "topo2-1", "rt6", "02", {}),
    ("nb-rpc-clear-adjacency2", "topo2-1", "rt6", "02", {}),
    ("nb-rpc-clear-database1", "topo2-1", "rt6", "02", {}),
    ("ibus-iface-update1", "topo2-1", "rt6", "02", {}),
    ("ibus-iface-update2", "topo2-1", "rt6", "02", {}),
    ("timeout-adj1", "topo2-1", "rt6", "02", {}),
    ("pdu-lsp-att-bit1", "topo1-2", "rt7", "02", {}),
    ("pdu-lsp-expiration1", "topo2-1", "rt6", "02", {}),
    ("pdu-lsp-overload1", "topo2-1", "rt6", "02", {}),
]

# L1 -> L2 summary routes of an L1/L2 router (holo-isis/src/route.rs:189-231): chains of steps, each
# diffed against the state before it.  (test, topology, router, [(step, configured summaries
# [[prefix, metric or null]], ...)]).
ISIS_SUMMARY_CHAINS = [
    ("nb-config-summary1", "topo1-2", "rt2", [("01", [["1.0.0.0/8", None]]), ("02", [["1.0.0.0/8", 100]]), ("04", [])]),
    ("nb-config-summary2", "topo1-2", "rt2", [("01", [["1.0.0.0/8", None]]), ("03", [["1.0.0.0/8", None]])]),
]


def snapshot(rt: Path, state_path: Path, overrides=None):
    """One snapshot dict from a northbound-state dump and the router's config.json."""
    overrides = overrides or {}
    o = isis_root(json.loads(state_path.read_text()))
    cfg = isis_root(json.loads((rt / "config.json").read_text()))
    # instance.config.is_af_enabled(af): enabled unless configured `enabled: false`
    afs = [af for af in ("ipv4", "ipv6")
           if all(a.get("enabled", True) for a in cfg.get("address-families", {}).get("address-family-list", [])
                  if a["address-family"] == af)]
    snap = {"topo": rt.parent.name, "rt": rt.name, "system_id": cfg["system-id"],
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
                "updown": _updown(l),
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
    for i in snap["interfaces"]:
        if i["name"] in overrides.get("iface_metric", {}):
            i["metric"] = overrides["iface_metric"][i["name"]]
    snap["att_ignore"] = bool(overrides.get("att_ignore", False))
    if "max_paths" in overrides:
        snap["max_paths"] = overrides["max_paths"]
    if "afs" in overrides:
        snap["afs"] = overrides["afs"]
    if "delete_iface" in overrides:
        snap["interfaces"] = [i for i in snap["interfaces"] if i["name"] != overrides["delete_iface"]]
        snap["adjacencies"] = [a for a in snap["adjacencies"] if a["iface"] != overrides["delete_iface"]]
        snap["deleted_ifaces"] = [overrides["delete_iface"]]
    return snap


def ibus_stream(path: Path):
    """Ordered [kind, prefix, metric, [[ifindex, addr]]] of a step's ibus output."""
    out = []
    if not path.exists():
        return out
    for line in path.read_text().replace("}{", "}\n{").splitlines():
        try:
            msg = json.loads(line)
        except json.JSONDecodeError:
            continue
        if "RouteIpAdd" in msg:
            r = msg["RouteIpAdd"]
            out.append(["add", r["prefix"], r["metric"],
                        [[n["Address"]["ifindex"], n["Address"].get("addr")] for n in r.get("nexthops", []) if "Address" in n]])
        elif "RouteIpDel" in msg:
            out.append(["del", msg["RouteIpDel"]["prefix"], None, []])
    return out


def extract_isis(ref: Path):
    base = ref / "holo-isis/tests/conformance/topologies"
    out = []
    for topo in sorted(p for p in base.iterdir() if p.is_dir()):
        for rt in sorted(p for p in topo.iterdir() if p.is_dir()):
            st = rt / "output" / "northbound-state.json"
            if not st.exists():
                continue
            snap = snapshot(rt, st)
            # what the instance sent to the RIB manager (update_global_rib, holo-isis/src/route.rs:255-314)
            from make_golden import ibus_routes, ifindex_map, step_outputs
            snap["steps"] = step_outputs(ref / "holo-isis/tests/conformance", "isis", topo.name, rt.name)
            snap["ibus_routes"] = ibus_routes(rt / "output" / "ibus.jsonl")
            snap["ifindex"] = ifindex_map(rt / "events.jsonl")
            snap["after"] = {}
            for (name, t, r, nn, ov) in ISIS_STEPS:
                if (t, r) != (topo.name, rt.name):
                    continue
                sd = ref / "holo-isis/tests/conformance" / name
                after = snapshot(rt, sd / f"{nn}-output-northbound-state.json", ov)
                after["ibus"] = ibus_stream(sd / f"{nn}-output-ibus.jsonl")
                snap["after"][name] = after
            snap["summary_chains"] = {}
            for (name, t, r, chain) in ISIS_SUMMARY_CHAINS:
                if (t, r) != (topo.name, rt.name):
                    continue
                sd = ref / "holo-isis/tests/conformance" / name
                steps = []
                for nn, summ in chain:
                    st_ = snapshot(rt, sd / f"{nn}-output-northbound-state.json")
                    st_["summaries"] = summ
                    st_["ibus"] = ibus_stream(sd / f"{nn}-output-ibus.jsonl")
                    steps.append(st_)
                snap["summary_chains"][name] = steps
            out.append(snap)
    return out
