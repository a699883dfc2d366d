"""Emit holo-replay input for the synthetic OSPFv2 LSDBs.

`holo-replay` (holo-tools/holo-replay/src/main.rs:11-27) feeds a recorded `events.jsonl` to a
protocol instance started from `config.json`; with this emitter anyone with a Rust toolchain can
run the REAL `run_area` (holo-ospf/src/spf.rs:587-729) on the LSDBs this repo benchmarks and
compare its `local-rib` with ours — the one check this image cannot do (no cargo).

What is emitted, in the reference's own record format (holo-ospf/tests/conformance/ospfv2/
topologies/topo2-1/rt4/events.jsonl is the model):

  * `Ibus` records: RouterIdUpdate, InterfaceUpd and InterfaceAddressAdd for the local interfaces;
  * per point-to-point neighbour the received packets that take the adjacency to Full: Hello,
    Hello listing us, the Database Description negotiation (the neighbour is master when its
    router-id is larger), an empty DbDesc exchange;
  * `NetRxPacket` LS Update records carrying EVERY LSA of the area image — `raw` (wire bytes with
    the Fletcher checksum), `hdr` and `body` exactly as the reference serialises `Lsa`;
  * `SpfDelayEvent` records (Igp, then DelayTimer) so that the instance runs its SPF.

Pinned here without Rust: the LSA encoder reproduces the `raw` bytes of 398 LSAs recorded in the
reference's own event files from their `hdr` / `body` (Router, Network, Summary, Router-Information
with SR TLVs, Extended-Prefix with Prefix-SIDs; tests/golden/ospfv2_lsa_vectors.json,
tests/test_replay.py).  The neighbour bring-up script follows the recorded exchanges; it has not
been replayed (no toolchain) — INTEGRATION.md says how to run it.
"""
from __future__ import annotations

import json
import struct
from pathlib import Path

import numpy as np

OPTION_BITS = {"E": 0x02, "MC": 0x04, "NP": 0x08, "L": 0x10, "DC": 0x20, "O": 0x40}
ROUTER_FLAG_BITS = {"B": 0x01, "E": 0x02, "V": 0x04, "NT": 0x10}
LINK_TYPES = {"PointToPoint": 1, "TransitNetwork": 2, "StubNetwork": 3, "VirtualLink": 4}
LINK_NAMES = {v: k for k, v in LINK_TYPES.items()}
RI_CAP_BITS = {"GR": 1 << 31, "GR_HELPER": 1 << 30, "STUB_ROUTER": 1 << 29, "TE": 1 << 28, "P2P_LAN": 1 << 27,
               "EXPERIMENTAL_TE": 1 << 26}
EXT_PREFIX_FLAG_BITS = {"A": 0x80, "N": 0x40}
PREFIX_SID_FLAG_BITS = {"NP": 0x40, "M": 0x20, "E": 0x10, "V": 0x08, "L": 0x04}
ROUTE_TYPES = {"Unspecified": 0, "IntraArea": 1, "InterArea": 3, "AsExternal": 5, "NssaExternal": 7}
ALGOS = {"Spf": 0, "StrictSpf": 1}


def _bits(s: str, table: dict) -> int:
    v = 0
    for name in (x.strip() for x in (s or "").split("|")):
        if name:
            v |= table[name]
    return v


def _flags_str(v: int, table: dict) -> str:
    return " | ".join(k for k, b in table.items() if v & b)


def ip(a) -> bytes:
    if isinstance(a, str):
        return bytes(int(x) for x in a.split("."))
    return struct.pack(">I", int(a))


def ip_str(a: int) -> str:
    return ".".join(str((int(a) >> s) & 0xFF) for s in (24, 16, 8, 0))


def mask_len(mask: int) -> int:
    return bin(int(mask) & 0xFFFFFFFF).count("1")


def fletcher16(data: bytes, cksum_off: int) -> int:
    """OSPF LSA checksum (RFC 905 Annex B / RFC 2328 12.1.7) over `data` (the LSA without its age
    field), checksum field at `cksum_off` inside `data`."""
    c0 = c1 = 0
    buf = bytearray(data)
    buf[cksum_off] = buf[cksum_off + 1] = 0
    for b in buf:
        c0 = (c0 + b) % 255
        c1 = (c1 + c0) % 255
    x = ((len(buf) - cksum_off - 1) * c0 - c1) % 255
    if x <= 0:
        x += 255
    y = 510 - c0 - x
    if y > 255:
        y -= 255
    return (x << 8) | y


def _tlv(t: int, val: bytes) -> bytes:
    pad = (-len(val)) % 4
    return struct.pack(">HH", t, len(val)) + val + b"\0" * pad


def _sid(s: dict) -> bytes:
    """SID/Label sub-TLV value: 3-byte label or 4-byte index."""
    if "Label" in s:
        return struct.pack(">I", s["Label"])[1:]
    return struct.pack(">I", s["Index"])


def encode_body(body: dict) -> bytes:
    kind = next(iter(body))
    b = body[kind]
    if kind == "Router":
        out = struct.pack(">BBH", _bits(b["flags"], ROUTER_FLAG_BITS), 0, len(b["links"]))
        for l in b["links"]:
            out += ip(l["link_id"]) + ip(l["link_data"]) + struct.pack(">BBH", LINK_TYPES[l["link_type"]], 0, l["metric"])
        return out
    if kind == "Network":
        return ip(b["mask"]) + b"".join(ip(r) for r in b["attached_rtrs"])
    if kind in ("SummaryNetwork", "SummaryRouter"):
        return ip(b["mask"]) + struct.pack(">I", b["metric"] & 0xFFFFFF)
    if kind == "OpaqueArea":
        sub = next(iter(b))
        o = b[sub]
        if sub == "RouterInfo":
            out = b""
            if o.get("info_caps") is not None:
                out += _tlv(1, struct.pack(">I", _bits(o["info_caps"], RI_CAP_BITS)))
            if o.get("func_caps") is not None:
                out += _tlv(2, struct.pack(">I", _bits(o["func_caps"], {})))
            if o.get("sr_algo") is not None:
                out += _tlv(8, bytes(ALGOS[a] for a in o["sr_algo"]))
            for typ, key in ((9, "srgb"), (14, "srlb")):
                for r in o.get(key) or []:
                    sid = _sid(r["first"])
                    # (the sub-TLV's padding is not counted in the enclosing TLV's length)
                    out += _tlv(typ, struct.pack(">I", r["range"])[1:] + b"\0" + struct.pack(">HH", 1, len(sid)) + sid)
            if o.get("msds") or o.get("srms_pref") is not None or o.get("unknown_tlvs"):
                raise NotImplementedError("RouterInfo: MSD / SRMS / unknown TLVs")
            return out
        if sub == "ExtPrefix":
            out = b""
            for p in o["prefixes"].values():
                addr, plen = p["prefix"].split("/")
                val = struct.pack(">BBBB", ROUTE_TYPES[p["route_type"]], int(plen), p["af"],
                                  _bits(p["flags"], EXT_PREFIX_FLAG_BITS)) + ip(addr)
                for s in p["prefix_sids"].values():
                    val += _tlv(2, struct.pack(">BBBB", _bits(s["flags"], PREFIX_SID_FLAG_BITS), 0, 0, ALGOS[s["algo"]]) +
                                _sid(s["sid"]))
                if p.get("unknown_tlvs"):
                    raise NotImplementedError("ExtPrefix: unknown TLVs")
                out += _tlv(1, val)
            return out
    raise NotImplementedError(kind)


def encode_lsa(hdr: dict, body: dict) -> bytes:
    """Wire bytes of an OSPFv2 LSA from the reference's JSON `hdr` / `body` (cksum and length of
    `hdr` are recomputed)."""
    bb = encode_body(body)
    length = 20 + len(bb)
    h = struct.pack(">HBB", hdr["age"], _bits(hdr["options"], OPTION_BITS), hdr["lsa_type"]) + ip(hdr["lsa_id"]) + \
        ip(hdr["adv_rtr"]) + struct.pack(">IHH", hdr["seq_no"], 0, length)
    raw = bytearray(h + bb)
    ck = fletcher16(bytes(raw[2:]), 14)
    raw[16:18] = struct.pack(">H", ck)
    return bytes(raw)


def lsa_record(age, options, lsa_type, lsa_id, adv_rtr, body, seq_no=0x80000001) -> dict:
    hdr = {"age": int(age), "options": options, "lsa_type": int(lsa_type), "lsa_id": ip_str(lsa_id),
           "adv_rtr": ip_str(adv_rtr), "seq_no": int(seq_no), "cksum": 0, "length": 0}
    raw = encode_lsa(hdr, body)
    hdr["cksum"] = struct.unpack(">H", raw[16:18])[0]
    hdr["length"] = len(raw)
    return {"raw": list(raw), "hdr": hdr, "body": body}


# ----------------------------------------------------------------------------- area image -> LSAs
def area_lsas(area) -> list:
    """Every LSA of an Ospfv2Area image (holo_b200/ospfv2.py) as the reference's `Lsa` JSON."""
    out = []
    for r in area.router_lsas:
        links = [{"link_type": LINK_NAMES[int(l["link_type"])], "link_id": ip_str(l["link_id"]),
                  "link_data": ip_str(l["link_data"]), "metric": int(l["metric"])}
                 for l in area.links[int(r["link_off"]): int(r["link_off"]) + int(r["n_links"])]]
        out.append(lsa_record(r["age"], _flags_str(int(r["options"]), OPTION_BITS), 1, r["lsa_id"], r["adv_rtr"],
                              {"Router": {"flags": _flags_str(int(r["flags"]), ROUTER_FLAG_BITS), "links": links}}))
    for n in area.network_lsas:
        att = [ip_str(a) for a in area.attached[int(n["att_off"]): int(n["att_off"]) + int(n["n_att"])]]
        out.append(lsa_record(n["age"], "E", 2, n["lsa_id"], n["adv_rtr"],
                              {"Network": {"mask": ip_str(n["mask"]), "attached_rtrs": att}}))
    for r in area.ri_lsas:
        srgb = [{"first": ({"Index": int(g["first"])} if g["first_is_index"] else {"Label": int(g["first"])}),
                 "range": int(g["range"])}
                for g in area.srgbs[int(r["srgb_off"]): int(r["srgb_off"]) + int(r["n_srgb"])]]
        algo = (["Spf"] if r["sr_algo_has_spf"] else []) if r["has_sr_algo"] else None
        out.append(lsa_record(r["age"], "E", 10, r["lsa_id"], r["adv_rtr"],
                              {"OpaqueArea": {"RouterInfo": {"info_caps": "GR_HELPER | STUB_ROUTER", "func_caps": None,
                                                             "sr_algo": algo, "srgb": srgb, "srlb": [], "msds": None,
                                                             "srms_pref": None, "unknown_tlvs": []}}}))
    per_rtr = {}
    for e in area.ext_prefixes:
        per_rtr.setdefault(int(e["adv_rtr"]), []).append(e)
    for adv, lst in per_rtr.items():
        for k, e in enumerate(lst):
            pfx = f"{ip_str(e['prefix'])}/{mask_len(e['mask'])}"
            sids = {}
            if e["has_sid"]:
                sids["Spf"] = {"flags": _flags_str(int(e["sid_flags"]), PREFIX_SID_FLAG_BITS), "algo": "Spf",
                               "sid": ({"Label": int(e["sid_value"])} if e["sid_is_label"] else {"Index": int(e["sid_value"])})}
            body = {"OpaqueArea": {"ExtPrefix": {"prefixes": {pfx: {
                "route_type": {1: "IntraArea", 3: "InterArea", 5: "AsExternal", 7: "NssaExternal"}.get(int(e["route_type"]), "Unspecified"),
                "af": 0, "flags": "N" if mask_len(e["mask"]) == 32 else "", "prefix": pfx, "prefix_sids": sids,
                "unknown_tlvs": []}}}}}
            out.append(lsa_record(e["age"], "E", 10, 0x07000000 + k, adv, body))
    return out


# ----------------------------------------------------------------------------- records
def _pkt(area_key, iface_key, src, kind, hdr_rid, body):
    pk = dict(body)
    pk["hdr"] = {"pkt_type": kind, "router_id": ip_str(hdr_rid), "area_id": "0.0.0.0"}
    return {"Protocol": {"NetRxPacket": {"area_key": {"Id": area_key}, "iface_key": {"Id": iface_key},
                                         "src": ip_str(src), "dst": "224.0.0.5",
                                         "packet": {"Ok": {kind: {"hdr": pk.pop("hdr"), **pk}}}}}}


def events(area, max_lsas_per_update: int = 20) -> list:
    """Event records for the local router of `area` (point-to-point interfaces only bring up an
    adjacency; all LSAs arrive over the first one)."""
    from . import ospfv2
    me = int(area.router_id)
    ev = [{"Ibus": {"RouterIdUpdate": ip_str(me)}}]
    p2p = []
    for i, f in enumerate(area.ifaces):
        name = area.ifnames[i] if i < len(area.ifnames) else f"eth{i}"
        ev.append({"Ibus": {"InterfaceUpd": {"ifname": name, "ifindex": int(f["ifindex"]), "mtu": 1500, "flags": "OPERATIVE"}}})
        for a in area.iface_addrs[int(f["addr_off"]): int(f["addr_off"]) + int(f["n_addrs"])]:
            ev.append({"Ibus": {"InterfaceAddressAdd": {"ifname": name, "addr": f"{ip_str(a['addr'])}/{mask_len(a['mask'])}",
                                                         "flags": ""}}})
        if int(f["if_type"]) == ospfv2.IF_P2P and int(f["n_nbrs"]) == 1:
            p2p.append((i + 1, area.nbrs[int(f["nbr_off"])], area.iface_addrs[int(f["addr_off"])]))
    ev.append({"Protocol": {"LsaOrigEvent": {"event": {"AreaStart": {"area_id": 1}}}}})
    for key, nbr, addr in p2p:
        rid, src = int(nbr["router_id"]), int(nbr["src"])
        hello = {"network_mask": ip_str(addr["mask"]), "hello_interval": 3, "options": "E", "priority": 1,
                 "dead_interval": 12, "dr": None, "bdr": None, "neighbors": [], "lls": None}
        ev.append(_pkt(1, key, src, "Hello", rid, hello))
        ev.append(_pkt(1, key, src, "Hello", rid, dict(hello, neighbors=[ip_str(me)])))
        # Database Description: the larger router-id is master; an empty database summary
        nbr_master = rid > me
        seq = 0x04040405
        ev.append(_pkt(1, key, src, "DbDesc", rid, {"mtu": 1500, "options": "E | O", "dd_flags": "MS | M | I",
                                                    "dd_seq_no": seq, "lsa_hdrs": [], "lls": None}))
        ev.append(_pkt(1, key, src, "DbDesc", rid, {"mtu": 1500, "options": "E | O",
                                                    "dd_flags": "MS" if nbr_master else "", "dd_seq_no": seq + 1,
                                                    "lsa_hdrs": [], "lls": None}))
    if p2p:
        key, nbr, _ = p2p[0]
        lsas = area_lsas(area)
        for i in range(0, len(lsas), max_lsas_per_update):
            ev.append(_pkt(1, key, int(nbr["src"]), "LsUpdate", int(nbr["router_id"]), {"lsas": lsas[i: i + max_lsas_per_update]}))
    ev.append({"Protocol": {"SpfDelayEvent": {"event": "Igp"}}})
    ev.append({"Protocol": {"SpfDelayEvent": {"event": "DelayTimer"}}})
    return ev


def config(area) -> dict:
    """config.json of the local router (ietf-ospf instance data, as in the conformance topologies)."""
    from . import ospfv2
    names = [area.ifnames[i] if i < len(area.ifnames) else f"eth{i}" for i in range(len(area.ifaces))]
    ifs = []
    for name, f in zip(names, area.ifaces):
        e = {"name": name}
        if int(f["if_type"]) == ospfv2.IF_P2P:
            e.update({"interface-type": "point-to-point", "hello-interval": 3, "dead-interval": 12})
        elif int(f["if_type"]) == ospfv2.IF_BROADCAST:
            e.update({"interface-type": "broadcast", "hello-interval": 3, "dead-interval": 12})
        ifs.append(e)
    return {
        "ietf-interfaces:interfaces": {"interface": [{"name": n, "type": "iana-if-type:ethernetCsmacd", "ietf-ip:ipv4": {}}
                                                     for n in names]},
        "ietf-routing:routing": {"control-plane-protocols": {"control-plane-protocol": [{
            "type": "ietf-ospf:ospfv2", "name": "test",
            "ietf-ospf:ospf": {"explicit-router-id": ip_str(area.router_id),
                               "areas": {"area": [{"area-id": "0.0.0.0", "interfaces": {"interface": ifs}}]}}}]}},
    }


def write(directory, area, max_lsas_per_update: int = 20):
    """Write <directory>/config.json and <directory>/events.jsonl; returns the number of LSAs."""
    d = Path(directory)
    d.mkdir(parents=True, exist_ok=True)
    (d / "config.json").write_text(json.dumps(config(area), indent=2))
    ev = events(area, max_lsas_per_update)
    with open(d / "events.jsonl", "w") as f:
        for e in ev:
            f.write(json.dumps(e, separators=(",", ":")) + "\n")
    return sum(len(e["Protocol"]["NetRxPacket"]["packet"]["Ok"]["LsUpdate"]["lsas"]) for e in ev
               if "Protocol" in e and "NetRxPacket" in e["Protocol"] and "LsUpdate" in e["Protocol"]["NetRxPacket"]["packet"]["Ok"])


if __name__ == "__main__":
    import argparse
    from . import ospfv2, synth
    ap = argparse.ArgumentParser(description="emit holo-replay input for a BASELINE config's LSDB")
    ap.add_argument("config", choices=["C1", "C2", "C5"])
    ap.add_argument("out")
    a = ap.parse_args()
    if a.config == "C1":
        t = synth.random_topology(100, 400, synth.SEED_BASE + 1)
        area = ospfv2.synth_area(t, root=0)
    elif a.config == "C2":
        t = synth.random_topology(10000, 40000, synth.SEED_BASE + 2)
        area = ospfv2.synth_area(t, root=0)
    else:
        t = synth.random_topology(10000, 40000, synth.SEED_BASE + 5, cost_choices=[10, 20], lan_fraction=0.05)
        area = ospfv2.synth_area(t, root=0, sr=True)
    n = write(a.out, area)
    print(f"{a.out}: {n} LSAs; replay with: holo-replay --protocol OSPFv2 {a.out}/events.jsonl")
