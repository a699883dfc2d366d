"""holo-replay emitter (holo_b200/replay.py): the LSA encoder against the wire
bytes the reference itself recorded, and the shape of the emitted records against the
reference's record files."""
import json
from pathlib import Path

import numpy as np
import pytest

from holo_b200 import ospfv2, replay, synth

VEC = json.loads((Path(__file__).parent / "golden" / "ospfv2_lsa_vectors.json").read_text())["vectors"]


def test_encoder_reproduces_the_recorded_wire_bytes():
    """398 LSAs recorded by the reference (Router, Network, Summary, Router-Information with SR
    TLVs, Extended-Prefix with Prefix-SIDs): header, body, length and Fletcher checksum."""
    n = 0
    for v in VEC:
        if v["kind"] == "OpaqueArea:ExtLink":
            continue            # adjacency SIDs: not on the SPF path, not emitted
        raw = replay.encode_lsa(v["hdr"], v["body"])
        assert list(raw) == v["raw"], (v["kind"], v["hdr"])
        n += 1
    assert n >= 390


def test_checksum_detects_corruption():
    v = VEC[0]
    raw = bytearray(v["raw"])
    assert replay.fletcher16(bytes(raw[2:]), 14) == v["hdr"]["cksum"]
    raw[25] ^= 1
    assert replay.fletcher16(bytes(raw[2:]), 14) != v["hdr"]["cksum"]


def _shape(x):
    if isinstance(x, dict):
        return {k: _shape(v) for k, v in x.items()}
    if isinstance(x, list):
        return [_shape(x[0])] if x else []
    return type(x).__name__


def test_emitted_records_have_the_reference_shape(tmp_path):
    """Every LS Update record is {raw, hdr{age..length}, body{...}} inside the NetRxPacket envelope;
    the LSAs of the image are all there, decodable from their own wire bytes."""
    t = synth.random_topology(60, 260, synth.SEED_BASE + 51, lan_fraction=0.1)
    area = ospfv2.synth_area(t, root=0, sr=True)
    n = replay.write(tmp_path, area)
    assert n == len(area.router_lsas) + len(area.network_lsas) + len(area.ri_lsas) + len(area.ext_prefixes)
    cfg = json.loads((tmp_path / "config.json").read_text())
    ospf = cfg["ietf-routing:routing"]["control-plane-protocols"]["control-plane-protocol"][0]["ietf-ospf:ospf"]
    assert ospf["explicit-router-id"] == replay.ip_str(area.router_id)
    kinds, seen = set(), 0
    ref_hdr_keys = set(VEC[0]["hdr"].keys())
    for line in open(tmp_path / "events.jsonl"):
        rec = json.loads(line)
        top = next(iter(rec))
        kinds.add(next(iter(rec[top])))
        if top == "Protocol" and "NetRxPacket" in rec[top]:
            p = rec[top]["NetRxPacket"]
            assert set(p) == {"area_key", "iface_key", "src", "dst", "packet"}
            kind = next(iter(p["packet"]["Ok"]))
            if kind == "LsUpdate":
                for l in p["packet"]["Ok"]["LsUpdate"]["lsas"]:
                    assert set(l) == {"raw", "hdr", "body"} and set(l["hdr"]) == ref_hdr_keys
                    assert l["hdr"]["length"] == len(l["raw"])
                    assert replay.fletcher16(bytes(l["raw"][2:]), 14) == l["hdr"]["cksum"]
                    assert list(replay.encode_lsa(l["hdr"], l["body"])) == l["raw"]
                    seen += 1
    assert seen == n
    assert {"RouterIdUpdate", "InterfaceUpd", "InterfaceAddressAdd", "NetRxPacket", "SpfDelayEvent"} <= kinds
    # body shapes equal the reference's for the kinds both have
    ours = {}
    for line in open(tmp_path / "events.jsonl"):
        rec = json.loads(line)
        try:
            lsas = rec["Protocol"]["NetRxPacket"]["packet"]["Ok"]["LsUpdate"]["lsas"]
        except (KeyError, TypeError):
            continue
        for l in lsas:
            k = next(iter(l["body"]))
            k += ":" + next(iter(l["body"][k])) if k.startswith("Opaque") else ""
            ours.setdefault(k, _shape(l["body"]))
    theirs = {}
    for v in VEC:
        b = v["body"]
        if v["kind"].startswith("OpaqueArea:RouterInfo") and not b["OpaqueArea"]["RouterInfo"]["srgb"]:
            continue
        if v["kind"] == "Router" and not b["Router"]["links"]:
            continue
        theirs.setdefault(v["kind"] if ":" in v["kind"] else v["kind"], _shape(b))
    for k, sh in ours.items():
        kk = k if ":" in k else k
        if kk in theirs:
            if kk == "OpaqueArea:ExtPrefix":      # the prefix is the map key: compare the value shapes
                a = next(iter(sh["OpaqueArea"]["ExtPrefix"]["prefixes"].values()))
                b = next(iter(theirs[kk]["OpaqueArea"]["ExtPrefix"]["prefixes"].values()))
                assert a == b
            else:
                assert sh == theirs[kk], kk
