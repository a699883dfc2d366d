#!/usr/bin/env python
"""Extract (hdr, body, raw) triples of OSPFv2 LSAs from the reference's recorded event files
(holo-ospf/tests/conformance/ospfv2/topologies/*/*/events.jsonl, NetRxPacket LsUpdate records)
into tests/golden/ospfv2_lsa_vectors.json: known-answer vectors for the LSA encoder of
holo_b200/replay.py (the emitter of holo-replay input).

    python tests/golden/make_lsa_vectors.py /root/reference

Unique triples only, at most CAP per body kind (the files repeat the same LSAs many times)."""
import collections
import glob
import json
import sys
from pathlib import Path

CAP = 120
ref = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
out = Path(__file__).resolve().parent / "ospfv2_lsa_vectors.json"
seen, per, vec = set(), collections.Counter(), []
for f in sorted(glob.glob(str(ref / "holo-ospf/tests/conformance/ospfv2/topologies/*/*/events.jsonl"))):
    for line in open(f):
        if '"LsUpdate"' not in line:
            continue
        try:
            lsas = json.loads(line)["Protocol"]["NetRxPacket"]["packet"]["Ok"]["LsUpdate"]["lsas"]
        except (KeyError, TypeError):
            continue
        for l in lsas:
            b = l["body"]
            kind = next(iter(b))
            if kind.startswith("Opaque"):
                kind += ":" + next(iter(b[kind]))
            key = bytes(l["raw"])
            if key in seen or per[kind] >= CAP:
                continue
            seen.add(key)
            per[kind] += 1
            vec.append({"kind": kind, "hdr": l["hdr"], "body": b, "raw": l["raw"]})
out.write_text(json.dumps({"source": "holo-ospf/tests/conformance/ospfv2/topologies/*/*/events.jsonl", "vectors": vec},
                          separators=(",", ":")))
print(len(vec), dict(per), out.stat().st_size, "bytes")
