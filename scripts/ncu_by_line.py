#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` export of spf_batch_kernel by source line and phase.

The source page lists SASS instructions without line numbers, so they are joined (by
position) with `nvdisasm -g -c` output of the same cubin, which carries `//## File ..., line N`
markers (the library is built with -lineinfo).

  cuobjdump -xelf all holo_b200/lib/libholo_spf.so            # -> hspf_capi.sm_100a.cubin
  nvdisasm -g -c hspf_capi.sm_100a.cubin > dis.txt
  ncu -i gpurun_out/<tag>_full.ncu-rep --page source --csv > src.csv
  python scripts/ncu_by_line.py dis.txt src.csv [kernel-substring [spf_kernel.cuh of that build]]
"""
import collections
import csv
import re
import sys
from pathlib import Path

dis_path, csv_path = sys.argv[1], sys.argv[2]
kname = sys.argv[3] if len(sys.argv) > 3 else "spf_batch_kernelItLb1ELb1"
SRC = Path(sys.argv[4]) if len(sys.argv) > 4 else Path(__file__).resolve().parent.parent / "holo_b200" / "csrc" / "spf_kernel.cuh"

cur, infunc, seq = None, False, []
for l in open(dis_path).read().split("\n"):
    if l.startswith(".text."):
        infunc = kname in l
        continue
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    if infunc and re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        seq.append((cur, l.strip()))
rows = list(csv.reader(open(csv_path)))
hdr, data = rows[1], rows[2:]
iS, iI, iT = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
print(f"{len(seq)} SASS instructions, {len(data)} ncu rows")
assert len(seq) == len(data), "disassembly and profile are of different builds"
agg = collections.defaultdict(lambda: [0, 0, 0])
for (key, _), d in zip(seq, data):
    a = agg[key]
    a[0] += int(d[iS]); a[1] += int(d[iI]); a[2] += int(d[iT])
tot = [sum(a[i] for a in agg.values()) for i in range(3)]
print(f"total samples {tot[0]}  warp instructions {tot[1]}")
src = SRC.read_text().split("\n")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    f, ln = key if key else ("?", 0)
    text = src[ln - 1].strip()[:90] if f.endswith("spf_kernel.cuh") and 0 < ln <= len(src) else f
    print(f"{100 * a[0] / tot[0]:5.1f}% samp {100 * a[1] / tot[1]:5.1f}% inst  thr/inst {a[2] / max(a[1], 1):4.1f}  L{ln:<4d} {text}")


def find(t):
    return [i + 1 for i, l in enumerate(src) if t in l][0]


marks = [("helpers", "__device__ __forceinline__ uint32_t sat_add"),
         ("prolog / job fetch / init", "spf_batch_kernel(const BatchArgs"),
         ("sssp rounds", "= phase 1: SSSP ="),
         ("bucket scan", "near bucket exhausted"),
         ("h0 / zeroing", "SSSP done:"),
         ("parents (packed, jump path)", "= phase 2: ECMP parents"),
         ("parents (general)", "in-edge range of the next vertex is fetched one iteration ahead"),
         ("dist write-back", "distances are final"),
         ("jump: hops", "= phase 3J: pointer jumping ="),
         ("jump: next hops", "// -- next hops."),
         ("kahn", "= phase 3K: Kahn push"),
         ("write-back", "= write-back =")]
pos = sorted([(n, find(t)) for n, t in marks], key=lambda x: x[1]) + [("end", len(src) + 1)]
print("---- by phase")
for (name, a), (_, b) in zip(pos, pos[1:]):
    s = [0, 0, 0]
    for key, v in agg.items():
        if key and key[0].endswith("spf_kernel.cuh") and a <= key[1] < b:
            for i in range(3):
                s[i] += v[i]
    print(f"{name:30s} samples {100 * s[0] / tot[0]:5.1f}%  inst {100 * s[1] / tot[1]:5.1f}%  thr/inst {s[2] / max(s[1], 1):.1f}")
oth = [0, 0, 0]
for key, v in agg.items():
    if not (key and key[0].endswith("spf_kernel.cuh")):
        for i in range(3):
            oth[i] += v[i]
print(f"{'inlined headers':30s} samples {100 * oth[0] / tot[0]:5.1f}%  inst {100 * oth[1] / tot[1]:5.1f}%")
