#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` export of spf_quad_kernel by source line and phase.

  cuobjdump -xelf all holo_b200/lib/libholo_spf.so            # -> hspf_capi.sm_100a.cubin
  nvdisasm -g -c hspf_capi.sm_100a.cubin > dis.txt
  ncu -i gpurun_out/<tag>_full.ncu-rep --page source --csv > src.csv
  python scripts/ncu_by_line_quad.py dis.txt src.csv [kernel-substring]
(the library must be the build that was profiled: -lineinfo joins SASS to source lines by position)
"""
import collections
import csv
import re
import sys
from pathlib import Path

dis_path, csv_path = sys.argv[1], sys.argv[2]
kname = sys.argv[3] if len(sys.argv) > 3 else "spf_quad_kernelILi512ELb0"
SRC = Path(__file__).resolve().parent.parent / "holo_b200" / "csrc" / "spf_quad.cuh"

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
N = int(sys.argv[4]) if len(sys.argv) > 4 else 45
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:N]:
    f, ln = key if key else ("?", 0)
    text = src[ln - 1].strip()[:88] if f.endswith("spf_quad.cuh") and 0 < ln <= len(src) else f
    print(f"{100 * a[0] / tot[0]:5.1f}% samp {100 * a[1] / tot[1]:5.1f}% inst  thr/inst {a[2] / max(a[1], 1):4.1f}  L{ln:<4d} {text}")


def find(t):
    return [i + 1 for i, l in enumerate(src) if t in l][0]


marks = [("prolog / job fetch / init", "spf_quad_kernel(const QuadArgs a)"),
         ("sssp: relax (lambda)", "auto relax = "),
         ("sssp: bucket loop, chunk claim, emission", "constexpr uint32_t kWarps_"),
         ("sssp: expansion loop", "expand: one quad per lane, two quads of a lane in flight"),
         ("sssp: idle / bucket end", "clean = 0;\n"),
         ("h0 / seeds", "hops-0 non-HOP heads of root edges"),
         ("parents", "= phase 2: ECMP parents"),
         ("jump: hops", "= phase 3: pointer jumping ="),
         ("jump: next hops", "// -- next hops."),
         ]
marks = [m for m in marks if any(m[1] in l for l in src)]
pos = sorted([(n, find(t)) for n, t in marks], key=lambda x: x[1]) + [("end", len(src) + 1)]
print("---- by phase")
for (name, a), (_, b) in zip(pos, pos[1:]):
    s = [0, 0, 0]
    for key, v in agg.items():
        if key and key[0].endswith("spf_quad.cuh") and a <= key[1] < b:
            for i in range(3):
                s[i] += v[i]
    print(f"{name:30s} samples {100 * s[0] / tot[0]:5.1f}%  inst {100 * s[1] / tot[1]:5.1f}%  thr/inst {s[2] / max(s[1], 1):.1f}")
oth = [0, 0, 0]
for key, v in agg.items():
    if not (key and key[0].endswith("spf_quad.cuh")):
        for i in range(3):
            oth[i] += v[i]
print(f"{'inlined headers':30s} samples {100 * oth[0] / tot[0]:5.1f}%  inst {100 * oth[1] / tot[1]:5.1f}%")
