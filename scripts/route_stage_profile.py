#!/usr/bin/env python
"""One hspf_ospfv2_run_area_batch on the C5 LSDB (for `ncu -k regex:route_cells_kernel`):
python scripts/route_stage_profile.py [n_roots]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from holo_b200 import capi, ospfv2, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
t = synth.random_topology(10000, 40000, synth.SEED_BASE + 5, cost_choices=[10, 20], lan_fraction=0.05)
area = ospfv2.synth_area(t, root=0, sr=True)
ctx = capi.Context(0)
rids = np.array([ospfv2.RID_BASE + i for i in range(n)], np.uint32)
ospfv2.run_area_batch(ctx, area, rids[:8])
b = ospfv2.run_area_batch(ctx, area, rids)
P = b.cells.shape[1]
print(f"roots {n} prefixes {P} refused {int((b.status != 0).sum())} spt_batch_ms {b.device_ms[0]:.3f} "
      f"route_kernel_ms {b.device_ms[1]:.3f} cells {b.cells.nbytes / 1e6:.1f} MB -> {b.cells.nbytes / b.device_ms[1] / 1e6:.0f} GB/s")
ctx.close()
