#!/usr/bin/env python
"""End-to-end tour of the OSPFv2 path on one B200 (python examples/whatif_ospfv2.py):

 1. an LSDB image (here synthetic: 2 000 routers, LANs, SR) -> flatten -> upload
 2. one batch: the SPT of EVERY router of the area as root, 16-bit planes staying in HBM
 3. the intra-area route table of every root on the device (route cells), the local router's cells
    decoded into interface next hops and SR labels, compared with the single-root LSDB-level call
 4. a what-if batch for the local router: job j raises the cost of adjacency j (per-job overrides)
 5. an interface cost change: the flat update names the changed CSR edges, the device image is patched in place

Everything goes through the C ABI of include/*.h (ctypes twins in holo_b200/); nothing here touches oracle/.
Each step is what a GPU test checks against the oracle (tests/test_ospfv2_gpu.py, tests/test_engine_gpu.py); this
script itself is a tour, written where no GPU was available, and is not part of the suite."""
import copy
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from holo_b200 import capi, ospfv2, synth  # noqa: E402
from holo_b200.build import build_all  # noqa: E402

build_all()
t = synth.random_topology(2000, 8000, synth.SEED_BASE + 7, cost_choices=[10, 20], lan_fraction=0.05)
area = ospfv2.synth_area(t, root=0, sr=True)                 # the LSDB + router 0's interfaces and neighbours
ctx = capi.Context(0)

# 1. flatten + upload
flat = ospfv2.Flat(area)
g = ctx.upload(flat.csr)
info = ctx.graph_info(g)
print(f"graph: {info['V']} vertices, {info['E']} edges, fast path: {info['fast_path']}")

# 2 + 3. every router as root: SPTs and route cells in one device batch
rids = flat.ids[flat.is_router.astype(bool)]
t0 = time.perf_counter()
b = ospfv2.run_area_batch(ctx, area, rids)
dt = time.perf_counter() - t0
rt = ospfv2.RouteTable(flat)
print(f"{len(rids)} roots x {rt.n_prefixes} prefixes: SPT batch {b.device_ms[0]:.2f} ms, route kernel {b.device_ms[1]:.3f} ms, "
      f"whole call {1e3 * dt:.0f} ms; refused roots: {int((b.status != 0).sum())}")
j = int(np.nonzero(rids == area.router_id)[0][0])
mine = ospfv2.routes_from_cells(area, rt, b.cells[j], *b.gather(j))
single = ospfv2.run_area(ctx, area)                          # flatten + one SPF + host route stage
keep = [n for n in mine.routes.dtype.names if n != "nh_off"]
assert mine.rc == 0 and np.array_equal(mine.routes[keep], single.routes[keep])
print(f"local router: {len(mine.routes)} routes, {int((mine.routes['n_nh'] > 1).sum())} with ECMP, "
      f"{int(mine.routes['has_sr_label'].sum())} with an SR label; identical to hspf_ospfv2_run_area")

# 4. what-if: job j multiplies the cost of the j-th edge out of the local router by 10
root = flat.router_vertex(area.router_id)
e0, e1 = int(flat.csr.row_ptr[root]), int(flat.csr.row_ptr[root + 1])
ov = [[(e, int(flat.csr.cost[e]) * 10)] for e in range(e0, e1)]
res = ctx.run(g, [root] * len(ov), overrides=ov)
base = ctx.run(g, [root])
moved = [(int((res.dist[k] != base.dist[0]).sum())) for k in range(len(ov))]
print(f"what-if: {len(ov)} scenarios; vertices whose distance moves per scenario: {moved}")

# 5. interface cost change: the Router-LSA of the local router comes back with other metrics
new = copy.deepcopy(area)
i = int(np.nonzero(new.router_lsas["adv_rtr"] == area.router_id)[0][0])
lo, n = int(new.router_lsas["link_off"][i]), int(new.router_lsas["n_links"][i])
for k in range(lo, lo + n):
    if new.links["link_type"][k] != ospfv2.LINK_STUB:
        new.links["metric"][k] = 33
kind, edges, costs = ospfv2.flat_update(flat, new, [(area.router_id, area.router_id, 0, 1, 0, (0, 0))])
assert kind == ospfv2.FLAT_COSTS
ctx.update_costs(g, edges, costs)                           # 4-byte costs, not the graph image
after = ctx.run(g, [root])
fresh = ospfv2.run_area(ctx, new)                           # the full path on the new LSDB, for comparison
assert np.array_equal(after.dist[0][after.dist[0] != 0xFFFFFFFF], fresh.vertices["distance"])
print(f"cost change: {len(edges)} CSR edges patched in place; SPT equals a fresh flatten + upload")
g.free()
ctx.close()
