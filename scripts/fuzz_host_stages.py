#!/usr/bin/env python
"""Long-running CPU fuzzes of the product's host stages against the oracle's restatements:
  python scripts/fuzz_host_stages.py {rib|isis|ospf|cells} <seconds>
rib : hspf_ospfv{2,3}_update_rib_full + rib_diff on random multi-area tables
isis: hspf_isis_routes_from_planes (with SR) on random levels
ospf: hspf_ospfv{2,3}_area_from_planes on random areas
cells: the device route kernel's body on the CPU + hspf_ospfv2_routes_from_cells on colliding prefixes
(round 1: 261 k / 149 k / 146 k instances without a mismatch; round 2: 70 k / 91 k / 88 k / 72 k roots, none)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
which = sys.argv[1]
sys.argv = [sys.argv[0]] + sys.argv[2:]

if which == 'rib':
    import sys, time
    import test_ospf_rib as T
    from holo_b200 import ospf_rib
    from oracle import pyoracle
    t_end=time.time()+float(sys.argv[1]); n=0; bad=0; seed=1000
    olib=pyoracle.lib()
    while time.time()<t_end:
        seed+=1
        for v3 in (False, True):
            rid,mp,areas,ext=(T.random_instance_v3 if v3 else T.random_instance)(seed)
            a=(ospf_rib.update_rib_full_v3 if v3 else ospf_rib.update_rib_full)(rid,mp,areas,ext)
            b=(pyoracle.ospfv3_update_rib_full if v3 else pyoracle.ospfv2_update_rib_full)(rid,mp,areas,ext)
            n+=1
            if a.routes.tobytes()!=b.routes.tobytes() or a.nexthops.tobytes()!=b.nexthops.tobytes():
                bad+=1; print('MISMATCH rib',seed,v3,flush=True)
            rid2,mp2,areas2,ext2=(T.random_instance_v3 if v3 else T.random_instance)(seed+7)
            c=(ospf_rib.update_rib_full_v3 if v3 else ospf_rib.update_rib_full)(rid2,mp2,areas2,ext2)
            a0,f0=ospf_rib.rib_diff(None,a,v3=v3)
            fn=olib.oracle_ospfv3_rib_diff if v3 else olib.oracle_ospfv2_rib_diff
            b0,g0=ospf_rib.call_rib_diff(fn,None,a,v3=v3)
            old=ospf_rib.Rib(f0,a.nexthops)
            a1,f1=ospf_rib.rib_diff(old,c,v3=v3); b1,g1=ospf_rib.call_rib_diff(fn,old,c,v3=v3)
            if a0.tobytes()!=b0.tobytes() or f0.tobytes()!=g0.tobytes() or a1.tobytes()!=b1.tobytes() or f1.tobytes()!=g1.tobytes():
                bad+=1; print('MISMATCH diff',seed,v3,flush=True)
    print('fuzzed',n,'instances; mismatches',bad)


if which == 'isis':
    import sys, time, numpy as np
    from holo_b200 import isis, synth
    from oracle import pyoracle
    from isis_synth import synth_instance
    def planes(csr, root):
        c=pyoracle.csr_spf(csr, root, vec_mode=1, nh_words=4); return c['dist'],c['hops']
    rng=np.random.default_rng(99); t_end=time.time()+float(sys.argv[1]); n=bad=0
    while time.time()<t_end:
        V=int(rng.integers(8,90)); E=int(V*rng.uniform(1.5,5))
        kw={}
        r=rng.random()
        if r<0.4: kw['cost_choices']=[int(rng.integers(1,15))]
        elif r<0.7: kw['cost_choices']=[int(x) for x in rng.integers(1,20,2)]
        else: kw['cost_lo'],kw['cost_hi']=1,int(rng.integers(2,40))
        if rng.random()<0.6: kw['lan_fraction']=float(rng.uniform(0.05,0.4))
        try: t=synth.random_topology(V,E,int(rng.integers(0,1<<30)),**kw)
        except Exception: continue
        for root in rng.choice(V, 3, replace=False):
            mt=int(rng.choice([isis.METRIC_WIDE,isis.METRIC_WIDE,isis.METRIC_BOTH]))
            inst=synth_instance(t,int(root),mt,int(rng.integers(0,4)),sr=bool(rng.random()<0.7))
            inst['max_paths']=int(rng.choice([1,2,4,16]))
            a=isis.routes_from_planes(inst, planes); b=pyoracle.isis_compute_routes(inst); n+=1
            if a.routes.tobytes()!=b.routes.tobytes() or a.nexthops.tobytes()!=b.nexthops.tobytes():
                bad+=1; print('MISMATCH',V,E,kw,int(root),mt,flush=True)
    print('fuzzed',n,'instances; mismatches',bad)


if which == 'ospf':
    import sys, time, numpy as np
    from holo_b200 import ospfv2, ospfv3, synth
    from oracle import pyoracle
    def planes(csr, root, nhw):
        c=pyoracle.csr_spf(csr, root, nh_words=nhw); return c['dist'],c['hops'],c['nh_mask']
    rng=np.random.default_rng(7); t_end=time.time()+float(sys.argv[1]); n=bad=skipped=0
    while time.time()<t_end:
        V=int(rng.integers(6,80)); E=int(V*rng.uniform(1.5,5))
        kw={}
        r=rng.random()
        if r<0.4: kw['cost_choices']=[int(rng.integers(1,15))]
        elif r<0.7: kw['cost_choices']=[int(x) for x in rng.integers(1,20,2)]
        else: kw['cost_lo'],kw['cost_hi']=1,int(rng.integers(2,40))
        if rng.random()<0.6: kw['lan_fraction']=float(rng.uniform(0.05,0.4))
        try: t=synth.random_topology(V,E,int(rng.integers(0,1<<30)),**kw)
        except Exception: continue
        for root in rng.choice(V, 3, replace=False):
            a2=ospfv2.synth_area(t,root=int(root),sr=bool(rng.random()<0.6),max_paths=int(rng.choice([1,2,16])))
            ref=pyoracle.ospfv2_run_area(a2)
            try: got=ospfv2.area_from_planes(a2, planes)
            except Exception as e: skipped+=1; continue
            n+=1
            for name in ('vertices','routers','routes','nexthops'):
                if getattr(got,name).tobytes()!=getattr(ref,name).tobytes(): bad+=1; print('MISMATCH v2',name,V,E,kw,int(root),flush=True); break
            a3=ospfv3.synth_area(t,root=int(root),max_links_per_fragment=int(rng.integers(0,4)))
            ref3=pyoracle.ospfv3_run_area(a3); got3=ospfv3.area_from_planes(a3, planes); n+=1
            for name in ('vertices','routers','routes','nexthops'):
                if getattr(got3,name).tobytes()!=getattr(ref3,name).tobytes(): bad+=1; print('MISMATCH v3',name,V,E,kw,int(root),flush=True); break
    print('fuzzed',n,'areas; mismatches',bad,'skipped',skipped)



if which == 'cells':
    # route_cell_eval (the device route kernel's body, run on the CPU through tests/native/route_cells_harness.cc)
    # + hspf_ospfv2_routes_from_cells against the faithful oracle's run_area routes, on LSDBs whose prefixes
    # collide in every way tests/test_ospfv2_route_cells.py::collide knows (round 2: 56 k roots, 19 k refused as
    # mixed-SID, 0 mismatches)
    import ctypes as C, subprocess, time, numpy as np
    import test_ospfv2_route_cells as T
    from holo_b200 import capi, synth
    from holo_b200.build import build_all
    build_all()
    so = os.path.join(ROOT, 'tests', '_build', 'libroute_cells_harness.so')
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.run(['g++', '-O2', '-std=c++17', '-fPIC', '-shared', '-I', os.path.join(ROOT, 'include'), '-o', so,
                    os.path.join(ROOT, 'tests', 'native', 'route_cells_harness.cc')], check=True)
    lib = C.CDLL(so)
    lib.harness_route_cells.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.harness_route_cells16.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    t_end = time.time() + float(sys.argv[1]); seed = 10000; n = bad = refused = 0
    while time.time() < t_end:
        seed += 1
        rng = np.random.default_rng(seed)
        V = int(rng.integers(12, 120))
        try:
            t = synth.random_topology(V, int(V * rng.uniform(1.5, 6)), seed, cost_choices=[int(x) for x in rng.choice([1, 5, 10, 10, 20], 2)],
                                      lan_fraction=float(rng.uniform(0.0, 0.5)))
        except Exception:
            continue
        sr = bool(rng.random() < 0.75); mp = int(rng.choice([1, 2, 3, 16])); ms = int(rng.integers(0, 1 << 30))
        for root in rng.choice(V, min(V, 5), replace=False):
            try:
                area, rt, cells, res, ref = T.check_root(lib, t, int(root), sr=sr, max_paths=mp,
                                                         mutate=lambda ar: T.collide(ar, np.random.default_rng(ms)))
            except AssertionError:
                continue      # the oracle refused the root (more than 64 atoms)
            n += 1
            if res.rc == capi.HSPF_E_UNSUPPORTED:
                refused += 1
                continue
            try:
                T.same_routes(res, ref)
            except AssertionError as e:
                bad += 1; print('MISMATCH', seed, int(root), str(e)[:300], flush=True)
    print('fuzzed', n, 'roots; refused', refused, 'mismatches', bad)
