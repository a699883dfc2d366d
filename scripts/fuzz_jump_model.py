#!/usr/bin/env python
"""Long-running CPU fuzz of the pointer-jumping next-hop algorithm (tests/jump_model.py, the
step-by-step model of kernel phase 3J) against the reference-faithful oracle on random graphs
(sizes 4..120, equal / few / ranged costs, LANs, IS-IS rules): `python scripts/fuzz_jump_model.py
<seconds>`; prints every mismatch and a final count."""
import sys, time, numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from holo_b200 import synth
from oracle import pyoracle
from jump_model import jump_phase
def nh_int(row):
    x=0
    for w,word in enumerate(row): x|=int(word)<<(64*w)
    return x
rng=np.random.default_rng(12345)
t_end=time.time()+float(sys.argv[1]) if len(sys.argv)>1 else time.time()+600
n=0; bad=0
while time.time()<t_end:
    V=int(rng.integers(4,120)); E=int(V*rng.uniform(1.2,6))
    kw={}
    r=rng.random()
    if r<0.3: kw['cost_choices']=[int(rng.integers(1,20))]
    elif r<0.6: kw['cost_choices']=[int(x) for x in rng.integers(1,30,int(rng.integers(2,4)))]
    else: kw['cost_lo'],kw['cost_hi']=1,int(rng.integers(2,60))
    if rng.random()<0.5: kw['lan_fraction']=float(rng.uniform(0.05,0.5))
    seed=int(rng.integers(0,1<<30))
    try:
        t=synth.random_topology(V,E,seed,**kw)
    except Exception as e:
        continue
    isis=bool(rng.random()<0.4)
    csr=synth.topology_csr(t,isis=isis)
    for root in rng.choice(csr.n_vertices, min(csr.n_vertices,12), replace=False):
        ref=pyoracle.csr_spf(csr,int(root),vec_mode=int(isis),nh_words=4)
        if ref['status']!=0: continue
        hops,nh,na,st=jump_phase(csr,int(root),ref['dist'],ref['first_parent'],ref['n_parents'])
        n+=1
        if not np.array_equal(hops,ref['hops']) or nh!=[nh_int(r) for r in ref['nh_mask']]:
            bad+=1; print('MISMATCH',V,E,kw,seed,isis,int(root),flush=True)
print('fuzzed',n,'roots; mismatches',bad)
