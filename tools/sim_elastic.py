"""Scaling of the region grid with elastic helpers: idle CTAs join the Hamiltonians that have run longest once the ticket queue is empty.
Input: profiles/work_map_r01_expm_onestep_256.npz -- liouville evaluations of every point of the config-4 grid
(tools/probe_work.py on one B200).  Model: list scheduling on 148 persistent CTAs per GPU, points sharded i mod G,
65 us per liouville per SM (measured), 30 liouville-equivalents of fixed cost per point."""
import heapq
import os
import numpy as np
W = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "work_map_r01_expm_onestep_256.npz"))["work"].astype(np.int64)
w=W
US=65.0e-6; OVER=30
SPEED = {1: 1.0, 2: 1.75, 4: 2.2}   # per-Hamiltonian speed-up with 2 / 4 CTAs: bracket x1.90 / x2.56 (plan for 24 virtual warps, tools/sim_team.py), 9.5 % of a Lie step on the owner alone; the cluster mode measured 1.44 / 1.87
def sim(ws, nsm=148, elastic=True, dt=2e-3):
    ws=[(x+OVER)*US for x in ws]
    q=0; t=0.0
    jobs=[]  # [remaining, ctas, done]
    idle=nsm
    while q<len(ws) or jobs:
        # hand out tickets
        while idle>0 and q<len(ws):
            jobs.append([ws[q],1,0.0]); q+=1; idle-=1
        if elastic and q>=len(ws) and idle>0:
            # give helpers to the jobs that have run longest so far
            for j in sorted(jobs,key=lambda j:-j[2]):
                if j[1]==1 and idle>=1: j[1]=2; idle-=1
            for j in sorted(jobs,key=lambda j:-j[2]):
                if j[1]==2 and idle>=2: j[1]=4; idle-=2
        # advance
        step=dt
        nxt=min(j[0]/SPEED[j[1]] for j in jobs)
        step=min(step,nxt) if (q>=len(ws)) else min(step,nxt)
        step=max(step,1e-9)
        t+=step
        for j in jobs:
            j[0]-=step*SPEED[j[1]]; j[2]+=step
        fin=[j for j in jobs if j[0]<=1e-12]
        for j in fin: idle+=j[1]
        jobs=[j for j in jobs if j[0]>1e-12]
    return t
ideal=lambda G:(w.sum()+OVER*65536)*US/148/G
for G in (2,4,8):
    t0=max(sim(w[r::G],elastic=False) for r in range(G))
    t1=max(sim(w[r::G],elastic=True) for r in range(G))
    print(G,'ticket %.2f eff %.3f | elastic %.2f eff %.3f'%(t0,ideal(G)/t0,t1,ideal(G)/t1))
