"""Round-2 study: pair helpers for the 8-GPU tail -- a CTA that finds its cluster partner running a
chain older than X liouvilles joins it (speed-up s for the pair) instead of drawing a ticket; `adaptive` = only when the
chain is older than the estimated remaining queue time.  Simulated with the measured work map.  Result: 0.75-0.83 at 8
GPUs for s = 1.44-1.9; the bar of 0.85 is not reachable with pairs (clusters of 4 strand 16 of 148 SMs)."""
import heapq, numpy as np
import os
d=np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'work_map_r01_expm_onestep_256.npz'))
w=d['work'].astype(np.int64)
US=65e-6; OVER=30
def sim(ws, s2, X, adaptive=False, nsm=148, dt=0.5e-3):
    """time-stepped pairs. CTA states: idle/solo/helping. A free CTA joins its partner if partner chain age>=X (liouvilles) [and adaptive cond]."""
    n=len(ws); q=0
    rem=np.zeros(nsm); age=np.zeros(nsm); helped=np.zeros(nsm,bool); helping=np.zeros(nsm,bool)
    t=0.0; done_work=0.0; done_pts=0
    while True:
        # assign free CTAs
        for c in range(nsm):
            if rem[c]>0 or helping[c]: continue
            p=c^1
            cond = rem[p]>0 and not helped[p] and age[p]>=X
            if cond and adaptive:
                avg = (done_work/max(done_pts,1)) if done_pts>20 else 600.0
                cond = age[p] >= (n-q)*avg/nsm
            if cond:
                helped[p]=True; helping[c]=True
            elif q<n:
                rem[c]=ws[q]+OVER; age[c]=0; q+=1
        run=rem>0
        if not run.any(): break
        speed=np.where(helped,s2,1.0)
        step=min(dt, (rem[run]*US/speed[run]).min()); step=max(step,1e-9)
        adv=step/US*speed
        rem[run]-=adv[run]; age[run]+=adv[run]
        t+=step
        fin=run&(rem<=1e-9)
        for c in np.flatnonzero(fin):
            rem[c]=0; done_work+=age[c]; done_pts+=1
            if helped[c]: helped[c]=False; helping[c^1]=False
    return t
ideal=lambda G:(w.sum()+OVER*65536)*US/148/G
import sys
for G in (8,4,1):
    if G==1 and len(sys.argv)<2: continue
    for s in (1.44,1.75,1.9):
        for X in (1000,2000,4000,8000):
            for ad in (False,True):
                t=max(sim(w[r::G],s,X,ad) for r in range(G))
                print('G',G,'s',s,'X',X,'adaptive',ad,'%.3f eff %.3f'%(t,ideal(G)/t))
