"""Round-2 study: boxes of 16 lanes with an independent source list per half-warp (two y
addresses per broadcast load).  Ideal 0.874 of the issued lane-FMAs useful; with the two halves of a warp padded to a
common list per level class 0.805 (adjacent halves) to 0.832 (halves paired by list length): -3 % issued work at best for
a second shared-memory wavefront per y load -- not built."""
import os
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'plan_lane_efficiency.py')).read().split("parts={k:boxes_from_key")[0])
from collections import Counter
def boxes_n(key,n):
    o=sorted(outs,key=key); return [o[i:i+n] for i in range(0,len(o),n)]
def plan(keys, n=16, pairing='adjacent'):
    P=[boxes_n(k,n) for k in keys]
    tabs=[[ [touched(bx,s) for bx in Pk] for s in src] for Pk in P]
    stage_of=[min(range(len(P)), key=lambda k: sum(1 for v in tabs[k][si] if v>0)) for si in range(len(src))]
    useful=0; issued=0; ideal16=0
    for st in range(len(P)):
        comps=[]
        for bi,bx in enumerate(P[st]):
            l=[si for si in range(len(src)) if stage_of[si]==st and tabs[st][si][bi]>0]
            for si in l: useful+=tabs[st][si][bi]*tf(src[si][2],J)
            cnt=Counter(lc(src[si][2]) for si in l)
            comps.append(cnt)
            ideal16+=sum(n*v*tf(c,J) for c,v in cnt.items())
        idx=list(range(len(comps)))
        if pairing=='sorted': idx.sort(key=lambda i:-sum(comps[i].values()))
        if len(idx)%2: idx.append(None)
        for a,b in zip(idx[0::2],idx[1::2]):
            ca=comps[a]; cb=comps[b] if b is not None else Counter()
            for c in (3,6,11): issued+=max(ca.get(c,0),cb.get(c,0))*2*n*tf(c,J)
    return useful,issued,ideal16
for name,keys in {'col+wide16':[lambda p:(p[1],p[0]), lambda p:(p[1]//16,p[0],p[1])],
                  'col+wide8':[lambda p:(p[1],p[0]), lambda p:(p[1]//8,p[0],p[1])],
                  'col+row':[lambda p:(p[1],p[0]), lambda p:(p[0],p[1])],
                  'tall8+wide8':[lambda p:(p[0]//8,p[1],p[0]), lambda p:(p[1]//8,p[0],p[1])],
                  }.items():
    for n in (16,8):
      for pairing in ('adjacent','sorted'):
        if n==8:
            # groups of 4 eighth... skip pairing model: treat quads
            continue
        u,i,i16=plan(keys,n,pairing)
        print(name,n,pairing,'eff %.4f (ideal half-box %.4f)  issued vs current %.4f'%(u/i,u/i16,i/4629120))
