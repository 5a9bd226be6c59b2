"""Scaling of the region grid under static ticket orders, an oracle longest-first order and strided-pilot orders.
Input: profiles/work_map_r01_expm_onestep_256.npz -- liouville evaluations of every point of the config-4 grid
(tools/probe_work.py on one B200).  Model: list scheduling on 148 persistent CTAs per GPU, points sharded i mod G,
65 us per liouville per SM (measured), 30 liouville-equivalents of fixed cost per point."""
import heapq
import os
import numpy as np
W = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "work_map_r01_expm_onestep_256.npz"))["work"].astype(np.int64)
w=W
W=W.reshape(256,256)   # [iy][ix]
US=65.0e-6   # s per liouville per SM
OVER=30      # fixed liouville-equivalents per point (setup)
def makespan(ws, nsm=148):
    h=[0.0]*nsm; heapq.heapify(h)
    end=0
    for x in ws:
        t=heapq.heappop(h); t+= (x+OVER)*US; end=max(end,t); heapq.heappush(h,t)
    return end
def run(order_fn, G):
    ts=[]
    for r in range(G):
        idx=np.arange(r,65536,G)
        ts.append(makespan(order_fn(idx)))
    return max(ts)
ideal=lambda G: (w.sum()+OVER*65536)*US/148/G
for G in (1,2,4,8):
    t=run(lambda idx: w[idx], G)
    tl=run(lambda idx: np.sort(w[idx])[::-1], G)
    print(G,'ticket %.2fs eff %.3f | LPT oracle %.2fs eff %.3f | ideal %.2f'%(t,ideal(G)/t,tl,ideal(G)/tl,ideal(G)))
print('max',w.max(),'mean',w.mean(), 'n>4000',(w>4000).sum(), 'n>10000',(w>10000).sum(),'sum>4000 share',w[w>4000].sum()/w.sum())

from scipy.ndimage import maximum_filter, uniform_filter
def pilot_sim(G, stride, B, mode='max', off=None):
    off = stride//2 if off is None else off
    pil=np.zeros((256,256),bool); pil[off::stride, off::stride]=True
    pw=np.minimum(W,B)       # pilot pass: capped work
    # pass 1: pilot points, cyclic over ranks in ticket order, each capped at B
    pidx=np.flatnonzero(pil.ravel())
    t1=max(makespan(pw.ravel()[pidx[r::G]]) for r in range(G))
    # prediction for all points from pilot values
    pv=np.where(pil, pw, 0).astype(float)
    if mode=='max':
        pred=maximum_filter(pv, size=2*stride-1, mode='nearest')
    else:
        cnt=uniform_filter(pil.astype(float), size=2*stride-1, mode='nearest')
        pred=uniform_filter(pv, size=2*stride-1, mode='nearest')/np.maximum(cnt,1e-9)
    done = pil & (W<=B)       # pilot points finished within budget need no second visit
    rest=np.flatnonzero(~done.ravel())
    # pass 2: remaining points (incl. capped pilot points, known heavy -> first), by predicted work desc
    key=pred.ravel()[rest].copy()
    key[pil.ravel()[rest]] = 1e18
    order=rest[np.argsort(-key, kind='stable')]
    t2=max(makespan(w[order[r::G]]) for r in range(G))
    return t1,t2
for G in (4,8):
    for stride in (2,3,4):
        for B in (1000,3000):
            for mode in ('max','mean'):
                t1,t2=pilot_sim(G,stride,B,mode)
                print(G,stride,B,mode,'pilot %.2f main %.2f total %.2f eff %.3f'%(t1,t2,t1+t2,ideal(G)/(t1+t2)))

print('--- static orders at 8 and 4 GPUs')
idx_all=np.arange(65536); iy=idx_all//256; ix=idx_all%256
orders={'ascending':idx_all,'descending':idx_all[::-1],
        'x-major':np.lexsort((iy,ix)),'x-major desc':np.lexsort((iy,ix))[::-1],
        'diag':np.argsort(ix/255*0.12/0.12+iy/255,kind='stable'),'diag desc':np.argsort(-(ix/255+iy/255),kind='stable')}
rng=np.random.default_rng(1)
for s_ in range(3): orders['random%d'%s_]=rng.permutation(65536)
for G in (4,8):
    for name,o in orders.items():
        t=max(makespan(w[o[r::G]]) for r in range(G))
        print(G,name,'%.2f eff %.3f'%(t,ideal(G)/t))
