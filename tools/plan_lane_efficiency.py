"""Round-2 study: lane efficiency (useful / issued lane-FMAs) of the half-space bracket at
L=J=10 for box shapes and for two or three partitions of the output modes.  The shipped pair (column-major + wide16)
gives 0.8055; the best third partition adds 0.010."""
import sys, itertools
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))
import numpy as np, rg_oracle as ro
L=J=10
c=ro.OracleRG(ro.default_params(L=L,J=J))
n1=2*L+1; M=n1*n1; c0=(M-1)//2
# centred Iminus: jm per centred mode
im=c.Iminus  # [J+1, n1, n1] wrapped
def cen(a): return np.roll(a,(L,L),axis=(-2,-1))
imc=cen(im)
jm=np.zeros((n1,n1),int)
for j in range(J+1): jm[imc[j]]=j+1
def tf(jm_,J):
    cc=0
    for j1 in range(J):
        if j1+1<jm_: cc+=J-j1+1
    for j1 in range(J+1):
        if j1<jm_: cc+=min(J,J-j1+1)
    return cc
def lc(jm_):
    la,lb=3,6
    return J+1 if jm_>lb else (lb if jm_>la else la)
src=[(a,b,jm[a,b]) for a in range(n1) for b in range(n1) if jm[a,b]>0]
outs=[(i//n1,i%n1) for i in range(c0+1)]   # half space
def boxes_from_key(key):
    o=sorted(outs,key=key)
    return [o[i:i+32] for i in range(0,len(o),32)]
def touched(box,s):
    a,b,_=s
    n=0
    for (x,y) in box:
        if 0<=x-a+L<=2*L and 0<=y-b+L<=2*L: n+=1
    return n
def evaluate(partitions):
    issued=0; useful=0
    tabs=[[ [touched(bx,s) for bx in P] for s in src] for P in partitions]
    for si,s in enumerate(src):
        best=min(range(len(partitions)), key=lambda k: sum(1 for v in tabs[k][si] if v>0))
        tv=tabs[best][si]
        nb=sum(1 for v in tv if v>0)
        issued+=nb*32*tf(lc(s[2]),J); useful+=sum(tv)*tf(s[2],J)
    return useful,issued,useful/issued
shapes={
 'colmajor(kind1)': lambda p:(p[1],p[0]),
 'wide16(kind2)': lambda p:(p[1]//16,p[0],p[1]),
 'compact4x8(kind0)': lambda p:(p[1]//8,p[0],p[1]),
 'rowmajor': lambda p:(p[0],p[1]),
 'tall16': lambda p:(p[0]//16,p[1],p[0]),
 '8x4': lambda p:(p[0]//8,p[1]//4,p[0],p[1]),
 '4x8b': lambda p:(p[0]//4,p[1]//8,p[0],p[1]),
 '2x16': lambda p:(p[0]//2,p[1]//16,p[0],p[1]),
 '16x2': lambda p:(p[0]//16,p[1]//2,p[0],p[1]),
}
parts={k:boxes_from_key(f) for k,f in shapes.items()}
for k in parts: print(k, '%.4f'%evaluate([parts[k]])[2])
print('current kind1+kind2', evaluate([parts['colmajor(kind1)'],parts['wide16(kind2)']]))
for a,b in itertools.combinations(parts,2):
    print(a,b,'%.4f'%evaluate([parts[a],parts[b]])[2])
best3=None
for tri in itertools.combinations(parts,3):
    e=evaluate([parts[t] for t in tri])[2]
    if best3 is None or e>best3[0]: best3=(e,tri)
print('best triple',best3)
