"""Prototype of the convolution schedule planner (evaluation only; the product planner is C++ in csrc/)."""
import sys, numpy as np, itertools
sys.path.insert(0,'oracle')
import rg_oracle as ro

def taylor_cost(jm, J):
    # conv1: j1 < jm-1 : (J+1-j1) FMAs ; conv2: j1 < jm : (J-j1) FMAs (j1=J gives 0)
    c1 = sum(J+1-j1 for j1 in range(0, max(jm-1,0)))
    c2 = sum(max(J-j1,0) if j1>0 else J for j1 in range(0, jm))
    return c1 + c2

def evaluate(L, J, jm, order, lanes=32, nout=1):
    """order: list of output coords (a,b) centred in [-L,L]; consecutive `lanes` per warp."""
    n1 = 2*L+1
    supp = [(a,b) for a in range(-L,L+1) for b in range(-L,L+1) if jm[a+L,b+L] > 0]
    cost = {s: taylor_cost(jm[s[0]+L, s[1]+L], J) for s in supp}
    warps = [order[i:i+lanes] for i in range(0, len(order), lanes)]
    tot_exec = 0; tot_useful = 0; wcost = []
    for w in warps:
        amin = min(a for a,b in w); amax = max(a for a,b in w)
        U = []
        cw = 0
        for s in supp:
            valid = [(abs(a-s[0])<=L and abs(b-s[1])<=L) for a,b in w]
            nv = sum(valid)
            if nv:
                cw += cost[s]; tot_useful += nv*cost[s]
        wcost.append(cw); tot_exec += cw*lanes
    wcost = np.array(wcost)
    return dict(nwarps=len(warps), lane_eff=tot_useful/tot_exec, max=wcost.max(), mean=wcost.mean(), total=wcost.sum(), useful=tot_useful)

def order_boxes(L, bh, bw):
    n1 = 2*L+1
    order = []
    for a0 in range(-L, L+1, bh):
        for b0 in range(-L, L+1, bw):
            for a in range(a0, min(a0+bh, L+1)):
                for b in range(b0, min(b0+bw, L+1)):
                    order.append((a,b))
    return order

if __name__ == "__main__":
    L=int(sys.argv[1]) if len(sys.argv)>1 else 10; J=L
    c = ro.OracleRG(ro.default_params(L=L,J=J))
    Ic = np.roll(c.Iminus,(L,L),axis=(1,2))
    jm = Ic.sum(axis=0)
    dense = np.full_like(jm, J+1)
    for name, jmm in (("sparse", jm), ("dense", dense)):
        for bh,bw in ((4,8),(8,4),(2,16),(3,11),(1,21),(6,6),(5,7),(7,5),(21,1)):
            r = evaluate(L,J,jmm,order_boxes(L,bh,bw))
            nw = r['nwarps']
            print(f"{name} box {bh}x{bw}: warps {nw} lane_eff {r['lane_eff']:.3f} balance(mean/max) {r['mean']/r['max']:.3f} total_iters_cost {r['total']} ideal/warp {r['useful']/32/nw:.0f} max {r['max']}  overall {r['useful']/32/nw/r['max']:.3f}")
