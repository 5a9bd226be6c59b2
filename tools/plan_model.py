"""Model of multi-output convolution tilings (evaluation only).
lane = base mode + K output offsets delta_k sharing one gathered record G[base - nu1];
warp-uniform loop over nu1 (virtual nu1 outside the grid allowed: Y[nu1]=0 but Y[nu1+delta_k] may be real)."""
import sys, itertools, numpy as np
sys.path.insert(0, 'oracle')
import rg_oracle as ro

def tfma(jm, J):
    c = sum(J - j1 + 1 for j1 in range(J) if j1 + 1 < jm)
    c += sum(min(J, J - j1 + 1) for j1 in range(J + 1) if j1 < jm)
    return c

def evaluate(L, J, jm, lanes_by_warp, deltas, verbose=False):
    """lanes_by_warp: list of warps, each a list of (base, [valid_k flags]) ; base=(a,b) centred."""
    n1 = 2 * L + 1
    def jmof(a, b):
        return jm[a + L, b + L] if -L <= a <= L and -L <= b <= L else 0
    tot_exec = tot_useful = 0; wcost = []; smem = 0
    K = len(deltas)
    for lanes in lanes_by_warp:
        # union of nu1 over lanes: nu1 = base - nu2, nu2 in range
        U = set()
        for (a, b), flags in lanes:
            for a2 in range(-L, L + 1):
                for b2 in range(-L, L + 1):
                    U.add((a - a2, b - b2))
        cw = 0
        for (a1, b1) in U:
            costs = [tfma(jmof(a1 + d[0], b1 + d[1]), J) for d in deltas]
            if sum(costs) == 0: continue
            nact = 0
            for (a, b), flags in lanes:
                if abs(a - a1) <= L and abs(b - b1) <= L:
                    nact += 1
                    tot_useful += sum(c for c, f in zip(costs, flags) if f)
            if nact == 0: continue
            cw += sum(costs)
            smem += 44 + 11 * sum(1 for c in costs if c > 0)
        wcost.append(cw); tot_exec += cw * 32
    wcost = np.array(wcost)
    return dict(nw=len(lanes_by_warp), lane_eff=tot_useful / tot_exec, bal=wcost.mean() / wcost.max(), useful=tot_useful,
                exec=tot_exec, smem_per_130fma=smem * 130.0 * 32 / tot_exec, overall=tot_useful / (32 * len(wcost) * wcost.max()))

def make_lanes(L, outputs, deltas, box, order_key=None):
    """greedy: bases = outputs not yet covered, scanning in box order; a base covers base+delta_k if in `outputs`."""
    outs = set(outputs); covered = set(); lanes = []
    def key(o):
        a, b = o
        return ((a + L) // box[0], (b + L) // box[1], (a + L) % box[0], (b + L) % box[1])
    for o in sorted(outs, key=key):
        if o in covered: continue
        flags = []
        for d in deltas:
            t = (o[0] + d[0], o[1] + d[1])
            ok = t in outs and t not in covered
            flags.append(ok)
            if ok: covered.add(t)
        lanes.append((o, flags))
    return [lanes[i:i + 32] for i in range(0, len(lanes), 32)]

if __name__ == "__main__":
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 10; J = L
    c = ro.OracleRG(ro.default_params(L=L, J=J))
    jm = np.roll(c.Iminus, (L, L), axis=(1, 2)).sum(axis=0)
    full = [(a, b) for a in range(-L, L + 1) for b in range(-L, L + 1)]
    half = [(a, b) for (a, b) in full if a > 0 or (a == 0 and b >= 0)]
    for oname, outs in (("full", full), ("half", half)):
        for dname, deltas, boxes in (("K1", [(0, 0)], [(4, 8), (8, 4), (3, 11), (6, 6)]),
                                     ("K2b", [(0, 0), (0, 1)], [(4, 16), (8, 8), (2, 22), (16, 4)]),
                                     ("K2a", [(0, 0), (1, 0)], [(8, 8), (16, 4), (4, 16)]),
                                     ("K4", [(0, 0), (0, 1), (1, 0), (1, 1)], [(8, 16), (16, 8), (22, 6), (6, 22)])):
            for box in boxes:
                lw = make_lanes(L, outs, deltas, box)
                r = evaluate(L, J, jm, lw, deltas)
                print(f"{oname} {dname} box{box}: warps {r['nw']} lane_eff {r['lane_eff']:.3f} bal {r['bal']:.3f} overall {r['overall']:.3f} smem/130fma {r['smem_per_130fma']:.1f} useful {r['useful']}")
