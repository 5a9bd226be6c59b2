"""Bracket speed-up of a team of k CTAs that shares a Hamiltonian piece by piece (elastic
helpers): the pieces of the real L=J=10 half-space schedule (rg_debug_schedule), costed with the planner's issue
model, dealt longest-first to the 12 k warps of the team.  CPU only."""
import os, sys, heapq
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
os.chdir(os.path.join(ROOT, "tests"))
import test_abi as ta
import rg_oracle as ro

J = 10
def tfma(lim):    # DFMAs of a product truncated to `lim` levels of y (rg_plan.hpp taylor_fma)
    return sum(J - j1 + 1 for j1 in range(J) if j1 + 1 < lim) + sum(min(J, J - j1 + 1) for j1 in range(J + 1) if j1 < lim)
def level_class(jm):
    return J + 1 if jm > 6 else (6 if jm > 3 else 3)
c = ro.OracleRG(ro.default_params(L=10, J=10))
s = ta.fetch_schedule(c, 2)
cost_e = np.array([int(2.65 * tfma(level_class(int(e[0]) >> 24))) + 67 for e in s["entries"]])
pieces = [(int(b), int(e), int(box), int(r)) for box, b, e, r in s["pieces"] if box >= 0]
pc = np.array([cost_e[b:e].sum() + 600 for b, e, _, _ in pieces])     # + prologue and combination of a piece
nw = s["nw"]
# the plan's own dealing: warp of a piece = its slot row
rows = [i // s["maxp"] % nw for i, (box, b, e, r) in enumerate(s["pieces"]) if box >= 0]
own = np.zeros(nw)
for w, cst in zip(rows, pc): own[w] += cst
print("pieces %d, longest %.1f %% of a warp's share, plan makespan %.0f (mean %.0f)" % (len(pc), 100 * pc.max() / own.mean(), own.max(), own.mean()))
for k in (1, 2, 4):
    h = [0.0] * (nw * k); heapq.heapify(h)
    for cst in sorted(pc, reverse=True):
        heapq.heappush(h, heapq.heappop(h) + cst)
    print("team of %d: longest-first makespan %.0f -> bracket speed-up x%.2f" % (k, max(h), own.max() / max(h)))
