"""Bracket speed-up of a team of k CTAs that shares a Hamiltonian piece by piece (elastic
helpers), for plans dealt to 12, 24 or 48 (virtual) warps: the pieces of the real L=J=10 half-space plan
(tools/plan_pieces.cpp = rg_plan.hpp on the host), costed with the planner's issue model plus PIECE_OVERHEAD
issue slots per piece (prologue + ordered combination, an assumption to be measured), dealt longest-first to the
12 k warps of the team.  A lone CTA runs (virtual warps / 12) shares back to back on every warp.  CPU only."""
import heapq, os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import rg_oracle as ro

PIECE_OVERHEAD = 600
tmp = tempfile.mkdtemp()
exe = os.path.join(tmp, "plan_pieces")
subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tools", "plan_pieces.cpp")])
c = ro.OracleRG(ro.default_params(L=10, J=10))
np.ascontiguousarray(c.Iminus.reshape(c.J + 1, -1), dtype=np.uint8).tofile(os.path.join(tmp, "iminus.bin"))
base = None
for nw in (12, 24, 48):
    out = subprocess.run([exe, "2", "10", "10", os.path.join(tmp, "iminus.bin"), "1", str(nw)], capture_output=True, text=True).stdout.splitlines()
    pc = np.array([int(l.split()[5]) + PIECE_OVERHEAD for l in out[1:]])
    lone = pc.sum() / 12
    base = base or lone
    line = "plan for %2d warps: %3d pieces, longest %5d, lone CTA %+.1f %%;" % (nw, len(pc), pc.max(), 100 * (lone / base - 1))
    for k in (2, 4):
        h = [0.0] * (12 * k); heapq.heapify(h)
        for x in sorted(pc, reverse=True):
            heapq.heappush(h, heapq.heappop(h) + x)
        line += "  team of %d: bracket x%.2f" % (k, lone / max(h))
    print(line)
