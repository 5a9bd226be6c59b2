"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/rg_b200.h declares, the ctypes structs match the C layout, compute calls fail loudly without
a GPU, and the host-side convolution schedule (rg_debug_schedule) replayed in numpy reproduces the
oracle's Poisson bracket."""
import ctypes
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import ROOT, golden, relmax

import rg_oracle as ro
from rg_b200 import _lib


def header_functions():
    src = open(os.path.join(ROOT, "include", "rg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.findall(r"^(?:int|void|const char\*)\s+(rg_\w+)\s*\(", src, flags=re.M)


def test_library_exports_every_declared_symbol():
    names = header_functions()
    assert len(names) >= 18
    L = _lib.lib()
    for n in names:
        assert hasattr(L, n), "librg_b200.so does not export %s" % n
        assert n in _lib.SIGNATURES, "no ctypes signature for %s" % n
    assert sorted(names) == sorted(_lib.SIGNATURES)


def test_struct_layouts_match_c():
    prog = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "rg_b200.h"
    int main(void) { printf("%zu %zu %zu %zu\n", sizeof(rg_params), sizeof(rg_stats),
                            offsetof(rg_params, omega0_nu), offsetof(rg_stats, kernel_ms)); return 0; }
    '''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        out = subprocess.check_output([os.path.join(d, "t")]).split()
    assert int(out[0]) == ctypes.sizeof(_lib.RGParams)
    assert int(out[1]) == ctypes.sizeof(_lib.RGStats)
    assert int(out[2]) == _lib.RGParams.omega0_nu.offset
    assert int(out[3]) == _lib.RGStats.kernel_ms.offset


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point must fail loudly (never fall back to the CPU)."""
    if _lib.device_count() > 0:
        pytest.skip("a GPU is visible")
    from rg_b200 import RG, make_param_dict
    c = RG(make_param_dict(L=5, J=5, SaveData=False))
    h = c.generate_1Hamiltonian  # building the host object is fine ...
    with pytest.raises(_lib.RGError):
        h([0.01, 0.02], symmetric=True)          # ... but sym() runs on the device
    with pytest.raises(_lib.RGError):
        c.norm(np.zeros(c.r_jl))
    with pytest.raises(_lib.RGError):
        c.points(np.zeros((4, 2)))


def product_imports_no_oracle():
    for dp, _, fns in os.walk(os.path.join(ROOT, "rg_b200")):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".hpp")):
                txt = open(os.path.join(dp, fn)).read()
                assert "rg_oracle" not in txt and "oracle/" not in txt, fn


def test_product_never_touches_the_oracle():
    product_imports_no_oracle()


# ------------------------------------------------------------------------------------------ schedule replay
def fetch_schedule(c, mode):
    """mode: 0 sparse (I^- support), 1 dense, 2 sparse half-space."""
    L_ = _lib.lib()
    nw, nb, ne, ns, nst, mp = (ctypes.c_int32() for _ in range(6))
    lane_out, pieces, supp = (ctypes.POINTER(ctypes.c_int32)() for _ in range(3))
    entries = ctypes.POINTER(ctypes.c_uint32)()
    cost = (ctypes.c_int64 * 4)()
    im = np.ascontiguousarray(c.Iminus.reshape(-1), dtype=np.uint8)
    _lib.check(L_.rg_debug_schedule(c.dim, c.L, c.J, im.ctypes.data, int(mode), ctypes.byref(nw), ctypes.byref(nb),
                                    ctypes.byref(lane_out), ctypes.byref(pieces), ctypes.byref(ne), ctypes.byref(entries),
                                    ctypes.byref(ns), ctypes.byref(supp), cost, ctypes.byref(nst), ctypes.byref(mp)))
    M = (2 * c.L + 1) ** c.dim
    MAXP = mp.value
    out = dict(nw=nw.value, nb=nb.value, ns=ns.value, nst=nst.value, cost=list(cost), maxp=MAXP,
               lane_out=np.ctypeslib.as_array(lane_out, (nb.value * 32,)).copy(),
               pieces=np.ctypeslib.as_array(pieces, (nst.value * nw.value * MAXP, 4)).copy(),
               entries=np.ctypeslib.as_array(entries, (max(ne.value, 1), 2)).copy()[:ne.value],
               supp=np.ctypeslib.as_array(supp, (M,)).copy())
    for p in (lane_out, pieces, entries, supp):
        L_.rg_debug_free(p)
    # Taylor levels of y present per centred mode (1 + highest I^- level, 0 = not in the support; dense: J + 2)
    im_c = np.roll(c.Iminus, c.dim * (c.L,), axis=tuple(range(1, c.dim + 1))).reshape(c.J + 1, -1)
    jm = np.where(im_c.any(axis=0), c.J + 1 - np.argmax(im_c[::-1], axis=0), 0)
    out["jm"] = np.full(M, c.J + 2) if mode == 1 else jm
    # word 0 of an entry: the range mask of the box's lanes; product-rule plans (BIG kernels) are flagged
    runs, nslots, pr = ctypes.POINTER(ctypes.c_int32)(), ctypes.c_int32(), ctypes.c_int32()
    _lib.check(L_.rg_debug_piece_runs(c.dim, c.L, c.J, im.ctypes.data, int(mode), ctypes.byref(runs), ctypes.byref(nslots), ctypes.byref(pr)))
    L_.rg_debug_free(runs)
    out["prule"] = bool(pr.value)
    return out


def centred(c, a):
    """reference layout [j][wrapped...] -> [centred flat nu][j]"""
    L = c.L
    a = np.roll(a, c.dim * (L,), axis=tuple(range(1, c.dim + 1)))
    return a.reshape(c.J + 1, -1).T.copy()


def replay(c, sched, y1, y2, g, Omega_nu):
    """Execute the schedule exactly as conv_all does (lane = output mode, warp-uniform nu1 list)."""
    L, J, dim, n1 = c.L, c.J, c.dim, 2 * c.L + 1
    M = n1 ** dim
    Y1, Y2, G = centred(c, y1), centred(c, y2), centred(c, g)
    Om = centred(c, np.broadcast_to(Omega_nu, (c.J + 1,) + Omega_nu.shape[1:]))[:, 0]
    B1 = Om[:, None] * G                                   # Omega_nu * g
    B2 = np.zeros_like(G)
    B2[:, :J] = G[:, 1:] * np.arange(1, J + 1)             # dA g
    B1 = np.vstack([B1, np.zeros((1, J + 1))])             # the all-zero record
    B2 = np.vstack([B2, np.zeros((1, J + 1))])
    rec_of = sched["supp"]
    nu_of_rec = np.zeros(sched["ns"], dtype=int)
    nu_of_rec[rec_of[rec_of >= 0]] = np.nonzero(rec_of >= 0)[0]
    strides = [n1 ** (dim - 1 - d) for d in range(dim)]
    coords = np.stack([(np.arange(M) // strides[d]) % n1 for d in range(dim)], axis=1)   # biased coords
    out = np.zeros((M, J + 1))
    written = np.zeros(M, dtype=int)
    MAXP = sched["maxp"]
    for sw in range(sched["nst"] * sched["nw"]):      # (stage, warp)
        seen_cont = False
        for q in range(MAXP):
            box, b, e, rank = sched["pieces"][sw * MAXP + q]
            if box < 0:
                continue
            assert not seen_cont, "a continuation must be the last piece of its warp"
            seen_cont = rank > 0
            lanes = sched["lane_out"][box * 32:(box + 1) * 32]
            acc = np.zeros((32, J + 1))
            for ent in sched["entries"][b:e]:
                mask, rec = int(ent[0]), int(ent[1]) >> 16
                nu1 = nu_of_rec[rec]
                assert (int(ent[1]) & 0xffff) == nu1
                jm = int(sched["jm"][nu1])
                for l, nu in enumerate(lanes):
                    if nu < 0:
                        assert not (mask >> l) & 1, "an idle lane is never in range"
                        continue
                    u = coords[nu] - coords[nu1] + L
                    ok = bool(((u >= 0) & (u <= 2 * L)).all())
                    assert ok == bool((mask >> l) & 1), "word 0 of an entry is the box's range mask"
                    gi = int((u * strides).sum()) if ok else M
                    for j1 in range(J + 1):
                        if j1 + 1 < jm:                  # y1[j1] = (j1+1) y[j1+1]; jm = J+2 keeps a caller's y1[J]
                            a = Y1[nu1, j1]
                            acc[l, j1:J + 1] += a * B1[gi, :J + 1 - j1]
                        if j1 < jm:
                            a = Y2[nu1, j1]
                            acc[l, j1:J + 1] -= a * B2[gi, :J + 1 - j1]
            for l, nu in enumerate(lanes):
                if nu >= 0:
                    out[nu] += acc[l]
                    written[nu] += (rank == 0 and sw < sched["nw"])
    owned = np.zeros(M, dtype=int)
    nb1 = sched["nb"] // sched["nst"]
    owned[sched["lane_out"][:nb1 * 32][sched["lane_out"][:nb1 * 32] >= 0]] = 1
    assert (written == owned).all(), "every owned output mode must have exactly one head piece in the storing stage"
    # back to the reference layout
    res = out.T.reshape((J + 1,) + dim * (n1,))
    return np.roll(res, dim * (-L,), axis=tuple(range(1, dim + 1)))


@pytest.mark.parametrize("name,over,mode", [("ops_golden_L5J5", dict(L=5, J=5), 0),
                                            ("ops_golden_L5J5", dict(L=5, J=5), 1),
                                            ("ops_golden_L5J5", dict(L=5, J=5), 2),
                                            ("ops_golden_L7J4", dict(L=7, J=4), 2),
                                            ("ops_spiral_L4J4", dict(L=4, J=4, torus="spiral"), 0),
                                            ("ops_spiral_L4J4", dict(L=4, J=4, torus="spiral"), 2)])
def test_schedule_replay_matches_oracle(name, over, mode):
    g = golden(name)
    c = ro.OracleRG(ro.default_params(**over))
    c.Omega_nu = g["op_Omega_nu"].copy()
    sched = fetch_schedule(c, mode)
    want_lanes = list(range(c.M)) if mode != 2 else list(range((c.M + 1) // 2))   # half space: centred index <= centre
    nb1 = sched["nb"] // sched["nst"]
    for st in range(sched["nst"]):
        lanes = sched["lane_out"][st * nb1 * 32:(st + 1) * nb1 * 32]
        assert sorted(lanes[lanes >= 0]) == want_lanes, "every output mode owned by exactly one lane of every stage"
    got = replay(c, sched, g["op_y1"], g["op_y2"], g["op_g"], g["op_Omega_nu"])
    yg = (0.0, g["op_y1"], g["op_y2"], g["op_y3"])
    want = c.liouville(yg, g["op_g"])          # y0 = 0: only the two convolutions
    if mode == 2:
        keep = np.zeros(c.M, dtype=bool)
        keep[:(c.M + 1) // 2] = True
        keep = np.roll(keep.reshape(c.dim * (2 * c.L + 1,)), c.dim * (-c.L,), axis=tuple(range(c.dim)))
        got, want = got * keep, want * keep
    assert relmax(got, want) < 1e-13


def test_schedule_geometry_config4():
    """The L=J=10 schedule: 8 warps (255 registers each, rgdev::KernelShape), I^- support of 252 modes."""
    c = ro.OracleRG(ro.default_params(L=10, J=10))
    for mode, nb in ((0, 14), (2, 7)):
        s = fetch_schedule(c, mode)
        assert s["nw"] == 8 and s["nst"] == 2 and s["nb"] == 2 * nb and s["ns"] == 252
        useful, issued, cmax, ctot = s["cost"]
        assert cmax * s["nw"] <= 1.04 * ctot, "per-warp list costs are balanced to 4%"
        # issued counts the zero-padded levels of the near-full entries that run the branch-free path
        assert useful / issued > 0.75, "two output partitions keep more than 75% of the issued lane-FMAs useful"
    # the half-space schedule does half the work of the full one
    full, half = fetch_schedule(c, 0)["cost"], fetch_schedule(c, 2)["cost"]
    assert abs(half[0] / full[0] - 0.5) < 0.01


def test_schedule_geometry_config5():
    """Spiral mean, n = 3, L = J = 6 (BASELINE config 5): 2197 modes = 69 boxes per partition on 12 warps -- more
    boxes than the four default piece slots per warp -- and two partitions (lines along the last axis, blocks of
    8 x 4 x 1).  Every source mode lives in exactly one stage, appears there exactly once in every box that has a
    lane in range for it, pieces of a box tile its list, and the half-space schedule does half the work."""
    c = ro.OracleRG(ro.default_params(L=6, J=6, torus="spiral"))
    L, n1, M = 6, 13, 13 ** 3
    coords = np.stack([(np.arange(M) // n1 ** (2 - d)) % n1 for d in range(3)], axis=1)
    cost = {}
    for mode, nout in ((0, M), (2, (M + 1) // 2)):
        s = fetch_schedule(c, mode)
        nb = (nout + 31) // 32
        assert s["nw"] == 12 and s["nst"] == 2 and s["nb"] == 2 * nb and s["maxp"] > 4
        lanes = s["lane_out"]
        for st in range(2):
            ls = lanes[st * nb * 32:(st + 1) * nb * 32]
            assert sorted(ls[ls >= 0]) == list(range(nout))
        seen = {}
        for sw in range(2 * s["nw"]):
            conts = 0
            for q in range(s["maxp"]):
                box, b, e, rank = s["pieces"][sw * s["maxp"] + q]
                if box < 0:
                    continue
                assert box // nb == sw // s["nw"], "a stage's pieces name that stage's boxes"
                conts += rank > 0
                seen.setdefault(int(box), []).append((int(rank), int(b), int(e)))
            assert conts <= 1, "at most one continuation per warp and stage"
        assert set(range(nb)) <= set(seen), "every box of the storing stage has a head"
        supp = np.nonzero(s["supp"] >= 0)[0]
        stage_of = {}
        for box, plist in seen.items():
            plist.sort()
            assert [r for r, _, _ in plist] == list(range(len(plist)))
            ents = np.concatenate([s["entries"][b:e] for _, b, e in plist])
            nu1 = (ents[:, 1] & 0xffff).astype(int)
            assert len(set(nu1)) == len(nu1)
            for v in nu1:
                assert stage_of.setdefault(int(v), box // nb) == box // nb, "a source mode is visited in one stage only"
        assert sorted(stage_of) == sorted(supp)
        for box in range(2 * nb):
            out = lanes[box * 32:(box + 1) * 32]
            out = out[out >= 0]
            u = coords[out][:, None, :] - coords[supp][None, :, :] + L          # [lane, source, d]
            need = ((u >= 0) & (u <= 2 * L)).all(axis=2).any(axis=0)
            mine = np.array([stage_of[int(v)] == box // nb for v in supp])
            have = set()
            for _, b, e in seen.get(box, []):
                have |= set((s["entries"][b:e, 1] & 0xffff).astype(int))
                # word 0 of an entry is the range mask of this box's 32 lanes (the kernel has no test of its own)
                assert s["prule"]
                all_l = lanes[box * 32:(box + 1) * 32]
                act = all_l >= 0
                v1 = (s["entries"][b:e, 1] & 0xffff).astype(int)
                uu = coords[np.where(act, all_l, 0)][:, None, :] - coords[v1][None, :, :] + L       # [lane, entry, d]
                ok = ((uu >= 0) & (uu <= 2 * L)).all(axis=2) & act[:, None]
                want_mask = (ok.astype(np.uint64) << np.arange(32, dtype=np.uint64)[:, None]).sum(axis=0).astype(np.uint32)
                assert np.array_equal(want_mask, s["entries"][b:e, 0].astype(np.uint32)), (box, b, e)
            assert have == set(supp[need & mine]), box
        cost[mode] = s["cost"]
        assert s["cost"][0] / s["cost"][1] > 0.57, "two partitions + level classes: useful/issued = lane efficiency (0.65) x 54/60 (this size runs the BIG kernel: product-rule form, 60 DFMAs per complete pair at J = 6 for 54 algorithmic ones)"
    assert abs(cost[2][0] / cost[0][0] - 0.5) < 0.01


def test_gathers_of_config4_are_bank_conflict_free():
    """A quarter-warp's LDS.128 over the gathered records G[nu - nu1] is conflict free iff the eight record indices
    are distinct mod 8 (odd record stride in 16-byte units).  nu - nu1 shifts every lane of a box alike, so the lanes
    of every quarter of every box -- the partial last box of a stage included -- must carry eight distinct residues
    of the centred index (rg_plan.hpp: residue-balanced boxes, sorted positions dealt to the quarters)."""
    c = ro.OracleRG(ro.default_params(L=10, J=10))
    for mode in (0, 2):
        s = fetch_schedule(c, mode)
        lanes = s["lane_out"].reshape(-1, 4, 8)
        for box in range(lanes.shape[0]):
            for q in range(4):
                act = lanes[box, q][lanes[box, q] >= 0]
                assert len(set(act & 7)) == len(act), (mode, box, q, sorted(act & 7))


def test_sass_tune_round_trip(tmp_path):
    """rg_b200/csrc/sass_tune.py (run by the Makefile on the linked library) on a copy of an object file of the build, which
    is still as ptxas wrote it: it finds yield-hinted DFMAs in the hot loops, patches control bits only, and its own
    re-disassembly check accepts the result.  On the linked (already tuned) library it finds nothing left to do."""
    import shutil
    import sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    obj = os.path.join(ROOT, "rg_b200", "csrc", "build", "inst_10_0.o")
    if not os.path.isfile(obj):
        pytest.skip("no object files (library built elsewhere)")
    so = tmp_path / "inst_10_0_copy.o"
    shutil.copy(obj, so)
    before = so.read_bytes()
    tool = os.path.join(ROOT, "rg_b200", "csrc", "sass_tune.py")
    r = subprocess.run([sys.executable, tool, str(so)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"sass_tune: (\d+) DFMA retuned", r.stdout)
    assert m and int(m.group(1)) >= 20, r.stdout
    after = so.read_bytes()
    assert len(after) == len(before)
    diff = [i for i in range(len(before)) if before[i] != after[i]]
    assert 0 < len(diff) <= 2 * int(m.group(1)), "only control bytes of the retuned instructions may change"
    # a second pass finds nothing left to do in the kernels it already tuned
    r2 = subprocess.run([sys.executable, tool, str(so), "--dry-run"], capture_output=True, text=True)
    assert re.search(r"sass_tune: 0 DFMA retuned", r2.stdout), r2.stdout
    # the linked library was tuned by the Makefile (SASS_TUNE=1 is the default)
    r3 = subprocess.run([sys.executable, tool, _lib.SO_PATH, "--dry-run"], capture_output=True, text=True)
    assert re.search(r"sass_tune: 0 DFMA retuned", r3.stdout), r3.stdout


def test_hot_loop_canary():
    """ptxas allocates the hot loop of conv_all together with the rest of the kernel: an edit anywhere in rg_device.cuh can
    change its schedule by several per cent either way (the epilogue's chunk size moved the bracket from 107 k to
    117 k clocks).  Static canary on the built L=J=10 object: the complete-product loop body has 130 DFMAs, no more than 45
    other instructions, no spills, a stall-count sum below 305 cycles and -- in the linked, retuned library -- at least 78
    adjacent operand-reuse hits.  A failure here means: measure before committing."""
    import shutil
    import sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_sched as S

    def main_block(path):
        funcs = S.disassemble(path, "rg_kernelILi10ELb0")
        blocks = [b for insts in funcs.values() for b in S.hot_blocks(insts)]
        full = [b for b in blocks if sum(1 for r in b if r[1].startswith("DFMA ")) == 130]
        assert len(full) == 1, [len(b) for b in blocks]
        ins = [S.parse(*r) for r in full[0]]
        hits = 0
        for p, x in zip(ins, ins[1:]):
            if p.dfma and x.dfma and any((p.hi >> (57 + s)) & 1 and S.dfma_key(p, s) == S.dfma_key(x, s) for s in (1, 2)):
                hits += 1
        return ins, hits
    ins, hits = main_block(_lib.SO_PATH)
    assert len(ins) - 130 <= 45, len(ins)
    assert sum(max(x.stall, 1) for x in ins) < 305
    assert not any(x.op.startswith(("LDL", "STL")) for x in ins), "the hot loop spills"
    assert hits >= 78, hits


def test_sass_sched_study_tool_keeps_its_hazard_rules(tmp_path):
    """tools/sass_sched.py (a study, not part of the build: DESIGN 4.1) on a copy of the L=J=10 object: it regroups the DFMAs
    of the four hot loop bodies, its independent hazard pass finds nothing (scoreboard waits, covered loads, fixed latencies,
    the two-cycle scoreboard visibility), the patched file disassembles to exactly the planned instructions, and the
    symbolic replay of the loop bodies computes the same registers and issues the same loads as ptxas's order."""
    import shutil
    import sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    obj = os.path.join(ROOT, "rg_b200", "csrc", "build", "inst_10_0.o")
    if not os.path.isfile(obj):
        pytest.skip("no object files (library built elsewhere)")
    so = tmp_path / "inst_10_0_copy.o"
    shutil.copy(obj, so)
    env = dict(os.environ, SCHED_WINDOW="20", SCHED_MAXWAIT="100", SCHED_STRICT="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_sched.py"), str(so), "--only=rg_kernelILi10ELb0", "--verbose"],
                       capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "HAZARD" not in r.stdout and "verification failed" not in r.stdout, r.stdout
    m = re.search(r"sass_sched: (\d+) loop bodies rescheduled, (\d+) left alone", r.stdout)
    assert m and int(m.group(1)) >= 3, r.stdout
    hits = [(int(a), int(b)) for a, b in re.findall(r"adjacent reuse hits (\d+) -> (\d+)", r.stdout)]
    assert all(b > a for a, b in hits), hits
    # data flow: replay ptxas's order and the new one symbolically (registers as hashes of their producers)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_sched as S
    for k, v in env.items():
        os.environ[k] = v
    try:
        funcs = S.disassemble(obj, "rg_kernelILi10ELb0")
        S.LAT_TABLE.update(S.latency_table(funcs))

        def replay(seq, trips=2):
            regs, loads = {}, []
            for it in range(trips):
                for x in seq:
                    if x.dfma:
                        d, a, b, c = (int(re.search(r"R(\d+)", o).group(1)) for o in x.ops)
                        neg = x.ops[1].startswith("-") != x.ops[2].startswith("-")
                        val = lambda r: (regs.get(("R", r), ("in", r)), regs.get(("R", r + 1), ("in", r + 1)))
                        v = hash(("fma", tuple(sorted([hash(val(a)), hash(val(b))])), neg, val(c)))
                        regs[("R", d)], regs[("R", d + 1)] = ("lo", v), ("hi", v)
                    else:
                        src = tuple(sorted((r, hash(regs.get(r, ("in", r)))) for r in x.reads))
                        v = hash((x.addr, src))
                        if S.is_load(x):
                            loads.append((x.addr, it, src))
                        for k2, r in enumerate(sorted(x.writes)):
                            regs[r] = (k2, v)
            return regs, sorted(loads)
        n = 0
        for insts in funcs.values():
            for b in S.hot_blocks(insts):
                res, info = S.schedule(b)
                assert res is not None, info
                before = [S.parse(*rec) for rec in b]
                after = [S.parse(x.addr, x.text, x.lo, x.hi) for x, lo, hi in res]
                assert replay(before) == replay(after)
                n += 1
        assert n >= 3
    finally:
        for k in env:
            if k.startswith("SCHED_"):
                os.environ.pop(k, None)


@pytest.mark.parametrize("dim,L,J", [(1, 3, 2), (1, 20, 4), (2, 2, 3), (2, 4, 6), (2, 6, 5), (2, 9, 4), (3, 2, 2), (3, 3, 4)])
def test_plans_of_random_supports_cover_every_pair_once(dim, L, J):
    """Planner invariants on geometries and I^- supports it was not tuned for (random supports, partial last boxes of
    every size): in every stage the lanes are a permutation of the output modes, and every (output mode, source mode)
    pair whose difference lies in the grid is visited exactly once over both stages."""
    rng = np.random.default_rng(1000 * dim + 10 * L + J)
    n1, M = 2 * L + 1, (2 * L + 1) ** dim

    class C:            # the attributes fetch_schedule reads
        pass
    c = C()
    c.dim, c.L, c.J = dim, L, J
    c.Iminus = rng.random((J + 1,) + dim * (n1,)) < 0.4
    c.Iminus[..., :] |= rng.random(c.Iminus.shape) < 0.05
    strides = [n1 ** (dim - 1 - d) for d in range(dim)]
    coords = np.stack([(np.arange(M) // strides[d]) % n1 for d in range(dim)], axis=1)
    for mode in (0, 2):
        s = fetch_schedule(c, mode)
        nout = M if mode == 0 else (M + 1) // 2
        nb1 = s["nb"] // s["nst"]
        for st in range(s["nst"]):
            lanes = s["lane_out"][st * nb1 * 32:(st + 1) * nb1 * 32]
            assert sorted(lanes[lanes >= 0]) == list(range(nout)), (mode, st)
        supp = np.nonzero(s["supp"] >= 0)[0]
        visits = np.zeros((nout, M), dtype=int)
        for box, b, e, rank in s["pieces"]:
            if box < 0:
                continue
            lanes = s["lane_out"][box * 32:(box + 1) * 32]
            lanes = lanes[lanes >= 0]
            nu1 = (s["entries"][b:e, 1] & 0xffff).astype(int)
            u = coords[lanes][:, None, :] - coords[nu1][None, :, :] + L
            ok = ((u >= 0) & (u <= 2 * L)).all(axis=2)
            for li, nu in enumerate(lanes):
                visits[nu, nu1[ok[li]]] += 1
        u = coords[:nout][:, None, :] - coords[supp][None, :, :] + L
        need = ((u >= 0) & (u <= 2 * L)).all(axis=2)
        want = np.zeros((nout, M), dtype=int)
        want[:, supp] = need
        assert (visits == want).all(), (mode, int(np.abs(visits - want).sum()))


@pytest.mark.parametrize("over,mode,big", [(dict(L=10, J=10), 2, 0), (dict(L=5, J=5), 0, 0), (dict(L=6, J=6, torus="spiral"), 2, 1),
                                           (dict(L=6, J=6, torus="spiral"), 0, 1), (dict(L=20, J=4), 2, 1), (dict(L=4, J=2), 0, 0)])
def test_piece_runs_are_class_sorted(over, mode, big):
    """Every piece's list is sorted by level class -- complete products, then those truncated to LIM_B, then LIM_A -- and
    the run ends the device loops use (rg_debug_piece_runs) cut it exactly there; plans of Hamiltonians too large for
    the two-derivative records (n = 3 at L = J = 6; 2-D at L = 20, J = 4) are costed for the product-rule form."""
    c = ro.OracleRG(ro.default_params(**over))
    s = fetch_schedule(c, mode)
    L_ = _lib.lib()
    runs = ctypes.POINTER(ctypes.c_int32)()
    n, pr = ctypes.c_int32(), ctypes.c_int32()
    im = np.ascontiguousarray(c.Iminus.reshape(-1), dtype=np.uint8)
    _lib.check(L_.rg_debug_piece_runs(c.dim, c.L, c.J, im.ctypes.data, int(mode), ctypes.byref(runs), ctypes.byref(n), ctypes.byref(pr)))
    r = np.ctypeslib.as_array(runs, (n.value, 2)).copy()
    L_.rg_debug_free(runs)
    assert n.value == len(s["pieces"]) and pr.value == big
    J = c.J
    lim_a = 3 if J >= 8 else max(J // 2, 2)
    lim_b = 6 if J >= 8 else J - 1
    def cls(jm):
        return J + 1 if jm > lim_b else (lim_b if jm > lim_a else lim_a)
    seen = 0
    for (box, b, e, rank), (e_full, e_mid) in zip(s["pieces"], r):
        if box < 0:
            continue
        seen += 1
        assert b <= e_full <= e_mid <= e
        jm = s["jm"][(s["entries"][b:e, 1] & 0xffff).astype(int)]
        k = np.array([cls(v) for v in jm])
        assert (k[:e_full - b] == J + 1).all() and (k[e_full - b:e_mid - b] == lim_b).all() and (k[e_mid - b:] == lim_a).all()
    assert seen > 0
