#!/usr/bin/env python
"""Freeze golden vectors by running the UNMODIFIED reference (cchandre/RG).

Runs only in the build container (needs /root/reference).  Writes small .npz
fixtures next to this file; they are committed and travel to the GPU box.

    python tests/golden/make_golden.py ops region5 region10 points10 line surface spiral

Every fixture stores the inputs (amplitudes / parameter overrides) it was
generated from, so a test can rebuild the same case without the reference.
"""
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import ref_loader  # noqa: E402

warnings.simplefilter("ignore")
NPROC = int(os.environ.get("GOLDEN_NPROC", str(os.cpu_count())))

_case = None
_counts = None
_pristine = None


def _mk_case(over):
    RG, RG_modules, RG_dict = ref_loader.load()
    return RG.RG(ref_loader.make_params(**over))


def _init_worker(over, seed):
    global _case, _counts
    np.random.seed(seed)
    _case = _mk_case(over)
    global _pristine
    _pristine = (list(_case.AmpInf), list(_case.AmpSup))
    _counts = {"liouville": 0, "generate_y": 0, "rg_map": 0}
    for name in list(_counts):
        orig = getattr(_case, name)

        def wrap(*a, _o=orig, _n=name, **k):
            _counts[_n] += 1
            return _o(*a, **k)
        setattr(_case, name, wrap)


def _point(xy):
    RG, RG_modules, RG_dict = ref_loader.load()
    for k in _counts:
        _counts[k] = 0
    t0 = time.time()
    (conv, count), err = RG_modules.point(xy[0], xy[1], _case)
    return conv, count, err, _counts["liouville"], _counts["generate_y"], _counts["rg_map"], time.time() - t0


def run_points(over, pts, seed=12345):
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(NPROC, initializer=_init_worker, initargs=(over, seed)) as pool:
        res = pool.map(_point, pts, chunksize=1)
    return np.array(res, dtype=np.float64)


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print("wrote", path, {k: np.asarray(v).shape for k, v in arrs.items()}, flush=True)


def region_fixture(name, over, xs, ys):
    pts = [(x, y) for y in ys for x in xs]
    t0 = time.time()
    r = run_points(over, pts)
    wall = time.time() - t0
    ny, nx = len(ys), len(xs)
    save(name, over=repr(over), xs=np.asarray(xs), ys=np.asarray(ys),
         conv=r[:, 0].astype(np.int64).reshape(ny, nx), count=r[:, 1].astype(np.int64).reshape(ny, nx),
         error=r[:, 2].astype(np.int64).reshape(ny, nx), liouvilles=r[:, 3].astype(np.int64).reshape(ny, nx),
         lie_steps=r[:, 4].astype(np.int64).reshape(ny, nx), rg_maps=r[:, 5].astype(np.int64).reshape(ny, nx),
         seconds=r[:, 6].reshape(ny, nx), wall=wall, nproc=NPROC)


def ops_fixture(name, over, amps, kmax=4, with_multiply=True):
    """Operator-level goldens on the Hamiltonian reached from `amps`."""
    RG, RG_modules, RG_dict = ref_loader.load()
    out = {"over": repr(over), "amps": np.asarray(amps, dtype=np.float64)}
    case = _mk_case(dict(over, CT="expm_onestep"))
    h0 = case.generate_1Hamiltonian(list(amps), symmetric=True)
    out["f0"] = h0.f.copy()
    out["Omega0"] = np.asarray(h0.Omega, dtype=np.float64)
    out["Iminus"] = case.Iminus.copy()
    out["omega0_nu"] = case.omega0_nu.copy()
    out["Eigenvalue"] = case.Eigenvalue
    # rg_map^k for every CT
    cts = ["expm_onestep", "expm_adapt"] + (["expm_multiply"] if with_multiply else [])
    for ct in cts:
        np.random.seed(12345)
        c = _mk_case(dict(over, CT=ct))
        h = c.generate_1Hamiltonian(list(amps), symmetric=True)
        fs, oms, errs, onus = [], [], [], []
        for k in range(kmax):
            h = c.rg_map(h)
            fs.append(h.f.copy()); oms.append(np.asarray(h.Omega, float).copy()); errs.append(h.error)
            onus.append(c.Omega_nu.copy())
        out["rgmap_f_" + ct] = np.array(fs)
        out["rgmap_Omega_" + ct] = np.array(oms)
        out["rgmap_err_" + ct] = np.array(errs)
        out["rgmap_Omega_nu_" + ct] = np.array(onus)
    # operator goldens at h2 = rg_map^2(h0) (dense in modes) with the Omega_nu state left by rg_map
    case = _mk_case(dict(over, CT="expm_onestep"))
    h = case.generate_1Hamiltonian(list(amps), symmetric=True)
    h = case.rg_map(case.rg_map(h))
    # de-symmetrise slightly to exercise the general (non-symmetric) path
    rng = np.random.RandomState(7)
    g = h.f * (1.0 + 1e-3 * rng.standard_normal(h.f.shape))
    out["op_f"] = h.f.copy()
    out["op_g"] = g
    out["op_Omega_nu"] = case.Omega_nu.copy()
    y = case.generate_y(h)
    out["op_y0"] = y[0]; out["op_y1"] = y[1]; out["op_y2"] = y[2]; out["op_y3"] = y[3]
    out["op_conv_y1_g"] = case.conv_product(y[1], g)
    out["op_conv_g_g"] = case.conv_product(g, g)
    out["op_liouville_f"] = case.liouville(y, h.f)
    out["op_liouville_g"] = case.liouville(y, g)
    hh = RG.Hamiltonian(h.Omega, g.copy())
    out["op_sym_g"] = case.sym(hh).f
    out["op_norm_int_g"] = case.norm_int(hh)
    for nc in ("sum", "max", "Euclidean", "Analytic"):
        c2 = _mk_case(dict(over, CT="expm_onestep", NormChoice=nc))
        out["op_norm_" + nc] = c2.norm(g)
    e1 = case.expm_onestep(h, y)
    out["op_onestep_f"] = e1.f; out["op_onestep_err"] = e1.error
    e1h = case.expm_onestep(h, y, 0.5)
    out["op_onestep_half_f"] = e1h.f
    e2 = case.expm_adapt(h, y)
    out["op_adapt_f"] = e2.f; out["op_adapt_err"] = e2.error
    if with_multiply:
        np.random.seed(12345)
        out["op_m_s"] = np.array(case.compute_m_s(y))
        out["op_pnorm1"] = case.pnorm(y)
        np.random.seed(12345)
        e3 = case.expm_multiply(h, y)
        out["op_multiply_f"] = e3.f; out["op_multiply_err"] = e3.error
    save(name, **out)


def main(argv):
    what = set(argv) or {"ops"}
    t00 = time.time()
    if "ops" in what:
        ops_fixture("ops_golden_L5J5", dict(L=5, J=5), (0.01, 0.02))
        ops_fixture("ops_golden_L5J5_b", dict(L=5, J=5), (0.03, 0.05), kmax=3)
        ops_fixture("ops_golden_L10J10", dict(L=10, J=10), (0.01, 0.02), kmax=3)
        ops_fixture("ops_golden_L7J4", dict(L=7, J=4), (0.02, 0.03), kmax=3)
    if "oddsizes" in what:
        # any (L, J), not only the four benchmark sizes: two odd ones (RG_dict.py:48-49 leaves both free), operators and
        # rg_map^k plus a small region map each; and the tau-mean torus of RG_dict.py:36-46 (3-D, N with a 5 in it)
        ops_fixture("ops_golden_L6J7", dict(L=6, J=7), (0.02, 0.03), kmax=3)
        ops_fixture("ops_golden_L8J3", dict(L=8, J=3), (0.01, 0.04), kmax=3)
        ops_fixture("ops_golden_L4J12", dict(L=4, J=12), (0.02, 0.02), kmax=2)
        for ct in ("expm_onestep", "expm_multiply"):
            region_fixture("region_L6J7_6x6_" + ct, dict(L=6, J=7, CT=ct), np.linspace(0, 0.06, 6), np.linspace(0.004, 0.08, 6))
        region_fixture("region_L8J3_6x6_expm_adapt", dict(L=8, J=3, CT="expm_adapt"), np.linspace(0, 0.06, 6), np.linspace(0.004, 0.08, 6))
        region_fixture("region_L9J2_5x5_expm_onestep", dict(L=9, J=2, CT="expm_onestep"), np.linspace(0.002, 0.06, 5), np.linspace(0.004, 0.08, 5))
    if "taumean" in what:
        over = dict(torus="taumean", L=3, J=4, CT="expm_onestep")
        ops_fixture("ops_taumean_L3J4", over, (3e-4, 2e-3, 0.01), kmax=3, with_multiply=False)
        # the RG_dict domain AmpInf..AmpSup = [0, 7e-4] x [0, 6e-3] (third amplitude fixed at AmpSup[2] = 0.01, RG_modules.py:122-123)
        region_fixture("region_taumean_L3J4_5x5_expm_onestep", over, np.linspace(0.0, 7e-4, 5), np.linspace(0.0, 6e-3, 5))
    if "ops3" in what:
        ops_fixture("ops_spiral_L4J4", dict(torus="spiral", L=4, J=4), (0.005, 0.02, 0.1), kmax=2)
    if "region5" in what:
        # G-A: 12x12, L=J=5
        for ct in ("expm_onestep", "expm_adapt", "expm_multiply"):
            region_fixture("region_L5J5_12x12_" + ct, dict(L=5, J=5, CT=ct),
                           np.linspace(0, 0.06, 12), np.linspace(0, 0.08, 12))
        # G-B: default domain 8x8
        region_fixture("region_L5J5_8x8_default_expm_onestep", dict(L=5, J=5, CT="expm_onestep"),
                       np.linspace(0, 0.12, 8), np.linspace(0, 0.35, 8))
    if "region10" in what:
        # G-C: 16x16 strided subsample of config 4 (ix,iy in {8,24,...,248} of the Nxy=256 grid)
        xs = np.linspace(0, 0.12, 256)[8::16]
        ys = np.linspace(0, 0.35, 256)[8::16]
        for ct in ("expm_onestep", "expm_multiply", "expm_adapt"):
            region_fixture("region_L10J10_sub16_" + ct, dict(L=10, J=10, CT=ct), xs, ys)
    if "region10b" in what:
        # heavier secondary grid (richer boundary), 12x12 on [0,0.06]x[0,0.08] at L=J=10
        region_fixture("region_L10J10_12x12_expm_onestep", dict(L=10, J=10, CT="expm_onestep"),
                       np.linspace(0, 0.06, 12), np.linspace(0, 0.08, 12))
    if "line" in what or "line10" in what:
        RG, RG_modules, RG_dict = ref_loader.load()
        for L, rd in ((5, 1e-8),) + (((10, 1e-8),) if "line10" in what else ()):
            case = _mk_case(dict(L=L, J=L, CT="expm_onestep", RelDist=rd))
            import io, contextlib
            with contextlib.redirect_stdout(io.StringIO()):
                hl = RG_modules.compute_line(case)
            save("line_L%dJ%d" % (L, L), L=L, RelDist=rd,
                 eps=np.array([[2 * h.f[k] for k in case.K] for h in hl]),
                 counts=np.array([h.count for h in hl]), errors=np.array([h.error for h in hl]),
                 f0=hl[0].f, f1=hl[1].f)
    if "surface" in what:
        import multiprocessing as mp
        ctx = mp.get_context("fork")
        over = dict(L=5, J=5, CT="expm_onestep", RelDist=1e-5)
        eps = np.linspace(0, 1, 8)
        with ctx.Pool(NPROC, initializer=_init_worker, initargs=(over, 12345)) as pool:
            res = pool.map(_cr, list(eps), chunksize=1)
        save("surface_L5J5_n8", over=repr(over), eps=eps, data=np.array(res, dtype=np.float64).T)
    if "surface10" in what:
        # BASELINE config 3 at a checkable size: 6 of the 128 columns of the Nxy=128, L=J=10 critical surface
        import multiprocessing as mp
        ctx = mp.get_context("fork")
        over = dict(L=10, J=10, CT="expm_onestep", RelDist=1e-7)
        eps = np.linspace(0, 1, 128)[[0, 13, 40, 77, 101, 127]]
        with ctx.Pool(NPROC, initializer=_init_worker, initargs=(over, 12345)) as pool:
            res = pool.map(_cr, list(eps), chunksize=1)
        save("surface_L10J10_cols", over=repr(over), eps=eps, data=np.array(res, dtype=np.float64).T)
    if "spiral" in what:
        over = dict(torus="spiral", L=6, J=6, CT="expm_multiply")
        pts = [(0.005, 0.02), (0.02, 0.1), (0.05, 0.25)]
        r = run_points(over, pts)
        save("points_spiral_L6J6_expm_multiply", over=repr(over), pts=np.array(pts), res=r)
        over = dict(torus="spiral", L=4, J=4, CT="expm_onestep")
        region_fixture("region_spiral_L4J4_6x6_expm_onestep", over,
                       np.linspace(0, 0.05, 6), np.linspace(0, 0.25, 6))
    if "iterates" in what:
        # BASELINE config 1: compute_iterates, L=J=5, expm_adapt (RG_dict defaults otherwise).  The reference dies in
        # case.eigenvalues after the loop (quirk Q1): replace it, and capture what save_data would have written.
        RG, RG_modules, RG_dict = ref_loader.load()
        import io, contextlib
        case = _mk_case(dict(L=5, J=5, CT="expm_adapt", RelDist=1e-7, Iterates=4))
        case.eigenvalues = lambda *a, **k: None
        captured = {}
        orig_save = RG_modules.save_data
        RG_modules.save_data = lambda name, data, timestr, case_, info=[], display=False: captured.update(data=np.array(data))
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                RG_modules.compute_iterates(case)
        finally:
            RG_modules.save_data = orig_save
        save("iterates_L5J5_adapt", over=repr(dict(L=5, J=5, CT="expm_adapt", RelDist=1e-7, Iterates=4)), data=captured["data"])
    if "norms" in what:
        # the three other norms of RG.py:92-97 and two stress settings of the tolerances, L=J=5 on the 8x8 grid of G-A's domain
        xs, ys = np.linspace(0, 0.06, 8), np.linspace(0, 0.08, 8)
        for tag, extra in (("max", dict(NormChoice="max")), ("Euclidean", dict(NormChoice="Euclidean")),
                           ("Analytic", dict(NormChoice="Analytic", NormAnalytic=1.0)),
                           ("maxlie3", dict(MaxLie=3)), ("tolmax1", dict(TolMax=1.0, TolMin=1e-6))):
            over = dict(L=5, J=5, CT="expm_onestep", **extra)
            region_fixture("region_L5J5_8x8_" + tag, over, xs, ys)
    if "spiral6" in what:
        # BASELINE config 5 at a checkable size: the 4x4 strided sub-grid (ix, iy in {16, 48, 80, 112}) of the
        # Nxy=128 spiral-mean region, L=J=6 -- about 30-150 s per point and core in the reference
        full_x, full_y = np.linspace(0, 0.05, 128), np.linspace(0, 0.25, 128)
        for ct in ("expm_onestep", "expm_multiply"):
            over = dict(torus="spiral", L=6, J=6, CT=ct)
            region_fixture("region_spiral_L6J6_sub4_" + ct, over, full_x[16::32], full_y[16::32])
    print("done in %.1fs" % (time.time() - t00))


def _cr(e):
    RG, RG_modules, RG_dict = ref_loader.load()
    # Pool-mode semantics (quirk Q2): every task sees a fresh AmpInf/AmpSup
    _case.AmpInf = list(_pristine[0])
    _case.AmpSup = list(_pristine[1])
    r = RG_modules.compute_cr(np.float64(e), _case)
    return [float(r[0]), float(r[1])]


if __name__ == "__main__":
    main(sys.argv[1:])
