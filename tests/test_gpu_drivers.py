"""Drivers of RG_modules on the GPU: compute_line / compute_surface thresholds agree with the reference
to RelDist, compute_region writes the reference's .mat schema."""
import os

import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu


def make(**kw):
    from rg_b200 import RG, make_param_dict
    kw.setdefault("SaveData", False)
    return RG(make_param_dict(**kw))


def test_compute_line_L5J5(capsys):
    """G-D: critical threshold along DirLine, RelDist=1e-8 (BASELINE config 2 at the CPU-checkable size)."""
    from rg_b200 import compute_line
    g = golden("line_L5J5")
    c = make(L=5, J=5, CanonicalTransformation="expm_onestep", RelDist=1e-8, Method="line")
    hl = compute_line(c)
    eps = np.array([[2 * h.f[k] for k in c.K] for h in hl])
    assert np.allclose(eps, g["eps"], rtol=2e-8, atol=0), (eps, g["eps"])
    assert [h.count for h in hl] == [int(v) for v in g["counts"]]
    assert [h.error for h in hl] == [int(v) for v in g["errors"]]
    c.close()


def test_compute_line_L10J10(capsys):
    """BASELINE config 2: critical threshold along DirLine, golden mean, L=J=10, RelDist=1e-8, single B200.
    A strictly sequential chain of ~30 converge calls on one Hamiltonian each: runs in cluster mode (4 CTAs)."""
    import time
    from rg_b200 import compute_line
    g = golden("line_L10J10")
    c = make(L=10, J=10, CanonicalTransformation="expm_onestep", RelDist=1e-8, Method="line")
    t0 = time.perf_counter()
    hl = compute_line(c)
    dt = time.perf_counter() - t0
    eps = np.array([[2 * h.f[k] for k in c.K] for h in hl])
    assert np.allclose(eps, g["eps"], rtol=2e-8, atol=0), (eps, g["eps"])
    assert [h.count for h in hl] == [int(v) for v in g["counts"]]
    print("compute_line L=J=10 RelDist=1e-8: %.2f s on one B200 (reference: 219 s on one CPU core)" % dt)
    c.close()


def test_approach_batch_matches_host_bisection():
    """rg_approach (the whole bisection on the device, every pair at its own pace) against RG.approach (host loop
    around batched converge launches): same brackets, counts and arrays for three different pairs."""
    c = make(L=5, J=5, CanonicalTransformation="expm_onestep", RelDist=1e-6)
    pairs = []
    for e in (0.0, 0.4, 1.0):
        lo = [e * 0.12, 0.0]
        hi = [e * 0.12, 0.35]
        pairs.append([c.generate_1Hamiltonian(lo, symmetric=True), c.generate_1Hamiltonian(hi, symmetric=True)])
    flags = c.converge_batch([h for p in pairs for h in p])
    assert flags == [True, False] * 3
    dev = c.approach_batch(pairs, 1e-6)
    assert sorted(set(c.last_levels.tolist())) != [0]
    for p, d in zip(pairs, dev):
        host = c.approach(p, reldist=1e-6)
        for a, b in zip(host, d):
            assert np.array_equal(a.f, b.f) and a.count == b.count and a.error == b.error
    # ... and against the oracle's bisection (RG.py:292-307 restated on the CPU): the same bracketing pair.  The midpoints
    # are exact averages, so the two bisections agree bit for bit as long as every converge() gives the same verdict; a
    # midpoint on the critical surface itself may be decided either way, which moves a bracket end by at most RelDist
    import rg_oracle as ro
    orc = ro.OracleRG(ro.default_params(L=5, J=5, CT="expm_onestep"))
    for p, d in zip(pairs, dev):
        want = orc.approach([ro.Hamiltonian(h.Omega, h.f.copy()) for h in p], 1e-6)
        for a, b in zip(want, d):
            assert np.abs(a.f - b.f).max() <= 2e-6 * np.abs(a.f).max(), "device bisection left the oracle's bracket"
    c.close()


def test_compute_surface_L5J5(tmp_path, monkeypatch):
    """G-E: critical surface eps2(eps1), Nxy=8, RelDist=1e-5 (BASELINE config 3 at the CPU-checkable size)."""
    from rg_b200 import compute_surface
    monkeypatch.chdir(tmp_path)
    g = golden("surface_L5J5_n8")
    c = make(L=5, J=5, CanonicalTransformation="expm_onestep", RelDist=1e-5, Nxy=8, Method="surface", SaveData=True)
    data = compute_surface(c)
    assert data.shape == (2, 8)
    assert np.allclose(data[0], g["data"][0], rtol=0, atol=1e-15)
    assert np.allclose(data[1], g["data"][1], rtol=2e-5, atol=0), (data[1], g["data"][1])
    files = [f for f in os.listdir(tmp_path) if f.startswith("RG_surface_") and f.endswith(".mat")]
    assert len(files) == 1
    c.close()


def test_compute_surface_L10J10_columns():
    """BASELINE config 3 (Nxy=128, L=J=10, RelDist=1e-7) on 6 of its 128 columns, frozen from the reference
    run through multiprocess.Pool (quirk Q2 semantics); the lock-step batched bisection must land on the
    same critical amplitudes to RelDist."""
    from rg_b200.rg_modules import compute_cr_batch
    g = golden("surface_L10J10_cols")
    c = make(L=10, J=10, CanonicalTransformation="expm_onestep", RelDist=1e-7, Nxy=128, Method="surface", SaveData=False)
    res = np.array(compute_cr_batch([np.float64(e) for e in g["eps"]], c), dtype=np.float64).T
    assert np.allclose(res[0], g["data"][0], rtol=0, atol=1e-15), (res[0], g["data"][0])
    assert np.allclose(res[1], g["data"][1], rtol=4e-7, atol=0), (res[1], g["data"][1])
    c.close()


def test_compute_region_mat_schema(tmp_path, monkeypatch):
    """compute_region writes RG_region_<timestamp>.mat with data int64 (Nxy,Nxy,2), info (Nxy,Nxy) and every
    param_dict key except Precision (RG_modules.py:163-173)."""
    import scipy.io
    from rg_b200 import compute_region
    monkeypatch.chdir(tmp_path)
    g = golden("region_L5J5_8x8_default_expm_onestep")
    c = make(L=5, J=5, CanonicalTransformation="expm_onestep", Nxy=8, Method="region", SaveData=True)
    data, info = compute_region(c)
    assert np.array_equal(data[:, :, 0], g["conv"]) and np.array_equal(data[:, :, 1], g["count"]) and np.array_equal(info, g["error"])
    files = [f for f in os.listdir(tmp_path) if f.startswith("RG_region_") and f.endswith(".mat")]
    assert len(files) == 1
    m = scipy.io.loadmat(os.path.join(tmp_path, files[0]))
    assert m["data"].shape == (8, 8, 2) and m["info"].shape == (8, 8)
    for k in c.DictParams:
        assert (k in m) == (k != "Precision"), k
    assert "author" in m and "date" in m
    c.close()


def test_compute_iterates_runs(tmp_path, monkeypatch, capsys):
    """BASELINE config 1 (iterates, L=J=5, expm_adapt): the loop runs and prints (diff, delta, lambda); the
    reference dies in `eigenvalues` afterwards (quirk Q1), we report it and still save."""
    from rg_b200 import compute_iterates
    monkeypatch.chdir(tmp_path)
    c = make(L=5, J=5, CanonicalTransformation="expm_adapt", RelDist=1e-6, Iterates=2, Method="iterates", SaveData=True)
    data = compute_iterates(c)
    assert len(data) == 2 and all(np.isfinite(v) for row in data for v in row)
    # SURVEY 3.4 probe: (diff, delta, lambda) = (0.274, 4.87, 4.05), (0.554, 3.53, 4.72)
    assert abs(data[0][0] - 0.274) < 5e-3 and abs(data[0][1] - 4.87) < 5e-2 and abs(data[0][2] - 4.05) < 5e-2
    assert abs(data[1][0] - 0.554) < 5e-3 and abs(data[1][1] - 3.53) < 5e-2 and abs(data[1][2] - 4.72) < 5e-2
    assert "eigenvalues unavailable" in capsys.readouterr().out
    c.close()


def test_compute_iterates_matches_reference(tmp_path, monkeypatch, capsys):
    """BASELINE config 1 frozen from the reference (with its crashing `eigenvalues` call stubbed out): four
    iterates of (|H - R(H)|, delta, lambda) at RelDist = 1e-7.  delta is a ratio of two distances of order
    RelDist, so the bisection must take the same decisions as the reference's for it to agree this well."""
    from rg_b200 import compute_iterates
    monkeypatch.chdir(tmp_path)
    g = golden("iterates_L5J5_adapt")
    c = make(L=5, J=5, CanonicalTransformation="expm_adapt", RelDist=1e-7, Iterates=4, Method="iterates", SaveData=False)
    data = np.array(compute_iterates(c))
    capsys.readouterr()
    assert data.shape == g["data"].shape
    assert np.allclose(data, g["data"], rtol=1e-5, atol=0), (data, g["data"])
    c.close()


def test_jacobian_columns_match_oracle():
    """the finite-difference Jacobian behind `eigenvalues` as one batch of rg_maps.
    dx = 1e-6 keeps the finite-difference noise (roundoff / dx) below the comparison tolerance."""
    import rg_oracle as ro
    c = make(L=5, J=5, CanonicalTransformation="expm_onestep")
    orc = ro.OracleRG(ro.default_params(L=5, J=5, CT="expm_onestep"))
    h = c.generate_1Hamiltonian([0.02, 0.02], symmetric=True)
    h = c.rg_map(h)
    cols = [12, 140, 300, 577]
    dx = 1e-6
    jac = c.jacobian(h, dx, columns=cols)
    ho = ro.Hamiltonian(h.Omega, h.f.copy())
    for q, col in enumerate(cols):
        d = np.zeros(int(c.vecjl)); d[col] = dx
        hd = ro.Hamiltonian(h.Omega, h.f + orc.sym(ro.Hamiltonian(h.Omega, d.reshape(c.r_jl))).f)
        row = (orc.rg_map(hd).f.reshape(-1) - ho.f.reshape(-1)) / dx
        assert np.abs(jac[q] - row).max() <= 1e-6 * max(1.0, np.abs(row).max())
    eigs = c.eigenvalues_batched(h, 1e-6, 3)
    assert len(eigs) == 3 and np.all(np.isfinite(eigs))
    c.close()
