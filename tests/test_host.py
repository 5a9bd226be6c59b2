"""CPU-only tests of the host-side mirror of the reference interface: parameter surface, tables of
RG.__init__ (against goldens frozen from the reference), Hamiltonian algebra, region amplitude
layout, cyclic sharding and the world_size-2 gather (gloo)."""
import os
import pickle
import socket
import subprocess
import sys

import numpy as np
import pytest
import scipy.linalg as la

from conftest import ROOT, golden

from rg_b200 import RG, Hamiltonian, make_param_dict, param_dict
from rg_b200 import rg_modules


def test_param_dict_surface():
    keys = ['Method', 'Iterates', 'Nxy', 'RelDist', 'N', 'omega0', 'Omega', 'K', 'ModesK', 'AmpInf', 'AmpSup',
            'CoordLine', 'ModesLine', 'DirLine', 'L', 'J', 'ChoiceIm', 'Sigma', 'Kappa', 'CanonicalTransformation',
            'MinStep', 'AbsTol', 'RelTol', 'TolMax', 'TolMin', 'MaxLie', 'NormChoice', 'NormAnalytic', 'Precision',
            'SaveData', 'PlotResults', 'Parallelization']
    assert list(param_dict.keys()) == keys          # RG_dict.py:76-108, same 32 keys in the same order
    assert param_dict['Method'] == 'compute_region' and param_dict['CanonicalTransformation'] == 'self.expm_multiply'
    assert param_dict['Precision'] is np.float64
    assert param_dict['ModesK'] == [(0, 1, 0), (0, 1, 1)]
    p = make_param_dict(Precision=32, ModesLine=(1, 0))
    assert p['Precision'] is np.float32 and p['ModesK'] == [(0, 1, 0)] and p['omega0'].dtype == np.float32


def test_precision_128_rejected():
    with pytest.raises(ValueError):
        make_param_dict(Precision=128)
    p = dict(make_param_dict())
    p['Precision'] = np.longdouble
    with pytest.raises(ValueError):
        RG(p)


@pytest.mark.parametrize("name,over", [("ops_golden_L5J5", dict(L=5, J=5)), ("ops_golden_L10J10", dict(L=10, J=10)),
                                        ("ops_golden_L7J4", dict(L=7, J=4))])
def test_tables_match_reference(name, over):
    g = golden(name)
    c = RG(make_param_dict(SaveData=False, **over))
    assert np.array_equal(c.Iminus, g["Iminus"])
    assert np.array_equal(c.omega0_nu, g["omega0_nu"])
    assert c.Eigenvalue == g["Eigenvalue"]
    assert c.r_jl == g["f0"].shape and c.vecjl == g["f0"].size
    # the device gather table reproduces f_[nu_mask] = f[N_nu_mask] (RG.py:214-215)
    f = np.random.RandomState(0).standard_normal(c.r_jl)
    f_ = np.zeros_like(f)
    f_[c.nu_mask] = f[c.N_nu_mask]
    src = c._reidx
    got = np.where(src >= 0, f.reshape(c.J + 1, -1)[:, np.maximum(src, 0)], 0.0).reshape(c.r_jl)
    assert np.array_equal(got, f_)
    # per-iterate frequency tables: Omega_k and Omega_k.nu exactly as RG.py:210-213 produces them
    for ct in ("expm_onestep",):
        Om = np.asarray(c.Omega)
        for k in range(g["rgmap_Omega_" + ct].shape[0]):
            Om_ = c.N.T.dot(Om)
            Om = Om_ / la.norm(Om_)   # scipy's norm (BLAS dnrm2) as RG.py:211-212; numpy's differs by 1 ulp
            assert np.array_equal(Om, g["rgmap_Omega_" + ct][k])


def test_tables_spiral():
    from rg_b200.rg_dict import SPIRAL_MEAN
    g = golden("ops_spiral_L4J4")
    c = RG(make_param_dict(L=4, J=4, SaveData=False, **SPIRAL_MEAN))
    assert c.dim == 3 and np.array_equal(c.Iminus, g["Iminus"]) and c.Eigenvalue == g["Eigenvalue"]


def test_hamiltonian_algebra_and_pickle():
    h1 = Hamiltonian(np.array([1.0, 0.0]), np.ones((2, 3)), error=3, count=-4)
    h2 = Hamiltonian(np.array([0.0, 1.0]), 2 * np.ones((2, 3)))
    s = (h1 + h2) * 0.5
    assert np.array_equal(s.f, 1.5 * np.ones((2, 3))) and s.error == 3 and s.count == -4 and s.Omega is h1.Omega
    assert np.array_equal((h1 - h2).f, -np.ones((2, 3)))
    with pytest.raises(TypeError):
        h1 * h2
    c = RG(make_param_dict(L=5, J=5))
    c2 = pickle.loads(pickle.dumps(c))      # the reference's drivers pickle `case` into Pool workers
    assert c2.L == 5 and c2._plans == {} and callable(c2.norm)


def test_eigenvalues_quirk_Q1():
    c = RG(make_param_dict(L=5, J=5))
    with pytest.raises(AttributeError):
        c.eigenvalues(Hamiltonian(c.Omega, np.zeros(c.r_jl)), 1e-10, 5)


def test_region_amps_layout():
    c = RG(make_param_dict(L=5, J=5))
    xs, ys = np.linspace(0, 0.12, 4), np.linspace(0, 0.35, 3)
    a = rg_modules.region_amps(c, xs, ys)
    assert a.shape == (12, 2)
    assert np.array_equal(a[:4, 0], xs) and np.all(a[:4, 1] == ys[0]) and np.all(a[4:8, 1] == ys[1])   # y outer, x inner
    assert list(c.K_flat()) == [0 * 121 + 1 * 11 + 0, 0 * 121 + 1 * 11 + 1]


def test_cyclic_sharding_covers_everything():
    for n, w in ((65536, 8), (10, 4), (7, 2), (3, 8)):
        seen = np.concatenate([rg_modules.shard_indices(n, r, w) for r in range(w)])
        assert sorted(seen) == list(range(n))


WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch.distributed as dist
from rg_b200 import rg_modules
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
n = 37
full = np.stack([np.arange(n), -np.arange(n), np.arange(n) %% 5], axis=1).astype(np.int32)
idx = rg_modules.shard_indices(n, rank, world)
got = rg_modules.gather_sharded(full[idx], n, rank, world)
assert np.array_equal(got, full), (rank, got[:5])
f64 = np.linspace(0, 1, 2 * n).reshape(n, 2)
got = rg_modules.gather_sharded(f64[idx].copy().view(np.int64), n, rank, world).view(np.float64)
assert np.array_equal(got, f64)
# a shard may be empty (fewer items than ranks: compute_surface with Nxy < world size)
one = np.array([[7, 8]], dtype=np.int64)
idx1 = rg_modules.shard_indices(1, rank, world)
got = rg_modules.gather_sharded(one[idx1].reshape(len(idx1), 2), 1, rank, world)
assert np.array_equal(got, one), (rank, got)
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_gather_world_size_2_gloo(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "w.py"
    script.write_text(WORKER % ROOT)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=300)
        assert p.returncode == 0, out.decode()[-2000:]


def test_bench_reference_sample_is_the_prescribed_subsample():
    """bench.py's CPU arm runs the 16x16 strided subsample BASELINE.md 3.2 prescribes (ix, iy in {8, 24, ..., 248}) in stock
    order (y outer, x inner, RG_modules.py:138-139) -- the points of golden G-C."""
    sys.path.insert(0, ROOT)
    import bench
    pts, flat, desc = bench.sub16_points(256)
    g = golden("region_L10J10_sub16_expm_onestep")
    assert len(pts) == 256 and flat[0] == 8 * 256 + 8 and flat[-1] == 248 * 256 + 248
    assert np.allclose([p[0] for p in pts[:16]], g["xs"]) and np.allclose([p[1] for p in pts[::16]], g["ys"])
    assert bench.reference_dir() is None or os.path.isfile(os.path.join(bench.reference_dir(), "RG_modules.py"))
