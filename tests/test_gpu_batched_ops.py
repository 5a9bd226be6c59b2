"""Stand-alone batched operators of the C ABI with B >> 1 (rg_norm in its three flavours, rg_sym, rg_reindex_rescale)
and RG.pnorm / compute_m_s arguments, against the reference's own numpy expressions (RG.py:92-97, 210-216, 244-253)
evaluated on the host with the tables RG.__init__ builds, and against the CPU oracle."""
import numpy as np
import pytest
import scipy.linalg as la

from conftest import golden, relmax

pytestmark = pytest.mark.gpu


def make(L, J, ct="expm_onestep", **kw):
    from rg_b200 import RG, make_param_dict
    return RG(make_param_dict(L=L, J=J, CanonicalTransformation=ct, SaveData=False, **kw))


def make_spiral(L, J, **kw):
    from rg_b200 import RG, make_param_dict
    sigma = 1.3247179572447460259
    return RG(make_param_dict(L=L, J=J, SaveData=False, N=[[0, 1, 0], [0, 0, 1], [1, 0, -1]],
                              omega0=[sigma, sigma ** 2, 1.0], Omega=[1, 1, -1],
                              K=((0, 1, 0, 0), (0, 0, 1, 0), (0, 0, 0, 1)), AmpInf=[0, 0, 0], AmpSup=[0.05, 0.25, 0.1], **kw))


def batch(c, B, seed):
    """B Hamiltonian-like arrays: decaying random coefficients, f[2][0] = 0.5 +- 10 % (the rescaling divides by it)."""
    rng = np.random.default_rng(seed)
    nu1 = sum(np.abs(n) for n in c.nu)
    f = rng.standard_normal((B,) + c.r_jl) * np.exp(-0.7 * nu1)[None, None] * np.exp(-0.5 * np.arange(c.J + 1)).reshape((1, -1) + c.dim * (1,))
    f[(slice(None), 2) + c.zero_] = 0.5 * (1 + 0.1 * rng.standard_normal(B))
    return f


CASES = [("golden", 10, 10), ("golden", 5, 5), ("spiral", 4, 4)]


def case_of(kind, L, J, **kw):
    return make(L, J, **kw) if kind == "golden" else make_spiral(L, J, **kw)


@pytest.mark.parametrize("kind,L,J", CASES)
@pytest.mark.parametrize("nc", ["sum", "max", "Euclidean", "Analytic"])
def test_norm_batch_all_flavours(kind, L, J, nc):
    """rg_norm(which = NORM / NORM_INT / NORM_IMINUS) on 300 arrays at once == RG.py:92-97, 244-247, 217-218."""
    from rg_b200.rg import WHICH_NORM, WHICH_NORM_INT, WHICH_NORM_IMINUS
    c = case_of(kind, L, J, NormChoice=nc)
    f = batch(c, 300, 1)
    ax = tuple(range(1, c.dim + 2))
    if nc == "sum":
        ref = lambda a: np.abs(a).sum(axis=ax)
    elif nc == "max":
        ref = lambda a: np.abs(a).max(axis=ax)
    elif nc == "Euclidean":
        ref = lambda a: np.sqrt((np.abs(a) ** 2).sum(axis=ax))
    else:
        w = np.exp(c.NormAnalytic * sum(np.abs(n) for n in c.nu))
        ref = lambda a: (np.abs(a) * w[None, None]).max(axis=ax)
    got = c.norm_batch(f, WHICH_NORM)
    assert np.allclose(got, ref(f), rtol=1e-13, atol=0)
    f_int = f.copy()
    f_int[(slice(None), slice(None)) + c.zero_] = 0.0          # norm_int zeroes the whole nu = 0 column (RG.py:245-246)
    assert np.allclose(c.norm_batch(f, WHICH_NORM_INT), ref(f_int), rtol=1e-13, atol=0)
    f_im = np.where(c.Iminus[None], f, 0.0)                    # norm(f[Iminus]) of RG.py:217-218
    assert np.allclose(c.norm_batch(f, WHICH_NORM_IMINUS), ref(f_im), rtol=1e-13, atol=0)
    c.close()


@pytest.mark.parametrize("kind,L,J", CASES)
def test_sym_batch(kind, L, J):
    """rg_sym on 300 arrays == (f + roll(flip(f), 1)) / 2 with f[0][0] = 0 (RG.py:249-253), bit-exact."""
    c = case_of(kind, L, J)
    f = batch(c, 300, 2)
    ref = (f + np.roll(np.flip(f, axis=tuple(a + 1 for a in c.axis_dim)), 1, axis=tuple(a + 1 for a in c.axis_dim))) / 2
    ref[(slice(None), 0) + c.zero_] = 0.0
    assert np.array_equal(c.sym_batch(f), ref)
    c.close()


@pytest.mark.parametrize("kind,L,J", CASES)
def test_reindex_rescale_batch(kind, L, J):
    """rg_reindex_rescale on 200 arrays with per-array table rows k0 == the head of RG.rg_map (RG.py:210-216)
    evaluated with the reference's own numpy expressions on the host tables."""
    c = case_of(kind, L, J)
    B = 200
    f = batch(c, B, 3)
    k0 = (np.arange(B) % 5).astype(np.int32)
    got = c.reindex_rescale_batch(f, c.Omega, k0)
    Om = np.asarray(c.Omega, dtype=float)
    oms = [Om]
    for _ in range(5):
        o = c.N.T.dot(oms[-1])
        oms.append(o / la.norm(o))
    for b in range(B):
        Omega_ = c.N.T.dot(oms[k0[b]])
        f2 = f[b][2][c.zero_]
        ren = (2 * la.norm(Omega_) / c.Eigenvalue * f2) ** (2 - np.arange(c.J + 1, dtype=float)) / (2 * f2)
        ref = np.zeros_like(f[b])
        ref[c.nu_mask] = f[b][c.N_nu_mask]
        ref *= ren.reshape(c.r_j1)
        assert relmax(got[b], ref) < 4e-16, (b, relmax(got[b], ref))
        assert np.array_equal(got[b] == 0.0, ref == 0.0)
    c.close()


def test_reindex_rescale_matches_oracle_rg_map_head():
    """The stand-alone operator reproduces the first step of the fused rg_map: with MaxLie = 0 the oracle's rg_map is
    re-index + rescale only (RG.py:210-218 with an empty Lie loop)."""
    import rg_oracle as ro
    g = golden("ops_golden_L10J10")
    c = make(10, 10)
    orc = ro.OracleRG(ro.default_params(L=10, J=10, CT="expm_onestep", MaxLie=0))
    ho = orc.generate_1Hamiltonian(list(g["amps"]), symmetric=True)
    ho2 = orc.rg_map(orc.rg_map(ho))
    h = c.generate_1Hamiltonian(list(g["amps"]), symmetric=True)
    f1 = c.reindex_rescale_batch(h.f[None], c.Omega, [0])[0]
    f2 = c.reindex_rescale_batch(f1[None], c.Omega, [1])[0]
    assert relmax(f2, ho2.f) < 1e-15
    c.close()


@pytest.mark.parametrize("p", [1, 2, 3])
def test_pnorm_powers_bracket_the_exact_norm(p):
    """RG.pnorm(y, p) (RG.py:119-123): the block 1-norm estimate of (L_y)^p is a lower bound of the exact 1-norm
    (largest absolute column sum, computed here column by column with the oracle's liouville).  It can be far
    below it: the reference hands scipy `rmatvec = -matvec` (RG.py:121), which is not the adjoint, so Higham-Tisseur's
    ascent has no guarantee -- the GPU estimator is pinned to the oracle's restatement of the same algorithm (own RNG
    stream each, as the reference's own estimate varies with its seed, quirk Q7)."""
    import rg_oracle as ro
    g = golden("ops_golden_L5J5")
    c = make(5, 5, "expm_multiply")
    c.Omega_nu = g["op_Omega_nu"].copy()
    yg = (g["op_y0"], g["op_y1"], g["op_y2"], g["op_y3"])
    orc = ro.OracleRG(ro.default_params(L=5, J=5, CT="expm_multiply"))
    orc.Omega_nu = g["op_Omega_nu"].copy()
    n = int(c.vecjl)
    exact = 0.0
    for i in range(n):
        e = np.zeros(n)
        e[i] = 1.0
        v = e.reshape(c.r_jl)
        for _ in range(p):
            v = orc.liouville(yg, v)
        exact = max(exact, np.abs(v).sum())
    est = c.pnorm(yg, p)
    eo = orc.pnorm(yg, p)
    assert 0 < est <= exact ** (1.0 / p) * (1 + 1e-12), (est, exact ** (1.0 / p))
    assert 0 < eo <= exact ** (1.0 / p) * (1 + 1e-12), (eo, exact ** (1.0 / p))
    assert eo / 5.0 <= est <= eo * 5.0, (est, eo, exact ** (1.0 / p))
    if p == 1:
        assert 0.2 * g["op_pnorm1"] <= est <= 5.0 * g["op_pnorm1"]      # the reference's own estimate (one seed)
    c.close()


def test_compute_m_s_honours_n0_and_ell():
    """n0 and ell only enter the test of RG.py:135 that chooses between the cheap branch (norm_ls / theta_m) and the
    alpha_p branch: with the golden y (norm_ls ~ 1e-2 << 115) the defaults take the cheap branch; a huge n0 makes the
    threshold tiny and forces the alpha_p branch, which must still return a valid (m*, s)."""
    g = golden("ops_golden_L5J5")
    c = make(5, 5, "expm_multiply")
    c.Omega_nu = g["op_Omega_nu"].copy()
    yg = (g["op_y0"], g["op_y1"], g["op_y2"], g["op_y3"])
    m0, s0 = c.compute_m_s(yg)
    assert (m0, s0) == tuple(int(v) for v in g["op_m_s"])
    assert c.compute_m_s(yg, n0=1, ell=2) == (m0, s0)
    m1, s1 = c.compute_m_s(yg, n0=1e9, ell=2)          # threshold 115/1e9 < norm_ls -> alpha_p branch (RG.py:141-149)
    theta = c._theta
    n1 = c.pnorm(yg, 1)
    assert 1 <= m1 <= len(theta) and s1 >= 1
    # alpha_p <= ||L_y||_1, so the alpha_p branch can never cost more than the cheap branch evaluated at norm_ls
    assert m1 * s1 <= min(m * max(1, int(np.ceil(n1 / theta[m - 1]))) for m in range(1, len(theta) + 1)) * 2
    c.close()


def test_fused_multiply_uses_the_same_m_s_as_the_stand_alone_estimator():
    """on the fused path (rg_map with CT = expm_multiply) the 1-norm estimator must apply L_y itself, not
    L_y composed with the mirror shortcut of the symmetric series terms.  One Lie step on the fused path
    (MaxLie = 1) must then give the array of the staged path generate_y -> expm_multiply -> sym (same y, same estimator
    stream for item 0) and spend the same liouville evaluations on the series: the totals may differ only by whole
    iterations of the estimator (4 evaluations each: A X and A^T S on two columns) -- the staged path runs the dense
    schedule, whose sums round differently, and sign(Y) of a roundoff-sized entry may flip."""
    g = golden("ops_golden_L5J5")
    c1 = make(5, 5, "expm_multiply", MaxLie=1)
    h = c1.generate_1Hamiltonian(list(g["amps"]), symmetric=True)
    hf = c1.rg_map(h)
    n_fused = c1.last_stats.liouville
    # staged: re-index + rescale, generate_y, expm_multiply, sym -- every step through its own ABI entry point
    c2 = make(5, 5, "expm_multiply")
    from rg_b200 import Hamiltonian
    f1 = c2.reindex_rescale_batch(h.f[None], c2.Omega, [0])[0]
    nrm, onu, oms, tab = c2._tables(c2.Omega)
    c2.Omega_nu = onu[1].reshape(c2.r_1l).copy()
    h1 = Hamiltonian(oms[1], f1)
    y = c2.generate_y(h1)
    e = c2.expm_multiply(h1, y)
    n_staged = c2.last_stats.liouville
    hs = c2.sym(e)
    assert (n_fused - n_staged) % 4 == 0 and abs(n_fused - n_staged) <= 8, (n_fused, n_staged)
    assert relmax(hf.f, hs.f) < 1e-13
    c1.close(); c2.close()
