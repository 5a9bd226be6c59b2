"""CPU oracle for the cchandre/RG hot path -- TEST INFRASTRUCTURE ONLY.

Python face of oracle/rg_oracle.c (a plain-C restatement with direct convolutions).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this
module; the product package rg_b200/ never does.

Parity pinning: tests/test_oracle.py checks this oracle against golden vectors frozen from the
live, unmodified reference (tests/golden/*.npz, made by tests/golden/make_golden.py).

The class mirrors the reference's `RG` (RG.py:44-307) method for method so that tests read like
reference code; table construction restates RG.py:51-100 with the same numpy/scipy calls.
"""
import copy
import ctypes
import os
import subprocess
import warnings

import numpy as xp
import scipy.linalg as la

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
N_TAB = 256  # iterates of Omega tabulated per starting Omega (reference converge has no cap, Q6)


def build(force=False):
    so = os.path.join(_HERE, "librg_oracle.so")
    src = os.path.join(_HERE, "rg_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "librg_oracle.so"])
    return so


class _Case(ctypes.Structure):
    _fields_ = [("dim", ctypes.c_int), ("L", ctypes.c_int), ("J", ctypes.c_int), ("M", ctypes.c_int),
                ("n1", ctypes.c_int), ("vecjl", ctypes.c_int),
                ("omega0_nu", ctypes.c_void_p), ("iminus", ctypes.c_void_p), ("reidx", ctypes.c_void_p),
                ("nu_l1", ctypes.c_void_p), ("eigenvalue", ctypes.c_double),
                ("norm_choice", ctypes.c_int), ("norm_analytic", ctypes.c_double),
                ("tol_min", ctypes.c_double), ("tol_max", ctypes.c_double), ("max_lie", ctypes.c_int),
                ("ct", ctypes.c_int), ("min_step", ctypes.c_double), ("abs_tol", ctypes.c_double),
                ("rel_tol", ctypes.c_double), ("theta", ctypes.c_void_p), ("n_theta", ctypes.c_int),
                ("n_tab", ctypes.c_int), ("nrm_tab", ctypes.c_void_p), ("Omega_nu_tab", ctypes.c_void_p),
                ("max_series", ctypes.c_int), ("max_iter", ctypes.c_int)]


class _State(ctypes.Structure):
    _fields_ = [("rng", ctypes.c_uint64), ("n_liouville", ctypes.c_long), ("n_lie", ctypes.c_long),
                ("n_rgmap", ctypes.c_long), ("Omega_nu", ctypes.c_void_p)]


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "librg_oracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        assert L.oracle_sizeof_case() == ctypes.sizeof(_Case), "ocase layout mismatch"
        assert L.oracle_sizeof_state() == ctypes.sizeof(_State), "ostate layout mismatch"
        for fn in ("oracle_norm", "oracle_norm_int", "oracle_generate_y", "oracle_pnorm"):
            getattr(L, fn).restype = ctypes.c_double
        _LIB = L
    return _LIB


def load_theta():
    path = os.path.join(_HERE, "..", "rg_b200", "data", "theta_taylor.txt")
    return xp.array([float(l) for l in open(path) if l.strip() and not l.startswith("#")])


class Hamiltonian:
    """RG.py:309-332"""

    def __init__(self, Omega, f, error=0, count=0):
        self.Omega = Omega
        self.f = f
        self.error = error
        self.count = count

    def __add__(self, other):
        return Hamiltonian(self.Omega, self.f + other.f, error=self.error, count=self.count)

    def __sub__(self, other):
        return Hamiltonian(self.Omega, self.f - other.f, error=self.error, count=self.count)

    def __mul__(self, other):
        if isinstance(other, float) or isinstance(other, int):
            return Hamiltonian(self.Omega, other * self.f, error=self.error, count=self.count)
        raise TypeError('multiplication only defined with a float or a int')


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _d(a):
    return xp.ascontiguousarray(a, dtype=xp.float64)


class OracleRG:
    CTS = {"expm_onestep": 0, "expm_adapt": 1, "expm_multiply": 2}
    NORMS = {"sum": 0, "max": 1, "Euclidean": 2, "Analytic": 3}

    def __init__(self, parameters, seed=12345):
        # RG.py:52-54
        for key in parameters:
            setattr(self, key, parameters[key])
        self.DictParams = parameters
        if self.Precision not in (xp.float64, xp.float32):
            raise ValueError("Precision=128 is not supported")
        self.N = xp.asarray(self.N, dtype=int)
        self.omega0 = xp.asarray(self.omega0)
        self.Omega = xp.asarray(self.Omega)
        # RG.py:55-67
        self.dim = len(self.omega0)
        self.zero_ = self.dim * (0,)
        self.r_jl = (self.J + 1,) + self.dim * (2 * self.L + 1,)
        self.r_1l = (1,) + self.dim * (2 * self.L + 1,)
        self.r_j1 = (self.J + 1,) + self.dim * (1,)
        self.vecjl = int(xp.prod(xp.asarray(self.r_jl)))
        self.M = (2 * self.L + 1) ** self.dim
        indx = self.dim * (xp.hstack((xp.arange(0, self.L + 1), xp.arange(-self.L, 0))),)
        self.nu = xp.meshgrid(*indx, indexing='ij')
        # RG.py:68-75
        eigenval, w_eig = la.eig(self.N.T)
        self.Eigenvalue = xp.real(eigenval[xp.abs(eigenval) < 1])[0]
        sgn = xp.sign(self.Eigenvalue).astype(int)
        N_nu = sgn * xp.einsum('ij,j...->i...', self.N, self.nu)
        self.omega0_nu = xp.einsum('i,i...->...', self.omega0, self.nu).reshape(self.r_1l)
        self.Omega_nu = xp.einsum('i,i...->...', self.Omega, self.nu).reshape(self.r_1l)
        mask = xp.prod(abs(N_nu) <= self.L, axis=0, dtype=bool)
        norm_nu = self.Precision(la.norm(self.nu, axis=0)).reshape(self.r_1l)
        self.J_ = xp.arange(self.J + 1, dtype=self.Precision).reshape(self.r_j1)
        # RG.py:77-86
        if self.ChoiceIm == 'AKW1998':
            comp_im = self.Sigma * xp.repeat(norm_nu, self.J + 1, axis=0) + self.Kappa * self.J_
        elif self.ChoiceIm == 'K1999':
            comp_im = xp.maximum(self.Sigma * xp.repeat(norm_nu, self.J + 1, axis=0), self.Kappa * self.J_)
        else:
            w_nu = xp.einsum('ij,j...->i...', w_eig, self.nu)
            norm_w_nu = la.norm(w_nu, axis=0).reshape(self.r_1l)
            comp_im = self.Sigma * xp.repeat(norm_w_nu, self.J + 1, axis=0) + self.Kappa * self.J_
        omega0_nu_ = xp.repeat(xp.abs(self.omega0_nu), self.J + 1, axis=0) / la.norm(self.omega0)
        self.Iminus = omega0_nu_ > comp_im
        # RG.py:87-91 as a gather table: f'[nu] = f[N_nu] on mask, else 0
        n1 = 2 * self.L + 1
        flat = xp.zeros(self.r_1l[1:], dtype=xp.int64)
        for d in range(self.dim):
            flat = flat * n1 + (N_nu[d] % n1)
        self._reidx = xp.where(mask, flat, -1).astype(xp.int32).reshape(-1)
        self._nu_l1 = _d(xp.sum(xp.abs(self.nu), axis=0).reshape(-1))
        self._omega0_nu = _d(self.omega0_nu.reshape(-1))
        self._iminus = xp.ascontiguousarray(self.Iminus.reshape(-1), dtype=xp.uint8)
        self._theta = load_theta()
        self.theta = dict(zip(range(1, len(self._theta) + 1), self._theta))
        self._tabs = {}
        self._seed = seed
        self._state = _State()
        self._state.rng = seed * 2654435761 + 1
        self._Omega_nu_buf = _d(self.Omega_nu.reshape(-1))
        self._state.Omega_nu = self._Omega_nu_buf.ctypes.data
        self.counters = self._state

    # ---------------------------------------------------------------- plumbing
    def _tables(self, Omega):
        """nrm_tab[k], Omega_nu_tab[k] from a starting Omega with the reference's calls (RG.py:210-213)."""
        key = xp.asarray(Omega, dtype=xp.float64).tobytes()
        if key not in self._tabs:
            Om = xp.asarray(Omega)
            nrm = xp.empty(N_TAB)
            onu = xp.empty((N_TAB + 1, self.M))
            oms = [Om]
            onu[0] = xp.einsum('i,i...->...', Om, self.nu).reshape(-1)
            for k in range(N_TAB):
                Omega_ = self.N.T.dot(Om)
                nrm[k] = la.norm(Omega_)
                Om = Omega_ / la.norm(Omega_)
                oms.append(Om)
                onu[k + 1] = xp.einsum('i,i...->...', Om, self.nu).reshape(-1)
            self._tabs[key] = (nrm, onu, oms)
        return self._tabs[key]

    def _case(self, Omega=None, ct=None):
        nrm, onu, _ = self._tables(self.Omega if Omega is None else Omega)
        c = _Case()
        c.dim, c.L, c.J, c.M, c.n1, c.vecjl = self.dim, self.L, self.J, self.M, 2 * self.L + 1, self.vecjl
        c.omega0_nu = self._omega0_nu.ctypes.data
        c.iminus = self._iminus.ctypes.data
        c.reidx = self._reidx.ctypes.data
        c.nu_l1 = self._nu_l1.ctypes.data
        c.eigenvalue = float(self.Eigenvalue)
        c.norm_choice = self.NORMS.get(self.NormChoice, 0)
        c.norm_analytic = float(self.NormAnalytic)
        c.tol_min, c.tol_max, c.max_lie = float(self.TolMin), float(self.TolMax), int(self.MaxLie)
        name = (self.CanonicalTransformation if ct is None else ct).replace("self.", "")
        c.ct = self.CTS[name]
        c.min_step, c.abs_tol, c.rel_tol = float(self.MinStep), float(self.AbsTol), float(self.RelTol)
        c.theta, c.n_theta = self._theta.ctypes.data, len(self._theta)
        c.n_tab, c.nrm_tab, c.Omega_nu_tab = N_TAB, nrm.ctypes.data, onu.ctypes.data
        c.max_series, c.max_iter = 100000, N_TAB
        c._keep = (nrm, onu)
        return c

    def _sync_Omega_nu(self):
        self._Omega_nu_buf = _d(xp.asarray(self.Omega_nu).reshape(-1))
        self._state.Omega_nu = self._Omega_nu_buf.ctypes.data

    def norm(self, f):
        f = _d(f).reshape(-1)
        assert f.size == self.vecjl
        return _lib().oracle_norm(ctypes.byref(self._case()), _p(f))

    def norm_int(self, h):
        return _lib().oracle_norm_int(ctypes.byref(self._case()), _p(_d(h.f)))

    def sym(self, h):
        h_ = copy.deepcopy(h)
        out = xp.empty(self.r_jl)
        _lib().oracle_sym(ctypes.byref(self._case()), _p(_d(h.f)), _p(out))
        h_.f = out
        return h_

    def conv_product(self, fun1, fun2):
        out = xp.empty(self.r_jl)
        _lib().oracle_conv_product(ctypes.byref(self._case()), _p(_d(fun1)), _p(_d(fun2)), _p(out))
        return out

    def generate_y(self, h):
        self._sync_Omega_nu()
        y1, y2, y3 = xp.empty(self.r_jl), xp.empty(self.r_jl), xp.empty(self.r_jl)
        a = _lib().oracle_generate_y(ctypes.byref(self._case()), ctypes.c_void_p(self._state.Omega_nu), _p(_d(h.f)), _p(y1), _p(y2), _p(y3))
        return a, y1, y2, y3

    def liouville(self, y, f):
        self._sync_Omega_nu()
        out = xp.empty(self.r_jl)
        _lib().oracle_liouville(ctypes.byref(self._case()), ctypes.byref(self._state), ctypes.c_double(y[0]),
                                _p(_d(y[1])), _p(_d(y[2])), _p(_d(f)), _p(out))
        return out

    def pnorm(self, y, p=1):
        self._sync_Omega_nu()
        return _lib().oracle_pnorm(ctypes.byref(self._case()), ctypes.byref(self._state), ctypes.c_double(y[0]),
                                   _p(_d(y[1])), _p(_d(y[2])), ctypes.c_int(p))

    def compute_m_s(self, y):
        self._sync_Omega_nu()
        m, s = ctypes.c_int(), ctypes.c_int()
        _lib().oracle_compute_m_s(ctypes.byref(self._case()), ctypes.byref(self._state), ctypes.c_double(y[0]),
                                  _p(_d(y[1])), _p(_d(y[2])), ctypes.byref(m), ctypes.byref(s))
        return m.value, s.value

    def _expm(self, which, h, y, step=1.0):
        self._sync_Omega_nu()
        h_ = copy.deepcopy(h)
        f = _d(h.f).copy()
        c = self._case()
        args = (ctypes.byref(c), ctypes.byref(self._state), _p(f))
        ya = (ctypes.c_double(y[0]), _p(_d(y[1])), _p(_d(y[2])), _p(_d(y[3])))
        if which == 0:
            h_.error = _lib().oracle_expm_onestep(*args, *ya, ctypes.c_double(step))
        elif which == 1:
            h_.error = _lib().oracle_expm_adapt(*args, ctypes.c_int(h.error), *ya, ctypes.c_double(step))
        else:
            h_.error = _lib().oracle_expm_multiply(*args, *ya)
        h_.f = f.reshape(self.r_jl)
        return h_

    def expm_onestep(self, h, y, step=1):
        return self._expm(0, h, y, float(step))

    def expm_adapt(self, h, y, step=1.0):
        return self._expm(1, h, y, float(step))

    def expm_multiply(self, h, y):
        return self._expm(2, h, y)

    def rg_map(self, h):
        """RG.py:207-228 (also leaves self.Omega_nu updated, RG.py:213)."""
        h_ = copy.deepcopy(h)
        c = self._case(Omega=h.Omega)
        nrm, onu, oms = self._tables(h.Omega)
        f = _d(h.f).copy()
        h_.error = _lib().oracle_rg_map(ctypes.byref(c), ctypes.byref(self._state), _p(f), ctypes.c_int(0))
        h_.f = f.reshape(self.r_jl)
        h_.Omega = oms[1]
        self.Omega_nu = onu[1].reshape(self.r_1l).copy()
        if h_.error == -2:
            warnings.warn('Maximum number of Lie transforms reached (MaxLie)')
        return h_

    def converge(self, h):
        c = self._case(Omega=h.Omega)
        count, err = ctypes.c_int(0), ctypes.c_int(h.error)
        conv = _lib().oracle_converge(ctypes.byref(c), ctypes.byref(self._state), _p(_d(h.f)), ctypes.c_int(0),
                                      ctypes.byref(count), ctypes.byref(err))
        h.count = count.value
        h.error = err.value
        return bool(conv)

    def generate_1Hamiltonian(self, amps, symmetric=False):
        f_ = xp.zeros(self.r_jl, dtype=self.Precision)
        for k, amp in zip(self.K, amps):
            f_[k] = amp
        h = Hamiltonian(self.Omega, f_.astype(xp.float64))
        h.f[2][self.zero_] = 0.5
        return self.sym(h) if symmetric else h

    def generate_2Hamiltonians(self, amps):
        h_list = [self.generate_1Hamiltonian(amp, symmetric=True) for amp in amps]
        if not self.converge(h_list[0]):
            h_list[0].error = 3
        if self.converge(h_list[1]):
            h_list[1].error = -3
        else:
            h_list[1].error = 0
        return h_list

    def approach(self, h_list, reldist):
        """RG.py:292-307"""
        h_list_ = copy.deepcopy(h_list)
        for h in h_list_:
            h.error = 0
        while self.norm_int(h_list_[0] - h_list_[1]) > reldist * self.norm(h_list_[0].f):
            h_mid = (h_list_[0] + h_list_[1]) * 0.5
            if self.converge(h_mid):
                h_list_[0] = copy.deepcopy(h_mid)
            else:
                h_list_[1] = copy.deepcopy(h_mid)
        h_list_[1].error = 0
        return h_list_

    # ---------------------------------------------------------------- drivers (RG_modules.py)
    def point(self, x, y):
        """RG_modules.py:121-125"""
        amp = list(self.AmpSup).copy()
        amp[0:2] = [x, y]
        h = self.generate_1Hamiltonian(amp, symmetric=True)
        return [int(self.converge(h)), h.count], h.error

    def K_flat(self):
        return xp.array([int(xp.ravel_multi_index(tuple(xp.mod(k[1:], 2 * self.L + 1)), self.r_jl[1:])) + k[0] * self.M
                         for k in (xp.asarray(k) for k in self.K)], dtype=xp.int32)

    def region(self, xs, ys, nthreads=0):
        """compute_region's data/info maps (RG_modules.py:127-151) over an explicit grid, OpenMP over points."""
        pts = [(x, y) for y in ys for x in xs]
        return self.points(pts, nthreads, (len(ys), len(xs)))

    def points(self, pts, nthreads=0, shape=None):
        n = len(pts)
        nK = len(self.K)
        amps = xp.tile(xp.asarray(self.AmpSup, dtype=self.Precision).astype(xp.float64), (n, 1))[:, :nK].copy()
        amps[:, 0] = [self.Precision(p[0]) for p in pts]
        amps[:, 1] = [self.Precision(p[1]) for p in pts]
        conv = xp.zeros(n, xp.int32); count = xp.zeros(n, xp.int32); err = xp.zeros(n, xp.int32)
        nl = xp.zeros(n, xp.int64)
        c = self._case()
        kf = self.K_flat()
        _lib().oracle_region(ctypes.byref(c), ctypes.c_int(n), _p(amps), _p(kf), ctypes.c_int(nK),
                             _p(conv), _p(count), _p(err), _p(nl), ctypes.c_int(nthreads))
        if shape is not None:
            conv, count, err, nl = (a.reshape(shape) for a in (conv, count, err, nl))
        return conv, count, err, nl

    def compute_cr(self, epsilon):
        """RG_modules.py:73-82 with Pool-mode (intended) semantics, quirk Q2."""
        amp_inf, amp_sup = list(self.AmpInf), list(self.AmpSup)
        amp_inf[0] = self.AmpInf[0] + epsilon * (self.AmpSup[0] - self.AmpInf[0])
        amp_sup[0] = amp_inf[0]
        h_list = self.generate_2Hamiltonians((amp_inf, amp_sup))
        if self.converge(h_list[0]) and (not self.converge(h_list[1])):
            h_list = self.approach(h_list, reldist=self.RelDist)
            return [2 * h_list[0].f[_] for _ in self.ModesK]
        return [2 * h_list[0].f[self.ModesK[0]], xp.nan]

    def compute_line(self):
        """RG_modules.py:84-90"""
        amps = tuple(coord * xp.asarray(self.ModesLine) * xp.asarray(self.DirLine)
                     + (1 - xp.asarray(self.ModesLine)) * xp.asarray(self.DirLine) for coord in self.CoordLine)
        h_list = self.generate_2Hamiltonians(amps)
        if self.converge(h_list[0]) and (not self.converge(h_list[1])):
            h_list = self.approach(h_list, reldist=self.RelDist)
        return h_list


def default_params(**over):
    """The reference's RG_dict.py golden-mean block (RG_dict.py:7-108) as a dict, with overrides.

    `CT` sets CanonicalTransformation (the 'self.' prefix of RG_dict.py:96 is added here);
    `torus='spiral'` switches to the spiral-mean block (RG_dict.py:25-34).
    """
    Precision = xp.float64
    p = {'Method': 'compute_region', 'Iterates': 20, 'Nxy': 500, 'RelDist': 1e-7,
         'N': xp.asarray([[1, 1], [1, 0]], dtype=int),
         'omega0': xp.asarray([(xp.sqrt(5) - 1) / 2, -1], dtype=Precision),
         'Omega': xp.asarray([1, 0], dtype=Precision),
         'K': ((0, 1, 0), (0, 1, 1)), 'AmpInf': [0, 0], 'AmpSup': [0.12, 0.35],
         'CoordLine': [0.0, 0.028], 'ModesLine': xp.asarray((1, 1)), 'DirLine': xp.asarray([1, 1]),
         'L': 5, 'J': 5, 'ChoiceIm': 'AK2000', 'Sigma': 0.6, 'Kappa': 0.1,
         'CanonicalTransformation': 'self.expm_multiply', 'MinStep': 0.05, 'AbsTol': 1e-2, 'RelTol': 1e-3,
         'TolMax': 1e+4, 'TolMin': 1e-10, 'MaxLie': 5000, 'NormChoice': 'sum', 'NormAnalytic': 1,
         'Precision': Precision, 'SaveData': False, 'PlotResults': False, 'Parallelization': (True, 100)}
    torus = over.pop("torus", "golden")
    if torus == "spiral":
        sigma = 1.3247179572447460259
        p.update(N=xp.asarray([[0, 1, 0], [0, 0, 1], [1, 0, -1]], dtype=int),
                 omega0=xp.asarray([sigma, sigma ** 2, 1.0], dtype=Precision),
                 Omega=xp.asarray([1, 1, -1], dtype=Precision),
                 K=((0, 1, 0, 0), (0, 0, 1, 0), (0, 0, 0, 1)),
                 AmpInf=[0, 0, 0], AmpSup=[0.05, 0.25, 0.1], CoordLine=[0, 0.05],
                 ModesLine=xp.asarray((1, 1, 0)), DirLine=xp.asarray([1, 5, 0.1]))
    elif torus == "taumean":   # RG_dict.py:36-46
        tau = 0.445041867912629
        p.update(N=xp.asarray([[0, 1, -1], [1, -1, 1], [0, -1, 2]], dtype=int),
                 omega0=xp.asarray([1.0, tau, 1.0 - tau - tau ** 2], dtype=Precision),
                 Omega=xp.asarray([1, 1, -1], dtype=Precision),
                 K=((0, 0, -1, 1), (0, 1, -1, -1), (0, 0, 0, 1)),
                 AmpInf=[0.0, 0.0, 0.01], AmpSup=[7e-4, 6e-3, 0.01], CoordLine=[0.0, 0.002],
                 ModesLine=xp.asarray((1, 1, 0)), DirLine=xp.asarray([1, 5, 0.01]))
    ct = over.pop("CT", None)
    if ct is not None:
        p['CanonicalTransformation'] = 'self.' + ct
    p.update(over)
    p['ModesK'] = [p['K'][_] for _ in xp.nonzero(xp.asarray(p['ModesLine']))[0]]
    return p
