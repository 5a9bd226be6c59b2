"""Parameter surface of the reference (RG_dict.py:7-108): the same 32-key `param_dict`.

`make_param_dict(**overrides)` takes the user-level names of RG_dict.py:7-70 (Method without the
'compute_' prefix, CanonicalTransformation without 'self.', Precision as 32/64) and applies the
reference's post-processing (RG_dict.py:75-108).  `param_dict` is the golden-mean default block.
"""
import numpy as xp

_DEFAULTS = dict(
    Method='region', Iterates=20, Nxy=500, RelDist=1e-7,
    # 2D -- golden mean (RG_dict.py:13-22)
    N=[[1, 1], [1, 0]], omega0=[(xp.sqrt(5) - 1) / 2, -1], Omega=[1, 0],
    K=((0, 1, 0), (0, 1, 1)), AmpInf=[0, 0], AmpSup=[0.12, 0.35],
    CoordLine=[0.0, 0.028], ModesLine=(1, 1), DirLine=[1, 1],
    L=5, J=5, ChoiceIm='AK2000', Sigma=0.6, Kappa=0.1,
    CanonicalTransformation='expm_multiply', MinStep=0.05, AbsTol=1e-2, RelTol=1e-3,
    TolMax=1e+4, TolMin=1e-10, MaxLie=5000, NormChoice='sum', NormAnalytic=1, Precision=64,
    SaveData=True, PlotResults=False, Parallelization=(True, 100))

# 3D blocks of RG_dict.py:24-46
SPIRAL_MEAN = dict(N=[[0, 1, 0], [0, 0, 1], [1, 0, -1]],
                   omega0=[1.3247179572447460259, 1.3247179572447460259 ** 2, 1.0], Omega=[1, 1, -1],
                   K=((0, 1, 0, 0), (0, 0, 1, 0), (0, 0, 0, 1)), AmpInf=[0, 0, 0], AmpSup=[0.05, 0.25, 0.1],
                   CoordLine=[0, 0.05], ModesLine=(1, 1, 0), DirLine=[1, 5, 0.1])
TAU_MEAN = dict(N=[[0, 1, -1], [1, -1, 1], [0, -1, 2]],
                omega0=[1.0, 0.445041867912629, 1.0 - 0.445041867912629 - 0.445041867912629 ** 2],
                Omega=[1, 1, -1], K=((0, 0, -1, 1), (0, 1, -1, -1), (0, 0, 0, 1)),
                AmpInf=[0.0, 0.0, 0.01], AmpSup=[7e-4, 6e-3, 0.01], CoordLine=[0.0, 0.002],
                ModesLine=(1, 1, 0), DirLine=[1, 5, 0.01])


def make_param_dict(**over):
    v = dict(_DEFAULTS)
    v.update(over)
    prec = v['Precision']
    if prec == 128:
        raise ValueError("Precision=128 is rejected: the B200 backend computes in fp64 (RG_dict.py:75 maps it to float128)")
    Precision = {32: xp.float32, 64: xp.float64}.get(prec, xp.float64)   # RG_dict.py:75
    ModesLine = v['ModesLine']
    K = v['K']
    return {'Method': 'compute_' + v['Method'],                            # RG_dict.py:76
            'Iterates': v['Iterates'], 'Nxy': v['Nxy'], 'RelDist': v['RelDist'],
            'N': xp.asarray(v['N'], dtype=int),
            'omega0': xp.asarray(v['omega0'], dtype=Precision),
            'Omega': xp.asarray(v['Omega'], dtype=Precision),
            'K': K, 'ModesK': [K[_] for _ in xp.nonzero(ModesLine)[0]],    # RG_dict.py:85
            'AmpInf': list(v['AmpInf']), 'AmpSup': list(v['AmpSup']), 'CoordLine': v['CoordLine'],
            'ModesLine': xp.asarray(ModesLine), 'DirLine': xp.asarray(v['DirLine']),
            'L': v['L'], 'J': v['J'], 'ChoiceIm': v['ChoiceIm'], 'Sigma': v['Sigma'], 'Kappa': v['Kappa'],
            'CanonicalTransformation': 'self.' + v['CanonicalTransformation'],   # RG_dict.py:96
            'MinStep': v['MinStep'], 'AbsTol': v['AbsTol'], 'RelTol': v['RelTol'],
            'TolMax': v['TolMax'], 'TolMin': v['TolMin'], 'MaxLie': v['MaxLie'],
            'NormChoice': v['NormChoice'], 'NormAnalytic': v['NormAnalytic'], 'Precision': Precision,
            'SaveData': v['SaveData'], 'PlotResults': v['PlotResults'], 'Parallelization': v['Parallelization']}


param_dict = make_param_dict()
