# Portions of this file (the bodies of compute_iterates, compute_line, save_data and plot_fun) reproduce lines of
# cchandre/RG, RG_modules.py, so that the drivers stay drop-in:
#   BSD 2-Clause License -- Copyright (c) 2021, Cristel Chandre.  All rights reserved.
#   Redistribution and use in source and binary forms, with or without modification, are permitted provided that the
#   following conditions are met: 1. Redistributions of source code must retain the above copyright notice, this
#   list of conditions and the following disclaimer.  2. Redistributions in binary form must reproduce the above
#   copyright notice, this list of conditions and the following disclaimer in the documentation and/or other
#   materials provided with the distribution.
#   THIS SOFTWARE IS PROVIDED BY THE COPYRIGHT HOLDERS AND CONTRIBUTORS "AS IS" AND ANY EXPRESS OR IMPLIED WARRANTIES,
#   INCLUDING, BUT NOT LIMITED TO, THE IMPLIED WARRANTIES OF MERCHANTABILITY AND FITNESS FOR A PARTICULAR PURPOSE ARE
#   DISCLAIMED.  IN NO EVENT SHALL THE COPYRIGHT HOLDER OR CONTRIBUTORS BE LIABLE FOR ANY DIRECT, INDIRECT, INCIDENTAL,
#   SPECIAL, EXEMPLARY, OR CONSEQUENTIAL DAMAGES (INCLUDING, BUT NOT LIMITED TO, PROCUREMENT OF SUBSTITUTE GOODS OR
#   SERVICES; LOSS OF USE, DATA, OR PROFITS; OR BUSINESS INTERRUPTION) HOWEVER CAUSED AND ON ANY THEORY OF LIABILITY,
#   WHETHER IN CONTRACT, STRICT LIABILITY, OR TORT (INCLUDING NEGLIGENCE OR OTHERWISE) ARISING IN ANY WAY OUT OF THE
#   USE OF THIS SOFTWARE, EVEN IF ADVISED OF THE POSSIBILITY OF SUCH DAMAGE.
"""Drivers of the reference (RG_modules.py:48-201) on top of the batched device path.

Same entry points and outputs: compute_iterates / compute_line / compute_surface / compute_region,
compute_cr, point, save_data, plot_fun.  The reference's only parallelism -- multiprocess.Pool.imap
over grid points (RG_modules.py:97-104,132-143) -- is replaced by one CTA per grid point on the GPU
and cyclic sharding of the points over the ranks of torch.distributed (one process per GPU) with a
single final all_gather; `Parallelization` is kept in the dict and ignored.
"""
import copy
import os
import time
from datetime import date

import numpy as xp

from .rg import Hamiltonian  # noqa: F401


# ---------------------------------------------------------------------------------------------- distributed
def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def shard_indices(n, rank, world):
    """Cyclic assignment i mod world (cost is spatially coherent)."""
    return xp.arange(rank, n, world)


def gather_sharded(local, n, rank, world, device=None):
    """All-gather equally padded int32 shards [n_local, C] and undo the cyclic interleave.
    The only collective of the path: results are 3 x int32 per point."""
    if world == 1:
        return local
    import torch
    import torch.distributed as dist
    n_max = (n + world - 1) // world
    C = local.shape[1]
    pad = xp.zeros((n_max, C), dtype=local.dtype)
    pad[:local.shape[0]] = local
    use_cuda = dist.get_backend() == "nccl"
    t = torch.from_numpy(pad)
    if use_cuda:
        t = t.cuda(device) if device is not None else t.cuda()
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    g = torch.stack(parts).cpu().numpy()
    full = xp.zeros((n, C), dtype=local.dtype)
    for r in range(world):
        idx = shard_indices(n, r, world)
        full[idx] = g[r, :len(idx)]
    return full


# ---------------------------------------------------------------------------------------------- RG_modules.py:48-71
def compute_iterates(case):
    print('\033[92m    {} -- iterates \033[00m'.format(case.__str__()))
    h_list = case.generate_2Hamiltonians((case.AmpInf, case.AmpSup))
    data = []
    if h_list[0].error == 0 and h_list[1].error == 0:
        timestr = time.strftime("%Y%m%d_%H%M")
        plot_fun(case, h_list[1].f[0])
        k = 0
        while k < case.Iterates and h_list[0].error == 0:
            k += 1
            start = time.time()
            h_list = case.approach(h_list, reldist=case.RelDist)
            h_list_ = [case.rg_map(h_list[0]), case.rg_map(h_list[1])]
            if k == 1:
                print('\033[96m          ε = {:.6f} \033[00m'.format(2.0 * h_list[0].f[case.ModesK[0]]))
            plot_fun(case, h_list_[0].f[0])
            diff = case.norm(xp.abs(h_list[0].f) - xp.abs(h_list_[0].f))
            delta = case.norm(xp.abs(h_list_[0].f) - xp.abs(h_list_[1].f)) / case.norm(h_list[0].f - h_list[1].f)
            lam = 2 * h_list_[0].f[2][case.zero_] * ((case.N.T.dot(h_list_[0].Omega))**2).sum() / xp.abs(case.Eigenvalue)
            data.append([diff, delta, lam])
            h_list = copy.deepcopy(h_list_)
            print('\033[96m          ' + '‖' + 'H-R(H)‖ = {:.3e}    δ = {:.7f}   λ = {:.7f}    (done in {:d} seconds) \033[00m'.format(diff, delta, lam, int(time.time() - start)))
        try:
            print(case.eigenvalues(h_list[0], 1e-10, 5))
        except AttributeError as e:
            # the reference dies here on every run (quirk Q1, RG.py:237) and never reaches save_data
            print('\033[90m        eigenvalues unavailable ({}) \033[00m'.format(e))
            if int(case.vecjl) <= 1500:   # the dense eigen-solve of the Jacobian is host work: keep it small
                print(case.eigenvalues_batched(h_list[0], 1e-10, 5))
        save_data('iterates', data, timestr, case, info='diff     delta     lambda', display=True)
    return data


# ---------------------------------------------------------------------------------------------- RG_modules.py:73-82
def _cr_amps(epsilon, case):
    # Pool-mode (intended) semantics of compute_cr: every call sees the pristine AmpInf/AmpSup (quirk Q2)
    amp_inf, amp_sup = list(case.AmpInf), list(case.AmpSup)
    amp_inf[0] = case.AmpInf[0] + epsilon * (case.AmpSup[0] - case.AmpInf[0])
    amp_sup[0] = amp_inf[0]
    return amp_inf, amp_sup


def compute_cr(epsilon, case):
    return compute_cr_batch([epsilon], case)[0]


def compute_cr_batch(epsilons, case):
    """compute_cr for many epsilon at once (RG_modules.py:73-82 per column).  One batched `converge` launch
    decides which columns are bracketed, then ONE `approach_batch` launch runs every column's whole bisection on
    the device, each at its own pace (no lockstep between columns; the per-column arithmetic is exactly
    RG.approach, RG.py:292-307).  The reference re-converges the two end points a second time (quirk Q9); the
    results are identical so that pass is skipped."""
    n = len(epsilons)
    pairs = []
    for e in epsilons:
        amp_inf, amp_sup = _cr_amps(e, case)
        pairs.append([case.generate_1Hamiltonian(amp_inf, symmetric=True), case.generate_1Hamiltonian(amp_sup, symmetric=True)])
    conv = case.converge_batch([h for p in pairs for h in p])
    out = [None] * n
    active = []
    for i, p in enumerate(pairs):
        c0, c1 = conv[2 * i], conv[2 * i + 1]
        if c0 and not c1 and p[0].error > -98 and p[1].error > -98:
            active.append(i)
        else:
            out[i] = [2 * p[0].f[case.ModesK[0]], xp.nan]
    done = case.approach_batch([pairs[i] for i in active], case.RelDist)
    for i, p in zip(active, done):
        if p[0].error <= -98:      # this column's bisection hit the work budget: undecided, like a column without a bracket
            out[i] = [2 * p[0].f[case.ModesK[0]], xp.nan]
        else:
            out[i] = [2 * p[0].f[_] for _ in case.ModesK]
    return out


# ---------------------------------------------------------------------------------------------- RG_modules.py:84-90
def compute_line(case):
    print('\033[92m    {} -- line \033[00m'.format(case.__str__()))
    amps = tuple(coord * case.ModesLine * case.DirLine + (1 - case.ModesLine) * case.DirLine for coord in case.CoordLine)
    h_list = case.generate_2Hamiltonians(amps)
    if case.converge(h_list[0]) and (not case.converge(h_list[1])):
        h_list = case.approach(h_list, reldist=case.RelDist, display=True)
    return h_list


# ---------------------------------------------------------------------------------------------- RG_modules.py:92-119
def compute_surface(case):
    print('\033[92m    {} -- surface \033[00m'.format(case.__str__()))
    timestr = time.strftime("%Y%m%d_%H%M")
    epsilon = xp.linspace(0, 1, case.Nxy, dtype=case.Precision)
    rank, world = _dist()
    idx = shard_indices(len(epsilon), rank, world)
    local = xp.array(compute_cr_batch([epsilon[i] for i in idx], case), dtype=xp.float64).reshape(len(idx), max(2, len(case.ModesK)))
    if world > 1:
        full = gather_sharded(local.view(xp.int64).reshape(len(idx), local.shape[1]), len(epsilon), rank, world,
                              device=case._device_index()).view(xp.float64)
    else:
        full = local
    data = xp.array(full).T
    if rank == 0:
        save_data('surface', data, timestr, case, display=True)
    if case.PlotResults and rank == 0:
        plt = _pyplot()
        fig, ax = plt.subplots(1, 1)
        ax.set_box_aspect(1)
        ax.plot(data[0, :], data[1, :], color='b', linewidth=2)
        ax.set_xlim(case.AmpInf[0], case.AmpSup[0])
        ax.set_ylim(case.AmpInf[1], case.AmpSup[1])
        ax.set_xlabel(r'$\epsilon_1$')
        ax.set_ylabel(r'$\epsilon_2$')
    return data


# ---------------------------------------------------------------------------------------------- RG_modules.py:121-161
def point(x, y, case):
    amp = list(case.AmpSup).copy()
    amp[0:2] = [x, y]
    h = case.generate_1Hamiltonian(amp, symmetric=True)
    return [int(case.converge(h)), h.count], h.error


def region_amps(case, x_vec, y_vec):
    """amps[iy*Nx+ix] = AmpSup with entries 0,1 replaced by (x, y) (RG_modules.py:122-123); y is the outer loop."""
    nK = len(case.K)
    base = xp.asarray(case.AmpSup, dtype=case.Precision).astype(xp.float64)[:nK]
    amps = xp.tile(base, (len(y_vec) * len(x_vec), 1))
    X, Y = xp.meshgrid(xp.asarray(x_vec, dtype=xp.float64), xp.asarray(y_vec, dtype=xp.float64))
    amps[:, 0] = X.reshape(-1)
    amps[:, 1] = Y.reshape(-1)
    return xp.ascontiguousarray(amps)


def region_maps(case, x_vec, y_vec):
    """(data[Ny,Nx,2], info[Ny,Nx]) of compute_region for an explicit grid, sharded over ranks."""
    amps = region_amps(case, x_vec, y_vec)
    n = amps.shape[0]
    rank, world = _dist()
    idx = shard_indices(n, rank, world)
    conv, count, err = case.points(amps[idx])
    local = xp.stack([conv, count, err], axis=1).astype(xp.int32)
    full = gather_sharded(local, n, rank, world, device=case._device_index())
    data = xp.stack([full[:, 0], full[:, 1]], axis=1).astype(xp.int64).reshape(len(y_vec), len(x_vec), 2)
    info = full[:, 2].astype(xp.int64).reshape(len(y_vec), len(x_vec))
    return data, info


def partial_save_rows(case):
    """Rows of the region grid between two partial saves.  The reference re-saves the partial `data`/`info` after every
    y-row (RG_modules.py:138-143); `PartialSaveRows` in the parameter dict (not an RG_dict key: optional) selects the
    cadence here -- 1 = exactly the reference's, 0 = only the final save; default: Nxy/16 rows per launch when
    SaveData is on (every launch ends with a tail of idle SMs, so the reference's cadence costs GPU throughput), one
    launch for the whole grid when it is off."""
    rows = getattr(case, 'PartialSaveRows', None)
    if rows is None:
        rows = max(1, case.Nxy // 16) if case.SaveData else 0
    return int(rows)


def compute_region(case):
    print('\033[92m    {} -- region \033[00m'.format(case.__str__()))
    timestr = time.strftime("%Y%m%d_%H%M")
    x_vec, y_vec = (xp.linspace(case.AmpInf[_], case.AmpSup[_], case.Nxy, dtype=case.Precision) for _ in range(2))
    rank, _ = _dist()
    rows = partial_save_rows(case)
    if rows <= 0 or rows >= len(y_vec):
        data, info = region_maps(case, x_vec, y_vec)
    else:
        # RG_modules.py:138-143: the partial maps (rows done so far) overwrite the same file after every block of rows
        blocks_d, blocks_i = [], []
        for r0 in range(0, len(y_vec), rows):
            d, i = region_maps(case, x_vec, y_vec[r0:r0 + rows])
            blocks_d.append(d)
            blocks_i.append(i)
            if rank == 0 and r0 + rows < len(y_vec):
                save_data('region', xp.concatenate(blocks_d).reshape((-1, case.Nxy, 2)), timestr, case,
                          xp.concatenate(blocks_i).reshape((-1, case.Nxy)))
        data, info = xp.concatenate(blocks_d), xp.concatenate(blocks_i)
    if rank == 0:
        save_data('region', data, timestr, case, info=info, display=True)
    if case.PlotResults and rank == 0:
        plt = _pyplot()
        import matplotlib.colors as colors
        divnorm = colors.TwoSlopeNorm(vmin=min(data[:, :, 1].min(), -1), vcenter=0.0, vmax=max(data[:, :, 1].max(), 1))
        fig, ax = plt.subplots(1, 1)
        ax.set_box_aspect(1)
        im = ax.pcolormesh(x_vec, y_vec, data[:, :, 1].astype(int), norm=divnorm)
        ax.set_xlim(case.AmpInf[0], case.AmpSup[0])
        ax.set_ylim(case.AmpInf[1], case.AmpSup[1])
        ax.set_xlabel(r'$\epsilon_1$')
        ax.set_ylabel(r'$\epsilon_2$')
        fig.colorbar(im)
    return data, info


# ---------------------------------------------------------------------------------------------- RG_modules.py:163-201
def save_data(name, data, timestr, case, info=[], display=False):
    if case.SaveData:
        from scipy.io import savemat
        mdic = case.DictParams.copy()
        del mdic['Precision']
        mdic.update({'data': data, 'info': info})
        date_today = date.today().strftime(" %B %d, %Y\n")
        mdic.update({'date': date_today, 'author': 'cristel.chandre@cnrs.fr'})
        name_file = type(case).__name__ + '_' + name + '_' + timestr + '.mat'
        savemat(name_file, mdic)
        if display:
            print('\033[90m        Results saved in {} \033[00m'.format(name_file))
        return name_file


def _pyplot():
    try:
        import matplotlib.pyplot as plt
        return plt
    except ImportError as e:
        raise RuntimeError("PlotResults=True needs matplotlib, which is not installed") from e


def plot_fun(case, fun):
    if not case.PlotResults:
        return
    plt = _pyplot()
    import matplotlib.colors as colors
    if case.dim == 2:
        fig, ax = plt.subplots(1, 1)
        im = ax.imshow(xp.abs(xp.roll(fun, (case.L, case.L), axis=(0, 1))).T, origin='lower',
                       extent=[-case.L, case.L, -case.L, case.L],
                       norm=colors.LogNorm(vmin=case.TolMin, vmax=xp.abs(fun).max()), cmap='hot_r')
        fig.colorbar(im, orientation='vertical')
        ax.set_xlim(-case.L, case.L)
        ax.set_ylim(-case.L, case.L)
        ax.set_xlabel('$k_1$')
        ax.set_ylabel('$k_2$')
    elif case.dim == 3:
        from matplotlib import cm
        Y, Z = xp.meshgrid(xp.arange(-case.L, case.L + 1), xp.arange(-case.L, case.L + 1))
        X = xp.zeros_like(Y)
        fig = plt.figure()
        ax = fig.add_subplot(projection='3d')
        norm_c = colors.LogNorm(vmin=case.TolMin, vmax=xp.abs(fun).max())
        for k_ in range(-case.L, case.L + 1):
            A = xp.abs(xp.roll(fun[k_, :, :], (case.L, case.L), axis=(0, 1)))
            ax.plot_surface(X + k_, Y, Z, rstride=1, cstride=1, facecolors=cm.hot(norm_c(A)), alpha=0.4,
                            linewidth=0.0, shade=False)
        ax.view_init(elev=20, azim=120)
    plt.pause(0.5)
