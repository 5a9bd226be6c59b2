#!/usr/bin/env python
"""Freeze `point` results of the UNMODIFIED reference for pixels of BASELINE config 4 (region Nxy=256, L=J=10,
golden mean) that are very expensive with expm_multiply (compute_m_s returns s ~ 1e5 on a diverging Hamiltonian,
RG.py:136-139,155-170): ~1 CPU-hour each.  Build container only (needs /root/reference).

    python tests/golden/make_heavy_multiply.py OUT.jsonl NPROC ix,iy [ix,iy ...]

One JSON line per finished pixel (appended as soon as it is done): ix, iy, x, y, conv, count, error, liouvilles,
rg_maps, seconds.  np.random.seed(12345) per worker (quirk Q7), as every other golden.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

OVER = dict(L=10, J=10, CT="expm_multiply", AmpInf=[0.0, 0.0], AmpSup=[0.12, 0.35], Nxy=256)


def _one(job):
    ix, iy = job
    xs = np.linspace(0.0, 0.12, 256)
    ys = np.linspace(0.0, 0.35, 256)
    conv, count, err, nl, ny, nm, sec = mg._point((xs[ix], ys[iy]))
    return dict(ix=ix, iy=iy, x=float(xs[ix]), y=float(ys[iy]), conv=int(conv), count=int(count), error=int(err),
                liouvilles=int(nl), rg_maps=int(nm), seconds=float(sec))


if __name__ == "__main__":
    out, nproc = sys.argv[1], int(sys.argv[2])
    jobs = [tuple(int(v) for v in a.split(",")) for a in sys.argv[3:]]
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(nproc, initializer=mg._init_worker, initargs=(OVER, 12345)) as pool:
        for r in pool.imap_unordered(_one, jobs, chunksize=1):
            with open(out, "a") as fh:
                fh.write(json.dumps(r) + "\n")
            print(r, flush=True)
