#!/usr/bin/env python
"""Thick-slice super-resolution end to end on the sm_100a kernels -- the recipe of the
reference's demos/demo_single_channel.ipynb / demo_multi_channel.ipynb (cell 4):

  ground truth (NIfTI files, or a synthetic phantom)  ->  simulate thick-slice observations
  x_c = A_c y_c + N(0, sd^2)  ->  trilinear initial estimate (_init_y_dat)  ->  fit (ADMM / CG,
  optional even/odd scaling and rigid updates)  ->  MSE against the ground truth, NIfTI output.

    python demos/sr_demo.py                                   # synthetic 3-channel 128^3 phantom
    python demos/sr_demo.py t1.nii.gz t2.nii.gz pd.nii.gz     # e.g. the BrainWeb volumes in data/
    options: --thick 4 --sd 25 --max-iter 60 --scaling --estimate --out out_dir

Hyper-parameters follow the reference: tau = 1/sd^2, lam0 = sqrt(1/C) / mu (unires/_core.py:
134-136, 279).  By default sd is the known simulation value and mu the mean foreground; with
--estimate the noise is added everywhere (like the notebooks) and tau, mu come from
`_core._estimate_hyperpar` (mixture fit to the intensity histogram, unires/_core.py:96-142).
What is NOT reproduced is the co-registration of unires/_core.py (needs nitorch.tools).
"""
import argparse
import math
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from unires_b200 import _core, _project, io, run, struct, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('files', nargs='*')
    ap.add_argument('--thick', type=int, default=4)
    ap.add_argument('--sd', type=float, default=25.0)
    ap.add_argument('--max-iter', type=int, default=60)
    ap.add_argument('--dim', type=int, default=128)
    ap.add_argument('--scaling', action='store_true')
    ap.add_argument('--estimate', action='store_true')
    ap.add_argument('--out', default=None)
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    if a.files:
        truth, mats = [], []
        for f in a.files:
            dat, dim, mat, *_ = io._read_image(f, device=dev)
            truth.append(dat)
            mats.append(mat.cpu())
    else:
        truth = [t.to(dev) for t in synth.phantom((a.dim,) * 3, 3, seed=0)]
        mats = [torch.eye(4, dtype=torch.float64)] * 3
    C = len(truth)
    dim_y, mat_y = tuple(truth[0].shape), mats[0]
    sett = struct.settings()
    sett.device, sett.method, sett.do_proj, sett.do_print = str(dev), 'super-resolution', True, 0
    sett.max_iter, sett.scaling, sett.unified_rigid, sett.clean_fov = a.max_iter, a.scaling, False, False
    g = torch.Generator().manual_seed(0)
    x, y = [], []
    for c in range(C):
        axis = c % 3  # each channel thick-sliced along a different axis
        scl = [1.0, 1.0, 1.0]
        scl[axis] = float(a.thick)
        mat_x = mats[c] @ torch.diag(torch.tensor(scl + [1.0], dtype=torch.float64))
        dim_x = tuple(int(math.floor(d / s)) for d, s in zip(dim_y, scl))
        po = _project._proj_info(dim_y, mat_y, dim_x, mat_x, prof_ip=sett.profile_ip,
                                 prof_tp=sett.profile_tp, gap=sett.gap, device=dev,
                                 scl=0.1 if a.scaling else 0.0)
        clean = _project._proj_apply('A', truth[c][None, None], po, method=sett.method)[0, 0]
        noise = (a.sd * torch.randn(tuple(clean.shape), generator=g)).to(dev)
        noisy = clean + noise if a.estimate else \
            torch.where(clean != 0, clean + noise, torch.zeros((), device=dev))
        obs = struct._input(dat=noisy.float().contiguous(), dim=dim_x, mat=mat_x.to(dev),
                            tau=torch.tensor(1.0 / a.sd ** 2, device=dev), sd=a.sd, ct=False)
        fg = obs.dat[obs.dat > 0]
        obs.mu = float(fg.mean())
        if a.estimate:
            _core._estimate_hyperpar([[obs]], sett)
            print('  channel %d: estimated sd %.2f (simulated %g; a Rician fit of Gaussian noise '
                  'with the negatives dropped reads sd / sqrt(2) at zero signal), mu %.1f'
                  % (c, float(obs.sd), a.sd, float(obs.mu)))
            obs.mu = float(obs.mu)
        if a.scaling:  # the fit starts from scl = 0 and has to find exp(+-0.1)
            po = _project._proj_info(dim_y, mat_y, dim_x, mat_x, prof_ip=sett.profile_ip,
                                     prof_tp=sett.profile_tp, gap=sett.gap, device=dev, scl=0.0)
        obs.po = po
        x.append([obs])
        rec = struct._output(dim=dim_y, mat=mat_y.to(dev))
        rec.lam0 = torch.tensor(math.sqrt(1.0 / C) / obs.mu, device=dev)
        y.append(rec)
    y = io._init_y_dat(x, y, sett)
    mse = lambda u, v: float(((u - v) ** 2).mean())
    mse0 = [mse(y[c].dat, truth[c]) for c in range(C)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dat_y, mat, *_ = run.fit(x, y, sett)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    mse1 = [mse(dat_y[..., c], truth[c]) for c in range(C)]
    print('grid %s, %d channel(s), thick x%d, sd %g: %d ADMM iterations in %.2f s'
          % ('x'.join(map(str, dim_y)), C, a.thick, a.sd, run.fit.last['n_iter'], dt))
    for c in range(C):
        print('  channel %d: MSE trilinear init %10.2f -> super-resolved %10.2f%s'
              % (c, mse0[c], mse1[c],
                 '  (scl = %.4f)' % float(x[c][0].po.scl) if a.scaling else ''))
    if a.out:
        os.makedirs(a.out, exist_ok=True)
        for c in range(C):
            io._write_image(dat_y[..., c], os.path.join(a.out, 'u_channel%d.nii.gz' % c), mat=mat)
    assert all(m1 < m0 for m0, m1 in zip(mse0, mse1)), 'super-resolution did not beat the init'


if __name__ == '__main__':
    main()
