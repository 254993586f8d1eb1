"""The outer loop of UniRes on the sm_100a kernels: `fit(x, y, sett)` with the control flow of
unires/run.py:24-207 -- coarse-to-fine regularisation schedule (unires/_core.py:288-307),
ADMM iterations, convergence test on the objective, clean-FOV mask and output clamp
(unires/_core.py:619-627).

`sett.scaling` runs the even/odd slice-scaling update (`_update._update_scaling`) after every
ADMM iteration (unires/run.py:115-122) and `sett.unified_rigid` the rigid Gauss-Newton update
(`_update._update_rigid`) every `rigid_mod` iterations (unires/run.py:127-135); nothing is
written to disk.  `init` / `preproc` (I/O, hyper-parameter estimation, co-registration) are out of
scope: the caller supplies x (observations with tau, mu, po) and y (recon with lam0, mat).
"""
import torch

from . import _lib
from ._update import (_admm_aux, _expm, _step_size, _update_admm, _update_rigid,
                      _update_scaling)
from .optim import get_gain
from .spatial import affine_grid


def _get_sched(N, sett):
    """Coarse-to-fine scaling of the regularisation: the sched_num powers of two above
    reg_scl, then reg_scl itself (e.g. 4 -> [32, 16, 8, 4]); a single level when N == 1."""
    if sett.sched_num < 0 or N == 1:
        sett.sched_num = 0
    if sett.rigid_mod < 1:
        sett.rigid_mod = 1
    scl = torch.as_tensor(sett.reg_scl, dtype=torch.float32).reshape(1).cpu()
    powers = 2.0 ** torch.arange(0, 32, dtype=torch.float32).flip(0)
    ix = int(torch.min((powers - scl).abs(), dim=0)[1])
    sett.reg_scl = torch.cat((powers[ix - sett.sched_num:ix], scl)).to(sett.device)
    return sett


def _clean_fov(x, y, sett):
    """Zero the recon outside every observation's field of view (unires/run.py:162-187)."""
    for xc, yc in zip(x, y):
        msk = torch.ones(tuple(yc.dim), dtype=torch.bool, device=yc.dat.device)
        for obs in xc:
            cpu = lambda t: torch.as_tensor(t).detach().to('cpu', torch.float64)
            M = torch.linalg.solve(cpu(yc.mat), cpu(obs.po.rigid) @ cpu(obs.mat)).inverse()
            grid = affine_grid(M.to(obs.dat.dtype).to(yc.dat.device), tuple(yc.dim)).materialize()
            for d in range(3):
                msk &= (grid[..., d] >= 0) & (grid[..., d] < obs.dim[d])
        yc.dat[~msk] = 0.0


def fit(x, y, sett):
    """Fit the model (denoising / super-resolution by ADMM).

    Returns (dat_y, mat_y, pth_y, R, label, pth_label) like the reference: dat_y is the
    reconstruction as float32 (X, Y, Z, C); pth_y is empty and label None (nothing is
    written); R holds the rigid matrix exp(sum q_i B_i) of every observation.
    `fit.last` keeps {'n_iter', 'obj', 'jtv', 'reg_scl'} of the run."""
    if getattr(sett, 'unified_rigid', False) and getattr(sett, 'rigid_basis', None) is None:
        raise ValueError('sett.unified_rigid needs sett.rigid_basis (the SE(3) Lie basis)')
    with torch.no_grad():
        N = sum(len(xc) for xc in x)
        sett = _get_sched(N, sett)
        cnt_scl = 0
        for yc in y:
            yc.lam = sett.reg_scl[cnt_scl] * yc.lam0
        obj = torch.zeros(sett.max_iter, 3, dtype=torch.float64, device=sett.device)
        tmp = torch.zeros_like(y[0].dat)
        n_done = 0
        if sett.max_iter > 0:
            rho = _step_size(x, y, sett, verbose=True)
            z, w = _admm_aux(y, sett)
        cnt_scl_iter = 0  # at least a fixed number of iterations at each scale
        countdown0 = countdown1 = 6
        for n_iter in range(sett.max_iter):
            y, z, w, tmp, obj = _update_admm(x, y, z, w, rho, tmp, obj, n_iter, sett)
            n_done = n_iter + 1
            # convergence of the objective (one host sync per ADMM iteration, like upstream)
            gain = get_gain(obj[:n_iter + 1, 0], monotonicity='decreasing')
            small = bool(gain.abs() < sett.tolerance)
            if cnt_scl >= sett.reg_scl.numel() - 1 and cnt_scl_iter > 20 and \
                    (small or n_iter >= sett.max_iter - 1):
                countdown0 -= 1
                if countdown0 == 0:
                    break
            else:
                countdown0 = 6
            # even/odd slice scaling (unires/run.py:115-122)
            if getattr(sett, 'scaling', False):
                x, _ = _update_scaling(x, y, sett, max_niter_gn=1, num_linesearch=6)
            # rigid alignment of every observation (unires/run.py:127-135)
            if getattr(sett, 'unified_rigid', False) and n_iter > 0 and \
                    n_iter % sett.rigid_mod == 0:
                x, _ = _update_rigid(x, y, sett, mean_correct=False, max_niter_gn=1,
                                     num_linesearch=6, samp=sett.rigid_samp)
            # coarse-to-fine: next regularisation level, new ADMM step size
            if cnt_scl + 1 < len(sett.reg_scl) and cnt_scl_iter > 16 and bool(gain.abs() < 1e-3):
                countdown1 -= 1
                if countdown1 == 0:
                    cnt_scl_iter = 0
                    cnt_scl += 1
                    for yc in y:
                        yc.lam = sett.reg_scl[cnt_scl] * yc.lam0
                    rho = _step_size(x, y, sett)
            else:
                countdown1 = 6
            cnt_scl_iter += 1
        if getattr(sett, 'clean_fov', False):
            _clean_fov(x, y, sett)
        # output: clamp to the observed intensity range, stack the channels
        chans = []
        for xc, yc in zip(x, y):
            mn = min(float(obs.dat.min()) for obs in xc)
            mx = max(float(obs.dat.max()) for obs in xc)
            yc.dat.clamp_(mn, mx)
            chans.append(yc.dat[..., None])
        dat_y = torch.cat(chans, dim=3)
        R = torch.eye(4, dtype=torch.float64, device=sett.device).repeat(N, 1, 1)
        if getattr(sett, 'rigid_basis', None) is not None:
            obs_all = [o for xc in x for o in xc]
            for k, o in enumerate(obs_all):
                if getattr(o, 'rigid_q', None) is not None:
                    R[k] = _expm(o.rigid_q, sett.rigid_basis).to(sett.device)
        fit.last = {'n_iter': n_done, 'obj': obj[:n_done].clone(), 'jtv': tmp,
                    'reg_scl': sett.reg_scl}
        return dat_y, y[0].mat, [], R, None, None


fit.last = None
