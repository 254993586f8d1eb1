"""The ADMM iteration of UniRes on the sm_100a kernels -- same function names,
arguments, in-place behaviour and return values as unires/_update.py:

    _admm_aux      unires/_update.py:17-32
    _update_scaling unires/_update.py:270-393 (even/odd slice scaling, Gauss-Newton)
    _update_rigid / _update_rigid_channel / _rigid_match   unires/_update.py:198-267, 448-710
    _step_size     unires/_update.py:35-64
    _compute_nll   unires/_update.py:396-427
    _update_admm   unires/_update.py:105-195

Per ADMM iteration the reference launches O(100) ATen kernels per channel and
synchronises the host in every CG iteration; here the y-update is one
device-resident CG solve per channel (`optim.cg_fused`), the objective is two
reductions and the z/w updates are ONE pass over all channels
(`ur_jtv_prox`).  `_update_admm_sharded` is the multi-GPU form: channels are
sharded over ranks and only the JTV coupling field and the objective scalars
cross NVLink.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, ptr, i3, f3, stream, require_cuda_f32
from ._project import LhsOperator, _proj, _floats, proj_struct
from .optim import cg, cg_fused, stop_rule
from .spatial import voxel_size

_hs = _lib.host_scalar  # host float of a (possibly CUDA) scalar tensor, cached: no repeated sync


def _admm_aux(y, sett):
    """z, w = zeros (C, 3, X, Y, Z) float32 on sett.device."""
    shape = (len(y), 3) + tuple(y[0].dim)
    z = torch.zeros(shape, dtype=torch.float32, device=sett.device)
    w = torch.zeros(shape, dtype=torch.float32, device=sett.device)
    return z, w


def _has_ct(x):
    return any(bool(obs.ct) for xc in x for obs in xc)


def _step_size(x, y, sett, verbose=False):
    """rho: sett.rho if given (1 with CT data), else rho_scl sqrt(mean tau) / mean lam."""
    rho = 1.0 if _has_ct(x) else sett.rho
    if rho is not None:
        return torch.tensor(rho, device=sett.device, dtype=torch.float32)
    lam = torch.tensor([_hs(yc.lam) for yc in y], dtype=torch.float32, device=sett.device)
    tau = torch.tensor([float(o.tau) for xc in x for o in xc], dtype=torch.float32,
                       device=sett.device)
    return sett.rho_scl * torch.sqrt(torch.mean(tau)) / torch.mean(lam)


def _ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))()
    for k, t in enumerate(tensors):
        arr[k] = t.data_ptr()
    return arr


def _geometry(y):
    dim = tuple(y[0].dim) if y[0].dim is not None else tuple(y[0].dat.shape)
    # voxel_size(y[0].mat).float() (unires/_update.py:111) from the cached host copy of mat
    m = _lib.host_values(y[0].mat)
    n = int(round(len(m) ** 0.5))
    vx = [float(np.float32(sum(m[i * n + a] ** 2 for i in range(n - 1)) ** 0.5))
          for a in range(n - 1)]
    return dim, vx


def _nll_terms(x, y, sett, out, prior_field=None, accumulate_prior=False):
    """Write nll_xy into out[1] and (unless prior_field is given) nll_y into out[2].

    out: float64 CUDA tensor with >= 3 elements.  With `prior_field` the
    per-voxel prior energy of these channels is accumulated there instead (the
    caller all-reduces it and finishes with ur_sqrt_sum)."""
    dim, vx = _geometry(y)
    n_vox = dim[0] * dim[1] * dim[2]
    out[1].zero_()
    for c in range(len(x)):
        for n, obs in enumerate(x[c]):
            dat = require_cuda_f32(obs.dat, 'x.dat')
            if not sett.do_proj:  # A = identity
                check(lib.ur_nll_data(ptr(dat), ptr(require_cuda_f32(y[c].dat, 'y.dat')),
                                      dat.numel(), _hs(obs.tau), ptr(out[1:2]), 1, stream()))
                continue
            # 0.5 tau sum_{x != 0} (x - A y)^2: one pass over y for lattice operators (A y is
            # never materialised), A y in the workspace + reduction otherwise
            s = proj_struct(obs.po, sett.method)
            nbytes = lib.ur_proj_workspace_bytes(C.byref(s)) + 4 * dat.numel() + 256
            ws = _lib.workspace(nbytes, dat.device, 'proj')
            check(lib.ur_nll_data_proj(C.byref(s), ptr(require_cuda_f32(y[c].dat, 'y.dat')),
                                       ptr(dat), _hs(obs.tau), ptr(out[1:2]), 1, ptr(ws),
                                       ws.numel(), stream()))
    ys = [require_cuda_f32(yc.dat, 'y.dat') for yc in y]
    lam = _lib.farr([_hs(yc.lam) for yc in y])
    field = prior_field
    if field is None:
        field = _lib.workspace(4 * n_vox, ys[0].device, 'prior').view(torch.float32)[:n_vox]
    for c0 in range(0, len(ys), _lib.UR_MAX_CHANNELS):
        chunk = ys[c0:c0 + _lib.UR_MAX_CHANNELS]
        acc = 1 if (c0 > 0 or accumulate_prior) else 0
        check(lib.ur_nll_prior_energy(_ptr_array(chunk), ptr(field), len(chunk),
                                      _lib.farr(list(lam)[c0:c0 + len(chunk)]), i3(dim), f3(vx),
                                      acc, stream()))
    if prior_field is None:
        check(lib.ur_sqrt_sum(ptr(field), n_vox, ptr(out[2:3]), stream()))


def _compute_nll(x, y, sett, rho, sum_dtype=torch.float64):
    """(nll, nll_xy, nll_y): negative log posterior / likelihood / prior, float64."""
    if sum_dtype != torch.float64:
        raise NotImplementedError('sums are float64')
    out = torch.zeros(3, dtype=torch.float64, device=y[0].dat.device)
    _nll_terms(x, y, sett, out)
    return out[1] + out[2], out[1], out[2]


def _rhs(xc, yc, z_c, w_c, rho, tmp, sett, dim, vx, lhs=None):
    """tmp = sum_n tau_n An' x_n - lam div(w_c - rho z_c)   (unires/_update.py:124-133).

    One fused pass when every observation is lattice aligned (`ur_admm_rhs_fused`);
    otherwise At through the general path, accumulated, then the divergence."""
    if lhs is not None:
        dats = [require_cuda_f32(obs.dat, 'x.dat') for obs in xc]
        rc = lib.ur_admm_rhs_fused(C.byref(lhs.c), _ptr_array(dats), ptr(tmp), ptr(w_c), ptr(z_c),
                                   _hs(yc.lam), _hs(rho), stream())
        if rc != _lib.UR_ERR_UNSUPPORTED:
            check(rc)
            return
    tmp.zero_()
    n_vox = tmp.numel()
    for n, obs in enumerate(xc):
        dat = require_cuda_f32(obs.dat, 'x.dat')
        if not sett.do_proj:
            check(lib.ur_axpy(ptr(tmp), ptr(dat), _hs(obs.tau), n_vox, stream()))
            continue
        s = proj_struct(obs.po, sett.method)
        ws = _lib.workspace(lib.ur_proj_workspace_bytes(C.byref(s)), tmp.device, 'proj')
        check(lib.ur_proj_accumulate(_lib.UR_OP_AT, C.byref(s), ptr(dat), ptr(tmp),
                                     _hs(obs.tau), ptr(ws), ws.numel(), stream()))
    check(lib.ur_admm_rhs(ptr(tmp), ptr(w_c), ptr(z_c), i3(dim), f3(vx), _hs(yc.lam),
                          _hs(rho), stream()))


def _backproject(xc, out, lhs, scale=None):
    """out = scale * sum_n An' x_n in one pass (`ur_backproject`; lattice-aligned observations).

    With scale = 1 / (A' 1) this is the normalised back-projection `HostPipeline` users form as
    the initial estimate of a freshly uploaded subject on the device (the role of
    unires/_core.py:371-399 `_init_y_dat`, without uploading the estimate).  Returns False --
    nothing launched -- when an observation is not lattice aligned (rotated operators: use
    `_proj_apply('At')`)."""
    dats = [require_cuda_f32(obs.dat, 'x.dat') for obs in xc]
    rc = lib.ur_backproject(C.byref(lhs.c), _ptr_array(dats), ptr(out),
                            ptr(scale) if scale is not None else None, stream())
    if rc == _lib.UR_ERR_UNSUPPORTED:
        return False
    check(rc)
    return True


def _solve_channel(xc, yc, z_c, w_c, rho, tmp, sett, dim, vx):
    """RHS + device-resident CG of one channel (in place on yc.dat).  Returns its CgInfo."""
    stop = getattr(sett, 'cgs_stop', 'max_gain')
    lhs = LhsOperator(xc, yc, method=sett.method, do=sett.do_proj, rho=rho, vx_y=vx,
                      bound=sett.bound, interpolation=sett.interpolation, diff=sett.diff)
    _rhs(xc, yc, z_c, w_c, rho, tmp, sett, dim, vx, lhs=lhs)
    yc.dat = require_cuda_f32(yc.dat, 'y.dat')
    cg(A=lhs, b=tmp, x=yc.dat, verbose=sett.cgs_verbose, max_iter=sett.cgs_max_iter,
       stop=stop, inplace=True, precond=None, tolerance=sett.cgs_tol)
    return cg.last


def _side_streams(device, n):
    key = torch.device(device).index
    pool = _channel_streams.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=device))
    return pool[:n]


_channel_streams = {}


def _n_streams(sett, n_channels):
    return max(1, min(int(getattr(sett, 'channel_streams', 1) or 1), n_channels))


def _rhs_buffer(dim, device):
    n_vox = dim[0] * dim[1] * dim[2]
    return _lib.workspace(4 * n_vox, device, 'rhs').view(torch.float32)[:n_vox].view(dim)


def _solve_y(x, y, z, w, rho, tmp, sett, dim, vx):
    """y-update: one device-resident CG per channel.  Returns the CgInfo handles.

    The channels are independent linear systems (unires/_update.py:122-150 loops over them
    sequentially); with `sett.channel_streams` > 1 they are enqueued round-robin on that many
    CUDA streams, so that the launch gaps and the ramp-down of one channel's kernels are
    filled by another channel's.  Each stream owns its right-hand side and CG workspace."""
    ns = _n_streams(sett, len(x))
    if ns == 1:
        return [_solve_channel(x[c], y[c], z[c], w[c], rho, tmp, sett, dim, vx)
                for c in range(len(x))]
    main = torch.cuda.current_stream()
    streams = _side_streams(tmp.device, ns)
    start = main.record_event()
    infos = []
    for c in range(len(x)):
        s = streams[c % ns]
        if c < ns:
            s.wait_event(start)
        with torch.cuda.stream(s):
            infos.append(_solve_channel(x[c], y[c], z[c], w[c], rho, _rhs_buffer(dim, tmp.device),
                                        sett, dim, vx))
    for s in streams:
        main.wait_stream(s)
    return infos


def solve_y_from_host(x, y, z, w, rho, tmp, sett, host_x, host_y, host_out, copy_stream=None):
    """The y-update with HOST (pinned) observations / initial estimates / results.

    host_x[c][n] -> x[c][n].dat and host_y[c] -> y[c].dat are uploaded, and y[c].dat ->
    host_out[c] downloaded, on a separate copy stream, per channel, so that the PCIe
    transfers of channel c+1 (and the download of channel c-1) overlap the CG solve of
    channel c.  Returns the CgInfo handles; the caller synchronises before reading host_out."""
    dim, vx = _geometry(y)
    main = torch.cuda.current_stream()
    cs = copy_stream if copy_stream is not None else _copy_stream(main.device)
    cs.wait_stream(main)  # earlier work on the destination buffers is finished
    ready = []
    with torch.cuda.stream(cs):
        for c in range(len(x)):
            for n, obs in enumerate(x[c]):
                obs.dat.copy_(host_x[c][n], non_blocking=True)
            y[c].dat.copy_(host_y[c], non_blocking=True)
            ready.append(cs.record_event())
    ns = _n_streams(sett, len(x))
    streams = _side_streams(main.device, ns) if ns > 1 else [main]
    start = main.record_event()
    infos = []
    for c in range(len(x)):
        s = streams[c % ns]
        if ns > 1 and c < ns:
            s.wait_event(start)
        s.wait_event(ready[c])
        with torch.cuda.stream(s):
            b = _rhs_buffer(dim, main.device) if ns > 1 else tmp
            infos.append(_solve_channel(x[c], y[c], z[c], w[c], rho, b, sett, dim, vx))
            solved = s.record_event()
        with torch.cuda.stream(cs):
            cs.wait_event(solved)
            host_out[c].copy_(y[c].dat, non_blocking=True)
    for s in streams:
        if s is not main:
            main.wait_stream(s)
    main.wait_stream(cs)
    return infos


class HostPipeline:
    """y-updates of a STREAM of subjects whose observations / initial estimates live in host
    (pinned) memory and whose reconstructions are wanted back on the host.

    `sets` are two or more device-side container sets [(x, y), ...] of identical geometry; subject
    k uses set k % len(sets).  Per subject and channel: upload on the H2D stream, CG solve on
    that channel's stream (`_solve_channel`), download on the D2H stream -- so the PCIe
    transfers of subject k+1 and k-1 overlap the solves of subject k in both directions.  A
    set is reused only after its previous download has completed (event, no host sync).
    Every byte of every subject crosses PCIe; `drain()` waits for the last result."""

    def __init__(self, sets, z, w, rho, sett, init_y=None):
        """init_y(x, y, c): optional device-side initial estimate of channel c from its freshly
        uploaded observations (then `host_y` of submit() may be None and is not uploaded)."""
        self.sets, self.z, self.w, self.rho, self.sett = sets, z, w, rho, sett
        self.init_y = init_y
        y0 = sets[0][1]
        self.dim, self.vx = _geometry(y0)
        dev = y0[0].dat.device
        self.dev = dev
        self.h2d = torch.cuda.Stream(device=dev)
        self.d2h = torch.cuda.Stream(device=dev)
        self.ns = max(2, _n_streams(sett, len(y0)))
        self.free = [None] * len(sets)
        self.k = 0
        self.h2d.wait_stream(torch.cuda.current_stream(dev))

    def submit(self, host_x, host_y, host_out):
        x, y = self.sets[self.k % len(self.sets)]
        slot = self.k % len(self.sets)
        self.k += 1
        streams = _side_streams(self.dev, self.ns)
        ready = []
        with torch.cuda.stream(self.h2d):
            if self.free[slot] is not None:
                self.h2d.wait_event(self.free[slot])
            for c in range(len(x)):
                for n, obs in enumerate(x[c]):
                    obs.dat.copy_(host_x[c][n], non_blocking=True)
                if host_y is not None:
                    y[c].dat.copy_(host_y[c], non_blocking=True)
                ready.append(self.h2d.record_event())
        infos = []
        for c in range(len(x)):
            s = streams[c % self.ns]
            s.wait_event(ready[c])
            with torch.cuda.stream(s):
                if self.init_y is not None:
                    self.init_y(x, y, c)
                infos.append(_solve_channel(x[c], y[c], self.z[c], self.w[c], self.rho,
                                            _rhs_buffer(self.dim, self.dev), self.sett, self.dim,
                                            self.vx))
                solved = s.record_event()
            with torch.cuda.stream(self.d2h):
                self.d2h.wait_event(solved)
                host_out[c].copy_(y[c].dat, non_blocking=True)
        self.free[slot] = self.d2h.record_event()
        return infos

    def drain(self):
        torch.cuda.current_stream(self.dev).wait_stream(self.d2h)
        torch.cuda.current_stream(self.dev).wait_stream(self.h2d)


_copy_streams = {}


def _copy_stream(device):
    key = torch.device(device).index
    if key not in _copy_streams:
        _copy_streams[key] = torch.cuda.Stream(device=device)
    return _copy_streams[key]


def _update_admm(x, y, z, w, rho, tmp, obj, n_iter, sett):
    """One ADMM iteration (y by CG, objective, JTV prox z, dual w).

    In place on y[c].dat, z, w and obj[n_iter]; returns (y, z, w, jtv, obj) where
    jtv is the shrink-factor image the reference returns in `tmp`."""
    dim, vx = _geometry(y)
    z = require_cuda_f32(z, 'z')
    w = require_cuda_f32(w, 'w')
    tmp = require_cuda_f32(tmp, 'tmp')
    rho_f = _hs(rho)
    _update_admm.last_cg = _solve_y(x, y, z, w, rho_f, tmp, sett, dim, vx)
    if sett.tolerance > 0:
        row = torch.zeros(3, dtype=torch.float64, device=tmp.device)
        _nll_terms(x, y, sett, row)
        row[0] = row[1] + row[2]
        obj[n_iter, :] = row.to(obj.device, obj.dtype)
    # z and w in one pass; the shrink factor lands in tmp (returned, as upstream)
    ys = [yc.dat for yc in y]
    if len(ys) > _lib.UR_MAX_CHANNELS:
        raise NotImplementedError('more than %d channels' % _lib.UR_MAX_CHANNELS)
    check(lib.ur_jtv_prox(_ptr_array(ys), ptr(z), ptr(w), ptr(tmp), len(ys),
                          _lib.farr([_hs(yc.lam) for yc in y]), i3(dim), f3(vx), rho_f,
                          float(sett.alpha), stream()))
    return y, z, w, tmp, obj


_update_admm.last_cg = None


def _update_admm_sharded(x, y, z, w, rho, tmp, obj, n_iter, sett, group=None):
    """Channel-sharded ADMM iteration: this rank holds a subset of the channels
    (x, y, z, w are the LOCAL lists / tensors).  Communication per iteration:
    ONE SUM all-reduce of a (2, X, Y, Z) buffer -- the prior-energy field of the objective and
    the JTV coupling field of the prox, formed back to back -- and one float64 scalar (the data
    term); with sett.tolerance == 0 only the coupling field travels.
    With a single rank it reduces to `_update_admm`."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return _update_admm(x, y, z, w, rho, tmp, obj, n_iter, sett)
    from . import parallel
    if len(y) == 0 or len(y) > _lib.UR_MAX_CHANNELS:
        # parallel.channel_shard rejects world_size > n_channels on every rank up front
        raise ValueError('_update_admm_sharded: every rank needs between 1 and %d channels '
                         '(world size > number of channels?)' % _lib.UR_MAX_CHANNELS)
    dim, vx = _geometry(y)
    n_vox = dim[0] * dim[1] * dim[2]
    rho_f = _hs(rho)
    _update_admm.last_cg = _solve_y(x, y, z, w, rho_f, tmp, sett, dim, vx)
    ys = [yc.dat for yc in y]
    lam = _lib.farr([_hs(yc.lam) for yc in y])
    alpha = float(sett.alpha)
    norm2 = lambda f: check(lib.ur_jtv_norm2(_ptr_array(ys), ptr(z), ptr(w), ptr(f), len(ys), lam,
                                             i3(dim), f3(vx), rho_f, alpha, 0, stream()))
    apply = lambda f: check(lib.ur_jtv_apply(_ptr_array(ys), ptr(z), ptr(w), ptr(f), ptr(tmp),
                                             len(ys), lam, i3(dim), f3(vx), rho_f, alpha, stream()))
    fields = _lib.workspace(8 * n_vox, tmp.device, 'shard_fields').view(torch.float32)[:2 * n_vox] \
        .view((2,) + tuple(dim))
    if sett.tolerance > 0:
        row = torch.zeros(3, dtype=torch.float64, device=tmp.device)

        def sqrt_sum(f):
            check(lib.ur_sqrt_sum(ptr(f), n_vox, ptr(row[2:3]), stream()))
            return row[2].clone()

        # prior-energy field and JTV coupling field in ONE all-reduce (2 N floats)
        parallel.coupled_objective_and_prox(
            row, fields, lambda r, f: _nll_terms(x, y, sett, r, prior_field=f), norm2, sqrt_sum,
            apply, group)
        obj[n_iter, :] = row.to(obj.device, obj.dtype)
    else:
        parallel.coupled_prox(fields[1], norm2, apply, group)
    return y, z, w, tmp, obj


def _update_scaling(x, y, sett, max_niter_gn=1, num_linesearch=4, verbose=0):
    """Gauss-Newton update of the even/odd slice scaling `po.scl` of every observation
    (unires/_update.py:270-393; derivation in derivations/scaling.m).  Returns (x, sll).

    Per observation: dat_y = A y with the current scaling (pull -> slice profile -> exp(+-s)),
    then over the voxels with x != 0 the five float64 sums of `ur_scaling_sums` give the
    log-likelihood 0.5 tau sum (x - y)^2, the gradient tau (sum_e y (x-y) - sum_o y (x-y)) and
    the Fisher Hessian tau (sum_e y^2 + sum_o y^2), where -- as upstream -- "odd" are the slices
    ::2 and "even" the slices 1::2 along the thick axis.  The line search rescales dat_y in
    place by exp(+-(s - s_old)) at every trial, cumulatively, exactly like the reference.
    One host synchronisation per trial (the reference's `if ll < old_ll`)."""
    import math
    from ._project import _proj_apply
    dev = y[0].dat.device
    sums = torch.zeros(5, dtype=torch.float64, device=dev)
    sll = 0.0
    for c in range(len(x)):
        for obs in x[c]:
            if getattr(obs, 'ct', False):
                continue
            po = obs.po
            thick = int(po.dim_thick)
            tau = float(np.float32(_hs(obs.tau)))
            scl = _hs(po.scl) if po.scl is not None else 0.0
            dat_x = require_cuda_f32(obs.dat, 'x.dat')
            dim_x = i3(tuple(dat_x.shape))
            dat_y = _proj_apply('A', y[c].dat[None, None, ...], po, method='super-resolution',
                                bound=sett.bound, interpolation=sett.interpolation)[0, 0]
            dat_y = require_cuda_f32(dat_y, 'A y')
            if scl == 0.0:
                dat_y = dat_y.clone()  # rescaled in place below; A may have returned a view

            def measure():
                check(lib.ur_scaling_sums(ptr(dat_x), ptr(dat_y), dim_x, thick, ptr(sums),
                                          stream()))
                return sums.tolist()

            ll = 0.0
            for _ in range(max_niter_gn):
                s = measure()
                ll = float(np.float32(0.5 * tau)) * s[0]
                gr = tau * (s[2] - s[1])
                hes = tau * (s[4] + s[3])
                update = gr / hes
                old_scl, old_ll, armijo = scl, ll, 1.0
                if num_linesearch == 0:
                    scl = old_scl - armijo * update
                else:
                    for _ls in range(num_linesearch):
                        scl = old_scl - armijo * update
                        # torch: exp(float64 0-dim) * float32 volume -> the factor rounded to float32
                        f0 = float(np.float32(math.exp(scl - old_scl)))
                        f1 = float(np.float32(math.exp(-(scl - old_scl))))
                        check(lib.ur_scale_slices(ptr(dat_y), ptr(dat_y), dim_x, f0, f1, thick,
                                                  stream()))
                        ll = float(np.float32(0.5 * tau)) * measure()[0]
                        if ll < old_ll:
                            break
                        scl, ll = old_scl, old_ll
                        armijo *= 0.5
            po.scl = torch.tensor(scl, dtype=torch.float64, device=dev)
            sll += ll
    return x, torch.tensor(sll, dtype=torch.float64, device=dev)


# ---------------------------------------------------------------------------
# rigid Gauss-Newton update  (unires/_update.py:198-267, 448-710)
# ---------------------------------------------------------------------------
def _expm(q, basis, grad_X=False):
    """exp(sum_i q_i B_i) (4x4, float64, host) and optionally d/dq_i, via the exact identity
    exp([[X, B], [0, X]]) = [[exp X, dexp_X(B)], [0, exp X]] (nitorch.core._linalg_expm)."""
    q = torch.as_tensor(q).detach().to('cpu', torch.float64)
    basis = torch.as_tensor(basis).detach().to('cpu', torch.float64)
    X = torch.einsum('k,kij->ij', q, basis)
    R = torch.linalg.matrix_exp(X)
    if not grad_X:
        return R
    n = X.shape[-1]
    dR = []
    for B in basis:
        blk = torch.zeros(2 * n, 2 * n, dtype=torch.float64)
        blk[:n, :n] = X
        blk[n:, n:] = X
        blk[:n, n:] = B
        dR.append(torch.linalg.matrix_exp(blk)[:n, n:])
    return R, torch.stack(dR)


def _rigid_match(dat_x, dat_y, po, tau, rigid, sett, CtC=None, diff=False, verbose=0):
    """Rigid matching term and (diff=True) its derivatives w.r.t. the sampling coordinates on
    the intermediate grid (unires/_update.py:448-538).  Returns (ll, gr (*dim, 3), res) where
    gr is the UNSCALED spatial gradient of the warped recon and res = C'(A y - x) (masked):
    the reference's gr_m = gr * res and Hes_m = (gr_a gr_b) CtC are formed inside
    `ur_rigid_sums`, which reduces them against d(coordinates)/dq in the same pass."""
    from ._project import _proj_apply, _slice_profile, proj_struct
    from .spatial import affine_grad
    sr = sett.method == 'super-resolution'
    trial = _lib.copy_bag(po)
    trial.rigid = rigid
    trial.__dict__.pop('_c_cache', None)
    Ay = _proj_apply('A', dat_y, trial, method=sett.method, bound=sett.bound,
                     interpolation=sett.interpolation)[0, 0]
    out = torch.zeros(1, dtype=torch.float64, device=Ay.device)
    check(lib.ur_nll_data(ptr(dat_x), ptr(Ay), dat_x.numel(), float(np.float32(tau)), ptr(out), 0,
                          stream()))
    if not diff:
        return out[0], None, None
    dim = tuple(po.dim_yx) if sr else tuple(po.dim_x)
    s = proj_struct(trial, sett.method)
    mat = torch.tensor(list(s.mat), dtype=torch.float32).reshape(3, 4)
    gr = affine_grad(dat_y[0, 0], mat, dim)
    res = Ay - dat_x
    res[(dat_x == 0) | (Ay == 0)] = 0
    if sr:
        res = _slice_profile(res, po, transpose=True)
    return out[0], gr, res


def _update_rigid_channel(xc, yc, sett, max_niter_gn=1, num_linesearch=4, verbose=0, samp=3, c=1):
    """Gauss-Newton update of rigid_q / po.rigid of every observation of one channel
    (unires/_update.py:541-710).  The 6 + 21 chain-rule sums run in one kernel
    (`ur_rigid_sums`); the 6x6 solve and the line-search logic stay on the host."""
    from ._project import _proj_info, _slice_profile
    from .spatial import affine_grid, grid_pull
    dev = yc.dat.device
    basis = torch.as_tensor(sett.rigid_basis).detach().to('cpu', torch.float64)
    num_q = basis.shape[0]
    if num_q != 6:
        raise NotImplementedError('rigid update: SE(3) basis (6 parameters) only')
    sr = sett.method == 'super-resolution'
    sums = torch.zeros(27, dtype=torch.float64, device=dev)
    cpu64 = lambda t: torch.as_tensor(t).detach().to('cpu', torch.float64)
    sll = 0.0
    for obs in xc:
        q = cpu64(obs.rigid_q)
        tau = _hs(obs.tau)
        armijo = 1.0
        po = _proj_info(obs.po.dim_y, obs.po.mat_y, obs.po.dim_x, obs.po.mat_x, rigid=obs.po.rigid,
                        prof_ip=sett.profile_ip, prof_tp=sett.profile_tp, gap=sett.gap,
                        device=dev, scl=obs.po.scl, samp=samp)
        dim = tuple(po.dim_yx) if sr else tuple(po.dim_x)
        src_mat = cpu64(po.mat_yx if sr else po.mat_x)
        mat_y = cpu64(po.mat_y)
        dat_y = yc.dat[None, None, ...]
        if samp > 0 and po.D_x is not None:
            grid = affine_grid(po.D_x.to(torch.float32), tuple(po.dim_x))
            dat_x = grid_pull(obs.dat[None, None, ...], grid[None, ...], bound='zero',
                              extrapolate=False, interpolation=0)[0, 0]
        else:
            dat_x = obs.dat
        dat_x = require_cuda_f32(dat_x, 'x.dat')
        CtC = None
        if sr:
            ones = torch.ones(dim, dtype=torch.float32, device=dev)
            CtC = _slice_profile(_slice_profile(ones, po), po, transpose=True)
        rigid, ll = cpu64(obs.po.rigid), 0.0
        for _ in range(max_niter_gn):
            rigid, d_rigid = _expm(q, basis, grad_X=True)
            dm = torch.stack([torch.linalg.solve(mat_y, d_rigid[i] @ src_mat)[:3, :]
                              for i in range(num_q)]).to(torch.float32)
            ll_t, gr, res = _rigid_match(dat_x, dat_y, po, tau, rigid, sett, CtC=CtC, diff=True)
            check(lib.ur_rigid_sums(ptr(gr), ptr(res), ptr(CtC) if CtC is not None else None,
                                    i3(dim), _lib.farr(dm.reshape(-1).tolist()), ptr(sums),
                                    stream()))
            vals = sums.tolist()
            ll = float(ll_t)
            g = torch.tensor(vals[:6], dtype=torch.float64).reshape(6, 1)
            H = torch.zeros(6, 6, dtype=torch.float64)
            iu = torch.triu_indices(6, 6)
            H[iu[0], iu[1]] = torch.tensor(vals[6:], dtype=torch.float64)
            H = torch.triu(H) + torch.triu(H, 1).T
            step = torch.linalg.solve(H, g)[:, 0]
            old_ll, old_q, old_rigid = ll, q.clone(), rigid.clone()
            if num_linesearch == 0:
                q = old_q - armijo * step
                rigid = _expm(q, basis)
            for _ls in range(num_linesearch):
                q = old_q - armijo * step
                rigid = _expm(q, basis)
                ll = float(_rigid_match(dat_x, dat_y, po, tau, rigid, sett)[0])
                if ll < old_ll:
                    armijo = min(1.25 * armijo, 1.0)
                    break
                ll, q, rigid = old_ll, old_q, old_rigid
                armijo *= 0.5
        obs.rigid_q = q.to(dev)
        obs.po.rigid = rigid.to(dev)
        sll += ll
    return xc, torch.tensor(sll, dtype=torch.float64, device=dev)


def _update_rigid(x, y, sett, mean_correct=True, max_niter_gn=1, num_linesearch=4, verbose=0,
                  samp=3):
    """Rigid registration parameters of every observation (unires/_update.py:198-267);
    returns (x, sll).  With mean_correct the mean of all q is subtracted afterwards."""
    dev = y[0].dat.device
    sll = torch.zeros((), dtype=torch.float64, device=dev)
    for c in range(len(x)):
        x[c], s = _update_rigid_channel(x[c], y[c], sett, max_niter_gn=max_niter_gn,
                                        num_linesearch=num_linesearch, verbose=verbose,
                                        samp=samp, c=c)
        sll = sll + s
    if mean_correct:
        basis = torch.as_tensor(sett.rigid_basis).detach().to('cpu', torch.float64)
        qs = [torch.as_tensor(o.rigid_q).detach().to('cpu', torch.float64) for xc in x for o in xc]
        mean_q = sum(qs) / float(len(qs))
        for xc in x:
            for o in xc:
                qn = torch.as_tensor(o.rigid_q).detach().to('cpu', torch.float64) - mean_q
                o.rigid_q = qn.to(dev)
                o.po.rigid = _expm(qn, basis).to(dev)
    return x, sll
