"""Independent CPU restatement of UniRes' ADMM/CG hot path (the travelling oracle).

TEST INFRASTRUCTURE, PARITY UNPINNED (see oracle/__init__.py).  The GPU box
has no /root/reference, so the parity tests there compare the CUDA path with
THIS port; in the build container the port itself is checked against the
reference's unmodified files run on the same shim primitives
(tests/test_oracle_vs_reference.py) and against tests/golden/.

Reference lines restated (relative to /root/reference):
  ProjGeometry / proj_info   unires/_project.py:193-297
  apply_scaling              unires/_project.py:9-24
  proj_apply                 unires/_project.py:99-190
  proj                       unires/_project.py:54-96
  dtd                        unires/_project.py:300-317
  admm_aux / step_size       unires/_update.py:17-64
  compute_nll                unires/_update.py:396-427
  update_admm                unires/_update.py:105-195
  even_odd / update_scaling  unires/_update.py:430-445, 270-393
  rigid_match / update_rigid_channel / update_rigid   unires/_update.py:448-538, 541-710, 198-267
  init_y_dat                 unires/_core.py:371-399
"""
import math
import types

import torch
from torch.nn import functional as F

from oracle.nitorch_shim import spatial as S
from oracle.nitorch_shim.core import kernels as K
from oracle.nitorch_shim.core import optim as O

F64 = torch.float64


# ----------------------------------------------------------------------------
# containers (same field names as unires/struct.py so that one scenario can be
# handed to the reference files, to this port and to the CUDA product)
# ----------------------------------------------------------------------------
def Observation(dat, mat, tau, po=None, mu=1.0, sd=1.0, ct=False):
    return types.SimpleNamespace(dat=dat, dim=tuple(dat.shape), mat=mat, tau=tau,
                                 po=po, mu=mu, sd=sd, ct=ct, rigid_q=None,
                                 label=None)


def Recon(dat, mat, lam):
    return types.SimpleNamespace(dat=dat, dim=tuple(dat.shape), mat=mat, lam=lam,
                                 lam0=lam, label=None)


def Settings(**kw):
    s = types.SimpleNamespace(alpha=1.0, bound='zero', cgs_max_iter=20,
                              cgs_tol=1e-3, cgs_verbose=False, device='cpu',
                              diff='forward', do_proj=True, do_print=0,
                              interpolation='linear',
                              method='super-resolution', rho=None, rho_scl=1.0,
                              tolerance=1e-4, profile_ip=2, profile_tp=0,
                              gap=0.0, max_iter=512, reg_scl=4.0, sched_num=3,
                              rigid_mod=1, clean_fov=False, scaling=False,
                              unified_rigid=False, rigid_samp=1, rigid_basis=None)
    for k, v in kw.items():
        setattr(s, k, v)
    return s


# ----------------------------------------------------------------------------
# operator geometry  (unires/_project.py:193-297)
# ----------------------------------------------------------------------------
def proj_info(dim_y, mat_y, dim_x, mat_x, rigid=None, prof_ip=0, prof_tp=0,
              gap=0.0, scl=0.0, samp=0):
    po = types.SimpleNamespace()
    mat_y = mat_y.to(F64)
    mat_x = mat_x.to(F64)
    nd = len(dim_y)
    po.dim_y = tuple(int(d) for d in dim_y)
    po.dim_x = tuple(int(d) for d in dim_x)
    po.mat_y, po.mat_x = mat_y, mat_x
    po.vx_y, po.vx_x = S.voxel_size(mat_y), S.voxel_size(mat_x)
    po.rigid = torch.eye(nd + 1, dtype=F64) if rigid is None else rigid.to(F64)
    # thick-slice axis: first maximum of the input voxel size (:241)
    thick = int(torch.max(po.vx_x, dim=0)[1])
    po.dim_thick = thick
    po.D_x = po.D_y = None
    if samp > 0:
        # sub-sampled observation grid for the rigid update (:245-264): keep every sk-th voxel,
        # sk = max(1, round(samp / vx_x)).  (Upstream's high-res branch compares vx_x with
        # itself and never fires, so D_y stays None.)
        sk = torch.clamp(torch.floor(samp / po.vx_x + 0.5), min=1.0)
        po.D_x = torch.diag(torch.cat([sk, torch.ones(1, dtype=F64)]))
        po.mat_x = mat_x = mat_x @ po.D_x
        po.dim_x = tuple(int(math.floor(d / k)) for d, k in zip(po.dim_x, sk.tolist()))
        po.vx_x = S.voxel_size(mat_x)
    profile = [prof_ip] * nd
    profile[thick] = prof_tp
    gaps = [0.0] * nd
    gaps[thick] = gap
    # integer decimation factor per axis (:266-268)
    y2x = torch.linalg.solve(mat_y, mat_x)[:nd, :nd]
    ratio = (y2x ** 2).sum(0).sqrt().ceil().clamp(1)
    po.ratio = tuple(int(r) for r in ratio.tolist())
    # intermediate grid: x orientation, voxel size vx_x / ratio (:269-271)
    shrink = torch.diag(torch.cat([1 / ratio, torch.ones(1, dtype=F64)]))
    mat_yx = mat_x @ shrink
    dim_yx = [(dx - 1) * r + 1 for dx, r in zip(po.dim_x, po.ratio)]
    # slice profile; ratio 1 -> dirac (:273-277)
    fwhm = [(1.0 - g) * r for g, r in zip(gaps, po.ratio)]
    profile = [-1 if r == 1 else p for p, r in zip(profile, po.ratio)]
    po.smo_ker = K.smooth(profile, fwhm, sep=False, dtype=torch.float32)
    # pad the intermediate grid so the valid strided conv lands on dim_x (:280-285)
    off = [-((k - 1) // 2) for k in po.smo_ker.shape[-nd:]]
    shift = torch.eye(nd + 1, dtype=F64)
    shift[:nd, -1] = torch.tensor(off, dtype=F64)
    po.mat_yx = mat_yx @ shift
    po.dim_yx = tuple(d + 2 * abs(o) for d, o in zip(dim_yx, off))
    po.scl = scl if isinstance(scl, torch.Tensor) else torch.tensor(scl, dtype=torch.float32)
    return po


def apply_scaling(dat, scl, dim):
    """exp(+scl) on even, exp(-scl) on odd slices along spatial axis dim (:9-24).
    The factors are formed in scl's own dtype (float64 once _update_scaling has run) and only
    then rounded to the data type, like torch's `0-dim tensor * volume`."""
    scl = torch.as_tensor(scl)
    n = dat.shape[dat.dim() - 3 + dim]
    factor = torch.exp(scl).to(dat.dtype).repeat(n)
    factor[1::2] = torch.exp(-scl).to(dat.dtype)
    shape = [1] * dat.dim()
    shape[dat.dim() - 3 + dim] = n
    return dat * factor.reshape(shape)


def proj_apply(operator, dat, po, method='super-resolution', bound='zero',
               interpolation='linear'):
    """dat is (1,1,X,Y,Z).  A = S.C.P ; At = P'.C'.S ; AtA = P'C' S^2 C P  (:99-190)."""
    if operator not in ('A', 'At', 'AtA', 'none'):
        raise ValueError('Undefined operator')
    if method not in ('denoising', 'super-resolution'):
        raise ValueError('Undefined method')
    if operator == 'none':
        return dat
    sr = method == 'super-resolution'
    src_mat, src_dim = (po.mat_yx, po.dim_yx) if sr else (po.mat_x, po.dim_x)
    vox = torch.linalg.solve(po.mat_y, po.rigid @ src_mat)  # :147,150
    grid = S.affine_grid(vox.to(dat.dtype), src_dim)[None]
    kw = dict(bound=bound, extrapolate=False, interpolation=interpolation)
    scl, thick = po.scl, int(po.dim_thick)
    ker, stride = po.smo_ker.to(dat.dtype), po.ratio

    def down(v):
        return F.conv3d(v, ker, stride=stride) if sr else v

    def up(v):
        return F.conv_transpose3d(v, ker, stride=stride) if sr else v

    def scale(v, s):
        return apply_scaling(v, s, thick) if (sr and scl != 0) else v

    if operator == 'A':
        return scale(down(S.grid_pull(dat, grid, **kw)), scl)
    if operator == 'At':
        return S.grid_push(up(scale(dat, scl)), grid, shape=po.dim_y, **kw)
    mid = scale(down(S.grid_pull(dat, grid, **kw)), 2 * scl)
    return S.grid_push(up(mid), grid, shape=po.dim_y, **kw)


def dtd(dat, vx_y, bound='zero', diff='forward'):
    return S.im_divergence(S.im_gradient(dat, vx=vx_y, bound=bound, which=diff),
                           vx=vx_y, bound=bound, which=diff)


def proj(operator, dat, x, y, method='super-resolution', do=True, rho=1, n=0,
         vx_y=None, interpolation='linear', bound='zero', diff='forward'):
    """x: list of observations of ONE channel, y: that channel's recon (:54-96)."""
    op = operator if do else 'none'
    kw = dict(method=method, bound=bound, interpolation=interpolation)
    if operator != 'AtA':
        return proj_apply(op, dat[None, None], x[n].po, **kw)[0, 0]
    out = None
    for obs in x:
        term = obs.tau * proj_apply(op, dat[None, None], obs.po, **kw)
        out = term if out is None else out + term
    out = out[0, 0]
    return out + rho * y.lam ** 2 * dtd(dat, vx_y, bound=bound, diff=diff)


# ----------------------------------------------------------------------------
# solver  (unires/_update.py)
# ----------------------------------------------------------------------------
def admm_aux(y):
    shape = (len(y), 3) + tuple(y[0].dim)
    return torch.zeros(shape), torch.zeros(shape)


def step_size(x, y, sett):
    """rho = rho_scl * sqrt(mean tau) / mean lam unless fixed (:35-64)."""
    rho = sett.rho
    if any(obs.ct for xc in x for obs in xc):
        rho = 1.0
    if rho is not None:
        return torch.tensor(rho, dtype=torch.float32)
    lam = torch.tensor([float(yc.lam) for yc in y], dtype=torch.float32)
    tau = torch.tensor([float(o.tau) for xc in x for o in xc], dtype=torch.float32)
    return sett.rho_scl * torch.sqrt(tau.mean()) / lam.mean()


def compute_nll(x, y, sett, rho=None):
    """(nll, nll_xy, nll_y) in float64 (:396-427)."""
    vx_y = S.voxel_size(y[0].mat).float()
    nll_xy = torch.zeros((), dtype=F64)
    prior = None
    for xc, yc in zip(x, y):
        for n, obs in enumerate(xc):
            fit = proj('A', yc.dat, xc, yc, n=n, method=sett.method, do=sett.do_proj,
                       bound=sett.bound, interpolation=sett.interpolation)
            msk = obs.dat != 0
            nll_xy = nll_xy + 0.5 * obs.tau * torch.sum((obs.dat[msk] - fit[msk]) ** 2, dtype=F64)
        g = yc.lam * S.im_gradient(yc.dat, vx=vx_y, bound=sett.bound, which=sett.diff)
        e = torch.sum(g ** 2, dim=0)
        prior = e if prior is None else prior + e
    nll_y = torch.sum(torch.sqrt(prior), dtype=F64)
    return nll_xy + nll_y, nll_xy, nll_y


def update_admm(x, y, z, w, rho, tmp, obj, n_iter, sett, cg_stop='max_gain',
                cg_record=None):
    """One ADMM iteration: y by CG per channel, objective, JTV prox z, dual w (:105-195).

    Returns (y, z, w, jtv, obj, cg_iters)."""
    vx_y = S.voxel_size(y[0].mat).float()
    alpha = float(sett.alpha)
    C = len(x)
    kw = dict(method=sett.method, do=sett.do_proj, bound=sett.bound,
              interpolation=sett.interpolation)
    cg_iters = []
    # ---- y ----
    for c in range(C):
        rhs = torch.zeros_like(tmp)
        for n, obs in enumerate(x[c]):
            rhs += obs.tau * proj('At', obs.dat, x[c], y[c], n=n, **kw)
        rhs -= y[c].lam * S.im_divergence(w[c] - rho * z[c], vx=vx_y,
                                          bound=sett.bound, which=sett.diff)

        def lhs(v, c=c):
            return proj('AtA', v, x[c], y[c], rho=rho, vx_y=vx_y, diff=sett.diff, **kw)

        rec = (lambda it, xi, c=c: cg_record(c, it, xi)) if cg_record else None
        O.cg(A=lhs, b=rhs, x=y[c].dat, verbose=sett.cgs_verbose,
             max_iter=sett.cgs_max_iter, stop=cg_stop, inplace=True,
             precond=lambda v: v, tolerance=sett.cgs_tol, record=rec)
        cg_iters.append(O.cg.last_n_iter)
    # ---- objective ----
    if sett.tolerance > 0:
        obj[n_iter, 0], obj[n_iter, 1], obj[n_iter, 2] = compute_nll(x, y, sett, rho)
    # ---- z (JTV prox) ----
    z_old = z.clone() if alpha != 1 else None

    def scaled_grad(c):
        g = y[c].lam * S.im_gradient(y[c].dat, vx=vx_y, bound=sett.bound, which=sett.diff)
        if alpha != 1:
            g = alpha * g + (1 - alpha) * z_old[c]
        return g

    nrm = torch.zeros_like(tmp)
    for c in range(C):
        nrm += torch.sum((w[c] / rho + scaled_grad(c)) ** 2, dim=0)
    nrm.sqrt_()
    one = torch.tensor(1, dtype=torch.float32)
    tiny = torch.tensor(1e-7, dtype=torch.float32)
    jtv = (nrm - one / rho).clamp_min(0) / (nrm + tiny)
    for c in range(C):
        g = scaled_grad(c)
        for d in range(3):
            z[c, d] = jtv * (w[c, d] / rho + g[d])
    # ---- w ----
    for c in range(C):
        w[c] += rho * (scaled_grad(c) - z[c])
    return y, z, w, jtv, obj, cg_iters


# ----------------------------------------------------------------------------
# outer loop  (unires/run.py:24-207, default path: no scaling / rigid updates;
# schedule from unires/_core.py:288-307, output clamp from unires/_core.py:619-627)
# ----------------------------------------------------------------------------
def get_sched(N, sett):
    """Coarse-to-fine regularisation scaling, e.g. reg_scl=4, sched_num=3 -> [32, 16, 8, 4]."""
    sched_num = 0 if (sett.sched_num < 0 or N == 1) else sett.sched_num
    scl = torch.as_tensor(sett.reg_scl, dtype=torch.float32).reshape(1)
    powers = 2.0 ** torch.arange(0, 32, dtype=torch.float32).flip(0)
    ix = int(torch.min((powers - scl).abs(), dim=0)[1])
    return torch.cat((powers[ix - sched_num:ix], scl))


def fit(x, y, sett):
    """Returns (dat_y (X,Y,Z,C), obj[:n_done], n_done, jtv)."""
    N = sum(len(xc) for xc in x)
    reg = get_sched(N, sett)
    cnt_scl = 0
    for yc in y:
        yc.lam = reg[cnt_scl] * yc.lam0
    rho = step_size(x, y, sett)
    z, w = admm_aux(y)
    obj = torch.zeros(sett.max_iter, 3, dtype=F64)
    tmp = torch.zeros_like(y[0].dat)
    cnt_scl_iter, countdown0, countdown1, n_done = 0, 6, 6, 0
    for n_iter in range(sett.max_iter):
        y, z, w, tmp, obj, _ = update_admm(x, y, z, w, rho, tmp, obj, n_iter, sett)
        n_done = n_iter + 1
        gain = O.get_gain(obj[:n_iter + 1, 0], monotonicity='decreasing')
        if cnt_scl >= reg.numel() - 1 and cnt_scl_iter > 20 and \
                (gain.abs() < sett.tolerance or n_iter >= sett.max_iter - 1):
            countdown0 -= 1
            if countdown0 == 0:
                break
        else:
            countdown0 = 6
        if getattr(sett, 'scaling', False):  # unires/run.py:115-122
            x, _ = update_scaling(x, y, sett, max_niter_gn=1, num_linesearch=6)
        if getattr(sett, 'unified_rigid', False) and n_iter > 0 and n_iter % sett.rigid_mod == 0:
            x, _ = update_rigid(x, y, sett, mean_correct=False, max_niter_gn=1, num_linesearch=6,
                                samp=sett.rigid_samp)  # unires/run.py:127-135
        if cnt_scl + 1 < len(reg) and cnt_scl_iter > 16 and gain.abs() < 1e-3:
            countdown1 -= 1
            if countdown1 == 0:
                cnt_scl_iter = 0
                cnt_scl += 1
                for yc in y:
                    yc.lam = reg[cnt_scl] * yc.lam0
                rho = step_size(x, y, sett)
        else:
            countdown1 = 6
        cnt_scl_iter += 1
    if getattr(sett, 'clean_fov', False):
        for xc, yc in zip(x, y):
            msk = torch.ones(yc.dim, dtype=torch.bool)
            for obs in xc:
                M = torch.linalg.solve(yc.mat, obs.po.rigid @ obs.mat).inverse()
                grid = S.affine_grid(M.to(obs.dat.dtype), yc.dim)
                for d in range(3):
                    msk &= (grid[..., d] >= 0) & (grid[..., d] < obs.dim[d])
            yc.dat[~msk] = 0.0
    out = []
    for xc, yc in zip(x, y):
        mn = min(float(obs.dat.min()) for obs in xc)
        mx = max(float(obs.dat.max()) for obs in xc)
        yc.dat.clamp_(mn, mx)
        out.append(yc.dat[..., None].clone())
    return torch.cat(out, dim=3), obj[:n_done], n_done, tmp


# ----------------------------------------------------------------------------
# even/odd slice-scaling update  (unires/_update.py:270-393, 430-445)
# ----------------------------------------------------------------------------
def even_odd(dat, which, dim):
    """Upstream's naming: 'odd' = slices 0, 2, 4, ...; 'even' = slices 1, 3, 5, ... (:430-445)."""
    start = 0 if which == 'odd' else 1
    index = [slice(None)] * 3
    index[dim] = slice(start, None, 2)
    return dat[tuple(index)]


def update_scaling(x, y, sett, max_niter_gn=1, num_linesearch=4):
    """One Gauss-Newton update (with backtracking) of every observation's po.scl.
    Returns (x, sll).  The operator's rigid matrix is po.rigid (upstream rebuilds it from
    rigid_q, which is what _proj_info stored there)."""
    sll = torch.tensor(0, dtype=F64)
    for c in range(len(x)):
        for obs in x[c]:
            if obs.ct:
                continue
            po = obs.po
            thick, tau, scl = int(po.dim_thick), obs.tau, po.scl
            dat_x = obs.dat
            msk = dat_x != 0
            m_odd, m_even = even_odd(msk, 'odd', thick), even_odd(msk, 'even', thick)
            x_odd, x_even = even_odd(dat_x, 'odd', thick)[m_odd], even_odd(dat_x, 'even', thick)[m_even]
            # reconstruction in observation space: pull -> slice profile -> scaling (:312-318)
            vox = torch.linalg.solve(po.mat_y, po.rigid @ po.mat_yx)
            grid = S.affine_grid(vox.to(torch.float32), po.dim_yx)[None]
            dat_y = S.grid_pull(y[c].dat[None, None], grid, bound=sett.bound, extrapolate=False,
                                interpolation=sett.interpolation)
            dat_y = F.conv3d(dat_y, po.smo_ker, stride=po.ratio)[0, 0]
            dat_y = apply_scaling(dat_y, scl, thick)

            def loglik(d):
                return 0.5 * tau * torch.sum((dat_x[msk] - d[msk]) ** 2, dtype=F64)

            ll = torch.tensor(0, dtype=F64)
            for _ in range(max_niter_gn):
                ll = loglik(dat_y)
                y_odd, y_even = even_odd(dat_y, 'odd', thick)[m_odd], even_odd(dat_y, 'even', thick)[m_even]
                grad = tau * (torch.sum(y_even * (x_even - y_even), dtype=F64)
                              - torch.sum(y_odd * (x_odd - y_odd), dtype=F64))
                hess = tau * (torch.sum(y_even ** 2, dtype=F64) + torch.sum(y_odd ** 2, dtype=F64))
                step = grad / hess
                old_scl, old_ll = scl.clone(), ll.clone()
                armijo = torch.tensor(1.0, dtype=old_scl.dtype)
                if num_linesearch == 0:
                    scl = old_scl - armijo * step
                for _ls in range(num_linesearch):
                    scl = old_scl - armijo * step
                    dat_y = apply_scaling(dat_y, scl - old_scl, thick)  # cumulative, as upstream
                    ll = loglik(dat_y)
                    if ll < old_ll:
                        break
                    scl, ll = old_scl, old_ll
                    armijo = armijo * 0.5
            po.scl = scl
            sll = sll + ll
    return x, sll


# ----------------------------------------------------------------------------
# rigid Gauss-Newton update  (unires/_update.py:198-267, 448-710)
# ----------------------------------------------------------------------------
def expm(q, basis, grad=False):
    from oracle.nitorch_shim.core._linalg_expm import _expm
    return _expm(q, basis, grad_X=grad)


def rigid_match(dat_x, dat_y, po, tau, rigid, sett, CtC=None, diff=False):
    """Matching term 0.5 tau sum_{x != 0} (x - A y)^2 for the rigid matrix `rigid` and, with
    diff=True, its first / second derivatives w.r.t. the sampling coordinates on the
    intermediate grid: (ll, gr (*dim, 3), Hes (*dim, 6))  (:448-538)."""
    sr = sett.method == 'super-resolution'
    dim, src_mat = (po.dim_yx, po.mat_yx) if sr else (po.dim_x, po.mat_x)
    vox = torch.linalg.solve(po.mat_y, rigid @ src_mat)
    grid = S.affine_grid(vox.to(torch.float32), dim)[None]
    kw = dict(bound=sett.bound, extrapolate=False, interpolation=sett.interpolation)
    warped = S.grid_pull(dat_y, grid, **kw)[0, 0]
    if sr:
        warped = F.conv3d(warped[None, None], po.smo_ker, stride=po.ratio)[0, 0]
        if po.scl != 0:
            warped = apply_scaling(warped, po.scl, int(po.dim_thick))
    msk = dat_x != 0
    ll = 0.5 * tau * torch.sum((dat_x[msk] - warped[msk]) ** 2, dtype=F64)
    if not diff:
        return ll, None, None
    g = S.grid_grad(dat_y, grid, **kw)[0, 0]
    res = warped - dat_x
    res[~(msk & (warped != 0))] = 0
    pairs = ((0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2))
    Hes = torch.stack([g[..., a] * g[..., b] for a, b in pairs], dim=-1)
    if sr:
        Hes = Hes * CtC[..., None]
        res = F.conv_transpose3d(res[None, None], po.smo_ker, stride=po.ratio)[0, 0]
    return ll, g * res[..., None], Hes


def update_rigid_channel(xc, yc, sett, max_niter_gn=1, num_linesearch=4, samp=3):
    """Gauss-Newton update of rigid_q (and po.rigid) of every observation of one channel
    (:541-710).  Returns (xc, sll)."""
    basis = sett.rigid_basis
    num_q = basis.shape[0]
    lkp = ((0, 3, 4), (3, 1, 5), (4, 5, 2))
    sr = sett.method == 'super-resolution'
    one = torch.tensor(1.0, dtype=F64)
    sll = torch.tensor(0, dtype=F64)
    for obs in xc:
        q, tau = obs.rigid_q, obs.tau
        armijo = torch.tensor(1, dtype=q.dtype)
        po = proj_info(obs.po.dim_y, obs.po.mat_y, obs.po.dim_x, obs.po.mat_x, rigid=obs.po.rigid,
                       prof_ip=sett.profile_ip, prof_tp=sett.profile_tp, gap=sett.gap,
                       scl=obs.po.scl, samp=samp)
        dim, src_mat = (po.dim_yx, po.mat_yx) if sr else (po.dim_x, po.mat_x)
        dat_y = yc.dat[None, None]
        if samp > 0 and po.D_x is not None:
            grid = S.affine_grid(po.D_x.to(torch.float32), po.dim_x)[None]
            dat_x = S.grid_pull(obs.dat[None, None], grid, bound='zero', extrapolate=False,
                                interpolation=0)[0, 0]
        else:
            dat_x = obs.dat
        CtC = None
        if sr:  # C'C 1: diagonal scaling of the Gauss-Newton Hessian
            CtC = F.conv3d(torch.ones((1, 1) + tuple(dim), dtype=torch.float32), po.smo_ker,
                           stride=po.ratio)
            CtC = F.conv_transpose3d(CtC, po.smo_ker, stride=po.ratio)[0, 0]
        ident = S.identity_grid(dim, dtype=torch.float32)
        rigid, ll = obs.po.rigid, torch.tensor(0, dtype=F64)
        for _ in range(max_niter_gn):
            rigid, d_rigid = expm(q, basis, grad=True)
            # derivative of the voxel-to-voxel matrix mat_y \ rigid mat w.r.t. q_i
            dM = [torch.linalg.solve(po.mat_y, d_rigid[i] @ src_mat) for i in range(num_q)]
            ll, gr_m, Hes_m = rigid_match(dat_x, dat_y, po, tau, rigid, sett, CtC=CtC, diff=True)
            # d(coordinate d) / d q_i at every voxel of the intermediate grid (float32, as torch
            # evaluates `float64 scalar * float32 volume`)
            dA = [[dM[i][d, 0] * ident[..., 0] + dM[i][d, 1] * ident[..., 1]
                   + dM[i][d, 2] * ident[..., 2] + dM[i][d, 3] for d in range(3)]
                  for i in range(num_q)]
            gr = torch.zeros(num_q, 1, dtype=F64)
            Hes = torch.zeros(num_q, num_q, dtype=F64)
            for d in range(3):
                for i in range(num_q):
                    gr[i] += torch.sum(gr_m[..., d] * dA[i][d], dtype=F64)
            for d1 in range(3):
                for d2 in range(3):
                    for i1 in range(num_q):
                        left = Hes_m[..., lkp[d1][d2]] * dA[i1][d1]
                        for i2 in range(i1, num_q):
                            Hes[i1, i2] += torch.sum(left * dA[i2][d2], dtype=F64)
            Hes = torch.triu(Hes) + torch.triu(Hes, 1).T
            step = torch.linalg.solve(Hes, gr)[:, 0]
            old_ll, old_q, old_rigid = ll.clone(), q.clone(), rigid.clone()
            if num_linesearch == 0:
                q = old_q - armijo * step
                rigid = expm(q, basis)
            for _ls in range(num_linesearch):
                q = old_q - armijo * step
                rigid = expm(q, basis)
                ll = rigid_match(dat_x, dat_y, po, tau, rigid, sett)[0]
                if ll < old_ll:
                    armijo = torch.min(1.25 * armijo, one)
                    break
                ll, q, rigid = old_ll, old_q, old_rigid
                armijo = armijo * 0.5
        obs.rigid_q = q
        obs.po.rigid = rigid
        sll = sll + ll
    return xc, sll


def update_rigid(x, y, sett, mean_correct=True, max_niter_gn=1, num_linesearch=4, samp=3):
    """Rigid update of every observation, optionally mean-corrected over all of them (:198-267)."""
    sll = torch.tensor(0, dtype=F64)
    for c in range(len(x)):
        x[c], s = update_rigid_channel(x[c], y[c], sett, max_niter_gn=max_niter_gn,
                                       num_linesearch=num_linesearch, samp=samp)
        sll = sll + s
    if mean_correct:
        qs = [obs.rigid_q for xc in x for obs in xc]
        total = torch.zeros(sett.rigid_basis.shape[0], dtype=F64)
        for q in qs:
            total += q
        mean_q = total / float(len(qs))
        for xc in x:
            for obs in xc:
                obs.rigid_q -= mean_q
                obs.po.rigid = expm(obs.rigid_q, sett.rigid_basis)
    return x, sll


# ----------------------------------------------------------------------------
# initial estimate  (unires/_core.py:371-399)
# ----------------------------------------------------------------------------
def init_y_dat(x, y, sett):
    """Trilinear pull of every observation into the recon grid, clamped to the observation's
    range, averaged over the repeats that are positive at a voxel."""
    dim_y, mat_y = tuple(y[0].dim), y[0].mat
    for c in range(len(x)):
        total = torch.zeros(dim_y, dtype=torch.float32)
        count = torch.zeros(dim_y, dtype=torch.float32)
        for obs in x[c]:
            dat = obs.dat[None, None]
            vox = torch.linalg.solve(obs.mat, mat_y)
            grid = S.affine_grid(vox.to(dat.dtype), dim_y)[None]
            lo, hi = torch.min(dat), torch.max(dat)
            pulled = S.grid_pull(dat, grid, bound='zero', extrapolate=False, interpolation=1)
            pulled[pulled < lo] = lo
            pulled[pulled > hi] = hi
            count = count + (pulled[0, 0] > 0)
            total = total + pulled[0, 0]
        count[count == 0] = 1.0
        y[c].dat = total / count
    return y


# ----------------------------------------------------------------------------
# hyper-parameter estimate  (unires/_core.py:96-142)
# ----------------------------------------------------------------------------
def estimate_hyperpar(x):
    """Noise sd / precision and mean foreground intensity of every observation from a two-class
    mixture fit to its histogram (non-negative voxels only unless the observation is CT)."""
    from oracle.nitorch_shim.tools.img_statistics import estimate_noise
    for xc in x:
        for obs in xc:
            dat = obs.dat if obs.ct else obs.dat[obs.dat >= 0]
            noise, rest = estimate_noise(dat, num_class=2)
            obs.sd = noise['sd'].float()
            obs.tau = 1 / noise['sd'].float() ** 2
            obs.mu = torch.abs(rest['mean'].float() - noise['mean'].float())
    return x
