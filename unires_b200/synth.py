"""Synthetic UniRes workloads (phantoms, geometries, simulated observations).

Used by bench.py and the tests: a deterministic piecewise-smooth head phantom
in BrainWeb's intensity range and the thick-slice forward simulation of the
reference's notebooks (demos/demo_multi_channel.ipynb cell 4: thick slices,
rect profile, Gaussian noise, tau = 1/sd^2, lam from the mean intensity as in
unires/_core.py:279).  Everything here is backend-agnostic: the projection
operators are supplied by the caller (`ops`), so the same scenario can be
built with the CUDA product, with the CPU oracle or with the reference's own
files.
"""
import math
import types

import torch

# BASELINE.json configs -> geometry.  dim_y: recon grid; fov: extent of the
# simulated scanner FOV inside it (None = whole grid); thick: per-channel
# (axis, factor) of the thick-slice direction.
CONFIGS = {
    # configs[1]: 3-channel BrainWeb-like (181x217x181 FOV) super-resolution on a 256^3 grid,
    # each channel thick-sliced x4 along a different axis (notebook recipe)
    'sr3_256': dict(dim_y=(256, 256, 256), fov=(181, 217, 181), vx_y=1.0,
                    thick=[(0, 4), (1, 4), (2, 4)]),
    # configs[1] with the notebook's rigid misalignment (demos/demo_multi_channel.ipynb cell 4:
    # translations within +-5 mm, rotations within +-0.1 rad): rotated operators, i.e. what
    # every real multi-channel run executes once `unified_rigid` has moved the scans
    'sr3_256_rigid': dict(dim_y=(256, 256, 256), fov=(181, 217, 181), vx_y=1.0,
                          thick=[(0, 4), (1, 4), (2, 4)],
                          rigid=[((4.1, -2.7, 3.3), (0.08, -0.05, 0.1)),
                                 ((-3.2, 4.6, -1.9), (-0.1, 0.07, 0.04)),
                                 ((2.4, 1.8, -4.4), (0.05, 0.1, -0.09))]),
    # configs[1], literal reading: 1 mm observations (181x217x181) on the 256^3 grid -- the
    # operator is an integer-shift crop / zero-pad embed, A'A a field-of-view mask
    'crop3_256': dict(dim_y=(256, 256, 256), fov=(181, 217, 181), vx_y=1.0, thick=[None] * 3),
    # configs[2]: 3-channel 2 mm -> 1 mm thick-slice (z x2), 256^3
    'thickz2_256': dict(dim_y=(256, 256, 256), fov=None, vx_y=1.0,
                        thick=[(2, 2), (2, 2), (2, 2)]),
    # configs[3]: 8-channel 384^3, z x2, one channel per GPU
    'thickz2_384x8': dict(dim_y=(384, 384, 384), fov=None, vx_y=1.0, thick=[(2, 2)] * 8),
    # configs[4]: 0.5 mm recon of 1 mm isotropic data, 512^3 (ratio 2 on every axis)
    'iso2_512': dict(dim_y=(512, 512, 512), fov=None, vx_y=0.5, thick=[None] * 3, vx_x=1.0),
    # one channel of configs[3] (what one GPU holds)
    'thickz2_384': dict(dim_y=(384, 384, 384), fov=None, vx_y=1.0, thick=[(2, 2)]),
    # other clinical slice ratios (tuning / coverage of the lean kernel's specialisations)
    'thick3_256': dict(dim_y=(256, 256, 256), fov=None, vx_y=1.0, thick=[(0, 3), (1, 3), (2, 3)]),
    'thick6_256': dict(dim_y=(256, 256, 256), fov=None, vx_y=1.0, thick=[(0, 6), (1, 6), (2, 6)]),
    'thick5_256': dict(dim_y=(256, 256, 256), fov=(181, 217, 181), vx_y=1.0,
                       thick=[(0, 5), (1, 5), (2, 5)]),
    # reduced copy of sr3_256 for quick CPU checks of bench.py
    'sr3_48': dict(dim_y=(48, 48, 48), fov=(34, 41, 34), vx_y=1.0, thick=[(0, 4), (1, 4), (2, 4)]),
    # configs[0]: single-channel denoise on the BrainWeb grid
    'denoise_181': dict(dim_y=(181, 217, 181), fov=None, vx_y=1.0, thick=[None], denoise=True),
}


def scaled(cfg, dim_y, n_channels=None):
    """The same geometry on a smaller grid (for tests)."""
    c = dict(cfg)
    f = [d / D for d, D in zip(dim_y, cfg['dim_y'])]
    c['dim_y'] = tuple(dim_y)
    if cfg.get('fov') is not None:
        c['fov'] = tuple(max(8, int(round(v * s))) for v, s in zip(cfg['fov'], f))
    if n_channels is not None:
        c['thick'] = list(cfg['thick'])[:n_channels]
        if cfg.get('rigid') is not None:
            c['rigid'] = list(cfg['rigid'])[:n_channels]
    return c


def phantom(dim, n_channels, seed=0, device='cpu', channels=None):
    """List of float32 (X,Y,Z) volumes: ellipsoid blobs inside a head-like
    ellipsoid, zero background shell, intensities ~ BrainWeb (0..~1200).

    The random parameters always come from the CPU generator (same phantom everywhere);
    `device` only says where the volumes are rasterised; `channels` selects a subset of the
    `n_channels` channels (multi-GPU: a rank builds only the channels it owns)."""
    g = torch.Generator().manual_seed(seed)
    ax = [torch.linspace(-1, 1, d).to(device) for d in dim]
    X, Y, Z = torch.meshgrid(*ax, indexing='ij')
    head = ((X / 0.82) ** 2 + (Y / 0.9) ** 2 + (Z / 0.8) ** 2) < 1
    n_blob = 20
    cen = torch.rand(n_blob, 3, generator=g) * 1.2 - 0.6
    rad = torch.rand(n_blob, 3, generator=g) * 0.35 + 0.08
    amp = torch.rand(n_channels, n_blob, generator=g) * 600 + 100
    base = torch.rand(n_channels, generator=g) * 300 + 200
    vols = []
    for c in (range(n_channels) if channels is None else channels):
        v = torch.full(dim, float(base[c]), device=device)
        v += 60 * (X + 0.5 * Y * Z)  # low-frequency ramp
        for k in range(n_blob):
            e = ((X - cen[k, 0]) / rad[k, 0]) ** 2 + ((Y - cen[k, 1]) / rad[k, 1]) ** 2 \
                + ((Z - cen[k, 2]) / rad[k, 2]) ** 2
            v = torch.where(e < 1, v * 0 + float(amp[c, k]), v)
        v = torch.where(head, v.clamp_min(1.0), torch.zeros((), device=device))
        vols.append(v.float().contiguous())
    return vols


def _translate(t):
    m = torch.eye(4, dtype=torch.float64)
    m[:3, 3] = torch.tensor(t, dtype=torch.float64)
    return m


def rigid_matrix(t, r):
    """Translation t (mm) and XYZ Euler rotation r (rad), SPM 'classic' order."""
    cx, cy, cz = [math.cos(a) for a in r]
    sx, sy, sz = [math.sin(a) for a in r]
    Rx = torch.tensor([[1, 0, 0], [0, cx, sx], [0, -sx, cx]], dtype=torch.float64)
    Ry = torch.tensor([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=torch.float64)
    Rz = torch.tensor([[cz, sz, 0], [-sz, cz, 0], [0, 0, 1]], dtype=torch.float64)
    m = torch.eye(4, dtype=torch.float64)
    m[:3, :3] = Rx @ Ry @ Rz
    return _translate(t) @ m


def geometry(cfg, c):
    """(dim_x, mat_x, dim_y, mat_y) of channel c of a CONFIGS entry."""
    dim_y = tuple(cfg['dim_y'])
    vx_y = float(cfg['vx_y'])
    mat_y = torch.diag(torch.tensor([vx_y] * 3 + [1.0], dtype=torch.float64))
    fov = cfg.get('fov') or dim_y
    # scanner FOV centred in the recon grid with an integer voxel offset, like the
    # pow-crop of unires/_core.py:251 (round((ndim - dim) / 2))
    off = [float(round((D - d) / 2)) for D, d in zip(dim_y, fov)]
    mat_fov = mat_y @ _translate(off)
    thick = cfg['thick'][c]
    if cfg.get('vx_x') is not None:  # isotropic coarser acquisition
        f = float(cfg['vx_x']) / vx_y
        scl = [f, f, f]
    else:
        scl = [1.0, 1.0, 1.0]
        if thick is not None:
            scl[thick[0]] = float(thick[1])
    mat_x = mat_fov @ torch.diag(torch.tensor(scl + [1.0], dtype=torch.float64))
    dim_x = tuple(int(math.floor(d / s)) for d, s in zip(fov, scl))
    return dim_x, mat_x, dim_y, mat_y


def make_scenario(cfg, ops, structs, device='cpu', seed=0, sd=25.0, reg_scl=4.0, rigid=None,
                  scl=0.0, truth=None, settings_kw=None, channels=None, phantom_device='cpu'):
    """Build (x, y, sett, rho, truth) for one ADMM problem.

    ops:     module/namespace with _proj_info(...) and _proj_apply(op, dat, po, method=...)
             (unires_b200._project, the oracle port, or the reference's _project)
    structs: namespace with _input, _output, settings classes/factories
    rigid:   optional list of 4x4 rigid matrices per channel (correctness cases)
    The simulated observation is x_c = A_c g_c + N(0, sd^2) on the non-zero FOV,
    tau = 1/sd^2, lam = reg_scl * sqrt(1/C) / mean(foreground).
    channels: build only these channels of the C-channel problem (a rank's shard; lam still uses
    the full C); phantom_device: where the ground truth is rasterised."""
    C = len(cfg['thick'])
    sel = list(range(C)) if channels is None else list(channels)
    denoise = bool(cfg.get('denoise'))
    dim_y = tuple(cfg['dim_y'])
    if truth is None:
        truth = phantom(dim_y, C, seed, device=phantom_device, channels=sel)
        truth = dict(zip(sel, truth))
    else:
        truth = dict(zip(sel, truth)) if len(truth) == len(sel) else dict(enumerate(truth))
    if rigid is None and cfg.get('rigid') is not None:
        rigid = [rigid_matrix(t, r) for t, r in cfg['rigid']]
    g = torch.Generator().manual_seed(seed + 1)
    sett = structs.settings()
    sett.device = device
    sett.method = 'denoising' if denoise else 'super-resolution'
    sett.do_proj = not denoise
    sett.do_print = 0
    for k, v in (settings_kw or {}).items():
        setattr(sett, k, v)
    x, y = [], []
    for c in sel:
        dim_x, mat_x, _, mat_y = geometry(cfg, c)
        gt = truth[c].to(device)
        obs = structs._input()
        if denoise:
            clean = gt
            obs.po = None
        else:
            obs.po = ops._proj_info(dim_y, mat_y.to(device), dim_x, mat_x.to(device),
                                    rigid=None if rigid is None else rigid[c],
                                    prof_ip=sett.profile_ip, prof_tp=sett.profile_tp,
                                    gap=sett.gap, device=device, scl=scl)
            clean = ops._proj_apply('A', gt[None, None], obs.po, method=sett.method)[0, 0]
        noise = (sd * torch.randn(tuple(clean.shape), generator=g)).to(device)
        obs.dat = torch.where(clean != 0, (clean + noise), torch.zeros((), device=device)) \
            .float().contiguous()
        obs.dim = tuple(obs.dat.shape)
        obs.mat = mat_x.to(device)
        obs.tau = torch.tensor(1.0 / sd ** 2, dtype=torch.float32, device=device)
        obs.sd = sd
        obs.ct = False
        fg = obs.dat[obs.dat > 0]
        obs.mu = float(fg.mean()) if fg.numel() else 1.0
        x.append([obs])
        rec = structs._output()
        rec.dim = dim_y
        rec.mat = mat_y.to(device)
        lam0 = math.sqrt(1.0 / C) / obs.mu
        rec.lam0 = torch.tensor(lam0, dtype=torch.float32, device=device)
        rec.lam = torch.tensor(reg_scl * lam0, dtype=torch.float32, device=device)
        # initial estimate: adjoint-normalised back-projection (cheap stand-in for the
        # trilinear initialisation of unires/_core.py:371-399)
        if denoise:
            rec.dat = obs.dat.clone()
        else:
            num = ops._proj_apply('At', obs.dat[None, None], obs.po, method=sett.method)[0, 0]
            den = ops._proj_apply('At', torch.ones_like(obs.dat)[None, None], obs.po,
                                  method=sett.method)[0, 0]
            rec.dat = (num / den.clamp_min(1e-3)).float().contiguous()
        y.append(rec)
    lam_mean = sum(float(r.lam) for r in y) / len(sel)
    tau_mean = sum(float(o[0].tau) for o in x) / len(sel)
    rho = torch.tensor(math.sqrt(tau_mean) / lam_mean, dtype=torch.float32, device=device)
    return types.SimpleNamespace(x=x, y=y, sett=sett, rho=rho, truth=[truth[c] for c in sel],
                                 cfg=cfg, channels=sel)
