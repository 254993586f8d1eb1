"""nitorch.core.kernels.smooth restated.

TEST INFRASTRUCTURE, PARITY UNPINNED (see oracle/__init__.py).  Call site:
unires/_project.py:277  ``smooth(profile, fwhm, sep=False, dtype=float32)``.

A slice profile (dirac / rect / triangle / gauss of a given FWHM, in voxels)
analytically convolved with the linear B-spline basis and sampled at integer
offsets (SPM ``spm_smoothkern``-style).  Odd length, symmetric.
"""
import math
import torch


def _tri_cdf(u):
    """CDF of the unit triangle T(u) = max(0, 1-|u|)."""
    u = u.clamp(-1.0, 1.0)
    neg = 0.5 * (u + 1) ** 2
    pos = 1 - 0.5 * (1 - u) ** 2
    return torch.where(u < 0, neg, pos)


def _dirac1d(w):
    return torch.ones(1, dtype=torch.float64)


def _rect1d(w):
    lim = int(math.floor((w + 2) / 2))
    x = torch.arange(-lim, lim + 1, dtype=torch.float64)
    return (_tri_cdf(x + w / 2) - _tri_cdf(x - w / 2)) / w


def _tri1d(w):
    # triangle of FWHM w (half-base w) convolved with the unit triangle;
    # evaluated by fine quadrature (not exercised by UniRes' defaults).
    lim = int(math.floor(w + 1))
    x = torch.arange(-lim, lim + 1, dtype=torch.float64)
    u = torch.linspace(-w, w, 20001, dtype=torch.float64)
    f = (1 - u.abs() / w).clamp_min(0) / w
    basis = (1 - (x[:, None] - u[None, :]).abs()).clamp_min(0)
    return torch.trapezoid(f[None, :] * basis, u, dim=1)


def _gauss1d(w):
    s = w / math.sqrt(8 * math.log(2)) + 1e-7
    lim = int(math.floor(4 * s + 1))
    x = torch.arange(-lim, lim + 1, dtype=torch.float64)
    a = 1 / (math.sqrt(2) * s)
    b = -0.5 / s ** 2
    c = s / math.sqrt(2 * math.pi)
    ker = 0.5 * (torch.erf(a * (x + 1)) * (x + 1)
                 + torch.erf(a * (x - 1)) * (x - 1)
                 - 2 * torch.erf(a * x) * x) \
        + c * (torch.exp(b * (x + 1) ** 2)
               + torch.exp(b * (x - 1) ** 2)
               - 2 * torch.exp(b * x ** 2))
    return ker.clamp_min(0)


_PROFILES = {-1: _dirac1d, 0: _rect1d, 1: _tri1d, 2: _gauss1d}


def smooth1d(profile, fwhm):
    """1-D factor (float64)."""
    return _PROFILES[int(profile)](float(fwhm))


def smooth(types, fwhm=1, basis=1, x=None, sep=True, dtype=None, device=None):
    if basis != 1 or x is not None:
        raise NotImplementedError
    if not isinstance(types, (list, tuple)):
        types = [types]
    fwhm = torch.as_tensor(fwhm, dtype=torch.float64).flatten().tolist()
    nd = max(len(types), len(fwhm))
    types = list(types) + [types[-1]] * (nd - len(types))
    fwhm = list(fwhm) + [fwhm[-1]] * (nd - len(fwhm))
    kers = [smooth1d(t, w) for t, w in zip(types, fwhm)]
    dtype = dtype or torch.get_default_dtype()
    if sep:
        out = []
        for d, k in enumerate(kers):
            shp = [1] * nd
            shp[d] = -1
            out.append(k.reshape([1, 1] + shp).to(dtype=dtype, device=device))
        return out
    full = kers[0]
    for k in kers[1:]:
        full = full[..., None] * k
    return full[None, None].to(dtype=dtype, device=device)
