"""Slice-profile kernels: drop-in for nitorch.core.kernels.smooth as called at
unires/_project.py:277 (``smooth(profile, fwhm, sep=False, dtype=float32)``).

Host-side float64 arithmetic (a handful of taps); the result is a small
tensor.  Profile codes: -1 dirac, 0 rect, 1 triangle, 2 gauss; each is the
slice profile of the given FWHM (voxels) convolved with the linear-
interpolation basis and sampled at integer offsets.
"""
import functools
import math

import torch


def _tri_cdf(u):
    u = min(1.0, max(-1.0, u))
    return 0.5 * (u + 1.0) ** 2 if u < 0 else 1.0 - 0.5 * (1.0 - u) ** 2


def _rect(w):
    half = int((w + 2) // 2)
    return [(_tri_cdf(x + w / 2) - _tri_cdf(x - w / 2)) / w for x in range(-half, half + 1)]


def _gauss(w):
    s = w / math.sqrt(8.0 * math.log(2.0)) + 1e-7
    half = int(math.floor(4.0 * s + 1.0))
    a, b, c = 1.0 / (math.sqrt(2.0) * s), -0.5 / (s * s), s / math.sqrt(2.0 * math.pi)
    taps = []
    for x in range(-half, half + 1):
        xm, xp = x - 1.0, x + 1.0
        v = 0.5 * (math.erf(a * xp) * xp + math.erf(a * xm) * xm - 2.0 * math.erf(a * x) * x) \
            + c * (math.exp(b * xp * xp) + math.exp(b * xm * xm) - 2.0 * math.exp(b * x * x))
        taps.append(max(v, 0.0))
    return taps


def _tri(w):
    half = int(math.floor(w + 1.0))
    n = 20000
    taps = []
    for x in range(-half, half + 1):
        acc = 0.0
        for k in range(n + 1):  # trapezoid rule over the triangle's support [-w, w]
            u = -w + 2.0 * w * k / n
            f = max(0.0, 1.0 - abs(u) / w) / w * max(0.0, 1.0 - abs(x - u))
            acc += f * (0.5 if k in (0, n) else 1.0)
        taps.append(acc * 2.0 * w / n)
    return taps


def smooth1d(profile, fwhm):
    """1-D factor as a list of Python floats (float64); memoised (the rigid update rebuilds the
    operator, hence its taps, for every observation on every step)."""
    return list(_smooth1d_cached(int(profile), float(fwhm)))


@functools.lru_cache(maxsize=256)
def _smooth1d_cached(profile, fwhm):
    if profile == -1:
        return (1.0,)
    if profile == 0:
        return tuple(_rect(fwhm))
    if profile == 1:
        return tuple(_tri(fwhm))
    if profile == 2:
        return tuple(_gauss(fwhm))
    raise ValueError('unknown slice profile %r' % (profile,))


def smooth(types, fwhm=1, basis=1, x=None, sep=True, dtype=None, device=None):
    """nitorch-compatible signature; only basis=1, x=None is implemented."""
    if basis != 1 or x is not None:
        raise NotImplementedError('smooth: only the linear basis on the default support')
    types = list(types) if isinstance(types, (list, tuple)) else [types]
    fwhm = torch.as_tensor(fwhm, dtype=torch.float64).flatten().tolist()
    nd = max(len(types), len(fwhm))
    types += [types[-1]] * (nd - len(types))
    fwhm += [fwhm[-1]] * (nd - len(fwhm))
    dtype = dtype or torch.get_default_dtype()
    factors = [torch.tensor(smooth1d(t, w), dtype=torch.float64) for t, w in zip(types, fwhm)]
    if sep:
        out = []
        for d, k in enumerate(factors):
            shape = [1, 1] + [1] * nd
            shape[2 + d] = k.numel()
            out.append(k.reshape(shape).to(dtype=dtype, device=device))
        return out
    dense = factors[0]
    for k in factors[1:]:
        dense = dense.unsqueeze(-1) * k
    return dense[None, None].to(dtype=dtype, device=device)


def separable_factors(smo_ker, tol=1e-5):
    """Split a dense (1,1,kx,ky,kz) outer-product kernel into its 1-D factors.

    UniRes builds smo_ker with sep=False (unires/_project.py:277) but it is an
    exact outer product; the CUDA path wants the three factors.  Raises
    NotImplementedError if the tensor is not rank-1 separable."""
    k = smo_ker.detach().to('cpu', torch.float64)
    k = k.reshape(k.shape[-3:])
    peak = [int(i) for i in torch.nonzero(k.abs() == k.abs().max())[0]]
    kp = k[peak[0], peak[1], peak[2]]
    if kp == 0:
        raise NotImplementedError('smo_ker is identically zero')
    lines = [k[:, peak[1], peak[2]], k[peak[0], :, peak[2]], k[peak[0], peak[1], :]]
    # k = l0 x l1 x l2 / kp^2: keep the longest line as is, divide the others by
    # kp so that length-1 (dirac) axes come out as exactly [1.0]
    keep = max(range(3), key=lambda a: lines[a].numel())
    f = [l if a == keep else l / kp for a, l in enumerate(lines)]
    rebuilt = f[0][:, None, None] * f[1][None, :, None] * f[2][None, None, :]
    if (rebuilt - k).abs().max() > tol * k.abs().max():
        raise NotImplementedError('smo_ker is not a separable outer product')
    return [t.tolist() for t in f]
