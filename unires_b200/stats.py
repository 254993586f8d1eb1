"""Noise / foreground statistics of an observed volume (the reference's hyper-parameter
estimate, unires/_core.py:96-142, which calls nitorch.tools.img_statistics.estimate_noise).

The volume-sized work -- range and 1024-bin histogram with the reference's selections
(dat >= 0 for MR, zeros and the maximum masked out) -- runs in two streaming CUDA passes
(csrc/stats.cu: ur_intensity_range, ur_histc).  The mixture fit then works on the 1024 bins on
the host in float64: Rician classes for non-negative data, Gaussian classes otherwise; the
M-step is the method of moments of SPM12's spm_rice_mixture.m (Koay & Basser inversion), which
nitorch ports.  nitorch is absent from the reference tree, so the details that are not visible
from UniRes' call site follow SPM's published code: PARITY UNPINNED; soft
pin = the reference's notebook logs (tests/test_hyperpar.py).
"""
import ctypes as C
import math

import torch

from . import _lib

_SQRT_HALF_PI = math.sqrt(math.pi / 2)


def intensity_histogram(dat, bins=1024, drop_negative=False):
    """(W, x, mn, mx): float64 CPU tensors of bin counts and bin positions
    linspace(mn, mx, bins), with mn = round(min), mx = round(max of the voxels that are
    neither zero nor the global maximum)."""
    if not dat.is_cuda:
        raise RuntimeError('unires_b200.stats: the volume must live on the GPU (no CPU fallback)')
    v = dat.detach().reshape(-1)
    if v.dtype != torch.float32:
        v = v.float()
    v = v.contiguous()
    n = v.numel()
    rng = (C.c_float * 2)()
    any_ = C.c_int32(0)
    _lib.check(_lib.lib.ur_intensity_range(_lib.ptr(v), n, int(drop_negative), 0, 0.0, rng,
                                           C.byref(any_), _lib.stream()))
    if not any_.value:
        raise ValueError('estimate_noise: no voxels selected')
    mn, gmax = float(round(rng[0])), float(rng[1])
    _lib.check(_lib.lib.ur_intensity_range(_lib.ptr(v), n, int(drop_negative), 1, gmax, rng,
                                           C.byref(any_), _lib.stream()))
    if not any_.value:
        raise ValueError('estimate_noise: the volume is constant')
    mx = float(round(rng[1]))
    counts = torch.empty(bins, dtype=torch.int64, device=v.device)
    _lib.check(_lib.lib.ur_histc(_lib.ptr(v), n, int(drop_negative), 1, gmax, mn, mx, bins,
                                 _lib.ptr(counts), _lib.stream()))
    W = counts.cpu().double()
    x = torch.linspace(mn, mx, steps=bins, dtype=torch.float64)
    return W, x, mn, mx


def _rice_log_pdf(x, nu, sig):
    """(N, K) log densities of K Rician classes at the N bin positions; Gaussian form where
    the Bessel form would overflow (spm_rice_mixture.m: ricepdf)."""
    sig2 = (sig * sig)[None, :]
    X = x[:, None]
    t = -(X * X + (nu * nu)[None, :]) / (2 * sig2)
    a = X * (nu[None, :] / sig2)
    ok = (t > -95) & (a < 85)
    zero = torch.zeros_like(t)
    rice = (X / sig2) * torch.exp(torch.where(ok, t, zero)) * torch.special.i0(torch.where(ok, a, zero))
    gauss = torch.exp((-0.5 / sig2) * (X - nu[None, :]) ** 2) / torch.sqrt(2 * math.pi * sig2)
    return torch.log(torch.where(ok, rice, gauss) + 1e-32)


def _moments_to_rice(mu1, mu2):
    """Rician (nu, sigma) from the mean and variance of a class."""
    r = mu1 / math.sqrt(mu2)
    theta = math.sqrt(math.pi / (4 - math.pi))
    if r <= theta:
        return 0.0, math.sqrt(0.5 * (mu1 * mu1 + mu2))
    xi = 1.0
    for _ in range(256):
        q = torch.tensor(theta * theta / 4, dtype=torch.float64)
        i0, i1 = float(torch.special.i0(q)), float(torch.special.i1(q))
        xi = 2 + theta ** 2 - math.pi / 8 * math.exp(-theta ** 2 / 2) * \
            ((2 + theta ** 2) * i0 + theta ** 2 * i1) ** 2
        g = math.sqrt(xi * (1 + r * r) - 2)
        if abs(theta - g) < 1e-6:
            break
        theta = g
    if not math.isfinite(xi):
        xi = 1.0
    sig = math.sqrt(mu2 / xi)
    return math.sqrt(mu1 * mu1 + (xi - 2) * sig * sig), sig


def _rice_mean(nu, sig):
    u = -nu * nu / (2 * sig * sig)
    if u <= -20:
        return nu
    q = torch.tensor(-u / 2, dtype=torch.float64)
    lag = math.exp(u / 2) * ((1 - u) * float(torch.special.i0(q)) - u * float(torch.special.i1(q)))
    return _SQRT_HALF_PI * sig * lag


def fit_mixture(W, x, num_class=2, rician=True, max_iter=10000, tol=1e-8):
    """Weighted EM over histogram bins.  Returns (mp, mean, sd, n_iter): mixing proportions,
    class expectations and class sigmas (float64 tensors of length num_class)."""
    K = num_class
    W, x = W.double(), x.double()
    mn, mx, sw = float(x.min()), float(x.max()), float(W.sum())
    lam = (float((x * W).sum()) / (sw * K)) ** 2  # variance regulariser
    mp = torch.full((K,), 1.0 / K, dtype=torch.float64)
    sig = torch.full((K,), (mx - mn) / (K * 10), dtype=torch.float64)
    k = torch.arange(K, dtype=torch.float64)
    nu = k * mx / (K + 1) if rician else mn + (k + 1) * (mx - mn) / (K + 1)
    lo = hi = last = None
    n_iter = 0
    for n_iter in range(1, max_iter + 1):
        if rician:
            logp = _rice_log_pdf(x, nu, sig)
        else:
            logp = -0.5 * torch.log(2 * math.pi * sig * sig)[None, :] - \
                0.5 * ((x[:, None] - nu[None, :]) / sig[None, :]) ** 2
        Z = torch.log(mp)[None, :] + logp
        lse = torch.logsumexp(Z, dim=1)
        lb = float((lse * W).sum())
        # relative gain of the lower bound over its range so far (increasing objective)
        lo = lb if lo is None else min(lo, lb)
        hi = lb if hi is None else max(hi, lb)
        if last is not None:
            gain = (lb - last) / (hi - lo) if hi > lo else 0.0
            if gain < tol:
                break
        last = lb
        R = torch.exp(Z - lse[:, None]) * W[:, None]
        ss0, ss1, ss2 = R.sum(0), (R * x[:, None]).sum(0), (R * (x * x)[:, None]).sum(0)
        mp = ss0 / sw
        mu1 = ss1 / ss0
        mu2 = (ss2 - ss1 * ss1 / ss0 + lam * 1e-3) / (ss0 + 1e-3)
        if rician:
            prm = [_moments_to_rice(float(mu1[c]), float(mu2[c])) for c in range(K)]
            nu = torch.tensor([p[0] for p in prm], dtype=torch.float64)
            sig = torch.tensor([p[1] for p in prm], dtype=torch.float64)
        else:
            nu, sig = mu1, torch.sqrt(mu2)
    if rician:
        mean = torch.tensor([_rice_mean(float(nu[c]), float(sig[c])) for c in range(K)],
                            dtype=torch.float64)
    else:
        mean = nu.clone()
    return mp, mean, sig, n_iter


def noise_from_mixture(mp, mu, sd, mu_noise=None):
    """Noise class = smallest sigma (or the class whose mean is closest to mu_noise); the
    remaining classes are pooled with their mixing proportions."""
    ix = int(torch.argmin(torch.abs(mu - mu_noise))) if mu_noise else int(torch.argmin(sd))
    rest = [c for c in range(mu.numel()) if c != ix]
    w = mp[rest] / mp[rest].sum()
    return (dict(sd=sd[ix], mean=mu[ix], mp=mp[ix]),
            dict(sd=(w * sd[rest]).sum(), mean=(w * mu[rest]).sum(), mp=mp[rest].sum()))


def estimate_noise(dat, show_fit=False, fig_num=1, num_class=2, mu_noise=None, max_iter=10000,
                   verbose=0, bins=1024, chi=False, drop_negative=False):
    """nitorch.tools.img_statistics.estimate_noise as UniRes calls it (unires/_core.py:122):
    returns (prm_noise, prm_not_noise), dicts with 0-dim float64 'sd', 'mean', 'mp'.
    `drop_negative=True` folds UniRes' `dat[dat >= 0]` (unires/_core.py:118) into the kernels."""
    if show_fit:
        raise NotImplementedError('estimate_noise: show_fit (plotting) is out of scope')
    if chi:
        raise NotImplementedError('estimate_noise: chi mixtures are not used by UniRes')
    W, x, mn, mx = intensity_histogram(dat, bins, drop_negative)
    mp, mu, sd, n_iter = fit_mixture(W, x, num_class, rician=mn >= 0, max_iter=max_iter)
    estimate_noise.last = dict(n_iter=n_iter, mn=mn, mx=mx, mp=mp, mean=mu, sd=sd)
    return noise_from_mixture(mp, mu, sd, mu_noise)
