"""`nitorch.tools.img_statistics.estimate_noise` restated.  TEST INFRASTRUCTURE (oracle).

UniRes calls it once per observation (unires/_core.py:122-124) to get the noise standard
deviation (-> tau = 1/sd^2) and the mean foreground intensity (-> mu, lam0 = sqrt(1/C)/mu).
nitorch (pinned at 8067d605, absent here) fits a two-class mixture to the 1024-bin intensity
histogram: Rician classes for non-negative (MR) data, Gaussian classes when the minimum is
negative (CT).  nitorch's implementation is a port of SPM12's published
`spm_noise_estimate.m` / `spm_rice_mixture.m` (Ashburner; method-of-moments M-step after Koay &
Basser 2006), which is what is restated here, with nitorch's histogram conventions
(zeros and the maximum masked out, `histc` over [round(min), round(max)], bin positions
`linspace(min, max, bins)`) and its relative-gain stop rule.

PARITY UNPINNED: no copy of nitorch to diff against.  Soft pin: the reference's notebook logs
(demos/demo_single_channel.ipynb cell 5: BrainWeb T1, thick slices x4 along z, even/odd scaling
0.1, N(0, 75^2) noise -> "sd=48.64 | mu=406.5"), reproduced statistically by
tests/test_hyperpar.py (the notebook's noise came from the CUDA RNG).
"""
import math

import torch


def _unavailable(*args, **kwargs):  # pragma: no cover
    raise NotImplementedError('nitorch.tools is out of scope (SURVEY.md section 2)')


atlas_crop = affine_align = atlas_align = reset_origin = _unavailable
estimate_fwhm = _unavailable
_bb_atlas = _mean_space = _unavailable


def _besseli(order, x):
    return torch.special.i0(x) if order == 0 else torch.special.i1(x)


def histogram(dat, bins=1024):
    """(W, x, mn, mx): counts and bin positions of the masked intensities (float64)."""
    dat = torch.as_tensor(dat).flatten().double()
    dat = torch.where(torch.isfinite(dat), dat, torch.zeros_like(dat))
    mn = dat.min().round()
    msk = (dat != 0) & (dat != dat.max())
    dat = dat[msk]
    mx = dat.max().round()
    W = torch.histc(dat, bins=bins, min=float(mn), max=float(mx)).double()
    x = torch.linspace(float(mn), float(mx), steps=bins, dtype=torch.float64)
    return W, x, float(mn), float(mx)


def _get_gain(lb):
    if len(lb) < 2:
        return float('inf')
    rng = max(lb) - min(lb)
    return (lb[-1] - lb[-2]) / rng if rng > 0 else 0.0


def _rice_log_pdf(x, nu, sig):
    """log of the Rician density, with SPM's switch to a Gaussian where the Bessel form
    would overflow (spm_rice_mixture.m: ricepdf)."""
    sig2 = sig * sig
    tmp = -(x * x + nu * nu) / (2 * sig2)
    msk = (tmp > -95) & (x * (nu / sig2) < 85)
    arg = torch.where(msk, x * (nu / sig2), torch.zeros_like(x))
    rice = (x / sig2) * torch.exp(torch.where(msk, tmp, torch.zeros_like(x))) * _besseli(0, arg)
    gauss = (1.0 / math.sqrt(2 * math.pi * sig2)) * torch.exp((-0.5 / sig2) * (x - nu) ** 2)
    return torch.log(torch.where(msk, rice, gauss) + 1e-32)


def _moments_to_rice(mu1, mu2):
    """(nu, sig) of a Rician from its mean and variance (Koay & Basser fixed point)."""
    r = mu1 / math.sqrt(mu2)
    theta = math.sqrt(math.pi / (4 - math.pi))
    if r > theta:
        xi = 1.0
        for _ in range(256):
            t2 = torch.tensor(theta * theta / 4, dtype=torch.float64)
            xi = 2 + theta ** 2 - math.pi / 8 * math.exp(-theta ** 2 / 2) * \
                ((2 + theta ** 2) * float(_besseli(0, t2)) + theta ** 2 * float(_besseli(1, t2))) ** 2
            g = math.sqrt(xi * (1 + r * r) - 2)
            if abs(theta - g) < 1e-6:
                break
            theta = g
        if not math.isfinite(xi):
            xi = 1.0
        sig = math.sqrt(mu2) / math.sqrt(xi)
        nu = math.sqrt(mu1 ** 2 + (xi - 2) * sig ** 2)
    else:
        nu = 0.0
        sig = (2 ** -0.5) * math.sqrt(mu1 ** 2 + mu2)
    return nu, sig


def _rice_mean(nu, sig):
    """E[x] of a Rician (Laguerre polynomial L_1/2), spm_noise_estimate.m."""
    x = -nu * nu / (2 * sig * sig)
    if x > -20:
        t = torch.tensor(-x / 2, dtype=torch.float64)
        lag = math.exp(x / 2) * ((1 - x) * float(_besseli(0, t)) - x * float(_besseli(1, t)))
        return math.sqrt(math.pi * sig * sig / 2) * lag
    return nu


def fit_mixture(W, x, num_class=2, rician=True, max_iter=10000, tol=1e-8):
    """EM on a weighted histogram.  Returns (mp, mean, sd) per class as float64 tensors;
    for Rician classes `sd` is the Rician sigma and `mean` the expectation of the class."""
    W = W.double()
    x = x.double()
    K = num_class
    mn, mx = float(x.min()), float(x.max())
    sw = float(W.sum())
    mp = torch.full((K,), 1.0 / K, dtype=torch.float64)
    lam = (float((x * W).sum()) / (sw * K)) ** 2
    if rician:
        nu = [k * mx / (K + 1) for k in range(K)]
        sig = [(mx - mn) / (K * 10)] * K
    else:
        nu = [mn + (k + 1) * (mx - mn) / (K + 1) for k in range(K)]
        sig = [(mx - mn) / (K * 10)] * K
    lb = []
    for _ in range(max_iter):
        Z = torch.stack([math.log(float(mp[k])) +
                         (_rice_log_pdf(x, nu[k], sig[k]) if rician else
                          -0.5 * math.log(2 * math.pi * sig[k] ** 2) - 0.5 * ((x - nu[k]) / sig[k]) ** 2)
                         for k in range(K)], dim=1)
        lse = torch.logsumexp(Z, dim=1)
        Z = torch.exp(Z - lse[:, None])
        lb.append(float((lse * W).sum()))
        if _get_gain(lb) < tol:
            break
        Z = Z * W[:, None]
        ss0 = Z.sum(0)
        ss1 = (Z * x[:, None]).sum(0)
        ss2 = (Z * (x * x)[:, None]).sum(0)
        mp = ss0 / sw
        for k in range(K):
            s0, s1, s2 = float(ss0[k]), float(ss1[k]), float(ss2[k])
            mu1 = s1 / s0
            mu2 = (s2 - s1 * s1 / s0 + lam * 1e-3) / (s0 + 1e-3)
            if rician:
                nu[k], sig[k] = _moments_to_rice(mu1, mu2)
            else:
                nu[k], sig[k] = mu1, math.sqrt(mu2)
    mean = [(_rice_mean(nu[k], sig[k]) if rician else nu[k]) for k in range(K)]
    return mp, torch.tensor(mean, dtype=torch.float64), torch.tensor(sig, dtype=torch.float64)


def noise_from_mixture(mp, mu, sd, mu_noise=None):
    """Noise class = smallest sd (or closest mean to `mu_noise`); the other classes are
    averaged with their mixing proportions."""
    K = mu.numel()
    if mu_noise:
        ix = int(torch.argmin(torch.abs(mu - mu_noise)))
    else:
        ix = int(torch.argmin(sd))
    rest = [k for k in range(K) if k != ix]
    w = mp[rest] / mp[rest].sum()
    prm_noise = dict(sd=sd[ix], mean=mu[ix], mp=mp[ix])
    prm_not_noise = dict(sd=(w * sd[rest]).sum(), mean=(w * mu[rest]).sum(), mp=mp[rest].sum())
    return prm_noise, prm_not_noise


def estimate_noise(dat, show_fit=False, fig_num=1, num_class=2, mu_noise=None, max_iter=10000,
                   verbose=0, bins=1024, chi=False):
    """Noise / not-noise statistics of an image from a mixture fit to its histogram.
    Returns (prm_noise, prm_not_noise): dicts with 'sd', 'mean', 'mp' (0-dim float64)."""
    W, x, mn, mx = histogram(dat, bins)
    mp, mu, sd = fit_mixture(W, x, num_class, rician=mn >= 0, max_iter=max_iter)
    return noise_from_mixture(mp, mu, sd, mu_noise)
