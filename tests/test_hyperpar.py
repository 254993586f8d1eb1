"""Hyper-parameter estimate (unires/_core.py:96-142 -> nitorch estimate_noise).

CPU: the oracle's restatement against known answers and the reference's notebook log (soft pin),
the product's host-side mixture fit against the oracle.  GPU: range / histogram kernels against
the oracle (torch.histc on the float64 copy), `_estimate_hyperpar` end to end."""
import math

import numpy as np
import pytest
import torch

from oracle.nitorch_shim.tools import img_statistics as S
from tests import _util as U


def _rice_sample(n, nu, sig, g):
    a = nu + sig * torch.randn(n, generator=g, dtype=torch.float64)
    b = sig * torch.randn(n, generator=g, dtype=torch.float64)
    return torch.sqrt(a * a + b * b)


def _rice_moments(nu, sig):
    """Exact mean / variance of a Rician (exponentially scaled Bessel functions: no overflow,
    no high-SNR shortcut like the oracle's `_rice_mean`)."""
    t = torch.tensor(nu * nu / (4 * sig * sig), dtype=torch.float64)
    lag = (1 + 2 * t) * torch.special.i0e(t) + 2 * t * torch.special.i1e(t)
    mean = float(sig * math.sqrt(math.pi / 2) * lag)
    return mean, 2 * sig * sig + nu * nu - mean * mean


@pytest.mark.parametrize('nu,sig', [(0.0, 10.0), (30.0, 20.0), (400.0, 60.0), (20.0, 50.0)])
def test_oracle_moment_inversion_round_trip(nu, sig):
    """Koay-Basser inversion: (mean, variance) of a Rician -> its (nu, sigma)."""
    mean, var = _rice_moments(nu, sig)
    nu2, sig2 = S._moments_to_rice(mean, var)
    assert abs(sig2 - sig) < 1e-3 * sig
    assert abs(nu2 - nu) < 1e-3 * max(nu, sig)


def test_oracle_rice_mean_limits():
    assert abs(S._rice_mean(0.0, 10.0) - 10.0 * math.sqrt(math.pi / 2)) < 1e-9  # Rayleigh
    assert abs(S._rice_mean(1000.0, 10.0) - 1000.0) < 0.1                        # high SNR


def test_oracle_recovers_a_known_two_class_rician_mixture():
    g = torch.Generator().manual_seed(0)
    dat = torch.cat([_rice_sample(600000, 0.0, 15.0, g), _rice_sample(400000, 400.0, 60.0, g)])
    noise, rest = S.estimate_noise(dat, num_class=2)
    assert abs(float(noise['sd']) - 15.0) < 0.5
    assert abs(float(noise['mp']) - 0.6) < 0.01
    # above nu^2 / (2 sig^2) = 20 the class mean is reported as nu itself (SPM's shortcut)
    assert abs(float(rest['mean']) - 400.0) < 3.0
    assert abs(float(rest['sd']) - 60.0) < 3.0


def test_oracle_uses_gaussian_classes_for_negative_data():
    g = torch.Generator().manual_seed(1)
    dat = torch.cat([-1000 + 5 * torch.randn(500000, generator=g, dtype=torch.float64),
                     40 + 30 * torch.randn(500000, generator=g, dtype=torch.float64)])
    noise, rest = S.estimate_noise(dat, num_class=2)
    assert abs(float(noise['sd']) - 5.0) < 0.5 and abs(float(noise['mean']) + 1000) < 1.0
    assert abs(float(rest['sd']) - 30.0) < 1.5 and abs(float(rest['mean']) - 40) < 1.5


def test_oracle_on_the_reference_demo_observation_soft_pin():
    """The histogram of the reference's own demo observation (fixture made by
    oracle/gen_golden_hyperpar.py from /root/reference/data): the oracle reproduces its stored
    estimate, which sits within 6 % of the notebook's log -- sd 46.4 vs 48.64, mu 393.7 vs 406.5.
    (The notebook drew its noise on the GPU; three CPU seeds move sd by +-0.1, so the 4.6 % is a
    real difference of an unseen detail of nitorch's fit -- most likely the iteration at which
    its EM stops: sd passes 48.5 three iterations before the first decrease of the bound, where
    this restatement, like SPM, stops.  Parity unpinned.)"""
    g = np.load(U.GOLDEN_DIR + '/hyperpar_t1.npz')
    W = torch.as_tensor(g['W']).double()
    x = torch.linspace(float(g['mn']), float(g['mx']), steps=W.numel(), dtype=torch.float64)
    mp, mu, sd = S.fit_mixture(W, x, 2, rician=True)
    assert np.allclose(mp.numpy(), g['mp'], rtol=1e-9)
    assert np.allclose(mu.numpy(), g['mean'], rtol=1e-9) and np.allclose(sd.numpy(), g['sd'], rtol=1e-9)
    noise, rest = S.noise_from_mixture(mp, mu, sd)
    assert abs(float(noise['sd']) - float(g['notebook_sd'])) < 0.06 * float(g['notebook_sd'])
    mu_fg = abs(float(rest['mean']) - float(noise['mean']))
    assert abs(mu_fg - float(g['notebook_mu'])) < 0.06 * float(g['notebook_mu'])


@pytest.mark.parametrize('rician', [True, False])
def test_product_host_fit_equals_oracle(rician):
    """unires_b200.stats.fit_mixture (vectorised over classes) against the oracle's loop form on
    the same histogram: same trip count, parameters to 1e-10."""
    from unires_b200 import stats
    if rician:
        g = np.load(U.GOLDEN_DIR + '/hyperpar_t1.npz')
        W = torch.as_tensor(g['W']).double()
        x = torch.linspace(float(g['mn']), float(g['mx']), steps=W.numel(), dtype=torch.float64)
    else:
        gen = torch.Generator().manual_seed(2)
        dat = torch.cat([-900 + 8 * torch.randn(200000, generator=gen, dtype=torch.float64),
                         60 + 25 * torch.randn(300000, generator=gen, dtype=torch.float64)])
        W, x, _, _ = S.histogram(dat, 1024)
    for K in (2, 3):
        mp_o, mu_o, sd_o = S.fit_mixture(W, x, K, rician=rician)
        mp, mu, sd, _ = stats.fit_mixture(W, x, K, rician=rician)
        assert torch.allclose(mp, mp_o, rtol=1e-10, atol=1e-12)
        assert torch.allclose(mu, mu_o, rtol=1e-10, atol=1e-10)
        assert torch.allclose(sd, sd_o, rtol=1e-10, atol=1e-10)
        a, b = stats.noise_from_mixture(mp, mu, sd), S.noise_from_mixture(mp_o, mu_o, sd_o)
        for k in ('sd', 'mean', 'mp'):
            assert abs(float(a[0][k]) - float(b[0][k])) < 1e-9 * (1 + abs(float(b[0][k])))
            assert abs(float(a[1][k]) - float(b[1][k])) < 1e-9 * (1 + abs(float(b[1][k])))


def _volume(seed, with_specials=True):
    g = torch.Generator().manual_seed(seed)
    v = (250 * torch.rand(61, 47, 53, generator=g)) ** 1.3
    v[:, :10] = 0.0                                    # zero background
    v += 30 * torch.randn(v.shape, generator=g)         # noise (negative voxels appear)
    v[5, 20:30, 7] = 0.0
    if with_specials:
        v[1, 2, 3] = float('nan')
        v[2, 3, 4] = float('inf')
        v[3, 4, 5] = float('-inf')
    return v


@pytest.mark.gpu
@pytest.mark.parametrize('drop_negative', [True, False])
def test_histogram_kernels_equal_oracle(cuda, drop_negative):
    from unires_b200 import stats
    v = _volume(3)
    ref_in = v[v >= 0] if drop_negative else v          # unires/_core.py:118
    W_o, x_o, mn_o, mx_o = S.histogram(ref_in, 1024)
    W, x, mn, mx = stats.intensity_histogram(v.to(cuda), 1024, drop_negative)
    assert (mn, mx) == (mn_o, mx_o)
    assert torch.equal(W, W_o) and torch.equal(x, x_o)
    assert float(W.sum()) > 0.5 * v.numel() * (0.4 if drop_negative else 0.7)


@pytest.mark.gpu
def test_estimate_hyperpar_equals_oracle(cuda):
    """_estimate_hyperpar on the GPU (MR and CT observations) against the reference's loop over
    the oracle's estimate_noise (unires/_core.py:112-136)."""
    from oracle import unires_port as P
    from unires_b200 import _core, struct
    vols = [_volume(4, False), _volume(5, False) - 400.0]
    cts = [False, True]
    x = [[struct._input(dat=v.to(cuda), ct=ct)] for v, ct in zip(vols, cts)]
    _core._estimate_hyperpar(x, None)
    # the port's control flow is pinned bitwise against the reference's own _estimate_hyperpar
    # (tests/test_oracle_vs_reference.py)
    xo = P.estimate_hyperpar([[P.Observation(v.clone(), torch.eye(4), tau=1.0, ct=ct)]
                              for v, ct in zip(vols, cts)])
    for xc, oc in zip(x, xo):
        for k in ('sd', 'tau', 'mu'):
            got, want = float(getattr(xc[0], k)), float(getattr(oc[0], k))
            assert abs(got - want) <= 1e-6 * abs(want), (k, got, want)
            assert getattr(xc[0], k).device.type == 'cuda'


@pytest.mark.gpu
def test_estimate_noise_rejects_unsupported_modes_and_cpu_tensors(cuda):
    from unires_b200 import stats
    with pytest.raises(NotImplementedError):
        stats.estimate_noise(_volume(6).to(cuda), show_fit=True)
    with pytest.raises(RuntimeError):
        stats.estimate_noise(_volume(6))
    with pytest.raises(ValueError):
        stats.estimate_noise(torch.zeros(8, 8, 8, device=cuda))
