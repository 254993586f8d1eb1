"""Generate tests/golden/hyperpar_t1.npz: the intensity histogram of the reference's own
single-channel demo observation (demos/demo_single_channel.ipynb cells 3-4: BrainWeb T1,
voxels x4 along z, rect profiles, even/odd scaling 0.1, N(0, 75^2) noise), and the oracle's
estimate on it.  Needs /root/reference/data (this container only); the notebook drew its noise
from the CUDA RNG, so its logged "sd=48.64 | mu=406.5" is a statistical (soft) pin.

    python -m oracle.gen_golden_hyperpar
"""
import os

import numpy as np
import torch

from oracle import unires_port as P
from oracle.nitorch_shim.tools import img_statistics as S
from unires_b200 import io

REF_T1 = '/root/reference/data/t1_icbm_normal_1mm_pn0_rf0.nii.gz'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden',
                   'hyperpar_t1.npz')
NOTEBOOK_LOG = dict(sd=48.64, mu=406.5, tau=0.0004227)  # demo_single_channel.ipynb cell 5 output


def simulate(seed=0):
    arr, mat = io.read_nifti(REF_T1)
    y = torch.as_tensor(arr).float()
    mat_y = torch.as_tensor(mat).double()
    mat_x = mat_y @ torch.diag(torch.tensor((1., 1., 4., 1.), dtype=torch.float64))
    dim_y = tuple(y.shape)
    dim_x = (dim_y[0], dim_y[1], dim_y[2] // 4)
    po = P.proj_info(dim_y, mat_y, dim_x, mat_x, prof_ip=0, prof_tp=0, scl=0.1)
    x = P.proj_apply('A', y[None, None], po)[0, 0]
    g = torch.Generator().manual_seed(seed)
    return x + 75 * torch.randn(x.shape, generator=g)


def main():
    x = simulate()
    dat = x[x >= 0]  # unires/_core.py:118
    W, pos, mn, mx = S.histogram(dat, 1024)
    mp, mu, sd = S.fit_mixture(W, pos, 2, rician=True)
    noise, rest = S.noise_from_mixture(mp, mu, sd)
    np.savez_compressed(OUT, W=W.numpy().astype(np.int64), mn=mn, mx=mx, mp=mp.numpy(),
                        mean=mu.numpy(), sd=sd.numpy(), sd_noise=float(noise['sd']),
                        mu=float(abs(rest['mean'] - noise['mean'])),
                        notebook_sd=NOTEBOOK_LOG['sd'], notebook_mu=NOTEBOOK_LOG['mu'])
    print('sd %.3f (notebook %.2f)  mu %.2f (notebook %.1f)' %
          (float(noise['sd']), NOTEBOOK_LOG['sd'], float(abs(rest['mean'] - noise['mean'])),
           NOTEBOOK_LOG['mu']))


if __name__ == '__main__':
    main()
