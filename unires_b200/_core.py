"""Pre-processing steps of unires/_core.py that feed the ADMM/CG path.

Built: `_estimate_hyperpar` (noise precision tau and mean foreground intensity mu of every
observation, unires/_core.py:96-142).  `_init_y_dat`, `_read_image` and `_write_image` live in
io.py.  Not built (they need nitorch.tools' co-registration and atlas code, out of scope):
`_fix_affine`, `_format_y`, `_crop_y`, `_init_reg` (unires/_core.py:145-368).
"""
import torch

from .stats import estimate_noise


def _estimate_hyperpar(x, sett=None):
    """Sets x[c][n].sd, .tau = 1 / sd^2 and .mu = |mean foreground - mean noise class| from a
    two-class mixture fit to the histogram of every observation (non-negative voxels only
    unless the observation is CT).  Returns x, like unires/_core.py:96-142."""
    for xc in x:
        for obs in xc:
            prm_noise, prm_not_noise = estimate_noise(obs.dat, num_class=2,
                                                      drop_negative=not getattr(obs, 'ct', False))
            dev = obs.dat.device  # the reference's estimates live on the data's device
            sd_bg = prm_noise['sd'].float().to(dev)
            obs.sd = sd_bg
            obs.tau = 1 / sd_bg ** 2
            obs.mu = torch.abs(prm_not_noise['mean'].float() - prm_noise['mean'].float()).to(dev)
    return x
