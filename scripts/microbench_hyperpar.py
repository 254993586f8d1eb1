"""Time the hyper-parameter estimate of one 256^3 observation (range x2 + histogram kernels,
host-side mixture fit)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unires_b200 import stats, synth  # noqa: E402

dev = torch.device('cuda:0')
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
v = synth.phantom((n, n, n), 1, seed=0)[0].to(dev)
v = v + 25 * torch.randn_like(v)
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    W, x, mn, mx = stats.intensity_histogram(v, 1024, True)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    mp, mu, sd, it = stats.fit_mixture(W, x, 2, rician=mn >= 0)
    t2 = time.perf_counter()
    noise, rest = stats.noise_from_mixture(mp, mu, sd)
    print('%d^3: histogram (3 passes, %.0f MB each) %.3f ms = %.0f GB/s; fit %d EM iterations %.1f ms; '
          'sd %.2f (true 25) mu %.1f' % (n, v.numel() * 4 / 1e6, (t1 - t0) * 1e3,
                                        3 * v.numel() * 4 / (t1 - t0) / 1e9, it, (t2 - t1) * 1e3,
                                        float(noise['sd']), float(abs(rest['mean'] - noise['mean']))))
