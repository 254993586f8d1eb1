"""Apply the CG left-hand side of channel 0 of a workload a few times (for ncu launch lists)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unires_b200 import _project, struct, synth  # noqa: E402

dev = torch.device('cuda:0')
workload = sys.argv[1] if len(sys.argv) > 1 else 'iso2_512'
cfg = dict(synth.CONFIGS[workload])
cfg['thick'] = cfg['thick'][:1]
sc = synth.make_scenario(cfg, _project, struct, device=dev, seed=0)
dim = tuple(sc.y[0].dim)
vx = [float(sc.cfg['vx_y'])] * 3
op = _project.LhsOperator(sc.x[0], sc.y[0], method=sc.sett.method, do=sc.sett.do_proj, rho=sc.rho, vx_y=vx)
v = torch.rand(dim, device=dev)
torch.cuda.synchronize()
print('MARK begin')
for _ in range(3):
    out = op(v)
torch.cuda.synchronize()
print('done', float(out.sum()))
