"""Debug: A'A only (rho = 0) through chained lean passes vs the general path vs the oracle,
random and smooth inputs, error localised by distance from each face."""
import numpy as np
import torch
from oracle import unires_port as P
from tests import _util as U
from unires_b200 import _lib, _project, struct

cuda = torch.device('cuda:0')


def faces(d):
    out = []
    for a in range(3):
        m = d.amax(dim=[b for b in range(3) if b != a])
        out.append(' '.join('%.1e' % float(t) for t in m[:6]) + ' .. ' +
                   ' '.join('%.1e' % float(t) for t in m[-6:]))
    return out


for dim_y, zoom in [((24, 24, 24), (2., 2., 2.)), ((24, 28, 32), (2., 2., 1.)), ((24, 28, 32), (2., 1., 2.)),
                    ((24, 30, 32), (2., 3., 1.)), ((24, 28, 32), (1., 2., 2.))]:
    mat_y = torch.eye(4, dtype=torch.float64)
    mat_x = mat_y @ torch.diag(torch.tensor(list(zoom) + [1.0], dtype=torch.float64))
    dim_x = tuple(int(d // z) for d, z in zip(dim_y, zoom))
    po_o = P.proj_info(dim_y, mat_y, dim_x, mat_x, prof_ip=2, prof_tp=0)
    po_g = _project._proj_info(dim_y, mat_y, dim_x, mat_x, prof_ip=2, prof_tp=0, device=cuda)
    print('== zoom', zoom, 'ker', tuple(po_o.smo_ker.shape[-3:]), 'dim_yx', po_o.dim_yx)
    obs_o = P.Observation(torch.zeros(dim_x), mat_x, tau=1.0, po=po_o)
    obs_g = struct._input(tau=1.0, po=po_g)
    rec_o = P.Recon(torch.zeros(dim_y), mat_y, lam=0.25)
    rec_g = struct._output(dim=dim_y, mat=mat_y, lam=0.25)
    vx = torch.ones(3)
    g = torch.Generator().manual_seed(5)
    ii = torch.stack(torch.meshgrid(*[torch.linspace(0, 1, n) for n in dim_y], indexing='ij'))
    inputs = {'rand': torch.rand(dim_y, generator=g),
              'smooth': 100 + 40 * torch.sin(3 * ii[0] + 2 * ii[1]) + 30 * ii[2]}
    op = _project.LhsOperator([obs_g], rec_g, rho=0.0, vx_y=vx)
    for name, v in inputs.items():
        ref = P.proj('AtA', v.double(), [P.Observation(torch.zeros(dim_x, dtype=torch.float64), mat_x, tau=1.0, po=po_o)],
                     P.Recon(torch.zeros(dim_y, dtype=torch.float64), mat_y, lam=0.25), rho=0.0, vx_y=vx.double()) \
            if False else P.proj('AtA', v, [obs_o], rec_o, rho=0.0, vx_y=vx)
        for variant in (1, 0):
            _lib.check(_lib.lib.ur_tune(b'lhs_variant', variant))
            out = op(v.to(cuda)).cpu()
            path = _lib.lib.ur_last_lhs_path()
            d = (out - ref).abs() / ref.abs().max()
            print(' ', name, 'variant', variant, 'path', path, 'rel_l2 %.2e' % U.rel_l2(out, ref),
                  'max rel %.2e' % float(d.max()), 'sum ratio %.9f' % float(out.double().sum() / ref.double().sum()))
            if variant == 0:
                for a, s in enumerate(faces(d)):
                    print('     axis', a, s)
    _lib.check(_lib.lib.ur_tune(b'lhs_variant', 0))
