"""Oracle parity AT THE BENCHMARKED SIZES (BASELINE.json configs): one CG matvec
`sum tau A'A v + rho lam^2 D'D v` per channel of every config against
`oracle.unires_port.proj('AtA', ...)` (unires/_project.py:73-87) on the same seeded input, with
the lean kernel's lock-step work split forced on and off, plus three CG iterates of the
headline workload.  Only the operators are built (no forward simulation), so the CPU side of a
case is one oracle matvec: seconds at 256^3, about a minute at 512^3."""
import ctypes as C
import math

import pytest
import torch

from oracle import unires_port as P
from oracle.adapters import port_ops, port_structs
from oracle.nitorch_shim.core import optim as OO
from tests import _util as U

pytestmark = pytest.mark.gpu

REL_MATVEC = 1e-5


def _operators(name, c):
    """Oracle-side containers of channel c of CONFIGS[name]: operator, tau, lam -- no data."""
    from unires_b200 import synth
    cfg = synth.CONFIGS[name]
    n_ch = len(cfg['thick'])
    sett = port_structs.settings()
    denoise = bool(cfg.get('denoise'))
    sett.method = 'denoising' if denoise else 'super-resolution'
    sett.do_proj = not denoise
    dim_x, mat_x, dim_y, mat_y = synth.geometry(cfg, c)
    obs = port_structs._input()
    obs.tau = torch.tensor(1.0 / 25.0 ** 2, dtype=torch.float32)
    obs.dim, obs.mat = dim_x, mat_x
    if not denoise:
        rigid = None
        if cfg.get('rigid') is not None:
            rigid = synth.rigid_matrix(*cfg['rigid'][c])
        obs.po = port_ops._proj_info(dim_y, mat_y, dim_x, mat_x, rigid=rigid,
                                     prof_ip=sett.profile_ip, prof_tp=sett.profile_tp,
                                     gap=sett.gap, scl=0.0)
    obs.dat = torch.zeros(dim_x)
    rec = port_structs._output()
    rec.dim, rec.mat = dim_y, mat_y
    rec.lam = torch.tensor(4.0 * math.sqrt(1.0 / n_ch) / 400.0, dtype=torch.float32)
    rec.dat = torch.zeros(1)
    rho = torch.tensor(math.sqrt(float(obs.tau)) / float(rec.lam), dtype=torch.float32)
    sc = type('Sc', (), {})()
    sc.x, sc.y, sc.sett, sc.rho, sc.cfg = [[obs]], [rec], sett, rho, cfg
    return sc


def _gpu_lhs(sc, cuda):
    from unires_b200 import _project
    x, y, sett = U.to_device(sc, cuda)
    vx = [float(sc.cfg['vx_y'])] * 3
    return _project.LhsOperator(x[0], y[0], method=sett.method, do=sett.do_proj, rho=sc.rho,
                                vx_y=vx), vx


def _tune(name, value):
    from unires_b200 import _lib
    _lib.check(_lib.lib.ur_tune(name.encode(), int(value)))


CASES = [('sr3_256', 0), ('sr3_256', 1), ('sr3_256', 2), ('thickz2_256', 0), ('thickz2_384', 0),
         ('denoise_181', 0), ('crop3_256', 0), ('sr3_256_rigid', 0), ('sr3_256_rigid', 1),
         ('sr3_256_rigid', 2), ('iso2_512', 0)]


@pytest.mark.parametrize('name,c', CASES)
def test_matvec_vs_oracle_at_full_size(cuda, name, c):
    sc = _operators(name, c)
    dim = tuple(sc.y[0].dim)
    g = torch.Generator().manual_seed(11 + c)
    # smooth + rough content: a constant offset exposes spurious identity terms of D'D
    v = torch.rand(dim, generator=g) + 3.0
    vx = torch.ones(3) * float(sc.cfg['vx_y'])
    ref = P.proj('AtA', v, sc.x[0], sc.y[0], rho=sc.rho, vx_y=vx, method=sc.sett.method,
                 do=sc.sett.do_proj)
    op, _ = _gpu_lhs(sc, cuda)
    vg = v.to(cuda)
    from unires_b200 import _lib
    try:
        for lock in (1, 0):
            _tune('fast_lock', lock)
            dot = torch.zeros(1, dtype=torch.float64, device=cuda)
            out = op(vg, dot=dot)
            err = U.rel_l2(out, ref)
            assert err < REL_MATVEC, (name, c, 'fast_lock', lock, err, _lib.lib.ur_last_lhs_path())
            want = torch.sum(v.double() * ref.double()).item()
            assert abs(dot.item() - want) < 1e-5 * abs(want)
    finally:
        _tune('fast_lock', 1)


@pytest.mark.parametrize('name', ['sr3_256', 'sr3_256_rigid'])
def test_cg_iterates_vs_oracle_at_full_size(cuda, name):
    """Three CG iterates (fixed trip count) of channel 0 from a seeded right-hand side."""
    from unires_b200 import optim
    sc = _operators(name, 0)
    dim = tuple(sc.y[0].dim)
    g = torch.Generator().manual_seed(5)
    # right-hand side and start in the operator's scale: b ~ tau * image, x0 ~ image
    img = 200.0 + 100.0 * torch.rand(dim, generator=g)
    b = float(sc.x[0][0].tau) * img
    x0 = img + 20.0 * torch.rand(dim, generator=g)
    vx = torch.ones(3) * float(sc.cfg['vx_y'])
    lhs_o = lambda v: P.proj('AtA', v, sc.x[0], sc.y[0], rho=sc.rho, vx_y=vx,
                             method=sc.sett.method, do=sc.sett.do_proj)
    iterates = {}
    OO.cg(A=lhs_o, b=b, x=x0.clone(), max_iter=3, tolerance=0, stop='max_gain',
          record=lambda it, xi: iterates.__setitem__(it, xi.clone()))
    op, _ = _gpu_lhs(sc, cuda)
    for k in (1, 2, 3):
        xk = x0.to(cuda)
        optim.cg(A=op, b=b.to(cuda), x=xk, max_iter=k, tolerance=0, stop='max_gain')
        assert optim.cg.last.n_iter == k
        err = U.rel_l2(xk, iterates[k])
        assert err < U.REL_TOL, (name, k, err)
