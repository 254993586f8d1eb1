"""HostPipeline (double-buffered host -> device -> host y-updates) against the in-place
y-update on resident data: identical results for every subject of the stream."""
import copy

import pytest
import torch

from tests import _util as U

pytestmark = pytest.mark.gpu


def test_host_pipeline_matches_resident_solves(cuda):
    from unires_b200 import _update
    _, recipe = U.load_golden('sr3_thick_xyz')
    sc = U.build(recipe, *U.port_namespaces())
    x, y, sett = U.to_device(sc, cuda)
    sett.cgs_tol = 0.0
    sett.cgs_max_iter = 6
    z, w = _update._admm_aux(y, sett)
    rho = float(sc.rho)
    dim, vx = _update._geometry(y)
    tmp = torch.zeros(dim, device=cuda)
    hx = [[o.dat.cpu().pin_memory() for o in xc] for xc in x]
    hy0 = [yc.dat.cpu().pin_memory() for yc in y]
    # reference: resident solve
    _update._solve_y(x, y, z, w, rho, tmp, sett, dim, vx)
    want = [yc.dat.clone() for yc in y]

    def clone_set():
        xb = [[copy.copy(o) for o in xc] for xc in x]
        for xc in xb:
            for o in xc:
                o.dat = torch.zeros_like(o.dat)
        yb = [copy.copy(yc) for yc in y]
        for yc in yb:
            yc.dat = torch.zeros_like(yc.dat)
        return xb, yb

    pipe = _update.HostPipeline([clone_set(), clone_set()], z, w, rho, sett)
    outs = [[torch.empty_like(t).pin_memory() for t in hy0] for _ in range(5)]
    for k in range(5):
        pipe.submit(hx, hy0, outs[k])
    pipe.drain()
    torch.cuda.synchronize()
    for k in range(5):
        for c in range(len(y)):
            assert torch.equal(outs[k][c].to(cuda), want[c]), (k, c)
