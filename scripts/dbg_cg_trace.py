import sys, torch
sys.path.insert(0, '' + __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))) + '')
from tests import _util as U
from oracle import unires_port as P
from oracle.nitorch_shim.core import optim as OO
from unires_b200 import _project, optim
cuda = torch.device('cuda:0')
_, recipe = U.load_golden('sr3_thick_xyz')
sc = U.build(recipe, *U.port_namespaces())
x, y, sett = U.to_device(sc, cuda)
vx = torch.ones(3)
for c in range(3):
    kw = dict(method=sc.sett.method, do=sc.sett.do_proj)
    b = sc.x[c][0].tau * P.proj('At', sc.x[c][0].dat, sc.x[c], sc.y[c], n=0, **kw)
    lhs = lambda v: P.proj('AtA', v, sc.x[c], sc.y[c], rho=sc.rho, vx_y=vx, **kw)
    its = {}
    xo = sc.y[c].dat.clone()
    OO.cg(A=lhs, b=b, x=xo, max_iter=20, tolerance=1e-3, stop='max_gain', record=lambda it, xi: its.__setitem__(it, xi.clone()))
    o = OO.cg.last_obj
    lhs_g = _project.LhsOperator(x[c], y[c], method=sett.method, do=sett.do_proj, rho=sc.rho, vx_y=vx)
    xg = y[c].dat.clone()
    optim.cg(A=lhs_g, b=b.to(cuda), x=xg, max_iter=20, tolerance=1e-3, stop='max_gain')
    info = optim.cg.last
    print('channel', c, 'oracle n', OO.cg.last_n_iter, 'cuda n', info.n_iter)
    print(' oracle obj', ['%.8e' % v for v in o.tolist()])
    print(' cuda   obj', ['%.8e' % v for v in info.obj])
    for k in (1, 2, 5, 10):
        xk = y[c].dat.clone()
        optim.cg(A=lhs_g, b=b.to(cuda), x=xk, max_iter=k, tolerance=0, stop='max_gain')
        print('  iterate', k, 'rel', U.rel_l2(xk, its[k]))
    # matvec check
    v = torch.rand(sc.y[c].dim)
    print('  matvec rel', U.rel_l2(lhs_g(v.to(cuda)), lhs(v)))
