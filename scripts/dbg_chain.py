"""Debug: chained lean passes vs general path vs oracle on the iso2_1ch fixture."""
import numpy as np
import torch
from oracle import unires_port as P
from oracle.nitorch_shim.core import optim as OO
from tests import _util as U
from tests.test_gpu_solver import _channel_problem
from unires_b200 import _lib, optim

cuda = torch.device('cuda:0')
_, recipe = U.load_golden('iso2_1ch')
sc = U.build(recipe, *U.port_namespaces())
b, lhs_o, lhs_g, x0 = _channel_problem(sc, 0, cuda)
g = torch.Generator().manual_seed(3)
v = torch.rand(sc.y[0].dim, generator=g)
ref = lhs_o(v)
ref64 = None
for variant in (1, 0):
    _lib.check(_lib.lib.ur_tune(b'lhs_variant', variant))
    out = lhs_g(v.to(cuda)).cpu()
    path = _lib.lib.ur_last_lhs_path()
    d = (out - ref).abs()
    idx = np.unravel_index(int(d.argmax()), d.shape)
    print('variant', variant, 'path', path, 'rel_l2', U.rel_l2(out, ref), 'max abs', d.max().item(),
          'at', idx, 'ref max', ref.abs().max().item())
    # error by shell distance from the boundary
    n = d.shape[0]
    ii = torch.arange(n)
    dist = torch.minimum(ii, n - 1 - ii)
    D = torch.minimum(torch.minimum(dist[:, None, None], dist[None, :, None]), dist[None, None, :])
    print('  max err by boundary distance', [float(d[D == k].max()) for k in range(8)])
    for stop in ('max_gain', 'residual'):
        xo = sc.y[0].dat.clone()
        OO.cg(A=lhs_o, b=b, x=xo, max_iter=20, tolerance=1e-3, stop=stop)
        n_ref, obj_ref = OO.cg.last_n_iter, OO.cg.last_obj.numpy()
        xg = x0.clone()
        optim.cg(A=lhs_g, b=b.to(cuda), x=xg, max_iter=20, tolerance=1e-3, stop=stop)
        info = optim.cg.last
        o = np.asarray(info.obj)
        m = min(len(o), len(obj_ref))
        print('  ', stop, 'n', info.n_iter, n_ref, 'x rel', U.rel_l2(xg, xo))
        print('   ratio to tol', np.round(np.abs(o[:m] - obj_ref[:m]) / (1e-4 * np.abs(obj_ref[:m]) + 1e-6 * abs(obj_ref[0])), 2))
        print('   gpu', o[-4:], 'ref', obj_ref[-4:])
_lib.check(_lib.lib.ur_tune(b'lhs_variant', 0))
