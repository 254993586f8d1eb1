"""Two devices in ONE process (the library keeps per-device state: opt-in to > 48 KB of
dynamic shared memory, occupancy cache, reduction scratch, SM count).  Skipped on a 1-GPU box;
run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi_device.py -m gpu`."""
import pytest
import torch

from tests import _util as U

pytestmark = pytest.mark.gpu


def test_second_device_in_the_same_process(cuda):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two CUDA devices')
    from unires_b200 import _update
    _, recipe = U.load_golden('sr3_thick_xyz')
    res = []
    for index in (0, 1, 0):
        dev = torch.device('cuda', index)
        with torch.cuda.device(dev):
            sc = U.build(recipe, *U.port_namespaces())
            x, y, sett = U.to_device(sc, dev)
            z, w = _update._admm_aux(y, sett)
            tmp = torch.zeros(y[0].dim, device=dev)
            obj = torch.zeros(2, 3, dtype=torch.float64, device=dev)
            for it in range(2):
                y, z, w, tmp, obj = _update._update_admm(x, y, z, w, sc.rho.to(dev), tmp, obj, it,
                                                         sett)
            res.append(([yc.dat.cpu() for yc in y], obj.cpu(),
                        [i.n_iter for i in _update._update_admm.last_cg]))
    for r in res[1:]:
        assert r[2] == res[0][2]
        assert torch.equal(r[1], res[0][1])
        for a, b in zip(r[0], res[0][0]):
            assert torch.equal(a, b)


def test_wrong_current_device_is_rejected(cuda):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two CUDA devices')
    from unires_b200 import spatial
    v = torch.rand(8, 8, 8, device='cuda:1')
    with torch.cuda.device(0):
        with pytest.raises(RuntimeError):
            spatial.im_gradient(v, vx=[1.0, 1.0, 1.0])
