"""Host-side logic of the product (no kernels launched): operator construction, slice
profiles, stop-rule parsing, lattice detection, error behaviour on CPU tensors."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import unires_port as P
from oracle.nitorch_shim.core import kernels as OK
from unires_b200 import _lib, _project, kernels, optim, struct, synth


@pytest.mark.parametrize('types,fwhm', [([-1, -1, 0], [1, 1, 2.0]), ([0, 2, 2], [2.0, 2, 2]),
                                        ([2, -1, -1], [3.0, 1, 1]), ([1, 0, 2], [2.0, 4, 1.5])])
def test_smooth_matches_oracle(types, fwhm):
    a = kernels.smooth(types, fwhm, sep=False, dtype=torch.float32)
    b = OK.smooth(types, fwhm, sep=False, dtype=torch.float32)
    assert a.shape == b.shape and torch.allclose(a, b, atol=1e-7, rtol=0)
    f = kernels.separable_factors(a)
    re = torch.tensor(f[0])[:, None, None] * torch.tensor(f[1])[None, :, None] * torch.tensor(f[2])
    assert torch.allclose(re.float(), a[0, 0], atol=1e-7)
    for t, fac in zip(types, f):
        if t == -1:
            assert fac == [1.0]


def test_separable_factors_rejects_dense_kernel():
    k = torch.rand(1, 1, 3, 3, 3)
    with pytest.raises(NotImplementedError):
        kernels.separable_factors(k)


@pytest.mark.parametrize('name', ['sr3_256', 'thickz2_256', 'iso2_512'])
def test_proj_info_matches_oracle(name):
    cfg = synth.scaled(synth.CONFIGS[name], (40, 36, 44))
    for c in range(len(cfg['thick'])):
        dim_x, mat_x, dim_y, mat_y = synth.geometry(cfg, c)
        rigid = synth.rigid_matrix((1.0, -2.0, 0.5), (0.02, 0.01, -0.03)) if c == 1 else None
        a = _project._proj_info(dim_y, mat_y, dim_x, mat_x, rigid=rigid, prof_ip=2, prof_tp=0,
                                scl=0.1)
        b = P.proj_info(dim_y, mat_y, dim_x, mat_x, rigid=rigid, prof_ip=2, prof_tp=0, scl=0.1)
        assert a.dim_yx == b.dim_yx and a.ratio == b.ratio and a.dim_x == b.dim_x
        assert int(a.dim_thick) == int(b.dim_thick)
        assert torch.allclose(a.mat_yx, b.mat_yx, atol=1e-12)
        assert torch.allclose(a.smo_ker, b.smo_ker, atol=1e-7)
        s = _project.proj_struct(a, 'super-resolution')
        vox = torch.linalg.solve(b.mat_y, b.rigid @ b.mat_yx).float()
        assert list(s.mat) == vox[:3].reshape(-1).tolist()
        assert _lib.lib.ur_proj_is_lattice(C.byref(s)) == (0 if c == 1 else 1)
        assert _project.proj_struct(a, 'super-resolution') is s  # cached
        a.scl = torch.tensor(0.2)
        assert _project.proj_struct(a, 'super-resolution') is not s  # invalidated


def test_stop_rule_parsing():
    assert optim.stop_rule('max_gain', 1e-3) == _lib.UR_STOP_ENERGY
    assert optim.stop_rule('E', 1e-3) == _lib.UR_STOP_RESIDUAL
    assert optim.stop_rule('residual', 1e-3) == _lib.UR_STOP_RESIDUAL
    assert optim.stop_rule('max_gain', 0) == _lib.UR_STOP_NONE


def test_get_gain():
    obj = torch.tensor([10.0, 4.0, 3.0], dtype=torch.float64)
    assert torch.isinf(optim.get_gain(obj[:1], 'decreasing'))
    assert optim.get_gain(obj[:2], 'decreasing').item() == 1.0
    with pytest.raises(ValueError):
        optim.get_gain(obj, 'sideways')


def test_no_cpu_fallback_and_reference_errors():
    po = _project._proj_info((8, 8, 8), torch.eye(4), (8, 8, 4),
                             torch.diag(torch.tensor([1, 1, 2, 1.0])))
    dat = torch.zeros(1, 1, 8, 8, 8)
    with pytest.raises(ValueError, match='Undefined operator'):
        _project._proj_apply('B', dat, po)
    with pytest.raises(ValueError, match='Undefined method'):
        _project._proj_apply('A', dat, po, method='sharpen')
    assert _project._proj_apply('none', dat, po) is dat
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        _project._proj_apply('A', dat, po)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        _project._DtD(torch.zeros(4, 4, 4), (1, 1, 1))
    from unires_b200 import _core, stats
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        stats.estimate_noise(torch.rand(4, 4, 4))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        _core._estimate_hyperpar([[struct._input(dat=torch.rand(4, 4, 4), ct=False)]])


def test_settings_defaults_match_reference_fields():
    s = struct.settings()
    assert (s.alpha, s.bound, s.cgs_max_iter, s.cgs_tol, s.diff) == (1.0, 'zero', 20, 1e-3, 'forward')
    assert (s.profile_ip, s.profile_tp, s.gap, s.reg_scl, s.tolerance) == (2, 0, 0.0, 4.0, 1e-4)
    po = struct._proj_op()
    for f in ('dim_x', 'mat_x', 'vx_x', 'dim_y', 'mat_y', 'vx_y', 'dim_yx', 'mat_yx', 'ratio',
              'smo_ker', 'rigid', 'scl', 'dim_thick', 'D_x', 'D_y'):
        assert hasattr(po, f)


def _colour_matrix(rot, scale=(1.0, 1.0, 1.0), shift=(3.3, -1.7, 0.45)):
    from unires_b200 import synth
    m = synth.rigid_matrix(shift, rot).numpy()
    m[:3, :3] = m[:3, :3] @ np.diag(scale)
    return m[:3, :4].astype(np.float32)


@pytest.mark.parametrize('rot,scale,want', [
    ((0.05, -0.1, 0.1), (1, 1, 1), 2),        # notebook-sized rotation: parity of i + j + k
    ((1e-5, 0.0, -2e-5), (1, 1, 1), 8),       # near identity: (1, 1, 0) maps within the float32
                                              # slack of a unit cell, so parities of i, j, k
    ((0.0, 0.0, 0.7853982), (1, 1, 1), 8),    # 45 degrees about z: (1, 0, 1) maps onto a unit cell
    ((0.5, -0.4, 0.7), (1, 1, 1), 8),
    ((0.05, 0.02, 0.0), (0.45, 1, 1), 0),     # strongly anisotropic map: per-voxel gather instead
])
def test_rot_cell_colouring_never_shares_a_cell(rot, scale, want):
    """The cell-coefficient adjoint of rotated operators (csrc/rot.cu) scatters intermediate
    voxels of ONE colour into shared-memory cells without atomics: ur_rot_cell_colours must only
    return a colouring under which two voxels of one colour never fall into the same unit cell
    of the recon grid.  Brute force over a 24^3 block of the intermediate lattice."""
    import ctypes as C
    from unires_b200 import _lib
    mat = _colour_matrix(rot, scale)
    ncol = _lib.lib.ur_rot_cell_colours(mat.ctypes.data_as(C.POINTER(C.c_float)))
    assert ncol == want
    if ncol == 0:
        return
    n = 24
    i, j, k = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing='ij')
    p = np.stack([i, j, k], -1).reshape(-1, 3).astype(np.float32)
    # the kernel's float32 expression: fma(m2, k, fma(m1, j, m0 * i)) + t
    c = (mat[:, 0] * p[:, :1] + mat[:, 1] * p[:, 1:2]) + mat[:, 2] * p[:, 2:3] + mat[:, 3]
    cell = np.floor(c).astype(np.int64)
    pi = p.astype(np.int64)
    if ncol == 2:
        colour = (pi.sum(1)) & 1
    else:
        colour = ((pi[:, 0] & 1) << 2) | ((pi[:, 1] & 1) << 1) | (pi[:, 2] & 1)
    key = ((cell[:, 0] + 64) * 4096 + (cell[:, 1] + 64)) * 4096 + (cell[:, 2] + 64)
    key = key * 8 + colour
    assert len(np.unique(key)) == len(key), 'two voxels of one colour share a cell'
    # and the colouring is needed: without it some cells do hold several voxels
    if rot[2] > 0.7:
        plain = key // 8
        assert len(np.unique(plain)) < len(plain)
