"""Outer loop `fit` (unires/run.py:24-207) on the GPU against the reference's fit fixture."""
import json

import numpy as np
import pytest
import torch

from oracle import gen_golden
from tests import _util as U

pytestmark = pytest.mark.gpu


def test_fit_matches_reference_fixture(cuda):
    from unires_b200 import run
    g = np.load(U.GOLDEN_DIR + '/fit_sr2.npz', allow_pickle=False)
    recipe = json.loads(str(g['recipe']))
    sc = gen_golden.prepare_fit(U.build(recipe, *U.port_namespaces()))
    x, y, sett = U.to_device(sc, cuda)
    for k in ('max_iter', 'tolerance', 'reg_scl', 'sched_num', 'clean_fov', 'scaling',
              'unified_rigid', 'rigid_mod'):
        setattr(sett, k, getattr(sc.sett, k))
    for c in range(len(y)):
        y[c].lam0 = torch.tensor(float(sc.y[c].lam0), device=cuda)
        for n, o in enumerate(x[c]):
            o.dim = tuple(sc.x[c][n].dat.shape)
            o.tau = torch.tensor(float(sc.x[c][n].tau), device=cuda)
    dat_y, mat, pth, R, label, pth_label = run.fit(x, y, sett)
    last = run.fit.last
    assert last['reg_scl'].tolist() == [32.0, 16.0, 8.0, 4.0]
    assert last['n_iter'] == int(g['n_iter'])
    obj = last['obj'].cpu().numpy()
    # The objective trajectory through the schedule changes.  The first rows agree to float32
    # summation noise.  From ADMM iteration ~8 on the warm-started CG solves stop on the
    # reference's |gain| < 1e-3 energy test at a point where the energy decrease per CG iteration
    # is of the order of the float32 noise of the energy itself: trip counts then vary by several
    # iterations with ANY change of rounding (measured on B200, gpurun_out/dbg_fit.log: two
    # roundings of the same diagonal constant give [14, 5] vs [14, 20] trips at iteration 43), the
    # objective moves by O(1e-4) relative, and a coarse-to-fine switch (six consecutive
    # |gain| < 1e-3 rows) can land a few iterations earlier or later.  So: rows 0-5 tight, every
    # later row within 2e-3 of the reference's trajectory allowing a shift of <= 3 iterations
    # (measured 8e-4 on the steep rows right after a shifted switch, 1.4e-4 without a shift),
    # the converged objective within 1e-3, the final image within 5e-3 (measured 6e-5 when the
    # switches land on the reference's iterations, 2.4e-3 when the last lands two later).
    ref = g['obj']
    assert np.allclose(obj[:6], ref[:6], rtol=1e-5)
    for i in range(len(obj)):
        win = ref[max(0, i - 3):i + 4, 0]
        assert np.min(np.abs(obj[i, 0] - win) / np.abs(win)) < 2e-3, (i, obj[i, 0], win)
    assert abs(obj[-1, 0] - ref[-1, 0]) < 1e-3 * abs(ref[-1, 0])
    assert tuple(dat_y.shape) == tuple(g['dat_y'].shape)
    assert U.rel_l2(dat_y, g['dat_y']) < 5e-3
    assert R.shape == (2, 4, 4) and pth == [] and label is None


def test_fit_with_scaling_matches_reference_fixture(cuda):
    """sett.scaling: the even/odd slice scaling exp(+-0.1) of the simulated data is recovered
    from 0 by the Gauss-Newton update interleaved with the ADMM iterations (unires/run.py:
    115-122); trajectory against the reference's own fit."""
    from unires_b200 import run
    g = np.load(U.GOLDEN_DIR + '/fit_scaling.npz', allow_pickle=False)
    recipe = json.loads(str(g['recipe']))
    sc = gen_golden.prepare_fit(U.build(recipe, *U.port_namespaces()), scaling=True)
    x, y, sett = U.to_device(sc, cuda)
    for k in ('max_iter', 'tolerance', 'reg_scl', 'sched_num', 'clean_fov', 'scaling',
              'unified_rigid', 'rigid_mod'):
        setattr(sett, k, getattr(sc.sett, k))
    for c in range(len(y)):
        y[c].lam0 = torch.tensor(float(sc.y[c].lam0), device=cuda)
        for n, o in enumerate(x[c]):
            o.dim = tuple(sc.x[c][n].dat.shape)
            o.tau = torch.tensor(float(sc.x[c][n].tau), device=cuda)
    dat_y = run.fit(x, y, sett)[0]
    last = run.fit.last
    assert last['n_iter'] == int(g['n_iter'])
    scl = [float(o.po.scl) for xc in x for o in xc]
    assert np.allclose(scl, g['scl'], rtol=2e-3), (scl, g['scl'].tolist())
    assert np.allclose(last['obj'].cpu().numpy(), g['obj'], rtol=1e-3)
    assert U.rel_l2(dat_y, g['dat_y']) < 1e-3


def test_fit_with_rigid_matches_reference_fixture(cuda):
    """sett.unified_rigid: observations whose operators start mis-registered by exp(+-q0) are
    re-aligned by the Gauss-Newton update interleaved with the ADMM iterations
    (unires/run.py:127-135); trajectory against the reference's own fit."""
    from unires_b200 import run
    g = np.load(U.GOLDEN_DIR + '/fit_rigid.npz', allow_pickle=False)
    recipe = json.loads(str(g['recipe']))
    sc = gen_golden.prepare_fit(U.build(recipe, *U.port_namespaces()), rigid=True)
    x, y, sett = U.to_device(sc, cuda)
    for k in ('max_iter', 'tolerance', 'reg_scl', 'sched_num', 'clean_fov', 'scaling',
              'unified_rigid', 'rigid_mod', 'rigid_samp', 'rigid_basis'):
        setattr(sett, k, getattr(sc.sett, k))
    for c in range(len(y)):
        y[c].lam0 = torch.tensor(float(sc.y[c].lam0), device=cuda)
        for n, o in enumerate(x[c]):
            o.dim = tuple(sc.x[c][n].dat.shape)
            o.tau = torch.tensor(float(sc.x[c][n].tau), device=cuda)
            o.rigid_q = sc.x[c][n].rigid_q.clone()
    dat_y, _, _, R, _, _ = run.fit(x, y, sett)
    last = run.fit.last
    assert last['n_iter'] == int(g['n_iter'])
    q = np.array([o.rigid_q.cpu().tolist() for xc in x for o in xc])
    # 29 interleaved Gauss-Newton steps with discrete decisions (line search, borderline CG
    # stops) amplify float32 rounding: the two deterministic adjoint kernels agree to 7e-8 on
    # every operator of this scenario (scripts/dbg_cell_ops.py), yet their trajectories end
    # 0.012 voxels apart on the smallest translation (the spread once measured between two runs
    # of the atomic scatter): per-voxel gather q within 2.5e-4 / image 1.1e-2 of the reference,
    # per-cell adjoint 1.2e-2 / 5.6e-2, objective within 7e-4 in both
    # (scripts/dbg_fit_rigid.py).  Hence a trajectory-level check; single steps agree to 2e-5
    # (test_update_rigid_vs_golden) and single operators to 1e-5 (test_gpu_ops.py).
    assert np.allclose(q, g['q'], atol=0.03), (q, g['q'])
    assert np.allclose(R.cpu().numpy(), g['R'], atol=0.03)
    assert np.allclose(last['obj'].cpu().numpy(), g['obj'], rtol=5e-3)
    assert U.rel_l2(dat_y, g['dat_y']) < 1e-1


def test_fit_unified_rigid_needs_a_basis(cuda):
    from unires_b200 import run, struct
    s = struct.settings()
    s.unified_rigid = True
    with pytest.raises(ValueError):
        run.fit([], [], s)
