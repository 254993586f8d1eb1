"""Generate tests/golden/*.npz by running the REFERENCE'S OWN control-flow files
(/root/reference/unires/{_project,_update,struct}.py, imported by path,
unmodified) on top of oracle/nitorch_shim.  TEST INFRASTRUCTURE.

    python -m oracle.gen_golden            # only works where /root/reference exists

Each fixture stores the scenario recipe (so the inputs can be regenerated from
the seed with unires_b200.synth), sha256 digests of the regenerated inputs, and
the reference outputs: operator applications on seeded random volumes and two
ADMM iterations (y, z/w digests, jtv, objective, CG trip counts).
PARITY UNPINNED at the nitorch boundary (see oracle/__init__.py); these
fixtures pin the reference's control flow over the restated primitives.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from unires_b200 import synth  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')
SAMPLE_STRIDE = 11

# name -> recipe (everything make_scenario needs)
RECIPES = {
    'denoise_1ch': dict(base='denoise_181', dim_y=(20, 24, 20), n_channels=1, sd=25.0,
                        scl=0.0, rigid=None, admm_iters=2),
    'sr3_thick_xyz': dict(base='sr3_256', dim_y=(28, 32, 24), n_channels=3, sd=25.0,
                          scl=0.0, rigid=None, admm_iters=2),
    'thickz2_scl': dict(base='thickz2_256', dim_y=(24, 20, 28), n_channels=2, sd=15.0,
                        scl=0.1, rigid=None, admm_iters=2),
    'sr2_rigid': dict(base='sr3_256', dim_y=(24, 28, 26), n_channels=2, sd=25.0, scl=0.05,
                      rigid=[((1.3, -0.7, 0.4), (0.03, -0.02, 0.05)),
                             ((-0.6, 0.9, 1.1), (-0.04, 0.03, 0.02))], admm_iters=2),
    'iso2_1ch': dict(base='iso2_512', dim_y=(24, 24, 24), n_channels=1, sd=20.0, scl=0.0,
                     rigid=None, admm_iters=2),
}


def digest(t):
    a = np.ascontiguousarray(t.detach().cpu().numpy())
    return hashlib.sha256(a.tobytes()).hexdigest()


def build(recipe, ops, structs, device='cpu'):
    cfg = synth.scaled(synth.CONFIGS[recipe['base']], recipe['dim_y'], recipe['n_channels'])
    rigid = None
    if recipe['rigid'] is not None:
        rigid = [synth.rigid_matrix(t, r) for t, r in recipe['rigid']]
    return synth.make_scenario(cfg, ops, structs, device=device, seed=0, sd=recipe['sd'],
                               scl=recipe['scl'], rigid=rigid)


def probe_inputs(sc, c):
    """Seeded random volumes for the operator-level golden values."""
    g = torch.Generator().manual_seed(100 + c)
    vy = torch.rand(tuple(sc.y[c].dim), generator=g)
    vx = torch.rand(tuple(sc.x[c][0].dat.shape), generator=g)
    return vy, vx


def main():
    from oracle.adapters import reference_namespaces
    from oracle.load_reference import load_reference
    from oracle.nitorch_shim.core import optim as shim_optim
    ref = load_reference()
    ops, structs = reference_namespaces()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name, recipe in RECIPES.items():
        torch.manual_seed(0)
        sc = build(recipe, ops, structs)
        out = {'recipe': json.dumps(recipe)}
        C = len(sc.x)
        for c in range(C):
            out['in_x%d_sha' % c] = digest(sc.x[c][0].dat)
            out['in_y%d_sha' % c] = digest(sc.y[c].dat)
        # operator-level goldens
        if sc.sett.do_proj:
            for c in range(C):
                vy, vx = probe_inputs(sc, c)
                po = sc.x[c][0].po
                kw = dict(method=sc.sett.method)
                out['A%d' % c] = ref._project._proj_apply('A', vy[None, None], po, **kw)[0, 0].numpy()
                out['At%d' % c] = ref._project._proj_apply('At', vx[None, None], po, **kw)[0, 0].numpy()
                out['AtA%d' % c] = ref._project._proj_apply('AtA', vy[None, None], po, **kw)[0, 0].numpy()
        vx_y = ref._update.voxel_size(sc.y[0].mat).float()
        for c in range(C):
            vy, _ = probe_inputs(sc, c)
            out['lhs%d' % c] = ref._project._proj(
                'AtA', vy, sc.x[c], sc.y[c], method=sc.sett.method, do=sc.sett.do_proj,
                rho=sc.rho, vx_y=vx_y).numpy()
        # ADMM iterations with the reference's _update_admm
        z, w = ref._update._admm_aux(sc.y, sc.sett)
        tmp = torch.zeros(tuple(sc.y[0].dim))
        n_it = recipe['admm_iters']
        obj = torch.zeros(n_it, 3, dtype=torch.float64)
        cg_iters = []
        # count CG trips by wrapping the shim's cg (the reference ignores its return value)
        orig_cg = ref._update.cg

        def counting_cg(*a, **k):
            r = orig_cg(*a, **k)
            cg_iters.append(shim_optim.cg.last_n_iter)
            return r

        ref._update.cg = counting_cg
        try:
            for it in range(n_it):
                y, z, w, tmp, obj = ref._update._update_admm(sc.x, sc.y, z, w, sc.rho, tmp, obj,
                                                             it, sc.sett)
                for c in range(C):
                    out['y%d_it%d' % (c, it)] = sc.y[c].dat.numpy().copy()
                out['jtv_it%d' % it] = tmp.numpy().copy()
                # z / w are 3C volumes each: keep their norms and a strided sample
                for nm, t in (('z', z), ('w', w)):
                    out['%s_norm_it%d' % (nm, it)] = np.float64(t.double().norm().item())
                    out['%s_sample_it%d' % (nm, it)] = t.flatten()[::SAMPLE_STRIDE].numpy().copy()
        finally:
            ref._update.cg = orig_cg
        out['obj'] = obj.numpy()
        out['cg_iters'] = np.array(cg_iters, dtype=np.int32).reshape(n_it, C)
        out['rho'] = np.float32(float(sc.rho))
        path = os.path.join(GOLDEN_DIR, name + '.npz')
        np.savez_compressed(path, **out)
        print(name, 'cg_iters', out['cg_iters'].tolist(), 'obj', obj[:, 0].tolist(),
              '%.0f KB' % (os.path.getsize(path) / 1024))


if __name__ == '__main__' and not {'--fit', '--scaling', '--rigid', '--mid'} & set(sys.argv):
    main()
    main_fit()
    main_fit(scaling=True)
    main_fit(rigid=True)
    main_scaling()
    main_rigid()


# ---------------------------------------------------------------------------
# outer loop (unires/run.py::fit) fixture
# ---------------------------------------------------------------------------
FIT_RECIPE = dict(base='sr3_256', dim_y=(16, 18, 16), n_channels=2, sd=25.0, scl=0.0, rigid=None,
                  max_iter=70)


FIT_SCALING_RECIPE = dict(base='thickz2_256', dim_y=(16, 14, 20), n_channels=2, sd=15.0, scl=0.1,
                          rigid=None, max_iter=40)


FIT_RIGID_RECIPE = dict(base='sr3_256', dim_y=(20, 24, 22), n_channels=2, sd=10.0, scl=0.0,
                        rigid=None, max_iter=30, q0=[0.6, -0.4, 0.3, 0.01, -0.015, 0.02])


def prepare_fit(sc, reference=False, scaling=False, rigid=False):
    """Settings of the fit fixture (and the fields only the reference's fit touches)."""
    s = sc.sett
    recipe = FIT_SCALING_RECIPE if scaling else (FIT_RIGID_RECIPE if rigid else FIT_RECIPE)
    s.max_iter, s.tolerance, s.reg_scl, s.sched_num = recipe['max_iter'], 1e-4, 4.0, 3
    s.clean_fov, s.scaling, s.unified_rigid, s.rigid_mod = True, scaling, False, 1
    s.rigid_samp = 1
    if scaling:  # the data carry exp(+-0.1); the fit starts from 0 and has to find it
        prepare_scaling(sc, 0.0)
    if rigid:  # the operators start mis-registered; the fit has to re-align them
        from oracle.unires_port import expm
        s.unified_rigid = True
        prepare_rigid(sc, FIT_RIGID_RECIPE['q0'], expm)
    if reference:
        from oracle.nitorch_shim.spatial import affine_basis
        s.rigid_basis = affine_basis(group='SE', dtype=torch.float64)
        s.write_out, s.write_jtv, s.show_jtv, s.plot_conv, s.mat, s.do_print = \
            False, False, False, False, None, 0
        for xc in sc.x:
            for o in xc:
                if not rigid:
                    o.rigid_q = torch.zeros(6, dtype=torch.float64)
                o.direc = o.nam = None
    return sc


def main_fit(scaling=False, rigid=False):
    from oracle.adapters import reference_namespaces
    from oracle.load_reference import load_reference
    ref = load_reference()
    ops, structs = reference_namespaces()
    recipe = FIT_SCALING_RECIPE if scaling else (FIT_RIGID_RECIPE if rigid else FIT_RECIPE)
    sc = prepare_fit(build(recipe, ops, structs), reference=True, scaling=scaling, rigid=rigid)
    obj_rows = []
    orig = ref.run._update_admm

    def spy(x, y, z, w, rho, tmp, obj, n_iter, sett):
        out = orig(x, y, z, w, rho, tmp, obj, n_iter, sett)
        obj_rows.append(out[4][n_iter].clone())
        return out

    ref.run._update_admm = spy
    try:
        fit_out = ref.run.fit(sc.x, sc.y, sc.sett)
        dat_y = fit_out[0]
    finally:
        ref.run._update_admm = orig
    out = {'recipe': json.dumps(recipe), 'dat_y': dat_y.numpy(),
           'obj': torch.stack(obj_rows).numpy(), 'n_iter': np.int32(len(obj_rows)),
           'scl': np.array([float(o.po.scl) for xc in sc.x for o in xc]),
           'q': np.array([o.rigid_q.tolist() for xc in sc.x for o in xc]),
           'R': fit_out[3].numpy()}
    path = os.path.join(GOLDEN_DIR, 'fit_scaling.npz' if scaling else
                        ('fit_rigid.npz' if rigid else 'fit_sr2.npz'))
    np.savez_compressed(path, **out)
    print(os.path.basename(path), 'scl', out['scl'].tolist(), 'n_iter', len(obj_rows), 'obj', obj_rows[0][0].item(), '->', obj_rows[-1][0].item(),
          '%.0f KB' % (os.path.getsize(path) / 1024))


# ---------------------------------------------------------------------------
# even/odd slice-scaling update (unires/_update.py::_update_scaling) fixture
# ---------------------------------------------------------------------------
SCALING_CASES = {
    # recipe name -> initial po.scl handed to the update (the data were simulated with recipe scl)
    'thickz2_scl': 0.0,
    'sr3_thick_xyz': 0.03,
}
SCALING_STEPS = 3


def prepare_scaling(sc, scl0):
    for xc in sc.x:
        for o in xc:
            o.po.scl = torch.tensor(scl0, dtype=torch.float32)
    return sc


def main_scaling():
    from oracle.adapters import reference_namespaces
    from oracle.load_reference import load_reference
    ref = load_reference()
    ops, structs = reference_namespaces()
    out = {}
    for name, scl0 in SCALING_CASES.items():
        torch.manual_seed(0)
        sc = prepare_fit(build(RECIPES[name], ops, structs), reference=True)
        sc.sett.max_iter = 512
        prepare_scaling(sc, scl0)
        scl, sll = [], []
        for _ in range(SCALING_STEPS):
            _, s = ref._update._update_scaling(sc.x, sc.y, sc.sett, max_niter_gn=1,
                                               num_linesearch=6, verbose=0)
            scl.append([float(o.po.scl) for xc in sc.x for o in xc])
            sll.append(float(s))
        out[name + '_scl'] = np.array(scl, dtype=np.float64)
        out[name + '_sll'] = np.array(sll, dtype=np.float64)
        print('scaling', name, scl, sll)
    np.savez_compressed(os.path.join(GOLDEN_DIR, 'scaling_update.npz'), **out)


# ---------------------------------------------------------------------------
# rigid Gauss-Newton update (unires/_update.py::_update_rigid) fixture
# ---------------------------------------------------------------------------
RIGID_CASES = {
    # name -> (recipe, samp, initial q of channel 0; channel c gets (-1)^c q)
    'sr2_lattice': (dict(base='sr3_256', dim_y=(24, 28, 26), n_channels=2, sd=10.0, scl=0.0,
                         rigid=None, admm_iters=2), 1, [0.6, -0.4, 0.3, 0.01, -0.015, 0.02]),
    'thickz2_samp2': (dict(base='thickz2_256', dim_y=(26, 22, 28), n_channels=2, sd=10.0, scl=0.05,
                           rigid=None, admm_iters=2), 2, [-0.5, 0.3, 0.4, -0.02, 0.01, 0.015]),
}
RIGID_STEPS = 3


def prepare_rigid(sc, q0, expm):
    """Mis-register every observation's operator by exp(+-q0) (the data are aligned)."""
    from oracle.nitorch_shim.spatial import affine_basis
    sc.sett.rigid_basis = affine_basis(group='SE', dtype=torch.float64)
    for c, xc in enumerate(sc.x):
        for o in xc:
            o.rigid_q = torch.tensor(q0, dtype=torch.float64) * (1 if c % 2 == 0 else -1)
            o.po.rigid = expm(o.rigid_q, sc.sett.rigid_basis)
    return sc


def main_rigid():
    from oracle.adapters import reference_namespaces
    from oracle.load_reference import load_reference
    ref = load_reference()
    ops, structs = reference_namespaces()
    out = {}
    for name, (recipe, samp, q0) in RIGID_CASES.items():
        torch.manual_seed(0)
        sc = prepare_fit(build(recipe, ops, structs), reference=True)
        prepare_rigid(sc, q0, ref._update._expm)
        qs, sll = [], []
        for k in range(RIGID_STEPS):
            _, s = ref._update._update_rigid(sc.x, sc.y, sc.sett, mean_correct=(k == RIGID_STEPS - 1),
                                             max_niter_gn=1, num_linesearch=6, verbose=0, samp=samp)
            qs.append([o.rigid_q.tolist() for xc in sc.x for o in xc])
            sll.append(float(s))
        out[name + '_q'] = np.array(qs, dtype=np.float64)
        out[name + '_sll'] = np.array(sll, dtype=np.float64)
        out[name + '_rigid'] = np.stack([o.po.rigid.numpy() for xc in sc.x for o in xc])
        print('rigid', name, sll, qs[-1][0])
    np.savez_compressed(os.path.join(GOLDEN_DIR, 'rigid_update.npz'), **out)


if __name__ == '__main__' and '--fit' in sys.argv:
    main_fit()
    main_fit(scaling=True)
    main_fit(rigid=True)
if __name__ == '__main__' and '--rigid' in sys.argv:
    main_rigid()
if __name__ == '__main__' and '--scaling' in sys.argv:
    main_scaling()


# ---------------------------------------------------------------------------
# mid-size fixtures: grids that span several row tiles, z tiles and column segments of the
# streaming kernels (the small fixtures above fit one tile).  Volumes are stored as a strided
# sample plus their float64 norm, so a fixture stays around 1 MB.
# ---------------------------------------------------------------------------
MID_STRIDE = 29
MID_RECIPES = {
    'mid_sr3': dict(base='sr3_256', dim_y=(40, 96, 260), n_channels=3, sd=25.0, scl=0.0,
                    rigid=None, admm_iters=1),
    'mid_sr3_rigid': dict(base='sr3_256', dim_y=(40, 96, 260), n_channels=3, sd=25.0, scl=0.04,
                          rigid=[((2.1, -1.7, 3.3), (0.08, -0.05, 0.1)),
                                 ((-3.2, 2.6, -1.9), (-0.1, 0.07, 0.04)),
                                 ((2.4, 1.8, -4.4), (0.05, 0.1, -0.09))], admm_iters=1),
}


def mid_sample(t):
    t = torch.as_tensor(t)
    return t.flatten()[::MID_STRIDE].numpy().copy(), np.float64(t.double().norm().item())


def main_mid():
    from oracle.adapters import reference_namespaces
    from oracle.load_reference import load_reference
    from oracle.nitorch_shim.core import optim as shim_optim
    ref = load_reference()
    ops, structs = reference_namespaces()
    for name, recipe in MID_RECIPES.items():
        torch.manual_seed(0)
        sc = build(recipe, ops, structs)
        out = {'recipe': json.dumps(recipe)}
        C = len(sc.x)

        def put(key, t):
            out[key + '_s'], out[key + '_n'] = mid_sample(t)

        for c in range(C):
            out['in_x%d_sha' % c] = digest(sc.x[c][0].dat)
            out['in_y%d_sha' % c] = digest(sc.y[c].dat)
            vy, vx = probe_inputs(sc, c)
            po = sc.x[c][0].po
            kw = dict(method=sc.sett.method)
            put('A%d' % c, ref._project._proj_apply('A', vy[None, None], po, **kw)[0, 0])
            put('At%d' % c, ref._project._proj_apply('At', vx[None, None], po, **kw)[0, 0])
            vx_y = ref._update.voxel_size(sc.y[0].mat).float()
            put('lhs%d' % c, ref._project._proj('AtA', vy, sc.x[c], sc.y[c], method=sc.sett.method,
                                                do=sc.sett.do_proj, rho=sc.rho, vx_y=vx_y))
        z, w = ref._update._admm_aux(sc.y, sc.sett)
        tmp = torch.zeros(tuple(sc.y[0].dim))
        obj = torch.zeros(1, 3, dtype=torch.float64)
        cg_iters = []
        orig_cg = ref._update.cg

        def counting_cg(*a, **k):
            r = orig_cg(*a, **k)
            cg_iters.append(shim_optim.cg.last_n_iter)
            return r

        ref._update.cg = counting_cg
        try:
            y, z, w, tmp, obj = ref._update._update_admm(sc.x, sc.y, z, w, sc.rho, tmp, obj, 0,
                                                         sc.sett)
        finally:
            ref._update.cg = orig_cg
        for c in range(C):
            put('y%d' % c, sc.y[c].dat)
        put('jtv', tmp)
        put('z', z)
        put('w', w)
        out['obj'] = obj.numpy()
        out['cg_iters'] = np.array(cg_iters, dtype=np.int32)
        out['rho'] = np.float32(float(sc.rho))
        path = os.path.join(GOLDEN_DIR, name + '.npz')
        np.savez_compressed(path, **out)
        print(name, 'cg_iters', cg_iters, 'obj', obj[0].tolist(),
              '%.0f KB' % (os.path.getsize(path) / 1024))


if __name__ == '__main__' and '--mid' in sys.argv:
    main_mid()
