"""Shared helpers for the parity tests."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from unires_b200 import synth  # noqa: E402
from oracle import gen_golden  # noqa: E402  (recipes + digest only; no reference access)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')
GOLDEN_NAMES = sorted(gen_golden.RECIPES)
REL_TOL = 1e-4  # north_star: <= 1e-4 relative L2 per CG iterate


def load_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False)
    recipe = json.loads(str(g['recipe']))
    return g, recipe


def build(recipe, ops, structs, device='cpu'):
    return gen_golden.build(recipe, ops, structs, device=device)


def rel_l2(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)


def product_namespaces():
    from unires_b200 import _project, struct
    return _project, struct


def port_namespaces():
    from oracle.adapters import port_ops, port_structs
    return port_ops, port_structs


def to_device(sc, device):
    """Move a CPU scenario (built with the oracle) onto the GPU product's containers."""
    from unires_b200 import struct, _project
    x, y = [], []
    for c in range(len(sc.x)):
        row = []
        for o in sc.x[c]:
            n = struct._input(dat=o.dat.to(device), dim=o.dim, mat=o.mat, tau=float(o.tau),
                              mu=o.mu, sd=o.sd, ct=False)
            if o.po is not None:
                n.po = _project._proj_info(o.po.dim_y, o.po.mat_y, o.po.dim_x, o.po.mat_x,
                                           rigid=o.po.rigid, prof_ip=sc.sett.profile_ip,
                                           prof_tp=sc.sett.profile_tp, gap=sc.sett.gap,
                                           device=device, scl=float(o.po.scl))
            row.append(n)
        x.append(row)
        y.append(struct._output(dat=sc.y[c].dat.to(device).clone(), dim=tuple(sc.y[c].dim),
                                mat=sc.y[c].mat, lam=float(sc.y[c].lam)))
    sett = struct.settings()
    for k in ('alpha', 'bound', 'cgs_max_iter', 'cgs_tol', 'cgs_verbose', 'diff', 'do_proj',
              'interpolation', 'method', 'rho', 'rho_scl', 'tolerance', 'profile_ip',
              'profile_tp', 'gap'):
        setattr(sett, k, getattr(sc.sett, k))
    sett.device = str(device)
    sett.do_print = 0
    return x, y, sett
