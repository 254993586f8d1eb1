"""Attribute bags of the hot path, field-compatible with unires/struct.py.

`_input` (:4-22), `_output` (:25-33), `_proj_op` (:36-54) and `settings`
(:57-111) carry the same attribute names and defaults as the reference so that
objects built for either implementation can be handed to the other.
"""


class _bag:
    _defaults = {}

    def __init__(self, **overrides):
        for name, value in self._defaults.items():
            setattr(self, name, value() if callable(value) else value)
        for name, value in overrides.items():
            setattr(self, name, value)

    def __repr__(self):
        body = ', '.join('%s=%r' % kv for kv in sorted(vars(self).items()))
        return '%s(%s)' % (type(self).__name__, body)


class _input(_bag):
    """One observed image: data, affine, noise precision tau, operator po."""
    _defaults = dict(dat=None, dim=None, ct=None, mat=None, mu=1.0, po=None, sd=1.0,
                     tau=1.0, file=None, fname=None, direc=None, nam=None,
                     rigid_q=None, label=None)


class _output(_bag):
    """One reconstructed channel: data, affine, regularisation lam."""
    _defaults = dict(dat=None, dim=None, lam=None, mat=None, label=None)


class _proj_op(_bag):
    """Projection operator description produced by _project._proj_info."""
    _defaults = dict(dim_x=None, mat_x=None, vx_x=None, dim_y=None, mat_y=None,
                     vx_y=None, dim_yx=None, mat_yx=None, ratio=None, smo_ker=None,
                     rigid=None, scl=None, dim_thick=None, D_x=None, D_y=None)


class settings(_bag):
    """Algorithm settings (same names and defaults as unires/struct.py:57-111)."""
    _defaults = dict(
        # solver (hot path)
        alpha=1.0, bound='zero', cgs_max_iter=20, cgs_tol=1e-3, cgs_verbose=False,
        device='cuda', diff='forward', do_proj=None, interpolation='linear',
        method=None, rho=None, rho_scl=1.0, tolerance=1e-4, max_iter=512,
        # slice profile
        gap=0.0, profile_ip=2, profile_tp=0,
        # regularisation schedule
        reg_scl=4.0, sched_num=3,
        # pipeline switches that live in the same bag upstream (not used here)
        atlas_rigid=False, bids=False, clean_fov=False,
        coreg_params=lambda: {'cost_fun': 'nmi', 'group': 'SE', 'samp': (1), 'fwhm': 7,
                              'mean_space': False},
        crop=False, common_output=False, ct=False, dir_out=None, do_coreg=True,
        do_atlas_align=False, do_print=1, do_res_origin=False, fix=0,
        force_inplane_res=False, fov='brain', label=None, mat=None, plot_conv=False,
        pow=0, prefix='u_', rigid_basis=None, rigid_mod=1, rigid_samp=1, scaling=False,
        show_hyperpar=False, show_jtv=False, unified_rigid=False, vx=1.0,
        write_jtv=False, write_out=True,
        # extension: which objective nitorch's cg uses for stop='max_gain'
        # (SURVEY.md Appendix A, Q1): 'energy' (nitorch's fall-through) or 'residual'
        cgs_stop='max_gain',
        # extension: number of CUDA streams the per-channel CG solves of one ADMM iteration
        # are spread over (the reference loops over channels sequentially; the systems are
        # independent, so this only changes scheduling, not results)
        channel_streams=3)
