"""Namespaces that let unires_b200.synth.make_scenario build one scenario for
the oracle port or for the reference's own files.  TEST INFRASTRUCTURE."""
import types

from oracle import unires_port as P


def _settings():
    return P.Settings()


def _input():
    return types.SimpleNamespace(dat=None, dim=None, ct=False, mat=None, mu=1.0, po=None,
                                 sd=1.0, tau=1.0, rigid_q=None, label=None)


def _output():
    return types.SimpleNamespace(dat=None, dim=None, lam=None, mat=None, label=None)


port_structs = types.SimpleNamespace(settings=_settings, _input=_input, _output=_output)


def _port_proj_info(dim_y, mat_y, dim_x, mat_x, rigid=None, prof_ip=0, prof_tp=0, gap=0.0,
                    device='cpu', scl=0.0, samp=0):
    return P.proj_info(dim_y, mat_y, dim_x, mat_x, rigid=rigid, prof_ip=prof_ip,
                       prof_tp=prof_tp, gap=gap, scl=scl)


port_ops = types.SimpleNamespace(_proj_info=_port_proj_info, _proj_apply=P.proj_apply)


def reference_namespaces():
    """(ops, structs) backed by the reference's unmodified files (build container only)."""
    from oracle.load_reference import load_reference
    ref = load_reference()

    def settings():
        s = ref.struct.settings()
        s.device = 'cpu'
        s.do_print = 0
        return s

    structs = types.SimpleNamespace(settings=settings, _input=ref.struct._input,
                                    _output=ref.struct._output)
    return ref._project, structs
