"""nitorch.core._linalg_expm._expm (unires/_update.py:8, unires/run.py:9): matrix exponential
of sum_i q_i B_i and its derivatives, float64 on the host (unires_b200._update._expm)."""
import torch


def _expm(q, basis, grad_X=False):
    from ..._update import _expm as impl
    dev = q.device if isinstance(q, torch.Tensor) else 'cpu'
    out = impl(q, basis, grad_X=grad_X)
    if grad_X:
        return out[0].to(dev), out[1].to(dev)
    return out.to(dev)
