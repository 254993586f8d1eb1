"""nitorch.core._linalg_expm._expm restated: exp(sum_i q_i B_i) (unires/run.py:199).
TEST INFRASTRUCTURE; only the value (not the derivatives the rigid update needs)."""
import torch


def _expm(q, basis, grad_X=False, hess_X=False):
    if grad_X or hess_X:  # pragma: no cover
        raise NotImplementedError('rigid update is out of scope (SURVEY.md 8f #2)')
    X = torch.einsum('k,kij->ij', q.to(basis.dtype), basis)
    return torch.linalg.matrix_exp(X)
