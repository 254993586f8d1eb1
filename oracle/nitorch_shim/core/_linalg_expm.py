"""nitorch.core._linalg_expm._expm restated: R = exp(sum_i q_i B_i) and, for the rigid
Gauss-Newton update (unires/_update.py:601), its derivative dR/dq_i.  TEST INFRASTRUCTURE.
[EXT-UNVERIFIED]: nitorch evaluates the derivative with its own series; the derivative of the
matrix exponential is unique, here it is the exact block-triangular identity
    exp([[X, B_i], [0, X]]) = [[exp X, d exp(X)[B_i]], [0, exp X]]."""
import torch


def _expm(q, basis, grad_X=False, hess_X=False):
    if hess_X:  # pragma: no cover
        raise NotImplementedError('second derivatives of expm are not used by UniRes')
    q = q.to(basis.dtype)
    X = torch.einsum('k,kij->ij', q, basis)
    R = torch.linalg.matrix_exp(X)
    if not grad_X:
        return R
    n = X.shape[-1]
    dR = []
    for B in basis:
        blk = torch.zeros(2 * n, 2 * n, dtype=X.dtype, device=X.device)
        blk[:n, :n] = X
        blk[n:, n:] = X
        blk[:n, n:] = B
        dR.append(torch.linalg.matrix_exp(blk)[:n, n:])
    return R, torch.stack(dR)  # (n, n), (num_q, n, n)
