"""nitorch.core.optim.{cg, get_gain} restated.

TEST INFRASTRUCTURE, PARITY UNPINNED (see oracle/__init__.py).  Call sites:
  cg        unires/_update.py:142-148 (stop='max_gain', inplace=True,
            max_iter=20, tolerance=1e-3, identity preconditioner)
  get_gain  unires/run.py:100

Stop rule: nitorch reduces ``stop`` to its first
letter; only 'e' (residual) selects sqrt(r.z); anything else -- including
UniRes' 'max_gain' -- selects the energy 0.5 x'Ax - b'x, which costs one
extra A(x) per iteration.  ``record`` is an oracle-only hook that collects
each iterate for the parity tests.
"""
import torch


def get_gain(obj, monotonicity='increasing'):
    if len(obj) <= 1:
        return torch.tensor(float('inf'), dtype=obj.dtype, device=obj.device)
    if monotonicity == 'increasing':
        gain = obj[-1] - obj[-2]
    elif monotonicity == 'decreasing':
        gain = obj[-2] - obj[-1]
    else:
        raise ValueError('Undefined monotonicity')
    return gain / (torch.max(obj) - torch.min(obj))


def plot_convergence(*args, **kwargs):  # pragma: no cover
    return None


def _stop_letter(stop):
    if stop == 'residual':
        stop = 'e'
    elif stop == 'norm':
        stop = 'a'
    return stop[0].lower()


def cg(A, b, x=None, precond=lambda y: y, max_iter=None, tolerance=1e-5,
       verbose=False, sum_dtype=torch.float64, inplace=True, stop='E',
       record=None):
    max_iter = max_iter or len(b) * 10
    if x is None:
        x = torch.zeros_like(b)
    elif not inplace:
        x = x.clone()
    if isinstance(A, torch.Tensor):
        A_mat = A
        A = lambda v: A_mat.mm(v)

    r = b - A(x)
    z = precond(r)
    rz = torch.sum(r * z, dtype=sum_dtype)
    p = z.clone()

    track = bool(tolerance or verbose)
    if track:
        stop = _stop_letter(stop)

        def objective():
            if stop == 'e':
                return torch.sqrt(rz)
            o = A(x).sub_(2 * b).mul_(x)
            return 0.5 * torch.sum(o, dtype=sum_dtype)

        obj = torch.zeros(max_iter + 1, dtype=sum_dtype, device=b.device)
        obj[0] = objective()
    n_done = 0
    for n_iter in range(1, max_iter + 1):
        Ap = A(p)
        alpha = rz / torch.sum(p * Ap, dtype=sum_dtype)
        x += alpha * p
        r -= alpha * Ap
        z = precond(r)
        rz0 = rz
        rz = torch.sum(r * z, dtype=sum_dtype)
        beta = rz / rz0
        p *= beta
        p += z
        n_done = n_iter
        if record is not None:
            record(n_iter, x)
        if track:
            obj[n_iter] = objective()
            gain = get_gain(obj[:n_iter + 1], monotonicity='decreasing')
            if verbose:
                print('{:3d} | {} = {:12.6g} | gain = {:12.6g}'.format(
                    n_iter, stop, obj[n_iter].item(), gain.item()))
            if gain.abs() < tolerance:
                break
    cg.last_n_iter = n_done
    cg.last_obj = obj[:n_done + 1].clone() if track else None
    return x
