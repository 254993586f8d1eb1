"""nitorch.core.utils.ceil_pow (unires/_core.py:17; only used by the pow-crop)."""
import math


def ceil_pow(t, p=2.0, l=2.0, mx=None):
    """Smallest p^k l >= t per element (unires/_core.py:249)."""
    import torch
    t = torch.as_tensor(t)
    out = []
    for v in t.reshape(-1).tolist():
        k = 0 if v <= l else math.ceil(math.log(v / l, p) - 1e-12)
        c = l * p ** k
        out.append(min(c, mx) if mx is not None else c)
    return torch.tensor(out, dtype=t.dtype, device=t.device).reshape(t.shape)
