"""nitorch.core.math.round (unires/_update.py:11, unires/_util.py:3, unires/_core.py:13)."""
import torch


def round(t, decimals=0):
    t = torch.as_tensor(t)
    return torch.round(t * 10 ** decimals) / (10 ** decimals)
