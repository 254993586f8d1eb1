"""nitorch.core.math.round stand-in (unires/_core.py:203-207 uses round(t, 3))."""
import torch


def round(t, decimals=0):
    return torch.round(t * 10 ** decimals) / (10 ** decimals)
