"""nitorch.io.map / savef (unires/_util.py:4) on the product's NIfTI-1 reader / writer
(unires_b200/io.py)."""
import os

import torch

from ..io import read_nifti, write_nifti


class _Mapped:
    """The slice of nitorch's BabelArray that unires/_util.py:156-160 touches."""

    def __init__(self, path):
        self._path = path
        self._dat, mat = read_nifti(path)
        self.affine = torch.as_tensor(mat, dtype=torch.float64)
        self.shape = tuple(self._dat.shape)

    def fdata(self, dtype=torch.float32, device='cpu', rand=False, cutoff=None):
        if rand or cutoff is not None:
            raise NotImplementedError('fdata(rand=, cutoff=)')
        return torch.as_tensor(self._dat).to(device=device, dtype=dtype)

    def filename(self):
        return self._path


def map(path):
    return _Mapped(os.fspath(path))


def savef(dat, fname, like=None, affine=None):
    write_nifti(dat, fname, mat=affine)
