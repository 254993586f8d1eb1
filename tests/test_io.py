"""NIfTI-1 reader / writer (unires_b200/io.py) -- host-side, no GPU needed."""
import gzip
import struct

import numpy as np
import pytest
import torch

from unires_b200 import io


def _raw_nifti(path, data, dtype_code, endian, slope, inter, sform=None, quatern=None, pixdim=None):
    hdr = bytearray(352)
    struct.pack_into(endian + 'i', hdr, 0, 348)
    struct.pack_into(endian + '8h', hdr, 40, 3, *data.shape, 1, 1, 1, 1)
    struct.pack_into(endian + 'h', hdr, 70, dtype_code)
    struct.pack_into(endian + '8f', hdr, 76, *(pixdim or (1, 1, 1, 1, 0, 0, 0, 0)))
    struct.pack_into(endian + 'f', hdr, 108, 352.0)
    struct.pack_into(endian + '2f', hdr, 112, slope, inter)
    if sform is not None:
        struct.pack_into(endian + '2h', hdr, 252, 0, 1)
        for r in range(3):
            struct.pack_into(endian + '4f', hdr, 280 + 16 * r, *sform[r])
    elif quatern is not None:
        struct.pack_into(endian + '2h', hdr, 252, 1, 0)
        struct.pack_into(endian + '6f', hdr, 256, *quatern)
    hdr[344:348] = b'n+1\x00'
    body = np.asfortranarray(data.astype(data.dtype.newbyteorder(endian))).tobytes(order='F')
    opener = gzip.open if str(path).endswith('.gz') else open
    with opener(path, 'wb') as f:
        f.write(bytes(hdr) + body)


@pytest.mark.parametrize('endian', ['<', '>'])
def test_read_int16_with_slope_and_sform(tmp_path, endian):
    """The layout of the BrainWeb volumes shipped with the reference: int16 + scl_slope, sform."""
    rng = np.random.default_rng(0)
    data = rng.integers(0, 3000, size=(5, 7, 6)).astype(np.int16)
    sform = [[-1, 0, 0, 90], [0, 1, 0, -126], [0, 0, 1, -72]]
    p = tmp_path / 'v.nii.gz'
    _raw_nifti(p, data, 4, endian, 0.0365198776, 0.0, sform=sform)
    arr, mat = io.read_nifti(str(p))
    assert arr.dtype == np.float32 and arr.shape == (5, 7, 6)
    assert np.array_equal(arr, data.astype(np.float32) * np.float32(0.0365198776))
    assert np.array_equal(mat[:3], np.array(sform, dtype=np.float64)) and mat[3].tolist() == [0, 0, 0, 1]


def test_read_qform_and_pixdim_fallback(tmp_path):
    data = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    p = tmp_path / 'q.nii'
    # 90 degree rotation about z: quaternion (b, c, d) = (0, 0, sin 45)
    _raw_nifti(p, data, 16, '<', 1.0, 0.0, quatern=(0, 0, np.sqrt(0.5), 10, 20, 30),
               pixdim=(1, 2.0, 3.0, 4.0, 0, 0, 0, 0))
    arr, mat = io.read_nifti(str(p))
    assert np.array_equal(arr, data)
    want = np.array([[0, -3, 0, 10], [2, 0, 0, 20], [0, 0, 4, 30], [0, 0, 0, 1]], dtype=np.float64)
    assert np.allclose(mat, want, atol=1e-6)
    p2 = tmp_path / 'p.nii'
    _raw_nifti(p2, data, 16, '<', 0.0, 0.0, pixdim=(1, 2.0, 3.0, 4.0, 0, 0, 0, 0))
    assert np.allclose(io.read_nifti(str(p2))[1], np.diag([2.0, 3.0, 4.0, 1.0]))


def test_write_read_round_trip_and_reference_signatures(tmp_path):
    g = torch.Generator().manual_seed(1)
    vol = torch.rand((6, 5, 7), generator=g)
    mat = torch.tensor([[0.0, -1.5, 0, 3], [1.2, 0, 0, -4], [0, 0, 2.0, 5], [0, 0, 0, 1]],
                       dtype=torch.float64)
    fname = io._write_image(vol, str(tmp_path / 'sub_T1w.nii.gz'), bids=True, mat=mat)
    assert fname.endswith('sub_space-unires_T1w.nii.gz')
    dat, dim, m, f, direc, nam, file, ct = io._read_image(fname)
    assert torch.equal(dat, vol) and dim == (6, 5, 7) and torch.allclose(m, mat) and ct is False
    assert nam == 'sub_space-unires_T1w.nii.gz' and direc == str(tmp_path)
    # [data, affine] input, non-finite values zeroed, singleton dims squeezed
    v2 = vol.clone()[None]
    v2[0, 0, 0, 0] = float('nan')
    dat2, dim2, m2, *_ = io._read_image([v2.numpy(), mat.numpy()])
    assert dim2 == (6, 5, 7) and dat2[0, 0, 0] == 0 and m2.dtype == torch.float64
    with pytest.raises(ValueError):
        io._read_image([torch.zeros(4, 4), mat])
