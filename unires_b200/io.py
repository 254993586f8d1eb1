"""NIfTI-1 in / out and the initial estimate, without nitorch.io / nibabel.

    _read_image    unires/_util.py:134-197   (file path, or [data, affine])
    _write_image   unires/_util.py:215-226
    _init_y_dat    unires/_core.py:371-399   (trilinear pull of every observation into the
                                              recon grid, clamped, averaged over repeats)

The reader handles single-file NIfTI-1 (.nii / .nii.gz, either byte order; uint8, int16, int32,
uint16, float32, float64; scl_slope / scl_inter; sform, else qform, else pixdim) -- enough for
the BrainWeb volumes shipped with the reference (int16 + scl_slope, sform) and for what the
writer produces (float32, sform = qform code 2).  Host-side code: parsing is numpy, the data
land on `device` as float32 like `file.fdata(dtype=float32, device=...)`.
"""
import gzip
import os
import struct

import numpy as np
import torch

_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 512: np.uint16,
           256: np.int8, 768: np.uint32}


def _open(path, mode):
    return gzip.open(path, mode) if str(path).endswith('.gz') else open(path, mode)


def _quatern_to_mat(b, c, d, qx, qy, qz, pixdim):
    a = np.sqrt(max(0.0, 1.0 - (b * b + c * c + d * d)))
    R = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                  [2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)],
                  [2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c]])
    qfac = -1.0 if pixdim[0] < 0 else 1.0
    S = np.diag([pixdim[1], pixdim[2], pixdim[3] * qfac])
    M = np.eye(4)
    M[:3, :3] = R @ S
    M[:3, 3] = [qx, qy, qz]
    return M


def read_nifti(path):
    """-> (data float32 numpy (X, Y, Z[, ...]) scaled by scl_slope/inter, affine float64 (4, 4))."""
    with _open(path, 'rb') as f:
        raw = f.read()
    if len(raw) < 352:
        raise ValueError('%s: not a NIfTI-1 file' % path)
    end = '<' if struct.unpack('<i', raw[0:4])[0] == 348 else '>'
    if struct.unpack(end + 'i', raw[0:4])[0] != 348 or raw[344:347] not in (b'n+1', b'ni1'):
        raise ValueError('%s: not a single-file NIfTI-1 volume' % path)
    if raw[344:347] == b'ni1':
        raise NotImplementedError('header/image pair (.hdr/.img) is not supported')
    dim = struct.unpack(end + '8h', raw[40:56])
    datatype = struct.unpack(end + 'h', raw[70:72])[0]
    pixdim = struct.unpack(end + '8f', raw[76:108])
    vox_offset = int(struct.unpack(end + 'f', raw[108:112])[0])
    slope, inter = struct.unpack(end + '2f', raw[112:120])
    qform_code, sform_code = struct.unpack(end + '2h', raw[252:256])
    if datatype not in _DTYPES:
        raise NotImplementedError('NIfTI datatype %d' % datatype)
    shape = tuple(int(d) for d in dim[1:1 + dim[0]])
    n = int(np.prod(shape))
    dt = np.dtype(_DTYPES[datatype]).newbyteorder(end)
    data = np.frombuffer(raw, dtype=dt, count=n, offset=vox_offset).reshape(shape, order='F')
    data = data.astype(np.float32)
    if slope != 0 and not (slope == 1 and inter == 0) and np.isfinite(slope):
        data = data * np.float32(slope) + np.float32(inter)
    if sform_code > 0:
        mat = np.eye(4)
        mat[0] = struct.unpack(end + '4f', raw[280:296])
        mat[1] = struct.unpack(end + '4f', raw[296:312])
        mat[2] = struct.unpack(end + '4f', raw[312:328])
    elif qform_code > 0:
        b, c, d, qx, qy, qz = struct.unpack(end + '6f', raw[256:280])
        mat = _quatern_to_mat(b, c, d, qx, qy, qz, pixdim)
    else:
        mat = np.diag([pixdim[1] or 1.0, pixdim[2] or 1.0, pixdim[3] or 1.0, 1.0])
    return np.ascontiguousarray(data), mat.astype(np.float64)


def write_nifti(dat, path, mat=None):
    """float32 single-file NIfTI-1 with sform (and a matching pixdim); .gz by extension."""
    a = np.asarray(torch.as_tensor(dat).detach().cpu().numpy(), dtype=np.float32)
    if a.ndim < 3 or a.ndim > 7:
        raise ValueError('write_nifti: 3 to 7 dimensions')
    M = np.eye(4) if mat is None else np.asarray(torch.as_tensor(mat).detach().cpu().numpy(),
                                                 dtype=np.float64)
    hdr = bytearray(352)
    struct.pack_into('<i', hdr, 0, 348)
    dims = [a.ndim] + list(a.shape) + [1] * (7 - a.ndim)
    struct.pack_into('<8h', hdr, 40, *dims)
    struct.pack_into('<h', hdr, 70, 16)   # float32
    struct.pack_into('<h', hdr, 72, 32)   # bitpix
    vx = np.sqrt((M[:3, :3] ** 2).sum(0))
    struct.pack_into('<8f', hdr, 76, 1.0, float(vx[0]), float(vx[1]), float(vx[2]), 1, 1, 1, 1)
    struct.pack_into('<f', hdr, 108, 352.0)
    struct.pack_into('<2f', hdr, 112, 1.0, 0.0)
    hdr[123] = 2                           # xyzt_units: mm
    struct.pack_into('<2h', hdr, 252, 0, 2)  # sform only (aligned)
    for r in range(3):
        struct.pack_into('<4f', hdr, 280 + 16 * r, *[float(v) for v in M[r]])
    hdr[344:348] = b'n+1\x00'
    with _open(path, 'wb') as f:
        f.write(bytes(hdr))
        f.write(np.asfortranarray(a).tobytes(order='F'))
    return path


def _read_image(data, device='cpu', is_ct=False):
    """(dat, dim, mat, fname, direc, nam, file, ct) like unires/_util.py:134-197."""
    if isinstance(data, str):
        arr, mat = read_nifti(data)
        dat = torch.from_numpy(arr).to(device)
        mat = torch.from_numpy(mat).to(device)
        fname = os.path.abspath(data)
        direc, nam = os.path.split(fname)
        file = fname
    else:
        dat = torch.as_tensor(data[0]).float().to(device)
        mat = torch.as_tensor(data[1]).double().to(device)
        file = fname = direc = nam = None
    dat = dat.squeeze()
    dim = tuple(dat.shape)
    if len(dim) != 3:
        raise ValueError('Input image dimension required to be 3D, recieved {:}D!'.format(len(dim)))
    dat = dat.contiguous()
    dat[~torch.isfinite(dat)] = 0.0
    return dat, dim, mat, fname, direc, nam, file, bool(is_ct)


def _write_image(dat, fname, bids=False, mat=torch.eye(4), file=None, dtype='float32',
                 do_print=False):
    """Write a volume to NIfTI (unires/_util.py:215-226; `file` / `dtype` kept for signature
    compatibility: the output is always float32 with the given affine)."""
    if bids:
        p, n = os.path.split(fname)
        s = n.split('_')
        fname = os.path.join(p, '_'.join(s[:-1] + ['space-unires'] + [s[-1]]))
    write_nifti(dat, fname, mat)
    if do_print:
        print('Output saved to: %s' % fname)
    return fname


def _init_y_dat(x, y, sett):
    """Initial reconstruction: every observation pulled trilinearly into the recon grid, clamped
    to its own intensity range, averaged over the repeats that are positive there
    (unires/_core.py:371-399).  Runs on the CUDA resampling kernel."""
    from .spatial import affine_grid, grid_pull
    dim_y, mat_y = tuple(y[0].dim), y[0].mat
    cpu = lambda t: torch.as_tensor(t).detach().to('cpu', torch.float64)
    for c in range(len(x)):
        acc = torch.zeros(dim_y, dtype=torch.float32, device=sett.device)
        cnt = torch.zeros_like(acc)
        for obs in x[c]:
            dat = obs.dat[None, None, ...]
            mat = torch.linalg.solve(cpu(obs.mat), cpu(mat_y))
            grid = affine_grid(mat.to(torch.float32), dim_y)
            mn, mx = torch.min(dat), torch.max(dat)
            pulled = grid_pull(dat, grid[None, ...], bound='zero', extrapolate=False,
                               interpolation=1)[0, 0]
            pulled = torch.minimum(torch.maximum(pulled, mn), mx)
            cnt = cnt + (pulled > 0)
            acc = acc + pulled
        cnt[cnt == 0] = 1.0
        y[c].dat = acc / cnt
    return y
