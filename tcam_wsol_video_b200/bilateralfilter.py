"""``bilateralfilter`` -- drop-in for the reference's SWIG module of the same name.

Call contract of ``bilateralfilter.i:21-25`` / ``bilateralfilter.hpp:10-12``: contiguous 1-D
float32 numpy inputs, the output array is filled in place, nothing is returned.  A wrong
dtype / non-contiguous array raises TypeError like the numpy.i typemaps do; a failure of
the CUDA path raises ``TcamCrfError`` (the reference would abort).
"""
from __future__ import annotations

import numpy as np

from . import _lib

__all__ = ['bilateralfilter', 'bilateralfilter_batch']


def _in_array(a, name):
    if not isinstance(a, np.ndarray) or a.dtype != np.float32 or a.ndim != 1 or not a.flags['C_CONTIGUOUS']:
        raise TypeError(f"{name}: array of type 'float' (contiguous, 1-D float32) required")
    return a


def _inplace_array(a, name):
    a = _in_array(a, name)
    if not a.flags['WRITEABLE']:
        raise TypeError(f"{name}: writeable array required")
    return a


def bilateralfilter(image, in_, out, H, W, sigmargb, sigmaxy):
    image, in_, out = _in_array(image, 'image'), _in_array(in_, 'in'), _inplace_array(out, 'out')
    lib = _lib.load()
    rc = lib.bilateralfilter(image.ctypes.data, image.size, in_.ctypes.data, in_.size, out.ctypes.data, out.size,
                             int(H), int(W), float(sigmargb), float(sigmaxy))
    _lib.check(rc, 'bilateralfilter')


def bilateralfilter_batch(images, ins, outs, N, K, H, W, sigmargb, sigmaxy):
    images, ins, outs = _in_array(images, 'images'), _in_array(ins, 'ins'), _inplace_array(outs, 'outs')
    N, K, H, W = int(N), int(K), int(H), int(W)
    if images.size < N * 3 * H * W or ins.size < N * K * H * W or outs.size < N * K * H * W:
        raise ValueError('array too short for N, K, H, W')
    lib = _lib.load()
    rc = lib.bilateralfilter_batch(images.ctypes.data, images.size, ins.ctypes.data, ins.size, outs.ctypes.data,
                                   outs.size, N, K, H, W, float(sigmargb), float(sigmaxy))
    _lib.check(rc, 'bilateralfilter_batch')
