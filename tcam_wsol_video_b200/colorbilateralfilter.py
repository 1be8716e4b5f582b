"""``colorbilateralfilter`` -- drop-in for the reference's SWIG module of the same name
(``colorbilateralfilter.i:21-25`` / ``colorbilateralfilter.hpp:10-16``).  See ``bilateralfilter.py``.
"""
from __future__ import annotations

from . import _lib
from .bilateralfilter import _in_array, _inplace_array

__all__ = ['colorbilateralfilter', 'colorbilateralfilter_batch']


def colorbilateralfilter(image, in_, out, H, W, sigmargb, DIM):
    image, in_, out = _in_array(image, 'image'), _in_array(in_, 'in'), _inplace_array(out, 'out')
    if image.size < int(DIM) * int(H) * int(W):
        raise ValueError('image too short for DIM, H, W')
    lib = _lib.load()
    rc = lib.colorbilateralfilter(image.ctypes.data, image.size, in_.ctypes.data, in_.size, out.ctypes.data, out.size,
                                  int(H), int(W), float(sigmargb), int(DIM))
    _lib.check(rc, 'colorbilateralfilter')


def colorbilateralfilter_batch(images, ins, outs, N, K, H, W, sigmargb, DIM):
    images, ins, outs = _in_array(images, 'images'), _in_array(ins, 'ins'), _inplace_array(outs, 'outs')
    N, K, H, W = int(N), int(K), int(H), int(W)
    if ins.size < N * K * H * W or outs.size < N * K * H * W:
        raise ValueError('array too short for N, K, H, W')
    lib = _lib.load()
    rc = lib.colorbilateralfilter_batch(images.ctypes.data, images.size, ins.ctypes.data, ins.size, outs.ctypes.data,
                                        outs.size, N, K, H, W, float(sigmargb), int(DIM))
    _lib.check(rc, 'colorbilateralfilter_batch')
