"""ColorDenseCRFLoss -- drop-in for ``dlib/crf/color_dense_crf_loss.py`` of the reference.

Colour-only lattice (no xy term); the lattice dimension is the number of image planes
(``nbr_p = images.shape[1]``, color_dense_crf_loss.py:47).  Value and gradient as in
color_dense_crf_loss.py:65-76.

Limits of this implementation (the reference has none, its int16 keys wrap silently instead):
1 <= C <= 6 image planes (vertex keys are packed into one 64-bit word: 20 bits per coordinate for C <= 3,
15 / 12 / 10 bits for C = 4 / 5 / 6), ``sigma_rgb`` large enough for 0..255 frames to stay inside those fields
(sigma_rgb >= 0.001 / 0.021 / 0.19 / 0.81 for C <= 3 / 4 / 5 / 6; the 5-D xy+RGB lattice of ``DenseCRFLoss`` needs
sigma_rgb >= 0.1.  Checked on the host -- ``tcamcrf_key_range_ok`` -- with a clear ``TcamCrfError`` instead of a NaN loss),
and a batch of N > 1 frames needs C == 3 (the reference strides the batch by three planes whatever C is,
colorbilateralfilter.cpp:50, which reads overlapping windows for any other C).
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib, ops
from .dense_crf_loss import _folded_weight, _scale_images, _scale_segs

__all__ = ['ColorDenseCRFLoss', 'ColorDenseCRFLossFunction']


class ColorDenseCRFLossFunction(Function):

    @staticmethod
    @torch.amp.custom_fwd(device_type='cuda')
    def forward(ctx, images, segmentations, sigma_rgb, weight=1.0):
        n = segmentations.shape[0]
        nbr_p = images.shape[1]
        if n > 1 and nbr_p != 3:
            # the reference's batch loop strides images by 3 planes whatever DIM is
            # (colorbilateralfilter.cpp:50), which is only meaningful for 3 planes
            raise _lib.TcamCrfError("ColorDenseCRFLoss with N > 1 needs 3 image planes (reference stride quirk)")
        cfg = _lib.make_config(ops.FEAT_COLOR, nbr_p, sigma_rgb, loss_weight=weight)
        _lib.require_key_range(cfg, segmentations.shape[2], segmentations.shape[3])
        as_t, loss, _ = ops.crf_forward(images, segmentations.detach(), cfg, want_loss=True, n_norm=float(n))
        ctx.AS = as_t
        ctx.N = n
        ctx.weight = float(weight)
        return loss

    @staticmethod
    @torch.amp.custom_bwd(device_type='cuda')
    def backward(ctx, grad_output):
        grad_segmentation = ops.crf_backward(ctx.AS, grad_output, float(ctx.N), ctx.weight)
        return None, grad_segmentation, None, None


class ColorDenseCRFLoss(nn.Module):
    def __init__(self, weight, sigma_rgb, scale_factor):
        """
        :param weight: float. lambda of the CRF loss.
        :param sigma_rgb: float. colour bandwidth of the kernel.
        :param scale_factor: float. images and segmentations are rescaled by it first.
        """
        super(ColorDenseCRFLoss, self).__init__()
        self.weight = weight
        self.sigma_rgb = sigma_rgb
        self.scale_factor = scale_factor

    def forward(self, images, segmentations):
        """
        :param images: N*C*H*W tensor with values in [0, 255]; 1 <= C <= 6 planes (C == 3 when N > 1, see above).
        :param segmentations: softmaxed logits, N*K*H*W, CUDA.
        :return: loss tensor of shape [1].
        """
        assert images.ndim == 4
        scaled_images = _scale_images(images, self.scale_factor)
        scaled_segs = _scale_segs(segmentations, self.scale_factor)
        w = _folded_weight(self.weight)
        if w is not None:   # weight * loss formed inside the loss kernels (same roundings, no extra launches)
            return ColorDenseCRFLossFunction.apply(scaled_images, scaled_segs, self.sigma_rgb, w)
        val = self.weight * ColorDenseCRFLossFunction.apply(scaled_images, scaled_segs, self.sigma_rgb)
        return val

    def extra_repr(self):
        return 'sigma_rgb={}, weight={}, scale_factor={}'.format(
            self.sigma_rgb, self.weight, self.scale_factor
        )
