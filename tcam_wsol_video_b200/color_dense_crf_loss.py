"""ColorDenseCRFLoss -- drop-in for ``dlib/crf/color_dense_crf_loss.py`` of the reference.

Colour-only lattice (no xy term); the lattice dimension is the number of image planes
(``nbr_p = images.shape[1]``, color_dense_crf_loss.py:47).  Value and gradient as in
color_dense_crf_loss.py:65-76.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib, ops
from .dense_crf_loss import _scale_images, _scale_segs

__all__ = ['ColorDenseCRFLoss', 'ColorDenseCRFLossFunction']


class ColorDenseCRFLossFunction(Function):

    @staticmethod
    @torch.amp.custom_fwd(device_type='cuda')
    def forward(ctx, images, segmentations, sigma_rgb):
        n = segmentations.shape[0]
        nbr_p = images.shape[1]
        if n > 1 and nbr_p != 3:
            # the reference's batch loop strides images by 3 planes whatever DIM is
            # (colorbilateralfilter.cpp:50), which is only meaningful for 3 planes
            raise _lib.TcamCrfError("ColorDenseCRFLoss with N > 1 needs 3 image planes (reference stride quirk)")
        cfg = _lib.make_config(ops.FEAT_COLOR, nbr_p, sigma_rgb)
        as_t, loss, _ = ops.crf_forward(images, segmentations.detach(), cfg, want_loss=True, n_norm=float(n))
        ctx.AS = as_t
        ctx.N = n
        return loss

    @staticmethod
    @torch.amp.custom_bwd(device_type='cuda')
    def backward(ctx, grad_output):
        grad_segmentation = ops.crf_backward(ctx.AS, grad_output, float(ctx.N))
        return None, grad_segmentation, None


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
        :param images: N*C*H*W tensor with values in [0, 255] (any number of planes C > 0).
        :param segmentations: softmaxed logits, N*K*H*W, CUDA.
        :return: loss tensor of shape [1].
        """
        assert images.ndim == 4
        scaled_images = _scale_images(images, self.scale_factor)
        scaled_segs = _scale_segs(segmentations, self.scale_factor)
        val = self.weight * ColorDenseCRFLossFunction.apply(scaled_images, scaled_segs, self.sigma_rgb)
        return val

    def extra_repr(self):
        return 'sigma_rgb={}, weight={}, scale_factor={}'.format(
            self.sigma_rgb, self.weight, self.scale_factor
        )
