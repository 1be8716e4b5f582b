"""DenseCRFFilter -- drop-in for dlib/crf/crf_post_processing.py:33-135 (mean-field CRF refinement).

The reference hands every image to pydensecrf (`DenseCRF2D` + `addPairwiseBilateral(compat=10, DIAG_KERNEL,
NORMALIZE_SYMMETRIC)` + `inference(itera)`), i.e. Kraehenbuehl & Koltun's mean field on the CPU, one image at a
time.  Here the whole batch runs on the GPU through ONE permutohedral lattice (ops.Lattice: built once from the
images, applied `itera + 1` times -- once for the normalisation, once per iteration):

    U    = -log(seg)                                        (crf_post_processing.py:86)
    norm = 1 / sqrt(A 1 + 1e-20)                            (densecrf pairwise.cpp, NORMALIZE_SYMMETRIC)
    Q    = softmax(-U)
    repeat itera times:  Q = softmax(-U + compat * norm * A(norm * Q))      (densecrf.cpp inference(), Potts)

pydensecrf is not part of this image (requirements.txt:64 pins pydensecrf@0d53acb): the algorithm above is restated
from the densecrf sources it wraps; tests/test_gpu_losses.py checks this module against a numpy restatement built on
the CPU oracle's filter.  Parity against the pydensecrf binary itself is unpinned.

Reference quirk kept: the image is handed to pydensecrf as `img.numpy().astype(uint8).transpose(2, 1, 0)`
(crf_post_processing.py:116-118) -- a [W,H,3] array where DenseCRF2D(w, h, k) expects [H,W,3].  The memory is read as
[H,W,3] anyway, so for square frames every pixel gets the colour of its mirror pixel (row and column swapped), and
for other frames the colours are the [W,H,3] memory re-read row-major.  `quirk_transposed_image=True` (default)
reproduces that; False uses the image as it is.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib, ops

__all__ = ['DenseCRFFilter']

_COMPAT = 10.0          # crf_post_processing.py:119
_MAX_FRAMES = 64        # frames per lattice (tcamcrf chunk_frames default)


class DenseCRFFilter(object):
    def __init__(self, sigma_rgb: float, sigma_xy: float, scale_factor: float, itera: int,
                 quirk_transposed_image: bool = True):
        """
        :param sigma_rgb: float. colour bandwidth of the appearance kernel (truncated to int like the reference).
        :param sigma_xy: float. spatial bandwidth of the appearance kernel (truncated to int).
        :param scale_factor: float. images and segmentations are rescaled by it first.
        :param itera: int. number of mean-field iterations.
        """
        super(DenseCRFFilter, self).__init__()
        self.sigma_rgb = int(sigma_rgb)
        self.sigma_xy = int(sigma_xy)
        self.scale_factor = scale_factor
        assert isinstance(itera, int)
        assert itera >= 0
        self.itera = itera
        self.quirk_transposed_image = quirk_transposed_image

    def __call__(self, images: torch.Tensor, segmentations: torch.Tensor) -> torch.Tensor:
        """
        :param images: [N,3,H,W], values in [0, 255].  CPU (like the reference) or CUDA.
        :param segmentations: [N,K,H,W] softmaxed logits.  CPU (like the reference) or CUDA.
        :return: refined segmentations, same shape and device as `segmentations` (after rescaling).
        """
        assert isinstance(images, torch.Tensor)
        assert isinstance(segmentations, torch.Tensor)
        assert images.ndim == 4
        assert segmentations.ndim == 4
        assert images.shape[0] == segmentations.shape[0]
        assert images.shape[2:] == segmentations.shape[2:]

        scaled_images = F.interpolate(images.float(), scale_factor=self.scale_factor, mode='nearest',
                                      recompute_scale_factor=False)
        scaled_segs = F.interpolate(segmentations, scale_factor=self.scale_factor, mode='bilinear',
                                    recompute_scale_factor=False, align_corners=False)
        if self.itera == 0:
            return scaled_segs
        assert scaled_images.shape[1] == 3

        out_device = segmentations.device
        if segmentations.is_cuda:
            dev = segmentations.device
        elif images.is_cuda:
            dev = images.device
        else:
            if not torch.cuda.is_available():
                raise _lib.TcamCrfError("DenseCRFFilter needs a CUDA device: this package has no CPU path")
            dev = torch.device('cuda', torch.cuda.current_device())
        n, k, h, w = scaled_segs.shape
        img = scaled_images.to(dev).to(torch.uint8)           # .astype(np.uint8), crf_post_processing.py:117
        if self.quirk_transposed_image:
            # [3,H,W] -> transpose(2,1,0) -> [W,H,3] contiguous, re-read as [H,W,3] (see the module docstring)
            img = img.permute(0, 3, 2, 1).contiguous().view(n, h, w, 3).permute(0, 3, 1, 2).contiguous()
        unary = -torch.log(scaled_segs.to(dev).float())        # crf_post_processing.py:86
        xy = int(self.sigma_xy * self.scale_factor)            # crf_post_processing.py:113
        cfg = _lib.make_config(ops.FEAT_XY_RGB, 3, float(self.sigma_rgb), float(xy))
        out = torch.empty_like(unary)
        per = min(_MAX_FRAMES, ops.lattice_capacity(cfg, unary.shape[1], h, w))   # fewer for large frames
        for n0 in range(0, n, per):
            n1 = min(n, n0 + per)
            out[n0:n1] = self._mean_field(img[n0:n1], unary[n0:n1], cfg)
        return out.to(out_device)

    def _mean_field(self, img: torch.Tensor, unary: torch.Tensor, cfg) -> torch.Tensor:
        n, k, h, w = unary.shape
        lattice = ops.Lattice(img, cfg, k)
        # NORMALIZE_SYMMETRIC: norm = 1/sqrt(A 1 + 1e-20); the K channels of A 1 are identical, channel 0 is used
        ones = torch.ones_like(unary)
        norm = torch.rsqrt(lattice.apply(ones)[:, :1] + 1e-20)
        q = torch.softmax(-unary, dim=1)
        for _ in range(self.itera):
            msg = lattice.apply(q * norm) * norm
            q = torch.softmax(_COMPAT * msg - unary, dim=1)
        return q

    def __str__(self):
        return '{}: (sigma_rgb={}, sigma_xy={}, scale_factor={})'.format(
            self.__class__.__name__, self.sigma_rgb, self.sigma_xy, self.scale_factor)
