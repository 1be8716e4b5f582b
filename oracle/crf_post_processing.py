"""oracle/crf_post_processing.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy restatement of what dlib/crf/crf_post_processing.py:99-128 asks pydensecrf for:
DenseCRF2D(w, h, k) + setUnaryEnergy(-log seg) + addPairwiseBilateral(sxy, srgb, rgbim, compat=10, DIAG_KERNEL,
NORMALIZE_SYMMETRIC) + inference(itera).  pydensecrf (requirements.txt:64, pydensecrf@0d53acb) is not installed and
not under /root/reference, so the mean-field update is restated from the densecrf sources it wraps (densecrf.cpp
inference(): Q = expAndNormalize(-unary); tmp1 = -unary - pairwise(Q); pairwise.cpp: out = norm * filter(norm * Q),
norm = 1/sqrt(filter(1) + 1e-20), PottsCompatibility: out = -compat * in).  The permutohedral filter itself is the
oracle's (same densecrf lineage as the reference's crfwrapper).  Parity against the pydensecrf binary: unpinned.
"""
from __future__ import annotations

import numpy as np


def _softmax0(x: np.ndarray) -> np.ndarray:
    m = x.max(axis=0, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=0, keepdims=True)


def quirk_image(img: np.ndarray) -> np.ndarray:
    """[3,H,W] -> the image pydensecrf actually sees (crf_post_processing.py:116-118): transpose(2,1,0) gives a
    [W,H,3] array whose memory DenseCRF2D(w, h, k) reads as [H,W,3]."""
    c, h, w = img.shape
    seen = np.ascontiguousarray(img.astype(np.uint8).transpose(2, 1, 0)).reshape(h, w, c)
    return np.ascontiguousarray(seen.transpose(2, 0, 1))


def mean_field(img: np.ndarray, seg: np.ndarray, sigma_rgb: int, sigma_xy: int, itera: int, filter_batch,
               compat: float = 10.0, quirk: bool = True) -> np.ndarray:
    """img [3,H,W] (0..255), seg [K,H,W] probabilities -> refined [K,H,W].  filter_batch is one of the oracle's
    *_bilateralfilter_batch functions."""
    k, h, w = seg.shape
    image = quirk_image(img) if quirk else img.astype(np.uint8)
    image = np.ascontiguousarray(image.astype(np.float32))[None]
    unary = -np.log(seg.astype(np.float32))

    def filt(x):
        return filter_batch(image, np.ascontiguousarray(x[None].astype(np.float32)), 1, k, h, w,
                            float(sigma_rgb), float(sigma_xy)).reshape(k, h, w)

    norm = 1.0 / np.sqrt(filt(np.ones_like(unary))[:1] + 1e-20)
    q = _softmax0(-unary)
    for _ in range(itera):
        q = _softmax0(compat * (norm * filt(q * norm)) - unary)
    return q.astype(np.float32)
