"""Seeded synthetic YTOv2.2-shaped inputs shared by tests/ and bench.py.

Two image distributions, because the lattice size M swings ~20x between them
(SURVEY.md §8d): ``noise`` = iid uniform integers in [0,255] (worst case,
M ~ 1.1 * H*W at 224^2) and ``natural`` = a smooth field plus +-4 noise
(M ~ 0.06 * H*W), which is closer to real video frames.  Segmentations follow the
reference's own self-test recipe, ``softmax(rand)`` (dlib/crf/dense_crf_loss.py:155-169).
"""
from __future__ import annotations

import numpy as np


def make_images(n: int, h: int, w: int, kind: str = "noise", seed: int = 0, channels: int = 3) -> np.ndarray:
    """float32 [n, channels, h, w] with integer values in [0, 255]."""
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 256, size=(n, channels, h, w)).astype(np.float32)
    if kind == "natural":
        y, x = np.mgrid[0:h, 0:w].astype(np.float64)
        out = np.empty((n, channels, h, w), dtype=np.float32)
        for i in range(n):
            for c in range(channels):
                base = 127.5 + 100.0 * np.sin(0.03 * x + 0.5 * c + i) * np.cos(0.02 * y)
                jitter = rng.integers(-4, 5, size=(h, w))
                out[i, c] = np.clip(np.rint(base + jitter), 0, 255)
        return out
    raise ValueError(f"unknown image kind {kind!r}")


def make_segs(n: int, k: int, h: int, w: int, seed: int = 0) -> np.ndarray:
    """float32 [n, k, h, w] = softmax over k of uniform(0,1) logits."""
    rng = np.random.default_rng(seed + 1)
    logits = rng.random(size=(n, k, h, w), dtype=np.float32)
    e = np.exp(logits - logits.max(axis=1, keepdims=True))
    return (e / e.sum(axis=1, keepdims=True)).astype(np.float32)


def make_low_res_cams(b: int, t: int, h: int = 28, w: int = 28, seed: int = 0) -> np.ndarray:
    """float32 [b, t, 1, h, w] in [0,1): the stored CAMs of the current + (t-1) neighbour frames."""
    rng = np.random.default_rng(seed + 2)
    return rng.random(size=(b, t, 1, h, w), dtype=np.float32)
