"""Temporal-CAM aggregation around the seeder: the loader/trainer-side pieces of SURVEY.md §8 rows a14-a15.

* frame pickers           dlib/datasets/wsol_loader.py:448-458 (_get_lef_knn / _get_right_knn), :544-569
* re_normalize_cam        dlib/datasets/wsol_loader.py:630-635
* temporal max            dlib/datasets/wsol_loader.py:585-600
* prepare_std_cams_disq   dlib/learning/train_wsol.py:417-432

The reference does the first three per sample inside a DataLoader worker on CPU tensors (one `.pt` load per frame);
here the per-frame low-resolution CAMs of a mini-batch are stacked on the GPU and reduced in one launch.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from . import ops

# dlib/configure/constants.py: sl_tc_knn_mode
TIME_INSTANT = 'instant'
TIME_BEFORE = 'before'
TIME_AFTER = 'after'
TIME_BEFORE_AFTER = 'before_and_after'


def get_left_knn(lframes: Sequence[str], frame: str, k: int) -> List[str]:
    """The k frames before `frame` in its shot (wsol_loader.py:448-452)."""
    assert frame in lframes
    idx = list(lframes).index(frame)
    return list(lframes[max(0, idx - k): idx])


def get_right_knn(lframes: Sequence[str], frame: str, k: int) -> List[str]:
    """The k frames after `frame` in its shot (wsol_loader.py:454-459).  Reference quirk kept: for the LAST frame
    of a shot the slice is lframes[n-1:n], i.e. the frame itself (harmless: the max is idempotent)."""
    assert frame in lframes
    idx = list(lframes).index(frame)
    n = len(lframes)
    return list(lframes[min(idx + 1, n - 1): min(idx + k + 1, n)])


def temporal_frames(lframes: Sequence[str], frame: str, k: int, mode: str) -> List[str]:
    """left + [frame] + right, by sl_tc_knn_mode (wsol_loader.py:544-557)."""
    left, right = [], []
    if mode in (TIME_BEFORE, TIME_BEFORE_AFTER):
        left = get_left_knn(lframes, frame, k)
    if mode in (TIME_AFTER, TIME_BEFORE_AFTER):
        right = get_right_knn(lframes, frame, k)
    return left + [frame] + right


def re_normalize_cam(cam: torch.Tensor, h: float) -> torch.Tensor:
    """exp((cam + 1e-6) * h) / max, nan_to_num(0, 1, 0) for ONE frame's CAM [1,h',w'] (wsol_loader.py:630-635)."""
    return ops.temporal_cam_max(cam.reshape(1, 1, -1).float().contiguous(), renorm_h=h).reshape(cam.shape)


def aggregate_temporal_cams(stack: torch.Tensor, knn_t: float = 0.0) -> torch.Tensor:
    """std_cam of every sample of a mini-batch from the CAMs of its temporal frames.

    stack [B,T,1,h',w'] or [B,T,h',w'] (CUDA float32; pad a shorter neighbourhood by repeating the sample's own
    frame: max is idempotent) -> [B,1,h',w'].  knn_t > 0 applies re_normalize_cam to every frame first, as the
    loader does when sl_tc_knn > 0 and sl_tc_knn_t > 0 (wsol_loader.py:591-600)."""
    if stack.ndim == 5:
        assert stack.shape[2] == 1
        stack = stack[:, :, 0]
    assert stack.ndim == 4
    return ops.temporal_cam_max(stack.float().contiguous(), renorm_h=knn_t).unsqueeze(1)


def prepare_std_cams_disq(std_cams: torch.Tensor, image_size: Tuple[int, int]) -> torch.Tensor:
    """(bsz,1,h',w') -> (bsz,1,H,W): nan_to_num, bilinear (align_corners=False), nan_to_num (train_wsol.py:417-432)."""
    assert std_cams.ndim == 4
    return ops.prepare_std_cams(std_cams.detach(), image_size)


__all__ = ["get_left_knn", "get_right_knn", "temporal_frames", "re_normalize_cam", "aggregate_temporal_cams",
           "prepare_std_cams_disq", "TIME_INSTANT", "TIME_BEFORE", "TIME_AFTER", "TIME_BEFORE_AFTER"]
