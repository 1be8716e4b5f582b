"""Multi-GPU harness of the CRF loss: frames are independent, so the batch is sharded over ranks and the
only exchange is one scalar.

One process per GPU (torchrun); ``torch.distributed`` is the plumbing.  Two conventions:

* ``"local"``  -- what the reference does under DDP: every rank divides by its LOCAL batch size
  (dlib/crf/dense_crf_loss.py:64) and DDP's gradient averaging turns that into the global mean.  No
  collective at all in the loss.
* ``"global"`` -- every rank divides by the GLOBAL batch size and the scalar losses are summed with one
  all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests).  ``loss.backward()`` on each rank then yields
  the gradient of the global mean for its own shard; nothing else crosses the link.

``RgbJointConRanFieldTcams`` couples the frames of one clip, so batches are sharded by clip
(``shard_by_clip``), never inside a clip.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

__all__ = ['shard_range', 'shard_by_clip', 'ShardedCRFLoss', 'all_reduce_scalar']


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of the batch axis owned by `rank`; sizes differ by at most one frame."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(int(n_total), world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_by_clip(seq_iter: Sequence[int], world: int) -> List[List[int]]:
    """Assigns whole clips to ranks (greedy, largest clip first, to the least loaded rank).
    Returns, per rank, the batch indices it owns (clips kept together, original order inside a clip)."""
    clips = {}
    for i, s in enumerate(list(seq_iter)):
        clips.setdefault(int(s), []).append(i)
    order = sorted(clips.items(), key=lambda kv: (-len(kv[1]), kv[0]))
    owned: List[List[int]] = [[] for _ in range(world)]
    load = [0] * world
    for _, idx in order:
        r = min(range(world), key=lambda k: (load[k], k))
        owned[r].extend(idx)
        load[r] += len(idx)
    return [sorted(o) for o in owned]


def all_reduce_scalar(value: torch.Tensor, group=None) -> torch.Tensor:
    """Sum of a 1-element tensor over the ranks, differentiable as the identity for the local term
    (d(sum)/d(local) = 1): the backward pass needs no collective."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return value
    total = value.detach().clone()
    dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return value + (total - value.detach())


class ShardedCRFLoss(torch.nn.Module):
    """Wraps a per-shard CRF loss module (``DenseCRFLoss`` / ``ColorDenseCRFLoss``; anything mapping
    (images, segmentations) -> 1-element loss that divides by its local batch size)."""

    def __init__(self, local_loss: torch.nn.Module, reduction: str = "global", group=None):
        super().__init__()
        if reduction not in ("local", "global"):
            raise ValueError(reduction)
        self.local_loss = local_loss
        self.reduction = reduction
        self.group = group

    def forward(self, images: torch.Tensor, segmentations: torch.Tensor, global_batch: Optional[int] = None):
        """images/segmentations: this rank's shard.  Returns the loss (global mean for ``"global"``)."""
        local = self.local_loss(images=images, segmentations=segmentations)
        if self.reduction == "local":
            return local
        n_local = segmentations.shape[0]
        if global_batch is None:
            n = torch.tensor([float(n_local)], device=segmentations.device)
            if dist.is_available() and dist.is_initialized():
                dist.all_reduce(n, group=self.group)
            global_batch = int(n.item())
        # local = -sum_local / n_local  ->  this rank's share of the global mean = local * n_local / N
        share = local * (float(n_local) / float(global_batch))
        return all_reduce_scalar(share, self.group)
