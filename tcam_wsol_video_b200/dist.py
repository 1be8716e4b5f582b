"""Multi-GPU harness of the CRF loss: frames are independent, so the batch is sharded over ranks and the
only exchange is one scalar.

One process per GPU (torchrun); ``torch.distributed`` is the plumbing.  Two conventions:

* ``"local"``  -- what the reference does under DDP: every rank divides by its LOCAL batch size
  (dlib/crf/dense_crf_loss.py:64) and DDP's gradient averaging turns that into the global mean.  No
  collective at all in the loss.
* ``"global"`` -- every rank divides by the GLOBAL batch size and the scalar losses are summed with one
  all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests).  ``loss.backward()`` on each rank then yields
  the gradient of the global mean for its own shard; nothing else crosses the link.

* ``"global_async"`` -- the same global mean, with the collective taken OFF the critical path: ``forward`` returns
  this rank's differentiable share (its gradient is already the gradient of the global mean: d(sum)/d(share) = 1),
  the 4-byte all-reduce is queued on a side stream behind the forward kernels, and ``global_loss()`` hands the
  reduced scalar out later (typically one step late, for logging).  Nothing on the compute stream ever waits for
  the network.

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
        if reduction not in ("local", "global", "global_async"):
            raise ValueError(reduction)
        self.local_loss = local_loss
        self.reduction = reduction
        self.group = group
        self._side = None        # side stream of the asynchronous reduction (CUDA only)
        self._pending = None     # (reduced tensor, event or work handle) of the last forward

    def _global_batch(self, n_local: int, device) -> int:
        n = torch.tensor([float(n_local)], device=device)
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(n, group=self.group)
        return int(n.item())

    def forward(self, images: torch.Tensor, segmentations: torch.Tensor, global_batch: Optional[int] = None):
        """images/segmentations: this rank's shard.  Returns the loss: the global mean for ``"global"``, this rank's
        share of it for ``"global_async"`` (same gradient; the reduced value comes from ``global_loss()``), the local
        mean for ``"local"``.  Pass ``global_batch`` to skip the all-reduce that counts the frames."""
        if self.reduction == "local":
            return self.local_loss(images=images, segmentations=segmentations)
        n_local = segmentations.shape[0]
        if global_batch is None:
            global_batch = self._global_batch(n_local, segmentations.device)
        if getattr(self.local_loss, "accepts_batch_size", False):
            # the loss kernels divide by the global batch directly: no extra multiply in forward or backward
            share = self.local_loss(images=images, segmentations=segmentations, batch_size=int(global_batch))
        else:
            # local = -sum_local / n_local  ->  this rank's share of the global mean = local * n_local / N
            local = self.local_loss(images=images, segmentations=segmentations)
            share = local * (float(n_local) / float(global_batch))
        if self.reduction == "global":
            return all_reduce_scalar(share, self.group)
        self._pending = self._reduce_async(share.detach())
        return share

    def _reduce_async(self, value: torch.Tensor):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return value.clone(), None
        if value.is_cuda:
            if self._side is None:
                self._side = torch.cuda.Stream(device=value.device)
            main = torch.cuda.current_stream(value.device)
            total = value.clone()                       # on the compute stream, behind the forward kernels
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                dist.all_reduce(total, op=dist.ReduceOp.SUM, group=self.group)
                done = torch.cuda.Event()
                done.record(self._side)
            total.record_stream(self._side)
            return total, done
        total = value.clone()
        return total, dist.all_reduce(total, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def global_loss(self) -> Optional[torch.Tensor]:
        """The all-reduced loss of the last ``"global_async"`` forward (None before the first).  The CURRENT stream
        waits for the side stream here -- call it where the value is needed (logging), not inside the step."""
        if self._pending is None:
            return None
        total, handle = self._pending
        if handle is not None:
            if isinstance(handle, torch.cuda.Event):
                torch.cuda.current_stream(total.device).wait_event(handle)
            else:
                handle.wait()
        return total
