"""CPU suite, part 3: the multi-GPU harness (tcam_wsol_video_b200/dist.py) on world_size-2 gloo.

The CUDA op cannot run here, so a small differentiable stand-in with the same contract as DenseCRFLoss
(1-element loss = -sum(S * A(S)) / N_local over its shard, frames independent) plays the local loss; what is
tested is the host logic: the shards cover the batch exactly once, the all-reduced loss equals the
single-process loss, and every rank's gradient equals its slice of the single-process gradient.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tcam_wsol_video_b200.dist import ShardedCRFLoss, all_reduce_scalar, shard_by_clip, shard_range


class StandInCRF(torch.nn.Module):
    """Same contract as DenseCRFLoss: per-frame bilinear form, divided by the local batch size."""

    def __init__(self, weight):
        super().__init__()
        self.weight = weight

    def forward(self, images, segmentations):
        n = segmentations.shape[0]
        gain = images.mean(dim=(1, 2, 3), keepdim=True) / 255.0              # per-frame, image dependent
        a_s = torch.roll(segmentations, 1, dims=3) * gain + segmentations     # a fixed linear "filter"
        return self.weight * (-(segmentations * a_s).sum().view(1) / n)


class StandInCRFWithBatchSize(StandInCRF):
    """The same stand-in with DenseCRFLoss's `batch_size=` extension (the loss divides by the global batch itself)."""
    accepts_batch_size = True

    def forward(self, images, segmentations, batch_size=None):
        n = segmentations.shape[0] if batch_size is None else batch_size
        return super().forward(images, segmentations) * (segmentations.shape[0] / n)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        images = torch.randint(0, 256, (n_total, 3, 8, 10), generator=g).float()
        segs = torch.softmax(torch.rand((n_total, 2, 8, 10), generator=g), dim=1)
        lo, hi = shard_range(n_total, rank, world)
        local_segs = segs[lo:hi].clone().requires_grad_(True)
        mod = ShardedCRFLoss(StandInCRF(weight=1e-3), reduction="global")
        loss = mod(images[lo:hi], local_segs)            # global batch found with an all-reduce
        loss.backward()
        loss2 = mod(images[lo:hi], local_segs.detach(), global_batch=n_total)
        local_only = ShardedCRFLoss(StandInCRF(weight=1e-3), reduction="local")(images[lo:hi], local_segs.detach())
        # "global_async": the forward returns this rank's share (same gradient), the reduced value arrives later
        seg_async = segs[lo:hi].clone().requires_grad_(True)
        mod_async = ShardedCRFLoss(StandInCRF(weight=1e-3), reduction="global_async")
        assert mod_async.global_loss() is None
        share = mod_async(images[lo:hi], seg_async, global_batch=n_total)
        share.backward()
        async_total = mod_async.global_loss().clone()
        # a local loss that takes the global batch itself (DenseCRFLoss does): no multiply outside, same numbers
        seg_bs = segs[lo:hi].clone().requires_grad_(True)
        mod_bs = ShardedCRFLoss(StandInCRFWithBatchSize(weight=1e-3), reduction="global")
        loss_bs = mod_bs(images[lo:hi], seg_bs, global_batch=n_total)
        loss_bs.backward()
        torch.save({"loss": loss.detach(), "loss2": loss2.detach(), "grad": local_segs.grad, "lo": lo, "hi": hi,
                    "local": local_only.detach(), "share": share.detach(), "async_total": async_total,
                    "async_grad": seg_async.grad, "loss_bs": loss_bs.detach(), "grad_bs": seg_bs.grad},
                   os.path.join(out_dir, f"rank{rank}.pt"))
        # all_reduce_scalar is the identity in the backward pass
        v = torch.tensor([float(rank + 1)], requires_grad=True)
        s = all_reduce_scalar(v * 2.0)
        s.backward()
        assert s.item() == 2.0 * sum(range(1, world + 1)) and v.grad.item() == 2.0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [6, 7])
def test_sharded_loss_matches_single_process(tmp_path, n_total):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    g = torch.Generator().manual_seed(0)
    images = torch.randint(0, 256, (n_total, 3, 8, 10), generator=g).float()
    segs = torch.softmax(torch.rand((n_total, 2, 8, 10), generator=g), dim=1).requires_grad_(True)
    ref = StandInCRF(weight=1e-3)(images, segs)
    ref.backward()
    covered = []
    for r in range(world):
        d = torch.load(os.path.join(str(tmp_path), f"rank{r}.pt"))
        assert torch.allclose(d["loss"], ref.detach(), rtol=1e-6)
        assert torch.allclose(d["loss2"], ref.detach(), rtol=1e-6)
        assert torch.allclose(d["grad"], segs.grad[d["lo"]:d["hi"]], rtol=1e-5, atol=1e-12)
        covered += list(range(d["lo"], d["hi"]))
        # the asynchronous reduction: same reduced value, same gradient, and the shares add up to the loss
        assert torch.allclose(d["async_total"], ref.detach(), rtol=1e-6)
        assert torch.equal(d["async_grad"], d["grad"])
        assert torch.allclose(d["loss_bs"], ref.detach(), rtol=1e-6)
        assert torch.allclose(d["grad_bs"], d["grad"], rtol=1e-6, atol=1e-12)
    assert covered == list(range(n_total))
    shares = [torch.load(os.path.join(str(tmp_path), f"rank{r}.pt"))["share"] for r in range(world)]
    assert torch.allclose(sum(shares), ref.detach(), rtol=1e-6)
    # "local" reduction is the reference's DDP convention: mean of the local means == global mean only for equal shards
    if n_total % world == 0:
        locs = [torch.load(os.path.join(str(tmp_path), f"rank{r}.pt"))["local"] for r in range(world)]
        assert torch.allclose(sum(locs) / world, ref.detach(), rtol=1e-6)


def test_shard_range_partitions_exactly():
    for n in (0, 1, 5, 32, 255, 256):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_shard_by_clip_never_splits_a_clip():
    seq = [5, 5, 5, 9, 9, 2, 2, 2, 2, 7, 1, 1]
    for world in (1, 2, 3, 4):
        owned = shard_by_clip(seq, world)
        assert sorted(i for o in owned for i in o) == list(range(len(seq)))
        for o in owned:
            for s in set(seq[i] for i in o):
                assert all(i in o for i, v in enumerate(seq) if v == s)
    # balanced: 256 frames of 64 clips of 4 frames over 8 ranks -> 32 frames each (BASELINE configs[4])
    seq = [c for c in range(64) for _ in range(4)]
    assert [len(o) for o in shard_by_clip(seq, 8)] == [32] * 8
