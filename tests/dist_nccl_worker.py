"""Worker of tests/test_gpu_dist.py: one rank of a torchrun job (NCCL, one process per GPU).

BASELINE configs[4]: the CRF loss over 256 frames sharded across the ranks with `dist.ShardedCRFLoss`; the only
exchange is the scalar loss.  Every rank checks its shard of the gradient, and the reduced loss, against the SAME
256 frames run on one GPU in this process (frames are independent, dense_crf_loss.py:56-74; under DDP the reference
divides by the local batch, :64).  Prints one JSON line per rank."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from tcam_wsol_video_b200 import synth  # noqa: E402
from tcam_wsol_video_b200.dense_crf_loss import DenseCRFLoss  # noqa: E402
from tcam_wsol_video_b200.dist import ShardedCRFLoss, shard_range  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    k, h, w = 2, 224, 224
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    try:
        img = torch.from_numpy(synth.make_images(n_total, h, w, "natural", seed=5))
        seg = torch.from_numpy(synth.make_segs(n_total, k, h, w, seed=5))
        lo, hi = shard_range(n_total, rank, world)
        weight = 2e-9
        out = {"rank": rank, "world": world, "lo": lo, "hi": hi}
        # single GPU, all frames: the reference result for this rank to compare its shard with
        full = seg.to(dev).requires_grad_(True)
        ref = DenseCRFLoss(weight, 15.0, 100.0, 1.0)(images=img.to(dev), segmentations=full)
        ref.backward()
        for reduction in ("global", "global_async", "local"):
            if reduction == "local" and n_total % world:
                continue        # mean of local means equals the global mean only for equal shards
            mine = seg[lo:hi].to(dev).requires_grad_(True)
            mod = ShardedCRFLoss(DenseCRFLoss(weight, 15.0, 100.0, 1.0), reduction=reduction)
            loss = mod(img[lo:hi].to(dev), mine, global_batch=n_total)
            loss.backward()
            if reduction == "global_async":
                loss = mod.global_loss()
            elif reduction == "local":
                # DDP convention: mean over ranks of the local means (equal shards) == global mean
                t = loss.detach().clone()
                dist.all_reduce(t)
                loss = t / world
                mine.grad.mul_((hi - lo) / float(n_total))     # DDP would average the gradients the same way
            want_g = full.grad[lo:hi]
            g_rel = float((mine.grad - want_g).abs().max() / want_g.abs().max())
            l_rel = float((loss.detach() - ref.detach()).abs() / ref.detach().abs())
            out[reduction] = {"loss_rel": l_rel, "grad_rel": g_rel}
            assert l_rel < 1e-5 and g_rel < 1e-5, (reduction, l_rel, g_rel)
        print("NCCL_WORKER " + json.dumps(out), flush=True)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
