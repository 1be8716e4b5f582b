"""Device time of the parts of the TCAM loss step (32 clips, natural frames, K=2), each captured in a CUDA graph:
python tools/r2_tcam_parts.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tcam_wsol_video_b200 import synth
from tcam_wsol_video_b200.dense_crf_loss import DenseCRFLossFromLogits
from tcam_wsol_video_b200.tcam_seeding import TCAMSeeder
dev = torch.device("cuda", 0)
N, H, W = 32, 224, 224
low = torch.from_numpy(synth.make_low_res_cams(N, 5, 28, 28, seed=3)).squeeze(2)
cams = torch.nn.functional.interpolate(low, size=(H, W), mode="bilinear", align_corners=False).to(dev)
roi = (cams.amax(dim=1, keepdim=True) >= 0.5).long()
img8 = torch.from_numpy(synth.make_images(N, H, W, "natural", seed=3).astype(np.uint8)).to(dev)
logits = torch.randn((N, 2, H, W), device=dev, requires_grad=True)
seeder = TCAMSeeder(seed_tech="seed_weighted", min_=1, max_=1, max_p=0.6, min_p=0.1, fg_erode_k=11, fg_erode_iter=0, ksz=3,
                    support_background=True, multi_label_flag=False, seg_ignore_idx=-255, cuda_id=0, roi_method="roi_all",
                    p_min_area_roi=0.05, use_roi=True, rng_parity=False)
crf = DenseCRFLossFromLogits(2e-9, 15.0, 100.0, 1.0)
seeds0, _ = seeder.forward_stack(cams, roi)
def graph_time(fn, reps=200):
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    logits.grad = None
    with torch.cuda.graph(g, stream=side):
        fn()
    for _ in range(10): g.replay()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): g.replay()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
def f_seed(): seeder.forward_stack(cams, roi)
def f_crf():
    logits.grad = None; crf(images=img8, logits=logits).backward()
def f_ce():
    logits.grad = None; torch.nn.functional.cross_entropy(logits, seeds0, ignore_index=-255).backward()
def f_both():
    logits.grad = None
    (crf(images=img8, logits=logits) + torch.nn.functional.cross_entropy(logits, seeds0, ignore_index=-255)).backward()
def f_all():
    logits.grad = None
    s, _ = seeder.forward_stack(cams, roi)
    (crf(images=img8, logits=logits) + torch.nn.functional.cross_entropy(logits, s, ignore_index=-255)).backward()
for name, fn in (("seeding", f_seed), ("CRF from logits fwd+bwd", f_crf), ("cross-entropy fwd+bwd", f_ce), ("CRF + CE", f_both), ("whole step", f_all)):
    print(f"{name}: {graph_time(fn):.4f} ms")
