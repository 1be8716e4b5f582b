"""Is the natural K=2 step bound by the CPU (python + launches) or by the GPU?  Eager, enqueue-only and CUDA-graph
timings of DenseCRFLoss fwd+bwd on device-resident inputs: python tools/r2_cpu_overhead.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tcam_wsol_video_b200 import synth
from tcam_wsol_video_b200.dense_crf_loss import DenseCRFLoss
dev = torch.device("cuda", 0)
for kind, k in (("natural", 2), ("natural", 10), ("noise", 2), ("noise", 10)):
    n, h, w = 32, 224, 224
    img = torch.from_numpy(synth.make_images(n, h, w, kind, seed=1)).to(dev)
    seg = torch.from_numpy(synth.make_segs(n, k, h, w, seed=1)).to(dev).requires_grad_(True)
    crf = DenseCRFLoss(2e-9, 15.0, 100.0, 1.0)
    def step():
        seg.grad = None
        crf(images=img, segmentations=seg).backward()
    for _ in range(20): step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(300): step()
    t_enq = (time.perf_counter() - t0) / 300
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / 300
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3): step()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    seg.grad = None
    with torch.cuda.graph(g, stream=side):
        step()
    for _ in range(10): g.replay()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(300): g.replay()
    torch.cuda.synchronize()
    t_graph = (time.perf_counter() - t0) / 300
    print(f"{kind} K={k}: eager {1e3*t_all:.4f} ms/step (CPU enqueue alone {1e3*t_enq:.4f}), CUDA graph replay {1e3*t_graph:.4f} ms/step")
