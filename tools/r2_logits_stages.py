"""Stage times of the CRF loss on stored probabilities vs from logits (fused softmax), natural frames K=2:
python tools/r2_logits_stages.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tcam_wsol_video_b200 import _lib, synth
from tcam_wsol_video_b200.dense_crf_loss import DenseCRFLoss, DenseCRFLossFromLogits
lib = _lib.load()
dev = torch.device("cuda", 0)
N, K, H, W = 32, 2, 224, 224
img8 = torch.from_numpy(synth.make_images(N, H, W, "natural", seed=3).astype(np.uint8)).to(dev)
logits = torch.randn((N, K, H, W), device=dev, requires_grad=True)
probs = torch.softmax(logits.detach(), dim=1).requires_grad_(True)
a = DenseCRFLoss(2e-9, 15.0, 100.0, 1.0)
b = DenseCRFLossFromLogits(2e-9, 15.0, 100.0, 1.0)
def run(name, fn, steps=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    total = e0.elapsed_time(e1) / steps
    lib.tcamcrf_profile_enable(1); lib.tcamcrf_profile_read(None, None, 1)
    for _ in range(steps): fn()
    torch.cuda.synchronize(); lib.tcamcrf_profile_enable(0)
    ms = (ctypes.c_double * len(_lib.STAGES))(); ln = (ctypes.c_longlong * len(_lib.STAGES))()
    lib.tcamcrf_profile_read(ms, ln, 1)
    print(name, f"{total:.4f} ms/step", {s: round(ms[i] / steps, 4) for i, s in enumerate(_lib.STAGES) if ln[i]})
def f_probs():
    probs.grad = None; a(images=img8, segmentations=probs).backward()
def f_logits():
    logits.grad = None; b(images=img8, logits=logits).backward()
def f_softmax_then_probs():
    logits.grad = None; a(images=img8, segmentations=torch.softmax(logits, dim=1)).backward()
run("stored probabilities      ", f_probs)
run("from logits (fused)       ", f_logits)
run("torch softmax + probs path", f_softmax_then_probs)
