"""Randomised bit-exactness run of the seeding path: TCAMSeeder (rng_parity=True: tcam_seed_fused with the caller's
draws, or the two-kernel path) against the torch restatement of the reference (oracle/seeding.py) on the same random
stream -- labels identical, generator at the same position afterwards.  python tools/r2_fuzz_seed.py [seconds] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import seeding as ref
from tcam_wsol_video_b200.tcam_seeding import TCAMSeeder

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
t0 = time.time()
cases = fails = 0
while time.time() - t0 < budget:
    b = int(rng.integers(1, 6)); h = int(rng.integers(3, 90)); w = int(rng.integers(3, 90)); t = int(rng.integers(1, 5))
    min_ = int(rng.integers(0, 7)); max_ = int(rng.integers(0, 7))
    if min_ + max_ == 0:
        continue
    cfg = dict(seed_tech=str(rng.choice(["seed_weighted", "seed_uniform"])), min_=min_, max_=max_,
               max_p=float(rng.choice([0.05, 0.2, 0.6, 1.0])), min_p=float(rng.choice([0.05, 0.1, 0.3, 1.0])),
               ksz=int(rng.integers(1, 7)), use_roi=bool(rng.random() < 0.7))
    g = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
    low = torch.rand((b, t, 6, 7), generator=g)
    cams = torch.nn.functional.interpolate(low, size=(h, w), mode="bilinear", align_corners=False).cuda()
    if rng.random() < 0.3:
        cams = torch.round(cams * 6) / 6                     # many ties
    if rng.random() < 0.3:
        cams[int(rng.integers(0, b))] = 0.5                  # a flat sample
    cam = cams.amax(dim=1, keepdim=True)
    roi = (cam >= cam.flatten(1).median(dim=1).values.view(b, 1, 1, 1)).long()
    if rng.random() < 0.1:
        roi[int(rng.integers(0, b))] = 0                     # an empty roi
    mod = TCAMSeeder(fg_erode_k=11, fg_erode_iter=0, support_background=True, multi_label_flag=False, seg_ignore_idx=-255,
                     cuda_id=0, roi_method="roi_all", p_min_area_roi=0.05, **cfg)
    seed = int(rng.integers(1 << 30))
    torch.manual_seed(seed)
    got, _ = mod.forward_stack(cams, roi)
    a = torch.rand(3, device="cuda")
    torch.manual_seed(seed)
    want = ref.tcam_seeder_forward(cam, roi, seed_tech=cfg["seed_tech"], min_=min_, max_=max_, min_p=cfg["min_p"],
                                   max_p=cfg["max_p"], ksz=cfg["ksz"], ignore_idx=-255, use_roi=cfg["use_roi"])
    same_stream = torch.equal(a, torch.rand(3, device="cuda"))
    cases += 1
    if not torch.equal(got, want) or not same_stream:
        fails += 1
        print("FAIL", dict(b=b, h=h, w=w, t=t, **cfg), "labels differ" if same_stream else "stream position differs", flush=True)
print(f"seed fuzz: {cases} cases in {time.time() - t0:.0f} s, {fails} failures")
sys.exit(1 if fails else 0)
