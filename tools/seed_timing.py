"""Where the time of TCAMSeeder.forward_stack goes (32 samples, T=5, 224x224): python tools/seed_timing.py"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tcam_wsol_video_b200 import synth, ops
from tcam_wsol_video_b200.tcam_seeding import TCAMSeeder
dev = torch.device("cuda", 0)
N, H, W = 32, 224, 224
low = torch.from_numpy(synth.make_low_res_cams(N, 5, 28, 28, seed=3)).squeeze(2)
cams = torch.nn.functional.interpolate(low, size=(H, W), mode="bilinear", align_corners=False).to(dev)
roi = (cams.amax(dim=1, keepdim=True) >= 0.5).long()
for parity in (True, False):
    seeder = TCAMSeeder(seed_tech="seed_weighted", min_=1, max_=1, max_p=0.6, min_p=0.1, fg_erode_k=11, fg_erode_iter=0,
                        ksz=3, support_background=True, multi_label_flag=False, seg_ignore_idx=-255, cuda_id=0,
                        roi_method="roi_all", p_min_area_roi=0.05, use_roi=True, rng_parity=parity)
    def timed(fn, reps=50):
        for _ in range(5): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps): fn()
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
    x_max = ops.temporal_cam_max(cams).unsqueeze(1)
    xm, r = seeder._prep(x_max, roi)
    counts, _ = seeder._candidate_counts(xm, r)
    print(f"rng_parity={parity}: forward_stack {timed(lambda: seeder.forward_stack(cams, roi)):.3f} ms | "
          f"temporal max {timed(lambda: ops.temporal_cam_max(cams)):.3f} | prep {timed(lambda: seeder._prep(x_max, roi)):.3f} | "
          f"counts (host sync) {timed(lambda: seeder._candidate_counts(xm, r)):.3f} | draws {timed(lambda: seeder._draws(counts, dev)):.3f} | "
          f"select+labels {timed(lambda: seeder._select(cams, r, counts)):.3f}")

# device time of the fused kernel alone (CUDA events)
seeder = TCAMSeeder(seed_tech="seed_weighted", min_=1, max_=1, max_p=0.6, min_p=0.1, fg_erode_k=11, fg_erode_iter=0,
                    ksz=3, support_background=True, multi_label_flag=False, seg_ignore_idx=-255, cuda_id=0,
                    roi_method="roi_all", p_min_area_roi=0.05, use_roi=True, rng_parity=False)
r = roi.long().contiguous()
for _ in range(5): seeder._select(cams, r, None)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(100): seeder._select(cams, r, None)
e1.record(); torch.cuda.synchronize()
print(f"_select (rng + fused kernel + allocations), device time: {e0.elapsed_time(e1) / 100:.4f} ms per call")
