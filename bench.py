#!/usr/bin/env python
"""bench.py -- CRF-loss fwd+bwd frames/s at 224^2 on N B200s (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA path
    python bench.py --impl reference [...]                         # the reference's CPU path

A step is one fwd+bwd of DenseCRFLoss over one batch of synthetic frames per GPU
(configs[1] of BASELINE.json: 32 frames x 10 classes x 224x224, sigma_rgb=15, sigma_xy=100):
lattice build + splat + blur + slice + loss reduction (forward) and the gradient kernel (backward).
For N > 1 every rank runs the same per-GPU batch (weak scaling; 8 GPUs = configs[4]'s 256 frames)
and the scalar loss is all-reduced over NCCL inside the timed region.

One JSON line is printed by rank 0; see DESIGN.md "Measurement" for every key.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "crf_loss_fwd_bwd_frames_per_sec_224"
UNIT = "frames/s"
H = W = 224
SIGMA_RGB, SIGMA_XY = 15.0, 100.0
D = 5


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--frames", type=int, default=32, help="frames per GPU per step")
    ap.add_argument("--classes", type=int, default=10)
    ap.add_argument("--kind", choices=["noise", "natural"], default="noise")
    ap.add_argument("--rotate", type=int, default=4, help="distinct input batches cycled through (defeats L2 reuse)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the short runs on the other input regimes")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 -> min(steps, 20)")
    return ap.parse_args()


# ---------------------------------------------------------------------------
# algorithmic bytes (restated in DESIGN.md "Roofline accounting")
# ---------------------------------------------------------------------------
def stage_bytes(P: int, d: int, K: int, M: float):
    """Algorithmic bytes per FRAME of each pipeline stage; they sum to SURVEY §8(d)'s B_frame."""
    dp1 = d + 1
    return {
        "build": 12 * P + 8 * dp1 * P,             # image in; offset + barycentric out
        "neighbour": 8 * dp1 * M,                  # neighbour table out
        "splat": 4 * K * P + 8 * dp1 * P + 4 * K * M,   # seg in; offset + bary in; value table out
        "blur": dp1 * (8 * M + 16 * K * M),        # per axis: neighbour ids in, 3 values in, 1 value out
        "slice": 8 * dp1 * P + 4 * K * M + 4 * K * P,   # offset + bary in; value table in; AS out
        "loss": 0,
        "backward": 4 * K * P,                     # gradient out
        "prepare": 0,                              # table clear: not algorithmic traffic
        "seed": 0,
    }


def frame_bytes(P, d, K, M):
    return float(sum(stage_bytes(P, d, K, M).values()))


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, device_index: int):
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        self._h = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self._nvml = pynvml
            try:
                uuid = str(torch.cuda.get_device_properties(device_index).uuid)
                if not uuid.startswith("GPU-"):
                    uuid = "GPU-" + uuid
                self._h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
            except Exception:
                self._h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._h = None

    def _loop(self):
        nv = self._nvml
        while not self._stop.is_set():
            try:
                mhz = int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                self.samples.append(mhz)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if os.environ.get("TCAMCRF_BENCH_NOCLOCK") == "1":
            return
        if self._h is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------
# CPU reference timing (oracle/_ref when built, else the C port)
# ---------------------------------------------------------------------------
def cpu_fwd_bwd_frames_per_sec(frames: int, K: int, kind: str, reps: int, warmup: int = 1):
    """Times the reference's CPU implementation of the same fwd+bwd on `frames` frames of the workload."""
    import oracle
    from tcam_wsol_video_b200 import synth

    fn, which = oracle.best_filter(color=False)
    img = synth.make_images(frames, H, W, kind, seed=0)
    seg = synth.make_segs(frames, K, H, W, seed=0)
    cores = len(os.sched_getaffinity(0))
    if which == "reference":
        # torchrun exports OMP_NUM_THREADS=1; give the reference every host core it is allowed to use
        # (its own code then takes min(max_threads, N), bilateralfilter.cpp:45-47)
        oracle.ref_set_threads(cores)
        threads = min(oracle.load_ref()[0].ref_omp_max_threads(), frames)
    else:
        threads = 1
    times = []
    for i in range(warmup + reps):
        t0 = time.perf_counter()
        oracle.densecrf_loss_fwd_bwd(img, seg, SIGMA_RGB, SIGMA_XY, 1.0, fn)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return frames / min(times), frames / (sum(times) / len(times)), which, threads, cores, times


def cpu_context_runs(K: int, kind: str):
    """SURVEY 8d's two other CPU arms, as context next to the all-cores number: (i) one thread, (iii) all cores with
    glibc malloc kept from returning the per-class buffers to the OS (the reference allocates and frees two
    [H*W] float arrays per class, bilateralfilter.cpp:31-37).  Small samples; (iii) needs the environment set
    before the process starts, so it runs in a child."""
    out = {}
    try:
        import oracle
        if not oracle.have_ref():
            return out
        cores = len(os.sched_getaffinity(0))
        fn, _ = oracle.best_filter(color=False)
        from tcam_wsol_video_b200 import synth
        img = synth.make_images(2, H, W, kind, seed=0)
        seg = synth.make_segs(2, K, H, W, seed=0)
        oracle.ref_set_threads(1)
        best = 1e30
        for _ in range(2):
            t0 = time.perf_counter()
            oracle.densecrf_loss_fwd_bwd(img, seg, SIGMA_RGB, SIGMA_XY, 1.0, fn)
            best = min(best, time.perf_counter() - t0)
        oracle.ref_set_threads(cores)
        out["one_thread"] = {"value": 2 / best, "unit": UNIT, "cores": 1, "sample": "2 frames, best of 2"}
        env = dict(os.environ, MALLOC_MMAP_THRESHOLD_="1073741824", MALLOC_TRIM_THRESHOLD_="1073741824",
                   MALLOC_TOP_PAD_="268435456", OMP_NUM_THREADS=str(cores))
        code = ("import sys, time; sys.path.insert(0, %r); import bench\n"
                "b, m, which, threads, cores, t = bench.cpu_fwd_bwd_frames_per_sec(%d, %d, %r, reps=2, warmup=1)\n"
                "print('MALLOC_TUNED', b, threads)" % (ROOT, min(32, max(cores, 2)), K, kind))
        res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
        for ln in res.stdout.splitlines():
            if ln.startswith("MALLOC_TUNED"):
                _, v, th = ln.split()
                out["malloc_tuned"] = {"value": float(v), "unit": UNIT, "cores": int(th),
                                       "sample": f"{min(32, max(cores, 2))} frames, best of 2 after 1 warm-up, "
                                                 "MALLOC_MMAP_THRESHOLD_/TRIM_THRESHOLD_=1 GiB, TOP_PAD_=256 MiB"}
    except Exception as exc:   # context only: never fail the bench line over it
        out["error"] = repr(exc)[:200]
    return out


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    frames = args.frames
    t0 = time.perf_counter()
    # bounded sample: keep the whole run within ~150 s whatever --steps is (calibrate on a few frames)
    cal = min(frames, max(2, len(os.sched_getaffinity(0))))
    _, cal_fps, *_ = cpu_fwd_bwd_frames_per_sec(cal, args.classes, args.kind, reps=1, warmup=0)
    budget_frames = int(150.0 * cal_fps / max(args.steps + args.warmup, 1))
    frames = max(1, min(frames, budget_frames))
    best, mean, which, threads, cores, times = cpu_fwd_bwd_frames_per_sec(frames, args.classes, args.kind,
                                                                          reps=max(args.steps, 1),
                                                                          warmup=max(args.warmup, 0))
    value = frames * len(times) / sum(times)
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": which, "host_cores": cores,
                         "sample": f"{frames} frames x K={args.classes} x {H}x{W} ({args.kind}), fwd+bwd, "
                                   f"{len(times)} steps, OpenMP over frames as shipped"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "wall_s": time.perf_counter() - t0,
    }
    emit(line)


def workload_config(args, world):
    return {
        "workload": f"BASELINE configs[1]: DenseCRFLoss fwd+bwd, {args.frames} frames/GPU x {args.classes} classes x "
                    f"{H}x{W} RGB, sigma_rgb={SIGMA_RGB:g}, sigma_xy={SIGMA_XY:g}, 5-D lattice, '{args.kind}' images",
        "frames_per_gpu": args.frames, "classes": args.classes, "height": H, "width": W,
        "image_kind": args.kind, "global_frames": args.frames * world,
        "parallelism": f"dp{world} (frames sharded, scalar loss all-reduce)" if world > 1 else "single GPU",
        "l2": f"{args.rotate} distinct input batches rotated; per-step working set (segs+AS+grad+tables) > 126 MB L2",
    }


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------
def run_ours(args, rank: int, local_rank: int, world: int):
    import torch
    import torch.distributed as dist

    from tcam_wsol_video_b200 import _lib, ops, synth
    from tcam_wsol_video_b200.dense_crf_loss import DenseCRFLoss

    assert torch.cuda.is_available(), "bench.py (impl=ours) needs a GPU; there is no CPU path"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = _lib.load()
    assert lib.tcamcrf_device_count() >= 1

    N, K = args.frames, args.classes
    P = H * W
    sets = []
    for r in range(max(args.rotate, 1)):
        seed = 1000 * rank + r
        img = torch.from_numpy(synth.make_images(N, H, W, args.kind, seed=seed))
        seg = torch.from_numpy(synth.make_segs(N, K, H, W, seed=seed))
        sets.append((img.pin_memory(), seg.pin_memory(), img.to(dev), seg.to(dev).requires_grad_(True)))
    crf = DenseCRFLoss(weight=2e-9, sigma_rgb=SIGMA_RGB, sigma_xy=SIGMA_XY, scale_factor=1.0).to(dev)

    def step(i):
        _, _, img_d, seg_d = sets[i % len(sets)]
        seg_d.grad = None
        loss = crf(images=img_d, segmentations=seg_d)
        loss.backward()
        if world > 1:
            lv = loss.detach()
            dist.all_reduce(lv)     # the path's only exchange: one scalar over NVLink
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # lattice size of this input (for the algorithmic-bytes formula)
    cfg = _lib.make_config(_lib.FEAT_XY_RGB, 3, SIGMA_RGB, SIGMA_XY)
    _, _, ws = ops.crf_forward(sets[0][2], sets[0][3].detach(), cfg, check=True)
    _, m_total = ops.workspace_status(ws)
    M = m_total / N

    for i in range(args.warmup):
        step(i)
    barrier()

    sampler = ClockSampler(local_rank)
    launches0 = lib.tcamcrf_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    barrier()
    ev0.record()
    for i in range(args.steps):
        step(i)
    ev1.record()
    barrier()
    sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    launches = lib.tcamcrf_launch_count() - launches0

    # per-stage times: the same steps again with CUDA events around every stage (tcamcrf_profile_*).  Kept out of
    # the timed region above because an event record between two kernels stops the next kernel's blocks from
    # becoming resident while the previous one drains (programmatic dependent launch).
    lib.tcamcrf_profile_enable(1)
    lib.tcamcrf_profile_read(None, None, 1)
    for i in range(args.steps):
        step(i)
    barrier()
    lib.tcamcrf_profile_enable(0)
    st_ms = (ctypes.c_double * len(_lib.STAGES))()
    st_ln = (ctypes.c_longlong * len(_lib.STAGES))()
    lib.tcamcrf_profile_read(st_ms, st_ln, 1)

    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * N * args.steps / (ms_total / 1e3)

    # ---- end to end through the host-pointer C ABI (pinned host buffers, copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        e2e_steps = args.e2e_steps or min(args.steps, 20)
        loss_h = torch.zeros(1).pin_memory()
        grad_h = torch.empty(N, K, H, W).pin_memory()
        cfg_h = _lib.make_config(_lib.FEAT_XY_RGB, 3, SIGMA_RGB, SIGMA_XY)

        def e2e_step(i):
            img_h, seg_h, _, _ = sets[i % len(sets)]
            rc = lib.tcamcrf_loss_fwd_bwd_host(ctypes.byref(cfg_h), img_h.data_ptr(), seg_h.data_ptr(),
                                               loss_h.data_ptr(), grad_h.data_ptr(), N, K, H, W, 2e-9)
            _lib.check(rc, "tcamcrf_loss_fwd_bwd_host")

        for i in range(3):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            e2e_step(i)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        e2e = {"value": world * N * e2e_steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(4 * N * 3 * P + 4 * N * K * P + 4),
               "d2h_bytes_per_step": int(4 * N * K * P + 4 + 4),
               "steps": e2e_steps, "ms_per_step": 1e3 * dt / e2e_steps,
               "api": "tcamcrf_loss_fwd_bwd_host (host pointers in, loss + gradient out)"}

        # context, not the headline: the call the reference's trainer makes (dlib/losses/tcam.py:113-115) -- frames
        # from pinned host memory (train_wsol.py:1128 keeps raw_img on the CPU), segmentations where the network
        # left them (device), loss read back; the gradient stays on the device for the backbone's backward pass
        def module_step(i):
            img_h, _, _, seg_d = sets[i % len(sets)]
            seg_d.grad = None
            loss = crf(images=img_h, segmentations=seg_d)
            loss.backward()
            return loss.item()

        for i in range(3):
            module_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            module_step(i)
        torch.cuda.synchronize()
        dtm = time.perf_counter() - t0
        t = torch.tensor([dtm], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dtm = float(t.item())
        e2e["trainer_call"] = {"value": world * N * e2e_steps / dtm, "unit": UNIT, "ms_per_step": 1e3 * dtm / e2e_steps,
                               "h2d_bytes_per_step": int(4 * N * 3 * P), "d2h_bytes_per_step": 4,
                               "api": "DenseCRFLoss(images=pinned host float32, segmentations=device).backward(); "
                                      "loss.item()"}

        # the same call with uint8 frames from the loader: a quarter of the bytes on the wire
        img8_h = [(s[0].to(torch.uint8)).pin_memory() for s in sets]

        def module_step_u8(i):
            seg_d = sets[i % len(sets)][3]
            seg_d.grad = None
            loss = crf(images=img8_h[i % len(sets)], segmentations=seg_d)
            loss.backward()
            return loss.item()

        for i in range(3):
            module_step_u8(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            module_step_u8(i)
        torch.cuda.synchronize()
        dtm = time.perf_counter() - t0
        t = torch.tensor([dtm], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dtm = float(t.item())
        e2e["trainer_call_u8"] = {"value": world * N * e2e_steps / dtm, "unit": UNIT, "ms_per_step": 1e3 * dtm / e2e_steps,
                                  "h2d_bytes_per_step": int(N * 3 * P), "d2h_bytes_per_step": 4,
                                  "api": "DenseCRFLoss(images=pinned host uint8, segmentations=device).backward(); "
                                         "loss.item()"}

    # ---- the same step on the other input regimes (short, device-resident; context for the headline number)
    extra = {}
    if world == 1 and not args.no_extra:
        for kind, k in (("natural", K), ("noise", 2), ("natural", 2)):
            if kind == args.kind and k == K:
                continue
            img = torch.from_numpy(synth.make_images(N, H, W, kind, seed=7)).to(dev)
            seg = torch.from_numpy(synth.make_segs(N, k, H, W, seed=7)).to(dev).requires_grad_(True)
            for _ in range(5):
                seg.grad = None
                crf(images=img, segmentations=seg).backward()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(30):
                seg.grad = None
                crf(images=img, segmentations=seg).backward()
            e1.record()
            torch.cuda.synchronize()
            extra[f"{kind}_k{k}"] = {"value": N * 30 / (e0.elapsed_time(e1) / 1e3), "unit": UNIT, "steps": 30}
            del img, seg

        # BASELINE configs[2] without the backbone: per step, for 32 clips -- temporal max over the current + 4
        # previous frames' CAMs fused with fg/bg seeding, CRF loss from logits (K=2, uint8 frames resident on the
        # GPU, fused softmax) + cross-entropy on the seeds, backward to the logits.
        try:
            from tcam_wsol_video_b200.dense_crf_loss import DenseCRFLossFromLogits
            from tcam_wsol_video_b200.tcam_seeding import TCAMSeeder
            low = torch.from_numpy(synth.make_low_res_cams(N, 5, 28, 28, seed=3)).squeeze(2)
            cams = torch.nn.functional.interpolate(low, size=(H, W), mode="bilinear", align_corners=False).to(dev)
            roi = (cams.amax(dim=1, keepdim=True) >= 0.5).long()
            img8 = torch.from_numpy(synth.make_images(N, H, W, "natural", seed=3).astype(np.uint8)).to(dev)
            logits = torch.randn((N, 2, H, W), device=dev, requires_grad=True)
            seeder = TCAMSeeder(seed_tech="seed_weighted", min_=1, max_=1, max_p=0.6, min_p=0.1, fg_erode_k=11,
                                fg_erode_iter=0, ksz=3, support_background=True, multi_label_flag=False,
                                seg_ignore_idx=-255, cuda_id=local_rank, roi_method="roi_all", p_min_area_roi=0.05,
                                use_roi=True, rng_parity=False)
            crf2 = DenseCRFLossFromLogits(2e-9, SIGMA_RGB, SIGMA_XY, 1.0)

            def tcam_step():
                logits.grad = None
                seeds, _ = seeder.forward_stack(cams, roi)
                loss = crf2(images=img8, logits=logits) + torch.nn.functional.cross_entropy(logits, seeds,
                                                                                             ignore_index=-255)
                loss.backward()

            for _ in range(5):
                tcam_step()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(30):
                tcam_step()
            e1.record()
            torch.cuda.synchronize()
            extra["tcam_seed_crf_step_natural_k2"] = {
                "value": N * 30 / (e0.elapsed_time(e1) / 1e3), "unit": UNIT, "steps": 30,
                "what": "temporal max (T=5) + seeding + CRF-from-logits + CE on seeds, fwd+bwd, no backbone"}
        except Exception as exc:  # context only; never fail the headline line
            extra["tcam_seed_crf_step_natural_k2"] = {"error": repr(exc)[:200]}

        # BASELINE configs[3]: lattice-size stress at 448x448 -- the 5-D colour lattice (x, y, r, g, b) and the 3-D
        # grayscale one (x, y, gray), 8 frames, K=2, noise frames (largest lattices), fwd+bwd like the headline.
        try:
            n4, k4, s4 = 8, 2, 448
            img4 = torch.from_numpy(synth.make_images(n4, s4, s4, "noise", seed=11)).to(dev)
            seg4 = torch.from_numpy(synth.make_segs(n4, k4, s4, s4, seed=11)).to(dev)
            for name, feat_channels in (("xyrgb_5d", 3), ("xygray_3d", 1)):
                cfg4 = _lib.make_config(_lib.FEAT_XY_RGB, feat_channels, SIGMA_RGB, SIGMA_XY)
                g_one = torch.ones(1, device=dev)

                def step4():
                    as4, _, _ = ops.crf_forward(img4, seg4, cfg4, want_loss=True, n_norm=float(n4))
                    ops.crf_backward(as4, g_one, float(n4))

                for _ in range(5):
                    step4()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(30):
                    step4()
                e1.record()
                torch.cuda.synchronize()
                _, ws4 = None, ops.crf_forward(img4, seg4, cfg4, check=True)[2]
                _, m4 = ops.workspace_status(ws4)
                extra[f"noise_448_{name}_k2"] = {"value": n4 * 30 / (e0.elapsed_time(e1) / 1e3), "unit": UNIT,
                                                 "steps": 30, "frames": n4, "vertices_per_frame": m4 / n4}
            del img4, seg4
        except Exception as exc:
            extra["noise_448"] = {"error": repr(exc)[:200]}

    if rank != 0:
        return

    # ---- roofline of the dominant kernel
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    sb = stage_bytes(P, D, K, M)
    stages = {}
    for i, name in enumerate(_lib.STAGES):
        if st_ln[i] > 0:
            per_launch_ms = st_ms[i] / st_ln[i]
            launches_per_step = st_ln[i] / args.steps
            bytes_per_launch = sb[name] * N / launches_per_step
            stages[name] = {"ms_per_step": st_ms[i] / args.steps, "launches_per_step": launches_per_step,
                            "ms_per_launch": per_launch_ms,
                            "gbs": bytes_per_launch / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else None}
    dom = max(stages, key=lambda s: stages[s]["ms_per_step"]) if stages else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if dom and os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(f"{dom}:{args.kind}:K{K}:N{N}")
        except Exception:
            traffic = None
    roofline = None
    if dom:
        roofline = {"bound": "hbm", "kernel": dom, "achieved": stages[dom]["gbs"], "peak": peak, "unit": "GB/s",
                    "frac": stages[dom]["gbs"] / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": sb[dom] * N / stages[dom]["launches_per_step"],
                    "ms_per_launch": stages[dom]["ms_per_launch"],
                    "pipeline": {"bytes_per_frame": frame_bytes(P, D, K, M), "vertices_per_frame": M,
                                 "achieved": frame_bytes(P, D, K, M) * value / world / 1e9,
                                 "frac": frame_bytes(P, D, K, M) * value / world / 1e9 / peak},
                    "stages": stages}
        if traffic and stages[dom]["ms_per_launch"] > 0:
            # the same launch against the DRAM bytes ncu measured for it: what the HBM really carried
            dram_gbs = traffic / (stages[dom]["ms_per_launch"] * 1e-3) / 1e9
            roofline["dram"] = {"achieved": dram_gbs, "frac": dram_gbs / peak, "unit": "GB/s",
                                "what": "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch (`traffic`) over "
                                        "the live per-launch time"}
        if roofline["frac"] > 1.0:
            roofline["note"] = ("frac > 1: SURVEY 8d's algorithmic bytes of this stage count the neighbour-row gathers, "
                                "which the L2 serves; `traffic` is the DRAM traffic per launch (ncu), "
                                "`pipeline.frac` the whole step against the HBM peak")

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        frames = min(N, 32)
        best, mean, which, threads, cores, times = cpu_fwd_bwd_frames_per_sec(frames, K, args.kind, reps=3, warmup=1)
        cpu = {"value": best, "unit": UNIT, "cores": threads, "kind": which, "host_cores": cores,
               "sample": f"{frames} frames x K={K} x {H}x{W} ({args.kind}), fwd+bwd, best of 3 after 1 warm-up, "
                         f"OpenMP over frames as shipped", "mean_value": mean,
               "context": cpu_context_runs(K, args.kind)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": sampler.summary(), "other_inputs": extra,
    }
    emit(line)


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The one JSON line goes to the real stdout; everything else (NCCL banners, warnings) to stderr."""
    text = json.dumps(line) + "\n"
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, text.encode())
    else:
        sys.stdout.write(text)
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)   # libraries that print to fd 1 (e.g. "NCCL version ...") must not pollute the JSON line
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, local_rank, world)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
