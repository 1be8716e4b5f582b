"""Pinned host<->device copy bandwidth of this box (context for the e2e numbers): python tools/pcie_bw.py
Under torchrun every rank measures its own GPU at the same time (the host ceiling with N GPUs busy)."""
import os
import torch
rank, world = int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(rank)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
n = 64 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_event(e0); s2.wait_event(e0)
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); torch.cuda.synchronize()
    return n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
run(True, True)
if world > 1:
    dist.barrier()
a, b, c = run(True, False), run(False, True), run(True, True)
print(f"rank {rank}/{world}: h2d alone {a:.1f} GB/s, d2h alone {b:.1f} GB/s, both at once {c:.1f} GB/s each", flush=True)
if world > 1:
    t = torch.tensor([a, b, c], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print(f"all {world} ranks at once, summed: h2d {t[0].item():.0f} GB/s, d2h {t[1].item():.0f} GB/s, "
              f"both ways {t[2].item():.0f} GB/s each way", flush=True)
    dist.destroy_process_group()
