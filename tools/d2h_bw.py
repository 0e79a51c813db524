"""Host-link bandwidth available to N ranks of one box AT THE SAME TIME: every rank copies a 49 MB block (the hidden states
of 32 x 10 s) device->host and a 20 MB block host->device, pinned memory, 20 repetitions, all ranks between barriers.
    python -m torch.distributed.run --nproc-per-node N tools/d2h_bw.py"""
import os, time, torch
import torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
d = torch.empty(32 * 499 * 768, dtype=torch.float32, device="cuda")
h = torch.empty(32 * 499 * 768, dtype=torch.float32, pin_memory=True)
u = torch.empty(32 * 160000, dtype=torch.float32, pin_memory=True)
ud = torch.empty(32 * 160000, dtype=torch.float32, device="cuda")
def bw(fn, nbytes, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9
a = bw(lambda: h.copy_(d, non_blocking=True), d.numel() * 4)
b = bw(lambda: ud.copy_(u, non_blocking=True), u.numel() * 4)
t = torch.tensor([a, b], dtype=torch.float64, device="cuda")
if world > 1:
    s = t.clone(); dist.all_reduce(s)
    m = t.clone(); dist.all_reduce(m, op=dist.ReduceOp.MIN)
else:
    s = m = t
if rank == 0:
    print(f"{world} ranks at once: D2H {s[0].item():.1f} GB/s aggregate ({m[0].item():.1f} slowest rank), H2D {s[1].item():.1f} GB/s aggregate ({m[1].item():.1f} slowest rank)")
if world > 1:
    dist.destroy_process_group()
