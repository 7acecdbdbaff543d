"""Per-rank host->device / device->host bandwidth of the e2e leg's buffers (134 MB pinned uint16 scene, 16.8 MB class
maps), alone and with all ranks copying at once.  Launch under torchrun with N ranks (or plain for one)."""
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
host = torch.empty((4, 4096, 4096), dtype=torch.int16).pin_memory()
devbuf = torch.empty_like(host, device=dev)
cls_dev = torch.empty((64, 512, 512), dtype=torch.uint8, device=dev)
cls_host = torch.empty((64, 512, 512), dtype=torch.uint8).pin_memory()


def bw(fn, nbytes, reps=5):
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return ms, nbytes / ms / 1e6


t0 = time.time()
h2d = bw(lambda: devbuf.copy_(host, non_blocking=True), host.numel() * 2)
d2h = bw(lambda: cls_host.copy_(cls_dev, non_blocking=True), cls_dev.numel())
print(f"rank {rank}/{world}: H2D 134 MB {h2d[0]:.2f} ms = {h2d[1]:.1f} GB/s | D2H 16.8 MB {d2h[0]:.2f} ms = {d2h[1]:.1f} GB/s "
      f"| cpu affinity {len(os.sched_getaffinity(0))} cores", flush=True)
if world > 1:
    dist.destroy_process_group()
