"""NCCL all-reduce / reduce timing of the two payloads the path exchanges (77 MB prompt-gradient buffer, 128 MB vote
canvas), with the transport NCCL picked (run with NCCL_DEBUG=INFO and grep 'via').  torchrun --nproc-per-node N."""
import os
import sys

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
for name, nbytes, op in [("prompt-gradient all-reduce (AVG)", 32 * 3 * 448 * 448 * 4, "allreduce"),
                         ("vote-canvas reduce (SUM, int32)", 8000 * 4000 * 4, "reduce")]:
    buf = torch.zeros(nbytes // 4, dtype=torch.float32 if op == "allreduce" else torch.int32, device=dev)
    for _ in range(3):
        if op == "allreduce":
            dist.all_reduce(buf, op=dist.ReduceOp.AVG)
        else:
            dist.reduce(buf, dst=0, op=dist.ReduceOp.SUM)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        if op == "allreduce":
            dist.all_reduce(buf, op=dist.ReduceOp.AVG)
        else:
            dist.reduce(buf, dst=0, op=dist.ReduceOp.SUM)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    if rank == 0:
        print(f"{world} GPUs: {name}: {nbytes / 1e6:.1f} MB in {ms:.3f} ms = {nbytes / ms / 1e6:.1f} GB/s (payload / time)")
dist.destroy_process_group()
