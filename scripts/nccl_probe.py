"""All-reduce latency of the data-parallel bucket sizes (fp32, in place) under the current NCCL_* environment.
torchrun --nproc-per-node N scripts/nccl_probe.py"""
import os, sys
import torch
import torch.distributed as dist

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
out = []
for mb in (2.5, 16.6):
    n = int(mb * 1e6 / 4)
    t = torch.ones(n, device="cuda")
    for _ in range(5):
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
    e1.record(); torch.cuda.synchronize()
    out.append(f"{mb} MB: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us")
if rank == 0:
    env = {k: v for k, v in os.environ.items() if k.startswith("NCCL_") and k != "NCCL_DEBUG"}
    print(env, " | ".join(out), flush=True)
dist.barrier()
os._exit(0)
