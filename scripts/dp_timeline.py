"""Device timeline of one data-parallel training step: where the bucketed NCCL all-reduces sit relative to the
backward kernels (torch.profiler / CUPTI kernel records of rank 0).  The step is the CUDA graph bench.py replays
(GraphedTrainStep), so the launches are not host-paced and the ranks do not drift apart under the profiler;
`--eager` profiles eager launches instead.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/dp_timeline.py > profiles/r02_dp_timeline_nN.txt
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NCCL_DEBUG", "WARN")
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile
from vit4hep_b200 import FusedAdamW, GraphedTrainStep, configs, dp

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
model = configs.build("ds2", "bf16").to(dev)
model.device, model.dtype = dev, torch.float32
with torch.no_grad():
    for n, p in model.net.named_parameters():
        if "adaLN" in n or n.endswith("bias") or n.startswith("final_layer.linear"):
            p.normal_(0, 0.02)
dp.enable_data_parallel(model.net)
opt = FusedAdamW(model.net, lr=1e-4, weight_decay=0.1, max_grad_norm=1000.0)
g = torch.Generator().manual_seed(1 + rank)
x = torch.randn(64, 1, 45, 16, 9, generator=g).to(dev)
c = torch.rand(64, 46, generator=g).to(dev)


def step():
    loss = model._batch_loss((x, c))
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()


for _ in range(5):
    step()
if "--eager" not in sys.argv:
    graphed = GraphedTrainStep(model, opt, x, c)
    step = lambda: graphed.step(x, c)
    for _ in range(5):
        step()
torch.cuda.synchronize(); dist.barrier()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(4):  # the first profiled steps absorb the ranks' profiler start-up skew; the last one is printed
        step()
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs = [e for e in evs if not e.name.startswith(("nccl:", "Optimizer.", "Memcpy", "Memset"))]
    evs.sort(key=lambda e: e.time_range.start)
    ends = [i for i, e in enumerate(evs) if "adamw_kernel" in e.name]
    evs = evs[ends[-2] + 1: ends[-1] + 1]       # the last step: after the previous step's optimizer kernel
    t0 = evs[0].time_range.start
    print(f"# one {'eager' if '--eager' in sys.argv else 'graph-replayed'} data-parallel training step on rank 0 of {world} (ds2, batch 64 per GPU); times in us from the first kernel")
    print(f"# {'start':>9s} {'dur':>8s}  kernel")
    nccl_busy, total_end = 0.0, 0.0
    for e in evs:
        name = e.name
        st, du = e.time_range.start - t0, e.time_range.end - e.time_range.start
        total_end = max(total_end, st + du)
        is_nccl = "nccl" in name.lower()
        if is_nccl:
            nccl_busy += du
        if is_nccl or du >= 8.0:
            print(f"  {st:9.1f} {du:8.1f}  {'>>> ' if is_nccl else ''}{name[:100]}")
    print(f"# step span {total_end:.1f} us, NCCL kernels busy {nccl_busy:.1f} us (overlapped with the kernels listed around them)")
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
