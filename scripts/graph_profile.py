"""Per-kernel-class device time INSIDE the replayed CUDA graph of the training step: the profiling scopes
record their events during capture (event-record nodes), so the elapsed times are device-side, free of the
host submission gaps that inflate the eager per-launch numbers of small kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from vit4hep_b200 import CaloChallengeCFM, FusedAdamW, GraphedTrainStep, ViT, _cabi

cfg = sys.argv[1] if len(sys.argv) > 1 else "ds2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda:0")
geom, param = bench.ds2_setup(cfg)
param["precision"] = "bf16"
torch.manual_seed(0)
net = ViT(param)
bench.rerandomise(net)
seg = geom.segments[0]
model = CaloChallengeCFM(net, list(seg.patch), 1, "uniform", "linear", dict(method="rk4", options=dict(step_size=0.05)),
                         shape=list(seg.shape)).to(dev)
model.device, model.dtype = dev, torch.float32
opt = FusedAdamW(model.net, lr=1e-4, weight_decay=0.1, max_grad_norm=1000.0)
x = torch.randn(B, *geom.sample_shape, device=dev)
c = torch.rand(B, param["condition_dim"], device=dev)
g0 = GraphedTrainStep(model, opt, x, c)  # warm-up + plain graph (timing reference)
for _ in range(5):
    g0.step(x, c)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    g0.step(x, c)
e1.record(); torch.cuda.synchronize()
print(f"plain graph: {e0.elapsed_time(e1) / 20:.3f} ms per step")
_cabi.profile_begin()
g1 = GraphedTrainStep(model, opt, x, c, warmup=0)  # captured with the profiling events as graph nodes
for _ in range(3):
    g1.step(x, c)
torch.cuda.synchronize()
prof = _cabi.profile_end(128)
tot = sum(e["ms"] for e in prof)
for e in sorted(prof, key=lambda e: -e["ms"]):
    print(f"{e['name']:14s} n={e['launches']:3d} ms={e['ms']:.4f} avg_us={e['ms'] / max(e['launches'], 1) * 1e3:7.1f} "
          f"tf={e['flops'] / (e['ms'] * 1e-3) / 1e12 if e['ms'] > 0 else 0:6.0f} share={e['ms'] / tot:.3f}")
print("sum of profiled classes", round(tot, 4), "ms")
