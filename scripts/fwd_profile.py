"""Per-kernel-class device time of forward-only network evaluations (the ODE sampler's inner call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import vit_oracle as vo
from vit4hep_b200 import ViT, _cabi
name = sys.argv[1] if len(sys.argv) > 1 else "ds2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
cfg = vo.CONFIGS[name]; param = dict(cfg["param"]); param["precision"] = "bf16"
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = ViT(param).to(dev)
with torch.no_grad():
    for p in net.parameters():
        p.copy_(torch.randn_like(p) * 0.02)
T, P = net.pos_z.numel(), param["patch_dim"]
x = torch.randn(B, T, P, device=dev); t = torch.rand(1, device=dev); c = torch.rand(B, param["condition_dim"], device=dev)
with torch.inference_mode():
    for _ in range(5):
        net(x, t, c, shared_t=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        net(x, t, c, shared_t=True)
    e1.record(); torch.cuda.synchronize()
    print(f"{name} B={B}: {e0.elapsed_time(e1) / 20:.3f} ms per evaluation (eager)")
    _cabi.profile_begin()
    for _ in range(3):
        net(x, t, c, shared_t=True)
    prof = _cabi.profile_end(128)
tot = sum(e["ms"] for e in prof)
for e in sorted(prof, key=lambda e: -e["ms"]):
    print(f"{e['name']:14s} n={e['launches'] / 3:5.1f} ms={e['ms'] / 3:.4f} avg_us={e['ms'] / max(e['launches'], 1) * 1e3:7.1f} "
          f"tf={e['flops'] / (e['ms'] * 1e-3) / 1e12 if e['ms'] > 0 else 0:6.0f} share={e['ms'] / tot:.3f}")
print("sum of classes", tot / 3, "ms")
