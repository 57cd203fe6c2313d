"""One small eager training step on the ds3 geometry (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import vit_oracle as vo
from vit4hep_b200 import CaloChallengeCFM, FusedAdamW, ViT
cfg = vo.CONFIGS["ds3"]; geom, param = cfg["geom"], dict(cfg["param"]); param["precision"] = "bf16"
if "--small" in sys.argv:
    param.update(hidden_dim=96, depth=2, num_heads=2)
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = ViT(param)
with torch.no_grad():
    for p in net.parameters():
        p.copy_(torch.randn_like(p) * 0.02)
seg = geom.segments[0]
model = CaloChallengeCFM(net, list(seg.patch), 1, "uniform", "linear", dict(method="rk4", options=dict(step_size=0.05)),
                         shape=list(seg.shape)).to(dev)
model.device, model.dtype = dev, torch.float32
opt = FusedAdamW(model.net, lr=1e-4, weight_decay=0.1, max_grad_norm=1000.0)
B = int(os.environ.get("B", "2"))
x = torch.randn(B, *geom.sample_shape, device=dev); c = torch.rand(B, param["condition_dim"], device=dev)
for i in range(2):
    opt.zero_grad(set_to_none=True)
    loss = model._batch_loss((x, c)); loss.backward(); opt.step()
    torch.cuda.synchronize(); print("step", i, loss.item(), flush=True)
