"""ds3 forward-only calls at a given batch under V4H_LAUNCH_SYNC (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import vit_oracle as vo
from vit4hep_b200 import ViT
cfg = vo.CONFIGS["ds3"]; param = dict(cfg["param"]); param["precision"] = "bf16"
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = ViT(param).to(dev)
with torch.no_grad():
    for p in net.parameters():
        p.copy_(torch.randn_like(p) * 0.02)
T, P = net.pos_z.numel(), param["patch_dim"]
for B in [int(v) for v in sys.argv[1:]]:
    x = torch.randn(B, T, P, device=dev); t = torch.rand(1, device=dev); c = torch.rand(B, param["condition_dim"], device=dev)
    with torch.inference_mode():
        for shared in (True, False):
            tt = t if shared else torch.rand(B, 1, device=dev)
            out = net(x, tt, c, shared_t=shared); torch.cuda.synchronize()
            print("B", B, "shared_t", shared, float(out.abs().mean()), flush=True)
