"""ds3 sampling loop (debug aid): repeat eager sample_batch calls at batch sizes 32 and 64."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import vit_oracle as vo
from vit4hep_b200 import CaloChallengeCFM, ViT
cfg = vo.CONFIGS[os.environ.get("CFG", "ds3")]; geom, param = cfg["geom"], dict(cfg["param"]); param["precision"] = "bf16"
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
for rep in range(int(os.environ.get("REPS", "3"))):
    for B in (32, 64):
        c = torch.rand(B, param["condition_dim"], device=dev)
        out = model.sample_batch(c); torch.cuda.synchronize()
        print("rep", rep, "B", B, float(out.abs().mean()), flush=True)
