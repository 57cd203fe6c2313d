"""Shows the error GraphedTrainStep raises when the autograd graph of an earlier eager step is still alive at capture
time (see the class docstring): python scripts/capture_rule_demo.py"""
import sys, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vit4hep_b200 as v4
from oracle import vit_oracle as vo
dev = torch.device("cuda:0")
cfg = vo.tiny_config("ds2", hidden_dim=96, depth=2, num_heads=2)
param = dict(cfg["param"]); param["precision"] = "bf16"
model = v4.CaloChallengeCFM(v4.ViT(param), [3, 16, 1], 1, "uniform", "linear", dict(method="rk4", options=dict(step_size=0.25)), shape=[45, 16, 9]).to(dev)
model.device, model.dtype = dev, torch.float32
opt = v4.FusedAdamW(model.net, lr=1e-3)
x, c = torch.randn(4, 1, 45, 16, 9, device=dev), torch.rand(4, 46, device=dev)
loss = model._batch_loss((x, c)); loss.backward(); opt.step()      # eager step, loss kept alive
try:
    v4.GraphedTrainStep(model, opt, x, c, warmup=1)
    print("captured (no stale graph effect)")
except RuntimeError as e:
    print("OK message:", str(e)[:120])
