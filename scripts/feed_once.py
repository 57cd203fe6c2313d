"""One pre-processing and one post-processing call on 25 600 synthetic ds2 showers (for an ncu capture)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit4hep_b200 import FusedForwardTransforms

dev = torch.device("cuda:0")
L, per = 45, 144
V, N = L * per, 25600
chain = {"NormalizeByElayer": {}, "ScaleTotalEnergy": {"n_layers": L, "factor": 0.35}, "CutValues": {"cut": 1.0e-7, "n_layers": L},
         "ExclusiveLogitTransform": {"delta": 1.0e-6, "rescale": True}, "GlobalStandardizeFromFile": {"model_dir": None, "eps": 1.0e-6},
         "LogEnergy": {}, "ScaleEnergy": {"e_min": 6.907755, "e_max": 13.815510}, "AddFeaturesToCond": {"split_index": V},
         "Reshape": {"shape": [1, V]}}
g = torch.Generator(device=dev).manual_seed(11)
raw = torch.exp(torch.randn(N, V, device=dev, generator=g) * 2 + 3) * (torch.rand(N, V, device=dev, generator=g) < 0.3)
e_inc = 10.0 ** (3 + 3 * torch.rand(N, 1, device=dev, generator=g))
raw = raw * (e_inc * 0.8 / raw.sum(1, keepdim=True))
fwd = FusedForwardTransforms(chain, range(0, V + 1, per))
x, cond = fwd(raw, e_inc)            # statistics pass: preprocess kernel + stats + standardise kernels
x, cond = fwd(raw, e_inc)            # one-kernel path
back, e = fwd.reverse()(x, cond)
torch.cuda.synchronize()
print("ok", float((back - raw).abs().max()))
