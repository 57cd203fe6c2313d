"""Activation workspace of the ViT plan per configuration (v4h_vit_workspace_bytes): what the saved activations cost."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit4hep_b200 import _cabi, configs

lib = _cabi.load()
torch.cuda.set_device(0)
for name, T in (("ds2", 135), ("ds3", 450)):
    net = configs.build(name, "bf16").to("cuda").net
    plan = net._plan(T)
    for B, save in ((64, 1), (64, 0), (256, 0)):
        print(f"{name} batch {B} {'training (activations saved)' if save else 'inference'}: "
              f"{lib.v4h_vit_workspace_bytes(plan, B, save) / 1e6:.1f} MB", flush=True)
