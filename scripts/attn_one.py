import sys, os
sys.path.insert(0, "/root/repo")
import torch
from vit4hep_b200 import _cabi
lib = _cabi.load(); dev = torch.device("cuda:0")
B,T,H,dh = [int(v) for v in sys.argv[1:5]]
g = torch.Generator().manual_seed(3)
qkv = torch.randn(B, T, 3, H, dh, generator=g).to(dev).to(torch.bfloat16)
d_o = torch.randn(B, T, H, dh, generator=g).to(dev).to(torch.bfloat16)
o = torch.zeros(B, T, H, dh, device=dev, dtype=torch.bfloat16); lse = torch.zeros(B, H, T, device=dev)
dqkv = torch.zeros_like(qkv); s = torch.cuda.current_stream().cuda_stream
_cabi.check(lib.v4h_test_attention_fwd(1, 1, qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), B, T, H, dh, s)); torch.cuda.synchronize(); print("fwd ok", flush=True)
_cabi.check(lib.v4h_test_attention_bwd(1, 1, qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), d_o.data_ptr(), dqkv.data_ptr(), B, T, H, dh, s)); torch.cuda.synchronize(); print("bwd ok", flush=True)
