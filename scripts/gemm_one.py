"""A few launches of ONE tcgen05 GEMM call site (kinds of v4h_debug_gemm, see scripts/gemm_bench.py) for
`ncu --set full --import-source on -k regex:gemm_umma`: python scripts/gemm_one.py <kind> <m> <n> <k> [launches]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit4hep_b200 import _cabi

kind, m, n, k = (int(v) for v in sys.argv[1:5])
launches = int(sys.argv[5]) if len(sys.argv) > 5 else 4
T = 135
lib = _cabi.load()
dev = torch.device("cuda:0")
bf = torch.bfloat16
g = torch.Generator().manual_seed(0)
A = (torch.randn((k, m) if kind == 5 else (m, k), generator=g) * 0.1).to(dev, bf)
B = (torch.randn((n, k) if kind <= 2 else (k, n), generator=g) * 0.1).to(dev, bf)
bias = torch.randn(n, generator=g).to(dev)
out = torch.zeros((m, n), device=dev, dtype=torch.float32 if kind == 5 else bf)
out2 = torch.zeros((m, n), device=dev, dtype=bf)
res_in = torch.randn(m, n, generator=g).to(dev) if kind == 2 else None
res_out = torch.empty(m, n, device=dev) if kind == 2 else None
gate = torch.randn((m + T - 1) // T, n, generator=g).to(dev) if kind == 2 else None
aux = torch.randn(m, n, generator=g).to(dev, bf) if kind == 3 else None
s = torch.cuda.current_stream().cuda_stream
ptr = lambda t: None if t is None else t.data_ptr()
for _ in range(launches):
    _cabi.check(lib.v4h_debug_gemm(kind, m, n, k, T, A.data_ptr(), B.data_ptr(), bias.data_ptr(), out.data_ptr(),
                                   out2.data_ptr(), ptr(res_in), ptr(res_out), ptr(gate), ptr(aux), None, s))
torch.cuda.synchronize()
print("ok", kind, m, n, k)
