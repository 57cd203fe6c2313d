import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit4hep_b200 import _cabi
lib = _cabi.load(); dev = torch.device("cuda:0")
def run(m, n, k, reps=30):
    g = torch.Generator().manual_seed(0)
    A = torch.randn(m, k, generator=g).to(dev, torch.bfloat16)
    B = torch.randn(n, k, generator=g).to(dev, torch.bfloat16)
    bias = torch.randn(n, generator=g).to(dev)
    C = torch.full((m, n), float("nan"), device=dev)
    s = torch.cuda.current_stream().cuda_stream
    try:
        for _ in range(reps):
            _cabi.check(lib.v4h_debug_gemm(6, m, n, k, 1, A.data_ptr(), B.data_ptr(), bias.data_ptr(), C.data_ptr(), None, None, None, None, None, None, s))
            torch.cuda.synchronize()
        want = A.double() @ B.double().t() + bias.double()
        print(f"ok   {m}x{n}x{k} rel {((C.double() - want).norm() / want.norm()).item():.2e}", flush=True)
    except Exception as e:
        print(f"FAIL {m}x{n}x{k}: {str(e)[-100:]}", flush=True); sys.exit(1)
for shape in [(14400, 90, 480), (20000, 96, 480), (28800, 48, 480), (28800, 90, 480)]:
    run(*shape)
