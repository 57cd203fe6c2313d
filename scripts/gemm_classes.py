"""Every tcgen05 GEMM call site of one ds2 transformer block, 4 launches each, in a fixed order -- for one ncu pass
that labels launches by position (ncu --metrics sm__pipe_tensor_cycles_active...,gpu__time_duration.sum):

    python scripts/gemm_classes.py [M]            # prints the class order: launch index -> class
    python scripts/ncu_gemm_table.py launches.csv order.txt > profiles/r02_gemm_classes.txt
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit4hep_b200 import _cabi

M = int(sys.argv[1]) if len(sys.argv) > 1 else 8640
D, H, T, REP = 480, 1920, 135, 4
lib = _cabi.load()
dev = torch.device("cuda:0")
bf = torch.bfloat16
s = torch.cuda.current_stream().cuda_stream
# (class, kind, m, n, k); kind 7 = gated residual + LayerNorm epilogue (v4h_debug_gemm_ln)
CLASSES = [("gemm.qkv", 1, M, 3 * D, D), ("gemm.proj+ln", 7, M, D, D), ("gemm.fc1", 0, M, H, D), ("gemm.fc2+ln", 7, M, D, H),
           ("gemm.proj", 2, M, D, D), ("gemm.fc2", 2, M, D, H),
           ("dgrad.fc2", 3, M, H, D), ("dgrad.fc1", 4, M, D, H), ("dgrad.proj", 4, M, D, D), ("dgrad.qkv", 4, M, D, 3 * D),
           ("wgrad.fc2", 5, D, H, M), ("wgrad.fc1", 5, H, D, M), ("wgrad.proj", 5, D, D, M), ("wgrad.qkv", 5, 3 * D, D, M)]
g = torch.Generator().manual_seed(0)
for name, kind, m, n, k in CLASSES:
    A = (torch.randn((k, m) if kind == 5 else (m, k), generator=g) * 0.1).to(dev, bf)
    B = (torch.randn((n, k) if kind in (0, 1, 2, 7) else (k, n), generator=g) * 0.1).to(dev, bf)
    bias = torch.randn(n, generator=g).to(dev)
    out = torch.zeros((m, n), device=dev, dtype=torch.float32 if kind == 5 else bf)
    out2 = torch.zeros((m, n), device=dev, dtype=bf)
    gated = kind in (2, 7)
    res_in = torch.randn(m, n, generator=g).to(dev) if gated else None
    res_out = torch.empty(m, n, device=dev) if gated else None
    nb = (m + T - 1) // T
    gate = torch.randn(nb, n, generator=g).to(dev) if gated else None
    aux = torch.randn(m, n, generator=g).to(dev, bf) if kind == 3 else None
    ptr = lambda t: None if t is None else t.data_ptr()
    if kind == 7:
        shift, scale = torch.randn(nb, n, generator=g).to(dev), torch.randn(nb, n, generator=g).to(dev)
        ln = torch.zeros(m, n + 8, device=dev, dtype=bf)
        stats = torch.zeros(m, 2, device=dev)
    for _ in range(REP):
        if kind == 7:
            _cabi.check(lib.v4h_debug_gemm_ln(m, n, k, T, A.data_ptr(), B.data_ptr(), bias.data_ptr(), out2.data_ptr(),
                                              res_in.data_ptr(), res_out.data_ptr(), gate.data_ptr(), shift.data_ptr(),
                                              scale.data_ptr(), ln.data_ptr(), n + 8, stats.data_ptr(), None, s))
        else:
            _cabi.check(lib.v4h_debug_gemm(kind, m, n, k, T, A.data_ptr(), B.data_ptr(), bias.data_ptr(), out.data_ptr(),
                                           out2.data_ptr(), ptr(res_in), ptr(res_out), ptr(gate), ptr(aux), None, s))
    torch.cuda.synchronize()
    print(f"{name} {m} {n} {k} {REP}", flush=True)
