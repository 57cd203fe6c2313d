"""Device time of the small GEMMs of the conditioning path, measured in a replayed CUDA graph (no host gaps),
with the per-role cycle counters of one launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit4hep_b200 import _cabi
lib = _cabi.load(); dev = torch.device("cuda:0")
NAMES = ["p.wait_empty", "p.issue", "m.wait_acc", "m.wait_full", "m.issue", "e.wait_acc", "e.ld", "e.wait_in", "e.math",
         "e.bar", "e.copy", "e.tail"]
def run(name, kind, m, n, k, T=135):
    bf = torch.bfloat16
    g = torch.Generator().manual_seed(0)
    A = (torch.randn((k, m) if kind == 5 else (m, k), generator=g) * 0.1).to(dev, bf)
    B = (torch.randn((n, k) if kind <= 2 or kind == 6 else (k, n), generator=g) * 0.1).to(dev, bf)
    bias = torch.randn(n, generator=g).to(dev)
    out = torch.zeros((m, n), device=dev, dtype=torch.float32 if kind in (5, 6) else bf)
    out2 = torch.zeros((m, n), device=dev, dtype=bf)
    aux = torch.randn(m, n, generator=g).to(dev, bf) if kind == 3 else None
    cnt = torch.zeros(16, dtype=torch.int64, device=dev)
    ptr = lambda t: None if t is None else t.data_ptr()
    def call(c, s):
        _cabi.check(lib.v4h_debug_gemm(kind, m, n, k, T, A.data_ptr(), B.data_ptr(), bias.data_ptr(), out.data_ptr(),
                                       out2.data_ptr(), None, None, None, ptr(aux), c, s))
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3): call(None, side.cuda_stream)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(20): call(None, torch.cuda.current_stream().cuda_stream)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    call(cnt.data_ptr(), torch.cuda.current_stream().cuda_stream); torch.cuda.synchronize()
    c = cnt.cpu().tolist()
    print(f"{name:12s} {m}x{n}x{k}: {us:6.2f} us/launch in a graph | raw cycles: " + " ".join(f"{nm}={v}" for nm, v in zip(NAMES, c)), flush=True)
run("tiny", 1, 128, 256, 64)
run("cond.fwd", 1, 64, 480, 480)
run("cond.fwd.f32", 6, 64, 480, 480)
run("cond.act", 0, 64, 480, 256)
run("cond.dgrad", 4, 64, 480, 480)
run("cond.dact", 3, 64, 480, 480)
run("cond.wgrad", 5, 480, 480, 64)
run("final", 6, 8640, 48, 480)
run("qkv", 1, 8640, 1440, 480)
