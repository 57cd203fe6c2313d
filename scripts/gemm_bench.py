"""Time the tcgen05 GEMM with the model's epilogues on the ds2 shapes and print where each warp role
spends its cycles (v4h_debug_gemm counters)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit4hep_b200 import _cabi

lib = _cabi.load()
dev = torch.device("cuda:0")
M, D, H, T = int(os.environ.get("GEMM_M", 8640)), 480, 1920, 135
NAMES = ["p.wait_empty", "p.issue", "m.wait_acc", "m.wait_full", "m.issue", "e.wait_acc", "e.ld", "e.wait_in", "e.math",
         "e.bar", "e.copy", "e.tail"]


def run(name, kind, m, n, k, iters=20):
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
    cnt = torch.zeros(16, dtype=torch.int64, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    ptr = lambda t: None if t is None else t.data_ptr()

    def call(c):
        _cabi.check(lib.v4h_debug_gemm(kind, m, n, k, T, A.data_ptr(), B.data_ptr(), bias.data_ptr(), out.data_ptr(),
                                       out2.data_ptr(), ptr(res_in), ptr(res_out), ptr(gate), ptr(aux), c, s))
    for _ in range(3):
        call(None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call(None)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    call(cnt.data_ptr()); torch.cuda.synchronize()
    c = cnt.cpu().tolist()
    tf = 2.0 * m * n * k / us / 1e6
    per = " ".join(f"{nm}={v / 148 / 1e3:.1f}k" for nm, v in zip(NAMES, c))
    per += f" clk={c[12]/148/1e3:.1f}k" if len(c) > 12 and c[12] else ""
    print(f"{name:10s} {m}x{n}x{k}: {us:7.1f} us {tf:7.1f} TF | cycles/CTA: {per}", flush=True)


if __name__ == "__main__":
    print("pairs:", os.environ.get("V4H_GEMM_PAIRS", "1"))
    if "--small" in sys.argv:
        run("tiny", 1, 128, 256, 64)
        run("tiny.k480", 1, 128, 256, 480)
        run("cond.fwd", 1, 64, 480, 480)
        run("cond.dgrad", 4, 64, 480, 480)
        run("cond.wgrad", 5, 480, 480, 64)
        run("adaln.fwd", 1, 64, 18240, 480)
        run("adaln.dgrad", 4, 64, 480, 18240)
        run("final", 1, M, 48, 480)
        sys.exit(0)
    run("fc1", 0, M, H, D)
    run("qkv", 1, M, 3 * D, D)
    run("proj", 2, M, D, D)
    run("fc2", 2, M, D, H)
    run("dgrad.fc2", 3, M, H, D)
    run("dgrad.fc1", 4, M, D, H)
    run("dgrad.qkv", 4, M, D, 3 * D)
    run("wgrad.fc2", 5, D, H, M)
    run("wgrad.fc1", 5, H, D, M)
    run("wgrad.qkv", 5, 3 * D, D, M)
