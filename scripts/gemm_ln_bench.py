"""Time the gated-residual GEMM with the LayerNorm epilogue (v4h_debug_gemm_ln) on the ds2 shapes and print where
the first epilogue thread of every CTA spends its cycles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit4hep_b200 import _cabi

lib = _cabi.load()
dev = torch.device("cuda:0")
T = 135
NAMES = ["wait_acc", "p1.ld", "p1.wait_box", "p1.math", "p1.store", "stats", "p2.math", "p2.store"]


def run(name, m, n, k, train=True, iters=20):
    bf = torch.bfloat16
    g = torch.Generator().manual_seed(0)
    A = (torch.randn(m, k, generator=g) * 0.1).to(dev, bf)
    W = (torch.randn(n, k, generator=g) * 0.1).to(dev, bf)
    bias = torch.randn(n, generator=g).to(dev)
    nb = (m + T - 1) // T
    res = torch.randn(m, n, generator=g).to(dev)
    gate, shift, scale = (torch.randn(nb, n, generator=g).to(dev) for _ in range(3))
    y = torch.zeros(m, n, device=dev, dtype=bf) if train else None
    res_out = torch.empty(m, n, device=dev)
    ln = torch.zeros(m, n + 8, device=dev, dtype=bf)
    stats = torch.zeros(m, 2, device=dev) if train else None
    cnt = torch.zeros(16, dtype=torch.int64, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    ptr = lambda t: None if t is None else t.data_ptr()

    def call(c):
        _cabi.check(lib.v4h_debug_gemm_ln(m, n, k, T, A.data_ptr(), W.data_ptr(), bias.data_ptr(), ptr(y), res.data_ptr(),
                                          res_out.data_ptr(), gate.data_ptr(), shift.data_ptr(), scale.data_ptr(),
                                          ln.data_ptr(), n + 8, ptr(stats), c, s))
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
    ctas = max(1, (m + 127) // 128 * (2 if m < 18944 else 1))
    per = " ".join(f"{nm}={v / ctas / 1e3:.1f}k" for nm, v in zip(NAMES, c))
    print(f"{name:14s} {m}x{n}x{k} train={int(train)}: {us:7.1f} us {2.0 * m * n * k / us / 1e6:7.1f} TF | cycles/CTA: {per}", flush=True)


if __name__ == "__main__":
    print("cluster:", os.environ.get("V4H_GEMM_LN_CLUSTER", "auto"))
    run("proj b64", 8640, 480, 480)
    run("fc2 b64", 8640, 480, 1920)
    run("proj b256", 34560, 480, 480, train=False)
    run("fc2 b256", 34560, 480, 1920, train=False)
