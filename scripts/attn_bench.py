"""Time the tcgen05 attention forward / backward (v4h_test_attention_*) on one shape and check them against
torch's fp32 attention.  python scripts/attn_bench.py [B T H dh]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit4hep_b200 import _cabi

lib = _cabi.load()
dev = torch.device("cuda:0")
B, T, H, dh = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (64, 450, 6, 80)
g = torch.Generator().manual_seed(3)
qkv = torch.randn(B, T, 3, H, dh, generator=g).to(dev).to(torch.bfloat16)
d_o = torch.randn(B, T, H, dh, generator=g).to(dev).to(torch.bfloat16)
o = torch.empty(B, T, H, dh, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, T, device=dev, dtype=torch.float32)
dqkv = torch.empty_like(qkv)
s = torch.cuda.current_stream().cuda_stream
fwd = lambda: _cabi.check(lib.v4h_test_attention_fwd(1, 1, qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), B, T, H, dh, s))
bwd = lambda: _cabi.check(lib.v4h_test_attention_bwd(1, 1, qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), d_o.data_ptr(),
                                                      dqkv.data_ptr(), B, T, H, dh, s))


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


fwd(); bwd(); torch.cuda.synchronize()
nb = min(B, 4)
ref = qkv[:nb].float().requires_grad_(True)
q, k, v = ref.permute(2, 0, 3, 1, 4)
want = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2)
want.backward(d_o[:nb].float())
rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm())
flops = 4.0 * B * H * T * T * dh
tf, tb = timed(fwd), timed(bwd)
env = {k: v for k, v in os.environ.items() if k.startswith("V4H_ATTN")}
print(f"B={B} T={T} H={H} dh={dh} {env}: fwd {tf:7.1f} us {flops / tf / 1e6:6.1f} TF | bwd {tb:7.1f} us {2.5 * flops / tb / 1e6:6.1f} TF"
      f" | rel err o {rel(o[:nb], want):.1e} dqkv {rel(dqkv[:nb], ref.grad):.1e}", flush=True)
if "--phases" in sys.argv or os.environ.get("ATTN_PHASES"):
    cnt = torch.zeros(10, dtype=torch.int64, device=dev)
    lib.v4h_debug_attention_counters(cnt.data_ptr())
    bwd(); torch.cuda.synchronize()
    lib.v4h_debug_attention_counters(None)
    nct = B * H * ((T + 127) // 128)
    names = ["prologue", "row_stats", "wait_scores", "wait_acc", "dS", "signal", "final_wait", "out", "teardown", "-"]
    print("dQ pipe phases, cycles per CTA:", " ".join(f"{n}={v / nct:.0f}" for n, v in zip(names, cnt.cpu().tolist())), flush=True)
