"""Per-output error report + timing of the tcgen05 attention kernels against an fp64 torch reference."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit4hep_b200 import _cabi

lib = _cabi.load()
dev = torch.device("cuda:0")


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def run(B, T, H, dh, engine, time_it=False):
    g = torch.Generator().manual_seed(3)
    qkv = torch.randn(B, T, 3, H, dh, generator=g).to(dev).to(torch.bfloat16)
    d_o = torch.randn(B, T, H, dh, generator=g).to(dev).to(torch.bfloat16)
    o = torch.zeros(B, T, H, dh, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, T, device=dev, dtype=torch.float32)
    dqkv = torch.zeros_like(qkv)
    s = torch.cuda.current_stream().cuda_stream
    fwd = lambda: _cabi.check(lib.v4h_test_attention_fwd(1, engine, qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), B, T, H, dh, s))
    bwd = lambda: _cabi.check(lib.v4h_test_attention_bwd(1, engine, qkv.data_ptr(), o.data_ptr(), lse.data_ptr(),
                                                          d_o.data_ptr(), dqkv.data_ptr(), B, T, H, dh, s))
    fwd(); bwd(); torch.cuda.synchronize()
    if B * T * T * H <= 64 * 135 * 135 * 6:
        ref = qkv.double().requires_grad_(True)
        q, k, v = ref.permute(2, 0, 3, 1, 4)
        sc = (q @ k.transpose(-1, -2)) * dh ** -0.5
        want = (torch.softmax(sc, -1) @ v).transpose(1, 2)
        want.backward(d_o.double())
        gq, gk, gv = ref.grad.unbind(2)
        dq, dk, dv = dqkv.unbind(2)
        print(f"B{B} T{T} H{H} dh{dh} eng{engine}: o {rel(o, want):.2e} lse {rel(lse, torch.logsumexp(sc, -1)):.2e} "
              f"dq {rel(dq, gq):.2e} dk {rel(dk, gk):.2e} dv {rel(dv, gv):.2e}", flush=True)
    if time_it:
        for name, fn in (("fwd", fwd), ("bwd", bwd)):
            for _ in range(3): fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): fn()
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 20 * 1e3
            fl = (4 if name == "fwd" else 10) * B * H * T * T * dh
            print(f"   {name} eng{engine}: {us:8.1f} us  {fl / us / 1e6:8.1f} TFLOP/s", flush=True)


def phases(B, T, H, dh):
    g = torch.Generator().manual_seed(3)
    qkv = torch.randn(B, T, 3, H, dh, generator=g).to(dev).to(torch.bfloat16)
    o = torch.zeros(B, T, H, dh, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, T, device=dev, dtype=torch.float32)
    cnt = torch.zeros(10, dtype=torch.int64, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    lib.v4h_debug_attention_counters(cnt.data_ptr())
    _cabi.check(lib.v4h_test_attention_fwd(1, 1, qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), B, T, H, dh, s))
    torch.cuda.synchronize()
    lib.v4h_debug_attention_counters(None)
    nct = B * H * ((T + 127) // 128)
    names = ["prologue", "issue_ld", "wait_ld", "publish", "mma_S", "softmax", "publish2", "mma_PV", "out", "teardown"]
    print("fwd phases, cycles per CTA:", " ".join(f"{n}={v / nct:.0f}" for n, v in zip(names, cnt.cpu().tolist())), flush=True)
    if T <= 160:
        d_o = torch.randn(B, T, H, dh, generator=g).to(dev).to(torch.bfloat16)
        dqkv = torch.zeros_like(qkv)
        cnt.zero_()
        lib.v4h_debug_attention_counters(cnt.data_ptr())
        _cabi.check(lib.v4h_test_attention_bwd(1, 1, qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), d_o.data_ptr(), dqkv.data_ptr(), B, T, H, dh, s))
        torch.cuda.synchronize()
        lib.v4h_debug_attention_counters(None)
        names = ["stage+stats", "wait_ld", "mma_S", "P", "mma_dP_dV", "dS", "mma_dQ_dK", "dQ_out", "dKdV_out", "teardown"]
        print("fused bwd phases, cycles per CTA:", " ".join(f"{n}={v / (B * H):.0f}" for n, v in zip(names, cnt.cpu().tolist())), flush=True)


if __name__ == "__main__":
    if "--phases" in sys.argv:
        phases(64, 135, 6, 80)
        phases(256, 135, 6, 80)
        sys.exit(0)
    shapes = [(1, 16, 1, 16), (1, 128, 1, 64), (2, 135, 6, 80), (1, 450, 6, 80), (3, 84, 2, 24), (1, 606, 6, 80), (2, 33, 4, 32),
              (1, 300, 3, 128)]
    for sh in shapes:
        run(*sh, 1)
    for eng in (0, 1):
        run(64, 135, 6, 80, eng, True)
    run(64, 450, 6, 80, 1, True)
    run(16, 606, 6, 80, 1, True)
