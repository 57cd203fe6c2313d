"""Kernel-level GPU tests through the C-ABI test hooks: the tcgen05 GEMM in its three operand layouts
(forward NT, dgrad NN, wgrad TN with split-K) and the attention kernels, against fp32 torch references
of the same op on the same bf16-rounded inputs."""
import pytest
import torch

from oracle import vit_oracle as vo

pytestmark = pytest.mark.gpu

NT, NN, TN = 0, 1, 2


def _gemm(engine, layout, A, B, m, n, k):
    from vit4hep_b200 import _cabi
    lib = _cabi.load()
    C = torch.full((m, n), float("nan"), device=A.device, dtype=torch.float32)
    _cabi.check(lib.v4h_test_gemm(engine, layout, A.data_ptr(), B.data_ptr(), C.data_ptr(), m, n, k,
                                  torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return C


def _operands(layout, m, n, k, dtype, dev, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    a_shape = (k, m) if layout == TN else (m, k)
    b_shape = (n, k) if layout == NT else (k, n)
    A = torch.randn(a_shape, generator=g).to(dev).to(dtype)
    B = torch.randn(b_shape, generator=g).to(dev).to(dtype)
    Af = A.double().t() if layout == TN else A.double()
    Bf = B.double().t() if layout == NT else B.double()
    return A, B, Af @ Bf


SHAPES = [
    (128, 64, 64), (128, 240, 64), (128, 256, 128), (256, 480, 480), (135, 144, 48), (8640, 1440, 480),
    (8640, 480, 1920), (1000, 48, 480), (64, 18240, 480), (77, 96, 200), (300, 1920, 480), (129, 272, 72),
]


@pytest.mark.parametrize("layout", [NT, NN, TN])
@pytest.mark.parametrize("m,n,k", SHAPES)
def test_umma_gemm_layouts(layout, m, n, k):
    dev = torch.device("cuda:0")
    if layout == TN and (m % 8 or n % 8):
        pytest.skip("TMA needs 16-byte row pitch")
    if layout == NN and n % 8:
        pytest.skip("TMA needs 16-byte row pitch")
    if layout != TN and k % 8:
        pytest.skip("TMA needs 16-byte row pitch")
    A, B, want = _operands(layout, m, n, k, torch.bfloat16, dev)
    got = _gemm(1, layout, A, B, m, n, k)
    assert torch.isfinite(got).all()
    err = vo.rel_l2(got, want)
    assert err < 2e-6, f"layout {layout} {m}x{n}x{k}: rel-L2 {err}"


@pytest.mark.parametrize("layout", [NT, NN, TN])
def test_simt_gemm_layouts(layout):
    dev = torch.device("cuda:0")
    for (m, n, k) in [(135, 48, 48), (64, 480, 46), (300, 77, 129)]:
        A, B, want = _operands(layout, m, n, k, torch.float32, dev)
        got = _gemm(0, layout, A, B, m, n, k)
        assert vo.rel_l2(got, want) < 1e-6


@pytest.mark.parametrize("precision,engine", [(0, 0), (1, 0), (1, 1)])
@pytest.mark.parametrize("B,T,H,dh", [(2, 135, 6, 80), (1, 450, 6, 80), (3, 84, 2, 24), (1, 606, 6, 80), (2, 33, 4, 32),
                                      (2, 128, 2, 64), (1, 300, 3, 128), (1, 1, 1, 8), (2, 17, 3, 40),
                                      # long sequences at other head sizes: two-CTA forward blocks and the pipelined
                                      # dQ / dK-dV kernels (TMA tiles: dh 32, 64), the cp.async path (dh 24, 40)
                                      (2, 200, 2, 64), (1, 257, 4, 32), (1, 180, 2, 24), (1, 161, 1, 40), (1, 1000, 1, 80)])
def test_attention_fwd_bwd(precision, engine, B, T, H, dh):
    """engine 0 = SIMT kernels (fp32 mode arithmetic), 1 = tcgen05 / TMEM kernels (bf16 mode)"""
    from vit4hep_b200 import _cabi
    lib = _cabi.load()
    dev = torch.device("cuda:0")
    dt = torch.bfloat16 if precision else torch.float32
    g = torch.Generator().manual_seed(3)
    qkv = torch.randn(B, T, 3, H, dh, generator=g).to(dev).to(dt)
    d_o = torch.randn(B, T, H, dh, generator=g).to(dev).to(dt)
    o = torch.empty(B, T, H, dh, device=dev, dtype=dt)
    lse = torch.empty(B, H, T, device=dev, dtype=torch.float32)
    dqkv = torch.empty_like(qkv)
    s = torch.cuda.current_stream().cuda_stream
    _cabi.check(lib.v4h_test_attention_fwd(precision, engine, qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), B, T, H, dh, s))
    _cabi.check(lib.v4h_test_attention_bwd(precision, engine, qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), d_o.data_ptr(),
                                           dqkv.data_ptr(), B, T, H, dh, s))
    ref = qkv.double().requires_grad_(True)
    q, k, v = ref.permute(2, 0, 3, 1, 4)
    sc = (q @ k.transpose(-1, -2)) * dh ** -0.5
    want = (torch.softmax(sc, -1) @ v).transpose(1, 2)
    want.backward(d_o.double())
    tol = 2e-2 if precision else 1e-5
    assert vo.rel_l2(o, want) < tol
    assert vo.rel_l2(lse, torch.logsumexp(sc, -1)) < (1e-5 if not precision else 2e-3)
    assert vo.rel_l2(dqkv, ref.grad) < tol


@pytest.mark.parametrize("m,n,k,T", [(8640, 480, 480, 135), (8640, 480, 1920, 135), (19000, 480, 480, 135),
                                     (34560, 480, 1920, 135), (19100, 256, 200, 88),
                                     (1000, 96, 96, 88), (300, 64, 200, 125), (540, 480, 480, 135), (4000, 224, 480, 450)])
def test_gate_residual_gemm_with_layernorm_epilogue(m, n, k, T):
    """csrc/gemm_umma.cu gemm_gate_res_ln against torch: a 2-CTA cluster splits the columns when the row tiles do not
    fill the GPU (m < 18944), a cta_group::2 pair owns 256 whole rows otherwise (m >= 19000, odd tile counts
    included); ragged m / k, several samples per tile."""
    from vit4hep_b200 import _cabi
    lib = _cabi.load()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(m + n + k)
    bf = torch.bfloat16
    nb = (m + T - 1) // T
    A = (torch.randn(m, k, generator=g) * 0.5).to(dev, bf)
    W = (torch.randn(n, k, generator=g) / k ** 0.5).to(dev, bf)
    bias = torch.randn(n, generator=g).to(dev)
    res = (torch.randn(m, n, generator=g) * 2 + 0.5).to(dev)
    gate, shift, scale = (torch.randn(nb, n, generator=g).to(dev) * s for s in (1.0, 0.5, 0.3))
    ld = n + 8
    y = torch.zeros(m, n, device=dev, dtype=bf)
    res_out = torch.zeros(m, n, device=dev)
    ln = torch.full((m, ld), 7.0, device=dev, dtype=bf)
    stats = torch.zeros(m, 2, device=dev)
    _cabi.check(lib.v4h_debug_gemm_ln(m, n, k, T, A.data_ptr(), W.data_ptr(), bias.data_ptr(), y.data_ptr(), res.data_ptr(),
                                      res_out.data_ptr(), gate.data_ptr(), shift.data_ptr(), scale.data_ptr(),
                                      ln.data_ptr(), ld, stats.data_ptr(), None, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    sample = torch.arange(m, device=dev) // T
    want_y = A.double() @ W.double().T + bias.double()
    want_h = res.double() + gate.double()[sample] * want_y
    mu = want_h.mean(1, keepdim=True)
    var = ((want_h - mu) ** 2).mean(1, keepdim=True)
    want_ln = (want_h - mu) / torch.sqrt(var + 1e-6) * (1 + scale.double()[sample]) + shift.double()[sample]
    assert vo.rel_l2(y, want_y) < 4e-3            # bf16 output rounding
    assert vo.rel_l2(res_out, want_h) < 1e-5      # fp32 residual stream: tensor-core accumulation only
    assert vo.rel_l2(ln[:, :n], want_ln) < 4e-3
    assert vo.rel_l2(stats[:, 0], mu.flatten()) < 1e-4 and vo.rel_l2(stats[:, 1], 1 / torch.sqrt(var.flatten() + 1e-6)) < 1e-4
    ones = torch.zeros(m, 8, device=dev, dtype=bf); ones[:, 0] = 1
    assert torch.equal(ln[:, n:], ones)           # the "ones" column of the wider pitch
