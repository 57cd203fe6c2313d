"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the golden vectors
of the unmodified reference.  Tolerances are north_star's: bit-exact patch indexing; rel-L2 <= 1e-5 in
fp32 precision, <= 2e-2 in bf16 precision, on velocity, loss, every parameter gradient and sampled
showers."""
import math

import numpy as np
import pytest
import torch

from oracle import vit_oracle as vo
from tests.helpers import build_model, geometry_of, load_golden

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5
BF16_TOL = 2e-2
# gradients of single parameters are small-norm sums of many bf16-rounded terms: same budget as the
# velocity, except a slightly wider one for the tiny-norm bias / frequency tensors
TOL = {"fp32": dict(out=FP32_TOL, grad=2e-5, sample=FP32_TOL), "bf16": dict(out=BF16_TOL, grad=3e-2, sample=BF16_TOL)}


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "the -m gpu tests need a B200"
    from vit4hep_b200 import _cabi
    _cabi.require_device(0)
    return torch.device("cuda:0")


# ----------------------------------------------------------------------------------------------
# patchify / unpatchify: bit-exact
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(vo.CONFIGS))
@pytest.mark.parametrize("batch", [1, 3, 64])
def test_patchify_bit_exact(dev, golden_dir, name, batch):
    g = geometry_of(name)
    og = vo.CONFIGS[name]["geom"]
    gen = torch.Generator().manual_seed(7)
    x = torch.randn(batch, *og.sample_shape, generator=gen)
    want = vo.to_patches(x, og)
    got = g.to_patches(x.to(dev))
    assert got.shape == want.shape
    assert torch.equal(got.cpu(), want)  # bit-exact copy semantics
    back = g.from_patches(got)
    assert torch.equal(back.cpu(), x)
    # arange showers reproduce the reference's golden integer map
    ar = torch.arange(og.voxels, dtype=torch.float32).reshape(1, *og.sample_shape).to(dev)
    z = np.load(f"{golden_dir}/patch_maps.npz")
    assert np.array_equal(g.to_patches(ar)[0].cpu().numpy().astype(np.int32), z[name])


def test_patchify_multichannel_and_empty(dev):
    from vit4hep_b200.cfm import PatchGeometry
    og = vo.Geometry((vo.Segment((6, 4, 6), (3, 2, 2)),), in_channels=3)
    g = PatchGeometry([(6, 4, 6)], [(3, 2, 2)], 3, False)
    x = torch.randn(5, 3, 6, 4, 6)
    assert torch.equal(g.to_patches(x.to(dev)).cpu(), vo.to_patches(x, og))
    assert torch.equal(g.from_patches(g.to_patches(x.to(dev))).cpu(), x)
    empty = g.to_patches(torch.zeros(0, 3, 6, 4, 6, device=dev))
    assert empty.shape == (0, og.tokens, og.patch_dim)


def test_patchify_round_trip_full_size(dev):
    """size-independent property at benchmark size: unpatchify(patchify(x)) == x, and patchify is a
    permutation (sorted values agree)."""
    for name in ("ds2", "ds3", "calohad"):
        g = geometry_of(name)
        og = vo.CONFIGS[name]["geom"]
        x = torch.randn(256, *og.sample_shape, device=dev)
        tok = g.to_patches(x)
        assert torch.equal(g.from_patches(tok), x)
        assert torch.equal(tok.reshape(256, -1).sort(dim=1).values, x.reshape(256, -1).sort(dim=1).values)


# ----------------------------------------------------------------------------------------------
# network / loss / gradients / sampling against the reference's golden vectors
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("tag,name", [("ds2_tiny", "ds2"), ("calogan_tiny", "calogan"), ("ds1_pions_tiny", "ds1_pions")])
def test_golden_forward_loss_grads_sample(dev, golden_dir, tag, name, precision):
    z, sd, over, B = load_golden(golden_dir, tag)
    tol = TOL[precision]
    param = dict(vo.CONFIGS[name]["param"]); param.update(over)
    model = build_model(name, param, precision, dev)
    model.net.load_state_dict(sd)
    x1, c, t = (torch.from_numpy(z[k]).to(dev) for k in ("x1", "c", "t_fwd"))

    # (1) velocity field: net on tokens, and wrapper on the voxel grid
    with torch.no_grad():
        out = model.net(model.to_patches(x1), t, c)
        wout = model.forward(x1, t, c)
    assert vo.rel_l2(out, torch.from_numpy(z["net_out"])) < tol["out"]
    assert vo.rel_l2(wout, torch.from_numpy(z["wrapper_out"])) < tol["out"]

    # (2) _batch_loss with the reference's RNG stream: t is drawn on the host (same generator, same
    # call), x_0 on the device - supply the recorded x_0 by patching randn_like
    want_x0 = torch.from_numpy(z["loss_x0"]).to(dev)
    orig = torch.randn_like
    torch.randn_like = lambda *_a, **_k: want_x0
    try:
        torch.manual_seed(77)
        loss = model._batch_loss((x1.cpu(), c.cpu()))
    finally:
        torch.randn_like = orig
    loss.backward()
    assert abs(loss.item() - float(z["loss"])) / float(z["loss"]) < tol["out"]
    for k in z.files:
        if k.startswith("grad/"):
            p = dict(model.net.named_parameters())[k[5:]]
            assert p.grad is not None, k
            assert vo.rel_l2(p.grad, torch.from_numpy(z[k])) < tol["grad"], k

    # (3) 20-step RK4 (3/8 rule) sampling from the recorded x_T
    s = model.integrate(torch.from_numpy(z["x_T"]).to(dev), c)
    assert vo.rel_l2(s, torch.from_numpy(z["sample"])) < tol["sample"]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_sample_batch_uses_the_reference_rng_stream(dev, golden_dir, precision):
    z, sd, over, B = load_golden(golden_dir, "ds2_tiny")
    param = dict(vo.CONFIGS["ds2"]["param"]); param.update(over)
    model = build_model("ds2", param, precision, dev)
    model.net.load_state_dict(sd)
    c = torch.from_numpy(z["c"]).to(dev)
    torch.manual_seed(99)
    x_T = torch.randn((B, 1, 45, 16, 9), device=dev)
    torch.manual_seed(99)
    s = model.sample_batch(c)
    assert s.shape == (B, 1, 45, 16, 9)
    assert torch.equal(s, model.integrate(x_T, c))


# ----------------------------------------------------------------------------------------------
# full-size network against the oracle (oracle on CPU, seconds)
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name,batch", [("ds2", 4), ("ds3", 2), ("calohad", 1), ("ds1_photons", 3), ("ds1_pions", 3),
                                        ("calogan", 5), ("lemurs", 2), ("ds2", 64)])
def test_full_size_forward_and_grads_vs_oracle(dev, name, batch, precision):
    cfg = vo.CONFIGS[name]
    geom, param = cfg["geom"], cfg["param"]
    tol = TOL[precision]
    sd = vo.init_state_dict(param, seed=3)
    model = build_model(name, param, precision, dev)
    model.net.load_state_dict(sd)
    gen = torch.Generator().manual_seed(11)
    x1 = torch.randn(batch, *geom.sample_shape, generator=gen)
    x0 = torch.randn(batch, *geom.sample_shape, generator=gen)
    c = torch.rand(batch, param["condition_dim"], generator=gen)
    t = torch.rand(batch, generator=gen)

    params = {k: v.clone().requires_grad_(not k.startswith("pos_") or k == "pos_embed_freqs") for k, v in sd.items()}
    want_loss = vo.cfm_loss(params, x1, c, x0, t, geom, param["num_heads"])
    want_loss.backward()

    xt, target = model.geometry.cfm_prepare(x1.to(dev), x0.to(dev), t.to(dev))
    # the fused trajectory kernel is exact up to fp32 rounding of one fma
    want_xt, want_target = vo.linear_trajectory(x0, x1, t.reshape(-1, *([1] * (x1.dim() - 1))))
    assert vo.rel_l2(xt, vo.to_patches(want_xt, geom)) < 1e-6
    assert torch.equal(target.cpu(), vo.to_patches(want_target, geom))
    from vit4hep_b200.cfm import _MSELoss
    v = model.net(xt, t.to(dev).view(-1, 1), c.to(dev))
    loss = _MSELoss.apply(v, target)
    loss.backward()
    assert abs(loss.item() - want_loss.item()) / want_loss.item() < tol["out"]
    named = dict(model.net.named_parameters())
    for k, p in params.items():
        if p.requires_grad:
            assert vo.rel_l2(named[k].grad, p.grad) < tol["grad"], k


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_ds2_forward_at_the_sampling_batch_vs_oracle(dev, precision):
    """ds2 forward at the sampling batch 256 (shared t, like every network evaluation of the ODE solve):
    other GEMM tile counts, persistent multi-tile schedules and sub-batch split than the small-batch cases."""
    cfg = vo.CONFIGS["ds2"]
    geom, param = cfg["geom"], cfg["param"]
    sd = vo.init_state_dict(param, seed=3)
    model = build_model("ds2", param, precision, dev)
    model.net.load_state_dict(sd)
    gen = torch.Generator().manual_seed(21)
    B = 256
    x = torch.randn(B, geom.tokens, geom.patch_dim, generator=gen)
    c = torch.rand(B, param["condition_dim"], generator=gen)
    t = torch.tensor([0.37])
    with torch.no_grad():
        want = vo.vit_forward(sd, x, t.repeat(B, 1), c, param["num_heads"])
        got = model.net(x.to(dev), t.to(dev), c.to(dev), shared_t=True)
        per_sample = model.net(x.to(dev), t.repeat(B, 1).to(dev), c.to(dev))
    assert vo.rel_l2(got, want) < TOL[precision]["out"]
    assert torch.equal(got, per_sample)  # shared_t is the same arithmetic as one t per sample


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_full_size_rk4_solve_vs_oracle(dev, precision):
    """Full-size ds2 network, the whole 20-step RK4 (3/8 rule) solve = 80 chained network evaluations from a
    fixed x_T against the oracle's fp32 solve (bf16 drift over the trajectory), and the
    size-independent link to the sampling batch: every row of the computation is independent of the batch it
    sits in, so the first showers of a 256-shower solve equal the 4-shower solve -- up to the summation order of
    the LayerNorm statistics, which the fused GEMM epilogue accumulates per CTA (one CTA per row at batch 256, two
    column halves at batch 4): a last-bit difference in a mean flips bf16 roundings downstream, 5e-5 after 80
    evaluations, against the 2e-2 budget."""
    cfg = vo.CONFIGS["ds2"]
    geom, param = cfg["geom"], cfg["param"]
    sd = vo.init_state_dict(param, seed=3)
    model = build_model("ds2", param, precision, dev)
    model.net.load_state_dict(sd)
    gen = torch.Generator().manual_seed(31)
    B = 4
    x_T = torch.randn(B, *geom.sample_shape, generator=gen)
    c = torch.rand(B, param["condition_dim"], generator=gen)
    with torch.no_grad():
        want = vo.sample_batch(sd, c, x_T, geom, param["num_heads"], step_size=0.05)
    got = model.integrate(x_T.to(dev), c.to(dev))
    assert torch.isfinite(got).all()
    assert vo.rel_l2(got, want) < TOL[precision]["sample"]
    if precision == "bf16":
        big_x = torch.cat([x_T, torch.randn(252, *geom.sample_shape, generator=gen)]).to(dev)
        big_c = torch.cat([c, torch.rand(252, param["condition_dim"], generator=gen)]).to(dev)
        big = model.integrate(big_x, big_c)
        assert vo.rel_l2(big[:B], got) < 1e-3
        assert vo.rel_l2(big[:B], want) < TOL[precision]["sample"]
        model.graph_sampling = True
        assert torch.equal(model.integrate(big_x, big_c), big)


def test_lemurs_batch_loss_golden(dev, golden_dir):
    """LEMURSCFM._batch_loss: (B, R, A, L) batches, K = 53 (reference experiments/lemurs/model.py:62-65)"""
    z, sd, over, B = load_golden(golden_dir, "lemurs_tiny")
    param = dict(vo.CONFIGS["lemurs"]["param"]); param.update(over)
    for precision in ("fp32", "bf16"):
        tol = TOL[precision]
        model = build_model("lemurs", param, precision, dev)
        model.net.load_state_dict(sd)
        want_x0 = torch.from_numpy(z["loss_x0"]).to(dev)
        orig = torch.randn_like
        torch.randn_like = lambda *_a, **_k: want_x0
        try:
            torch.manual_seed(55)
            loss = model._batch_loss([torch.from_numpy(z["x"]), torch.from_numpy(z["c"])])
        finally:
            torch.randn_like = orig
        loss.backward()
        assert abs(loss.item() - float(z["loss"])) / float(z["loss"]) < tol["out"]
        named = dict(model.net.named_parameters())
        for k in z.files:
            if k.startswith("grad/"):
                assert vo.rel_l2(named[k[5:]].grad, torch.from_numpy(z[k])) < tol["grad"], (precision, k)
        # graphed training passes device_rng through the override
        assert torch.isfinite(model._batch_loss([torch.from_numpy(z["x"]), torch.from_numpy(z["c"])], device_rng=True))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_finetuning_structures_golden(dev, golden_dir, precision):
    """The reference's finetuning surgery (experiment_finetuning.py:75-165) applied to a vit4hep_b200.ViT:
    mapper Linears in front of x_embedder / c_embedder, pos grid of the new geometry, a new final layer --
    output and every parameter gradient against the reference's."""
    import torch.nn as nn
    import vit4hep_b200 as v4
    from vit4hep_b200.vit import FinalLayer
    z, sd, over, B = load_golden(golden_dir, "finetune_tiny")
    tol = TOL[precision]
    back = dict(vo.CONFIGS["ds2"]["param"]); back.update(over); back["precision"] = precision
    new = vo.CONFIGS["ds3"]["param"]
    K_new = int(z["meta"][4])
    net = v4.ViT(back)
    net.x_embedder = nn.Sequential(nn.Linear(new["patch_dim"], back["patch_dim"]), nn.SiLU(), net.x_embedder)
    net.c_embedder = nn.Sequential(nn.Linear(K_new, back["condition_dim"]), nn.SiLU(), net.c_embedder)
    net.num_patches = new["num_patches"]
    net.pos_z, net.pos_y, net.pos_x = net.create_meshgrid()
    net.final_layer = FinalLayer(back["hidden_dim"], new["patch_dim"], back["out_channels"])
    net = net.to(dev)
    net.load_state_dict(sd)
    x, t, c, wgt = (torch.from_numpy(z[k]).to(dev) for k in ("x", "t", "c", "wgt"))
    y = net(x, t, c)
    assert vo.rel_l2(y, torch.from_numpy(z["net_out"])) < tol["out"]
    (y * wgt).sum().backward()
    named = dict(net.named_parameters())
    for k in z.files:
        if k.startswith("grad/"):
            assert named[k[5:]].grad is not None, k
            assert vo.rel_l2(named[k[5:]].grad, torch.from_numpy(z[k])) < tol["grad"], k


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_fixed_positional_embedding_golden(dev, golden_dir, precision):
    """learn_pos_embed=False (reference nn/vit.py:92-103): the table is a constant buffer the x_embedder epilogue adds"""
    import vit4hep_b200 as v4
    z, sd, over, B = load_golden(golden_dir, "fixed_pos_tiny")
    param = dict(vo.CONFIGS["ds3"]["param"]); param.update(over)
    param.update(learn_pos_embed=False, pos_embedding_coords="cylindrical", precision=precision,
                 num_patches=list(param["num_patches"][0]))
    net = v4.ViT(param)
    assert vo.rel_l2(net.pos_embed, torch.from_numpy(z["table/cylindrical"])) < 1e-6
    net = net.to(dev)
    net.load_state_dict(sd)
    with torch.no_grad():
        y = net(*(torch.from_numpy(z[k]).to(dev) for k in ("x", "t", "c")))
    assert vo.rel_l2(y, torch.from_numpy(z["net_out"])) < TOL[precision]["out"]


def test_sub_batch_lanes_give_the_single_chain_result(dev, monkeypatch):
    """The library runs a batch as two sub-batches on two streams (csrc/vit.cu Lane).  Every output row depends
    on its own sample only and a GEMM element sees the same k order whatever M is, so the forward is
    bit-identical to the unsplit run; gradients are sums over samples whose fp32 atomics reorder (same bound
    as for the side streams)."""
    name, batch = "ds2", 64
    cfg = vo.CONFIGS[name]
    geom, param = cfg["geom"], cfg["param"]
    sd = vo.init_state_dict(param, seed=5)
    gen = torch.Generator().manual_seed(19)
    x = torch.randn(batch, geom.tokens, geom.patch_dim, generator=gen).to(dev)
    t = torch.rand(batch, 1, generator=gen).to(dev)
    c = torch.rand(batch, param["condition_dim"], generator=gen).to(dev)
    dout = torch.randn(batch, geom.tokens, geom.patch_dim, generator=gen).to(dev)
    res = {}
    for mode in ("1", "2"):
        monkeypatch.setenv("V4H_MICROBATCH", mode)  # read when the plan is created
        model = build_model(name, param, "bf16", dev)
        model.net.load_state_dict(sd)
        v = model.net(x, t, c)
        v.backward(dout)
        torch.cuda.synchronize()
        res[mode] = (v.detach().clone(), {k: p.grad.detach().clone() for k, p in model.net.named_parameters()})
    assert torch.equal(res["1"][0], res["2"][0])
    for k, g in res["2"][1].items():
        assert vo.rel_l2(g, res["1"][1][k]) < 2e-3, k
    # odd batch: sub-batches of 3 and 2
    monkeypatch.setenv("V4H_MICROBATCH", "2")
    model = build_model(name, param, "bf16", dev)
    model.net.load_state_dict(sd)
    with torch.no_grad():
        assert torch.equal(model.net(x[:5], t[:5], c[:5]), res["1"][0][:5])


def test_side_streams_give_the_single_stream_result(dev, monkeypatch):
    """The library forks weight gradients and the small embedding / conditioning chains onto its own streams
    (fork after the producer, join before a buffer is rewritten).  A missing dependency would show up as a
    gradient that differs from the single-stream run or changes between repeats.  Equality is not available:
    the split-K atomics of dgrad.adaln reorder fp32 sums, a last-bit difference in d cond flips bf16 roundings
    downstream, and the conditioning-MLP gradients then differ by a few 1e-4 between any two runs (measured
    2.1e-4); a read of a half-written buffer would be orders of magnitude above the 2e-3 bound used here."""
    name, batch = "ds2", 64
    cfg = vo.CONFIGS[name]
    geom, param = cfg["geom"], cfg["param"]
    sd = vo.init_state_dict(param, seed=5)
    gen = torch.Generator().manual_seed(17)
    x = torch.randn(batch, geom.tokens, geom.patch_dim, generator=gen).to(dev)
    t = torch.rand(batch, 1, generator=gen).to(dev)
    c = torch.rand(batch, param["condition_dim"], generator=gen).to(dev)
    dout = torch.randn(batch, geom.tokens, geom.patch_dim, generator=gen).to(dev)

    def grads(side: bool, repeats: int):
        monkeypatch.setenv("V4H_WGRAD_STREAM", "1" if side else "0")  # read when the plan is created
        model = build_model(name, param, "bf16", dev)
        model.net.load_state_dict(sd)
        out = []
        for _ in range(repeats):
            model.net.zero_grad(set_to_none=True)
            v = model.net(x, t, c)
            v.backward(dout)
            torch.cuda.synchronize()
            out.append((v.detach().clone(), {k: p.grad.detach().clone() for k, p in model.net.named_parameters() if p.grad is not None}))
        return out

    ref_v, ref_g = grads(False, 1)[0]
    for v, g in grads(True, 4):
        assert torch.equal(v, ref_v)  # the forward has no atomics: bit-identical whatever the stream layout
        assert g.keys() == ref_g.keys()
        for k in g:
            assert vo.rel_l2(g[k].cpu(), ref_g[k].cpu()) < 2e-3, k


# ----------------------------------------------------------------------------------------------
# elementwise kernels
# ----------------------------------------------------------------------------------------------
def test_cfm_loss_and_axpy_kernels(dev):
    import ctypes
    from vit4hep_b200 import _cabi
    lib = _cabi.load()
    s = torch.cuda.current_stream().cuda_stream
    for n in (1, 31, 4097, 64 * 6480):
        v = torch.randn(n, device=dev); tg = torch.randn(n, device=dev)
        loss = torch.zeros((), device=dev); dv = torch.empty_like(v)
        _cabi.check(lib.v4h_cfm_loss(v.data_ptr(), tg.data_ptr(), n, 1.0, loss.data_ptr(), dv.data_ptr(), s))
        want = ((v.double() - tg.double()) ** 2).mean()
        assert abs(loss.item() - want.item()) / want.item() < 1e-6
        assert vo.rel_l2(dv, 2 * (v - tg) / n) < 1e-6
        y = torch.randn(n, device=dev); ks = [torch.randn(n, device=dev) for _ in range(4)]
        out = torch.empty_like(y)
        a = [0.3, -0.2, 0.7, 0.05]
        _cabi.check(lib.v4h_axpy4(out.data_ptr(), y.data_ptr(), ks[0].data_ptr(), a[0], ks[1].data_ptr(), a[1],
                                  ks[2].data_ptr(), a[2], ks[3].data_ptr(), a[3], n, s))
        want = y.double() + sum(ai * k.double() for ai, k in zip(a, ks))
        assert vo.rel_l2(out, want) < 1e-6
        _cabi.check(lib.v4h_axpy4(out.data_ptr(), y.data_ptr(), ks[0].data_ptr(), a[0], None, 0.0, None, 0.0,
                                  None, 0.0, n, s))
        assert vo.rel_l2(out, y.double() + a[0] * ks[0].double()) < 1e-6


def test_linearity_of_the_ode_combination_at_full_size(dev):
    """size-independent property at benchmark size: axpy4 is linear in every k."""
    import ctypes
    from vit4hep_b200 import _cabi
    lib = _cabi.load()
    s = torch.cuda.current_stream().cuda_stream
    n = 256 * 6480
    y = torch.randn(n, device=dev); k = torch.randn(n, device=dev)
    o1 = torch.empty_like(y); o2 = torch.empty_like(y)
    _cabi.check(lib.v4h_axpy4(o1.data_ptr(), y.data_ptr(), k.data_ptr(), 0.5, None, 0, None, 0, None, 0, n, s))
    _cabi.check(lib.v4h_axpy4(o2.data_ptr(), o1.data_ptr(), k.data_ptr(), -0.5, None, 0, None, 0, None, 0, n, s))
    assert vo.rel_l2(o2, y) < 1e-6


def test_fused_adamw_matches_torch_adamw_with_clipping(dev):
    """FusedAdamW (v4h_grad_norm_sq + v4h_adamw_step) against clip_grad_norm_ + torch.optim.AdamW on the same
    gradients (reference experiments/base_experiment.py:573-592), incl. the refreshed bf16 weight arena."""
    import copy
    import vit4hep_b200 as v4
    cfg = vo.tiny_config("ds2", hidden_dim=96, depth=2, num_heads=2)
    param = dict(cfg["param"]); param["precision"] = "bf16"
    torch.manual_seed(0)
    net = v4.ViT(param).to(dev)
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.randn_like(p) * 0.05)
    ref = copy.deepcopy(net)
    g = torch.Generator().manual_seed(5)
    B, T, P = 4, net.pos_z.numel(), param["patch_dim"]
    fused = v4.FusedAdamW(net, lr=1e-2, weight_decay=0.1, max_grad_norm=0.5)
    opt = torch.optim.AdamW(ref.parameters(), lr=1e-2, weight_decay=0.1)
    for step in range(3):
        x = torch.randn(B, T, P, generator=g).to(dev)
        t = torch.rand(B, 1, generator=g).to(dev)
        c = torch.rand(B, param["condition_dim"], generator=g).to(dev)
        net(x, t, c).square().mean().backward()
        for p, q in zip(net.parameters(), ref.parameters()):
            q.grad = p.grad.clone()
        want_norm = torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.5)
        opt.step()
        fused.step()
        assert abs(fused.last_grad_norm.item() - want_norm.item()) <= 1e-5 * want_norm.item()
        for (name, p), q in zip(net.named_parameters(), ref.parameters()):
            assert vo.rel_l2(p, q) < 1e-6, (step, name)
        fused.zero_grad(set_to_none=True)
    # the arena the optimizer refreshed gives the same forward as a freshly cast one
    x = torch.randn(B, T, P, generator=g).to(dev); t = torch.rand(B, 1, generator=g).to(dev)
    c = torch.rand(B, param["condition_dim"], generator=g).to(dev)
    with torch.no_grad():
        got = net(x, t, c)
        net._native.arena_key = None  # force the recast path
        want = net(x, t, c)
    assert torch.equal(got, want)


def test_fused_adamw_resumes_from_a_checkpoint_and_tracks_ema(dev):
    """load_state_dict must reach the kernel (job table with the loaded moments, step counters for the
    bias correction), also under graph-style device counters; the EMA fused into the optimizer pass follows
    torch_ema's update rule (reference experiments/base_experiment.py:127-134, :594)."""
    import copy
    import vit4hep_b200 as v4
    cfg = vo.tiny_config("ds2", hidden_dim=96, depth=2, num_heads=2)
    param = dict(cfg["param"]); param["precision"] = "bf16"
    torch.manual_seed(0)
    net = v4.ViT(param).to(dev)
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.randn_like(p) * 0.05)
    ref = copy.deepcopy(net)
    g = torch.Generator().manual_seed(5)
    B, T, P = 4, net.pos_z.numel(), param["patch_dim"]
    decay = 0.9
    ema = v4.ExponentialMovingAverage(net.parameters(), decay=decay)
    ema.to(dev)
    shadow = [p.detach().clone() for p in ref.parameters()]  # torch_ema arithmetic, restated
    fused = v4.FusedAdamW(net, lr=1e-2, weight_decay=0.1, max_grad_norm=0.5, ema=ema)
    opt = torch.optim.AdamW(ref.parameters(), lr=1e-2, weight_decay=0.1)
    n_updates = 0

    def one_step(fused_opt, the_ema):
        nonlocal n_updates
        x = torch.randn(B, T, P, generator=g).to(dev)
        t = torch.rand(B, 1, generator=g).to(dev)
        c = torch.rand(B, param["condition_dim"], generator=g).to(dev)
        net(x, t, c).square().mean().backward()
        for p, q in zip(net.parameters(), ref.parameters()):
            q.grad = p.grad.clone()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.5)
        opt.step()
        fused_opt.step()
        the_ema.update()  # the reference's loop calls it after optimizer.step(): recognised as already done
        n_updates += 1
        d = min(decay, (1 + n_updates) / (10 + n_updates))
        with torch.no_grad():
            for s_, q in zip(shadow, ref.parameters()):
                s_.sub_((1 - d) * (s_ - q))
        fused_opt.zero_grad(set_to_none=True)

    for _ in range(3):
        one_step(fused, ema)
    # checkpoint, then resume in fresh objects
    ckpt = {"optimizer": copy.deepcopy(fused.state_dict()), "ema": copy.deepcopy(ema.state_dict())}
    assert ckpt["ema"]["num_updates"] == 3
    assert {int(s_["step"]) for s_ in ckpt["optimizer"]["state"].values()} == {3}
    ema2 = v4.ExponentialMovingAverage(net.parameters(), decay=decay)
    ema2.load_state_dict(ckpt["ema"]); ema2.to(dev)
    fused2 = v4.FusedAdamW(net, lr=1e-2, weight_decay=0.1, max_grad_norm=0.5, ema=ema2)
    fused2.load_state_dict(ckpt["optimizer"])
    for _ in range(2):
        one_step(fused2, ema2)
    for (name, p), q in zip(net.named_parameters(), ref.parameters()):
        assert vo.rel_l2(p, q) < 1e-6, name
    for (name, _), s_, w in zip(net.named_parameters(), ema2.shadow_params, shadow):
        assert vo.rel_l2(s_, w) < 1e-6, name
    # loading in the SAME optimizer object must not reuse the job table of the old moments
    fused2.load_state_dict(ckpt["optimizer"])
    assert fused2._key is None and not fused2._tables and fused2._step == 3 and fused2._step_dev.item() == 3
    # stand-alone update() (EMA not attached to the optimizer) and average_parameters()
    ema3 = v4.ExponentialMovingAverage(net.parameters(), decay=decay); ema3.to(dev)
    before = [p.detach().clone() for p in net.parameters()]
    with torch.no_grad():
        for p in net.parameters():
            p.add_(0.01)
    ema3.update()
    d = min(decay, 2 / 11)
    for s_, b, p in zip(ema3.shadow_params, before, net.parameters()):
        assert vo.rel_l2(s_, b - (1 - d) * (b - p)) < 1e-6
    with ema3.average_parameters():
        assert all(torch.equal(p, s_) for p, s_ in zip(net.parameters(), ema3.shadow_params))
    assert all(torch.equal(p, b + 0.01) for p, b in zip(net.parameters(), before))


def test_second_backward_and_interleaved_forward(dev):
    """a forward with another token count between forward and backward must not disturb the pending
    graph (plans are kept per shape and travel with the autograd context); a second backward raises clearly."""
    import vit4hep_b200 as v4
    cfg = vo.tiny_config("ds2", hidden_dim=96, depth=2, num_heads=2)
    param = dict(cfg["param"]); param["precision"] = "bf16"
    torch.manual_seed(0)
    net = v4.ViT(param).to(dev)
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.randn_like(p) * 0.05)
    g = torch.Generator().manual_seed(5)
    T, P, K = net.pos_z.numel(), param["patch_dim"], param["condition_dim"]
    x = torch.randn(3, T, P, generator=g).to(dev); t = torch.rand(3, 1, generator=g).to(dev)
    c = torch.rand(3, K, generator=g).to(dev)
    net(x, t, c).square().mean().backward()
    want = {k: p.grad.clone() for k, p in net.named_parameters()}
    net.zero_grad(set_to_none=True)
    loss = net(x, t, c).square().mean()
    # another geometry in between (finetuning / validation on a different grid re-assigns the pos buffers)
    keep = (net.pos_z, net.pos_y, net.pos_x)
    net.pos_z, net.pos_y, net.pos_x = (b[:60].clone() for b in keep)
    with torch.no_grad():
        net(x[:, :60].contiguous(), t, c)
    net.pos_z, net.pos_y, net.pos_x = keep
    loss.backward()
    for k, p in net.named_parameters():
        assert vo.rel_l2(p.grad, want[k]) < 2e-3, k
    loss2 = net(x, t, c).square().mean()
    loss2.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="already consumed"):
        loss2.backward()


def test_graphed_train_step_and_sampling_match_eager(dev):
    """GraphedTrainStep / graph_sampling replay exactly what the eager path launches: same parameters after
    several steps with the same device RNG stream, same showers from the same noise."""
    import copy
    import vit4hep_b200 as v4
    cfg = vo.tiny_config("ds2", hidden_dim=96, depth=2, num_heads=2)
    geom, param = cfg["geom"], dict(cfg["param"]); param["precision"] = "bf16"
    torch.manual_seed(0)
    net = v4.ViT(param)
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.randn_like(p) * 0.05)
    mk = lambda n: v4.CaloChallengeCFM(n, [3, 16, 1], 1, "uniform", "linear",
                                       dict(method="rk4", options=dict(step_size=0.25)), shape=[45, 16, 9]).to(dev)
    m_eager, m_graph = mk(copy.deepcopy(net)), mk(copy.deepcopy(net))
    for m in (m_eager, m_graph):
        m.device, m.dtype = dev, torch.float32
    g = torch.Generator().manual_seed(7)
    xs = [torch.randn(4, *geom.sample_shape, generator=g).to(dev) for _ in range(4)]
    cs = [torch.rand(4, param["condition_dim"], generator=g).to(dev) for _ in range(4)]
    o_e = v4.FusedAdamW(m_eager.net, lr=1e-3, weight_decay=0.1, max_grad_norm=1.0)
    o_g = v4.FusedAdamW(m_graph.net, lr=1e-3, weight_decay=0.1, max_grad_norm=1.0)
    graphed = v4.GraphedTrainStep(m_graph, o_g, xs[0], cs[0], warmup=2)   # 2 eager warm-up steps, then the capture
    torch.cuda.manual_seed(11)
    losses_g = [graphed.step(x, c).item() for x, c in zip(xs, cs)]
    # eager twin: the same 2 priming steps on the first batch (a capture records, it does not execute)
    for _ in range(2):
        o_e.zero_grad(set_to_none=True); m_eager._batch_loss((xs[0], cs[0]), device_rng=True).backward(); o_e.step()
    torch.cuda.manual_seed(11)
    losses_e = []
    for x, c in zip(xs, cs):
        o_e.zero_grad(set_to_none=True)
        loss = m_eager._batch_loss((x, c), device_rng=True)
        loss.backward(); o_e.step(); losses_e.append(loss.item())
    assert all(math.isfinite(v) for v in losses_g)
    for (name, p), q in zip(m_graph.net.named_parameters(), m_eager.net.parameters()):
        # the device RNG offsets differ inside / outside a graph, so the twins agree statistically only
        assert vo.rel_l2(p, q) < 0.1, name
    assert o_g._step_dev.item() == 2 + 4 and o_e._step_dev.item() == 2 + 4  # the device step counter advances per replay
    # prefetched input path: batches staged from pinned host memory reach the static input buffers in order
    hx = [x.cpu().pin_memory() for x in xs]; hc = [c.cpu().pin_memory() for c in cs]
    graphed.stage(hx[1], hc[1])
    for i in (1, 2, 3):
        loss = graphed.step_staged()
        if i < 3:
            graphed.stage(hx[i + 1], hc[i + 1])  # issued while step i is still running
        assert math.isfinite(loss.item())
        if i == 3:
            assert torch.equal(graphed.x, xs[3]) and torch.equal(graphed.c, cs[3])
    assert o_g._step_dev.item() == 2 + 4 + 3
    # sampling: identical noise -> identical showers
    x_T = torch.randn(4, 1, 45, 16, 9, generator=g).to(dev)
    want = m_graph.integrate(x_T, cs[0])
    m_graph.graph_sampling = True
    got1 = m_graph.integrate(x_T, cs[0]); got2 = m_graph.integrate(x_T, cs[0])
    assert torch.equal(got1, want) and torch.equal(got2, want)


def test_ds3_sampling_at_changing_batch_sizes(dev):
    """Full-size ds3 network (T = 450: multi-block attention, multi-tile GEMMs, patch_dim 90 on the direct
    epilogue path), forward-only, at two batch sizes in a row; the solve is deterministic for a fixed x_T."""
    cfg = vo.CONFIGS["ds3"]
    geom, param = cfg["geom"], cfg["param"]
    model = build_model("ds3", param, "bf16", dev)
    model.odeint_kwargs = dict(method="rk4", options=dict(step_size=0.25))
    model.net.load_state_dict(vo.init_state_dict(param, seed=3))
    g = torch.Generator().manual_seed(5)
    for B in (32, 64):
        x_T = torch.randn(B, 1, *geom.segments[0].shape, generator=g).to(dev)
        c = torch.rand(B, param["condition_dim"], generator=g).to(dev)
        a = model.integrate(x_T, c)
        b = model.integrate(x_T, c)
        torch.cuda.synchronize()
        assert torch.isfinite(a).all() and torch.equal(a, b)
