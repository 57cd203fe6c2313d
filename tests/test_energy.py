"""Energy-ratio velocity network: oracle pinned by the reference's golden vectors (CPU),
the native forward and its ODE sampling against both (GPU).  Tolerances as for the ViT path: rel-L2 <= 1e-5 in
fp32 precision, <= 2e-2 in bf16 precision."""
import os

import numpy as np
import pytest
import torch

from oracle import energy_oracle as eo
from oracle import ref_stubs
from oracle import vit_oracle as vo

TOL = {"fp32": 1e-5, "bf16": 2e-2}
KEYS = ("dims_in", "dims_c", "dim_embedding", "encode_t_dim", "nhead", "num_encoder_layers", "num_decoder_layers",
        "dim_feedforward")


def _golden(golden_dir, dims_c):
    z = np.load(os.path.join(golden_dir, f"energy_tiny_c{dims_c}.npz"))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    param = dict(zip(KEYS, (int(v) for v in z["meta"])))
    param.update(dropout=0.0, activation="relu", embeds=True, encode_t_scale=30)
    return z, sd, param


@pytest.mark.parametrize("dims_c", [1, 3])
def test_oracle_matches_the_reference_golden(golden_dir, dims_c):
    z, sd, param = _golden(golden_dir, dims_c)
    x, t, c = (torch.from_numpy(z[k]) for k in ("x", "t", "c"))
    assert vo.rel_l2(eo.energy_forward(sd, x, t, c, param["nhead"]), torch.from_numpy(z["velocity"])) < 1e-5
    s = eo.sample_batch(sd, c, torch.from_numpy(z["x_T"]), param["nhead"])
    assert vo.rel_l2(s, torch.from_numpy(z["sample"])) < 1e-5


@pytest.mark.skipif(not ref_stubs.reference_available(), reason="live reference not present")
def test_oracle_matches_the_live_reference_at_full_size():
    ref_stubs.install()
    import importlib
    tc = importlib.import_module("nn.cfm.transformer_cfm")
    torch.manual_seed(0)
    net = tc.ParallelTransformer(eo.DS2_ENERGY)
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(1)
    x = torch.randn(4, 45, generator=g); t = torch.rand(4, 1, generator=g); c = torch.rand(4, 1, generator=g)
    with torch.no_grad():
        assert vo.rel_l2(eo.energy_forward(sd, x, t, c, 4), net(x, t, c)) < 1e-5
    mine = eo.init_state_dict(eo.DS2_ENERGY)
    assert {k: tuple(v.shape) for k, v in mine.items()} == {k: tuple(v.shape) for k, v in sd.items()}


def test_module_mirrors_the_reference_state_dict_and_refuses_cpu_and_training():
    import vit4hep_b200 as v4
    net = v4.ParallelTransformer(eo.DS2_ENERGY)
    sd = eo.init_state_dict(eo.DS2_ENERGY)
    assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    assert sum(p.numel() for p in net.parameters()) == 1958817  # the reference's count for configs/model/cfm/cfm_ds2_energy.yaml
    assert net.state_dict()["layer.weight"].data_ptr() == net.state_dict()["layers.0.weight"].data_ptr()
    x, t, c = torch.zeros(2, 45), torch.zeros(2, 1), torch.zeros(2, 1)
    with pytest.raises(NotImplementedError, match="forward-only"):
        net(x, t, c)
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        net(x, t, c)
    with pytest.raises(NotImplementedError):
        v4.ParallelTransformer(dict(eo.DS2_ENERGY, embeds=False))
    from vit4hep_b200 import configs
    model = configs.instantiate(dict(_target_="models.base_model.CFM", shape=[45], time_distribution="uniform",
                                     trajectory="linear", odeint_kwargs=dict(method="rk4", options=dict(step_size=0.05)),
                                     net=dict(_target_="nn.cfm.transformer_cfm.ParallelTransformer", param=eo.DS2_ENERGY)),
                                remap=True)
    assert isinstance(model, v4.CFM) and isinstance(model.net, v4.ParallelTransformer) and model.geometry is None


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("dims_c", [1, 3])
def test_native_forward_and_sampling_match_the_reference_golden(golden_dir, dims_c, precision):
    import vit4hep_b200 as v4
    dev = torch.device("cuda:0")
    z, sd, param = _golden(golden_dir, dims_c)
    net = v4.ParallelTransformer(dict(param, precision=precision)).to(dev)
    net.load_state_dict(sd)
    x, t, c = (torch.from_numpy(z[k]).to(dev) for k in ("x", "t", "c"))
    with torch.no_grad():
        v = net(x, t, c)
    assert vo.rel_l2(v, torch.from_numpy(z["velocity"])) < TOL[precision]
    model = v4.CFM(net, "uniform", "linear", dict(method="rk4", options=dict(step_size=0.05)), shape=[param["dims_in"]])
    s = model.integrate(torch.from_numpy(z["x_T"]).to(dev), c)
    assert vo.rel_l2(s, torch.from_numpy(z["sample"])) < TOL[precision]
    torch.manual_seed(5)
    x_T = torch.randn(c.shape[0], param["dims_in"], device=dev)
    torch.manual_seed(5)
    assert torch.equal(model.sample_batch(c), model.integrate(x_T, c))  # the reference's noise draw


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("dims_c,batch", [(1, 64), (3, 7), (1, 256)])
def test_full_size_forward_vs_oracle(dims_c, batch, precision):
    """the shipped ds2 / LEMURS energy configuration (d_model 128, 4 + 4 layers, 45 tokens)"""
    import vit4hep_b200 as v4
    dev = torch.device("cuda:0")
    param = dict(eo.DS2_ENERGY, dims_c=dims_c)
    sd = eo.init_state_dict(param, seed=2)
    net = v4.ParallelTransformer(dict(param, precision=precision)).to(dev)
    net.load_state_dict(sd)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(batch, 45, generator=g); t = torch.rand(batch, 1, generator=g); c = torch.rand(batch, dims_c, generator=g)
    with torch.no_grad():
        want = eo.energy_forward(sd, x, t, c, param["nhead"])
        got = net(x.to(dev), t.to(dev), c.to(dev))
        shared = net(x.to(dev), t[:1].to(dev), c.to(dev), shared_t=True)
        want_shared = eo.energy_forward(sd, x, t[:1].repeat(batch, 1), c, param["nhead"])
    assert vo.rel_l2(got, want) < TOL[precision]
    assert vo.rel_l2(shared, want_shared) < TOL[precision]
    # a new condition tensor is re-encoded; the same one is not (and gives the same result)
    with torch.no_grad():
        c2 = torch.rand(batch, dims_c, generator=g)
        got2 = net(x.to(dev), t.to(dev), c2.to(dev))
        assert vo.rel_l2(got2, eo.energy_forward(sd, x, t, c2, param["nhead"])) < TOL[precision]


@pytest.mark.gpu
def test_full_size_sampling_graphed_and_sharded():
    """20-step RK4 solve of the full-size energy network against the oracle, CUDA-graph replay == eager, and the
    size-independent property that a shower does not depend on the batch it is sampled in"""
    import vit4hep_b200 as v4
    dev = torch.device("cuda:0")
    param = eo.DS2_ENERGY
    sd = eo.init_state_dict(param, seed=4)
    g = torch.Generator().manual_seed(12)
    B = 6
    x_T = torch.randn(B, 45, generator=g); c = torch.rand(B, 1, generator=g)
    with torch.no_grad():
        want = eo.sample_batch(sd, c, x_T, param["nhead"])
    for precision in ("fp32", "bf16"):
        net = v4.ParallelTransformer(dict(param, precision=precision)).to(dev)
        net.load_state_dict(sd)
        model = v4.CFM(net, "uniform", "linear", dict(method="rk4", options=dict(step_size=0.05)), shape=[45])
        got = model.integrate(x_T.to(dev), c.to(dev))
        assert vo.rel_l2(got, want) < TOL[precision]
        if precision == "bf16":
            bx = torch.cat([x_T, torch.randn(250, 45, generator=g)]).to(dev)
            bc = torch.cat([c, torch.rand(250, 1, generator=g)]).to(dev)
            big = model.integrate(bx, bc)
            assert vo.rel_l2(big[:B], got) < 1e-6
            model.graph_sampling = True
            assert torch.equal(model.integrate(bx, bc), big)
            bc2 = torch.rand(256, 1, generator=g).to(dev)  # the replay must re-encode the new conditions
            with torch.inference_mode():
                want2 = model._integrate(bx, bc2)
            assert torch.equal(model.integrate(bx, bc2), want2)
