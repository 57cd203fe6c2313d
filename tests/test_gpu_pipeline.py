"""The drop-ins working together on the GPU, in the order the reference's CaloChallenge experiment uses them
(experiments/calochallenge/experiment.py): raw showers -> dataset with the forward transforms -> CFM training steps
(fused optimizer + EMA, eager and as a CUDA graph) -> energy-ratio sampling -> shape sampling -> reverse transforms.
Small networks, ds2 geometry; every stage is checked against what the previous one handed over."""
import math

import pytest
import torch

from oracle import energy_oracle as eo
from oracle import vit_oracle as vo
from tests.test_postprocess import CHAIN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _raw_showers(n, seed):
    g = torch.Generator().manual_seed(seed)
    raw = torch.exp(torch.randn(n, 6480, generator=g) * 1.5 + 3.0) * (torch.rand(n, 6480, generator=g) < 0.4)
    e_inc = 10.0 ** (3.0 + 3.0 * torch.rand(n, 1, generator=g))
    return raw * (e_inc * (0.5 + 0.4 * torch.rand(n, 1, generator=g)) / raw.sum(1, keepdim=True)), e_inc


def test_dataset_training_sampling_postprocessing(dev, tmp_path):
    import vit4hep_b200 as v4
    chain = dict(CHAIN)
    chain["GlobalStandardizeFromFile"] = {"model_dir": str(tmp_path), "eps": 1.0e-6}
    chain["AddFeaturesToCond"] = {"split_index": 6480}
    chain["Reshape"] = {"shape": [1, 45, 16, 9]}
    bounds = list(range(0, 6481, 144))
    raw, e_inc = _raw_showers(192, 21)

    # ---- data feed: statistics computed on the device, saved like the reference does, picked up by a second object
    fwd = v4.FusedForwardTransforms(chain, bounds)
    train = v4.ShowerDataset.from_arrays(raw, e_inc, train_val_frac=[0.75, 0.25], transform=fwd, split="training", device=dev)
    assert (tmp_path / "means.npy").exists() and (tmp_path / "stds.npy").exists()
    val = v4.ShowerDataset.from_arrays(raw, e_inc, train_val_frac=[0.75, 0.25], transform=v4.FusedForwardTransforms(chain, bounds),
                                       split="validation", device=dev)
    assert (len(train), len(val)) == (144, 48) and tuple(train.layers.shape[1:]) == (1, 45, 16, 9)
    assert torch.isfinite(train.layers).all() and torch.isfinite(train.energy).all()

    # ---- shape model: a small ViT behind the reference wrapper, fused optimizer with EMA
    cfg = vo.tiny_config("ds2", hidden_dim=96, depth=2, num_heads=2)
    param = dict(cfg["param"]); param["precision"] = "bf16"
    torch.manual_seed(0)
    model = v4.CaloChallengeCFM(v4.ViT(param), [3, 16, 1], 1, "uniform", "linear",
                                dict(method="rk4", options=dict(step_size=0.25)), shape=[45, 16, 9]).to(dev)
    model.device, model.dtype = dev, torch.float32
    with torch.no_grad():   # adaLN-Zero leaves the blocks inert at initialisation: give them something to learn from
        for name, p in model.net.named_parameters():
            if "adaLN" in name or name.startswith("final_layer.linear"):
                p.normal_(0, 0.02)
    ema = v4.ExponentialMovingAverage(model.net.parameters(), decay=0.9)
    opt = v4.FusedAdamW(model.net, lr=2e-3, weight_decay=0.0, max_grad_norm=10.0, ema=ema)
    gen = torch.Generator(device=dev).manual_seed(5)
    first, last = [], []
    for epoch in range(6):
        for x, cond in train.batches(48, shuffle=True, drop_last=True, generator=gen):
            opt.zero_grad(set_to_none=True)
            loss = model._batch_loss((x, cond))
            loss.backward()
            opt.step()
            (first if epoch == 0 else last if epoch == 5 else []).append(loss.item())
    assert all(math.isfinite(v) for v in first + last)
    assert sum(last) / len(last) < 0.9 * sum(first) / len(first), (first, last)   # it learns
    # the same step as one CUDA graph, fed from the dataset (no autograd graph of the eager steps may be alive: its
    # gradient-accumulation nodes are bound to the default stream)
    del loss
    x0, c0 = next(train.batches(48, shuffle=False))
    graphed = v4.GraphedTrainStep(model, opt, x0, c0, warmup=1)
    for x, cond in train.batches(48, shuffle=True, drop_last=True, generator=gen):
        assert math.isfinite(graphed.step(x, cond).item())
    with torch.no_grad():
        val_loss = sum(model._batch_loss(b).item() for b in val.batches(48, shuffle=False))
    assert math.isfinite(val_loss)

    # ---- sampling pipeline: energy ratios for given incident energies, then showers, then detector energies
    en = v4.CFM(v4.ParallelTransformer(dict(eo.DS2_ENERGY, precision="bf16")), "uniform", "linear",
                dict(method="rk4", options=dict(step_size=0.25)), shape=[45]).to(dev)
    en.device, en.dtype = dev, torch.float32
    e_cond = val.energy[:, -1:].contiguous()                         # scaled log incident energies
    with torch.inference_mode():
        u = en.sample_batch(e_cond)                                  # untrained: only shapes / finiteness matter here
    assert tuple(u.shape) == (48, 45) and torch.isfinite(u).all()
    cond = torch.cat([val.energy[:, :-1], e_cond], dim=1)            # the validation showers' own u features
    with ema.average_parameters():                                   # sample with the averaged weights, as the reference does
        with torch.inference_mode():
            showers = model.sample_batch(cond)
    assert tuple(showers.shape) == (48, 1, 45, 16, 9) and torch.isfinite(showers).all()
    energies, e_out = fwd.reverse()(showers.squeeze(1), cond)
    assert tuple(energies.shape) == (48, 6480) and (energies >= 0).all() and torch.isfinite(energies).all()
    # post-processing restores the incident energies and shares E_inc * u_0 out over the layers
    assert vo.rel_l2(e_out.cpu().flatten(), e_inc[-48:].flatten()) < 1e-4
    e_tot = raw[-48:].sum(1)
    assert vo.rel_l2(energies.sum(1).cpu(), e_tot) < 1e-3
    # the dataset's own showers come back through the reverse chain
    back, _ = fwd.reverse()(val.layers.squeeze(1), val.energy)
    keep = raw[-48:] > 1e-5 * raw[-48:].reshape(48, 45, 144).sum(-1).repeat_interleave(144, dim=1)
    assert vo.rel_l2(back.cpu()[keep], raw[-48:][keep]) < 1e-3
