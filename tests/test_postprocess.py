"""Post-processing of sampled showers: the oracle restatement against the golden output
of the reference's own transform objects (CPU), and the fused CUDA kernel against both (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import transforms_oracle as to
from oracle import vit_oracle as vo

CHAIN = {  # reference configs/calochallenge/cfm/calochallenge_ds2.yaml:15-28 (paths replaced by None)
    "NormalizeByElayer": {"ptype": None, "xml_file": "electron"},
    "ScaleTotalEnergy": {"n_layers": 45, "factor": 0.35},
    "CutValues": {"cut": 1.0e-7, "n_layers": 45},
    "ExclusiveLogitTransform": {"delta": 1.0e-6, "rescale": True},
    "GlobalStandardizeFromFile": {"model_dir": None, "eps": 1.0e-6},
    "LogEnergy": {},
    "ScaleEnergy": {"e_min": 6.907755, "e_max": 13.815510},
    "AddFeaturesToCond": {"split_index": 540},
    "Reshape": {"shape": [1, 45, 4, 3]},
}
PARAMS = dict(delta=1.0e-6, cut=1.0e-7, factor=0.35, e_min=6.907755, e_max=13.815510)


def _golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "postprocess_ds2.npz"))
    return z, [int(b) for b in z["bounds"]], float(z["mean"]), float(z["std"])


def _close(got, want, tol=1e-5):
    """rel-L2 on the values, and the zero pattern (the cuts) may only differ where the kept value is tiny"""
    got, want = got.detach().cpu().double(), torch.as_tensor(want).double()
    assert vo.rel_l2(got, want) < tol
    flip = (got == 0) != (want == 0)
    assert flip.float().mean() < 1e-4
    assert (got[flip].abs().max() if flip.any() else 0.0) <= 1e-6 * want.abs().max()


def test_oracle_reverse_chain_matches_the_reference(golden_dir):
    z, bounds, mean, std = _golden(golden_dir)
    showers, e = to.reverse_chain(torch.from_numpy(z["samples"]), torch.from_numpy(z["cond"]), bounds, mean=mean, std=std,
                                  **PARAMS)
    assert torch.equal((showers == 0), torch.from_numpy(z["showers"] == 0))
    assert vo.rel_l2(showers, torch.from_numpy(z["showers"])) < 1e-6
    assert vo.rel_l2(e, torch.from_numpy(z["energies"])) < 1e-6
    # energy bookkeeping of NormalizeByElayer: the layers share out E_inc * u_0 exactly
    u0 = (torch.sigmoid(torch.from_numpy(z["cond"])[:, 0] * std + mean) - 1e-6) / (1 - 2e-6) / 0.35
    full = showers.sum(1) / (e.flatten() * u0)
    assert (full[showers.sum(1) > 0] - 1).abs().max() < 1e-3


def test_fused_chain_validates_its_configuration():
    from vit4hep_b200.postprocess import FusedReverseTransforms
    bounds = list(range(0, 541, 12))
    FusedReverseTransforms(CHAIN, bounds, -7.5, 2.25)
    with pytest.raises(NotImplementedError):
        FusedReverseTransforms({k: CHAIN[k] for k in list(CHAIN)[1:]}, bounds, 0.0, 1.0)
    bad = dict(CHAIN); bad["ExclusiveLogitTransform"] = {"delta": 1e-6, "rescale": False}
    with pytest.raises(NotImplementedError):
        FusedReverseTransforms(bad, bounds, 0.0, 1.0)
    with pytest.raises(ValueError):
        FusedReverseTransforms(CHAIN, list(range(0, 529, 12)), 0.0, 1.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FusedReverseTransforms(CHAIN, bounds, -7.5, 2.25)(torch.zeros(2, 540), torch.zeros(2, 46))


@pytest.mark.gpu
def test_fused_chain_matches_the_reference_golden(golden_dir):
    from vit4hep_b200.postprocess import FusedReverseTransforms
    dev = torch.device("cuda:0")
    z, bounds, mean, std = _golden(golden_dir)
    fused = FusedReverseTransforms(CHAIN, bounds, mean, std)
    samples = torch.from_numpy(z["samples"]).to(dev)
    showers, e = fused(samples.squeeze(1), torch.from_numpy(z["cond"]).to(dev))
    assert showers.shape == (64, 540) and e.shape == (64, 1)
    _close(showers, z["showers"])
    assert vo.rel_l2(e, torch.from_numpy(z["energies"])) < 1e-6
    empty, e0 = fused(samples[:0], torch.from_numpy(z["cond"]).to(dev)[:0])
    assert empty.shape == (0, 540) and e0.shape == (0, 1)


@pytest.mark.gpu
@pytest.mark.parametrize("name,per_layer", [("ds2", 144), ("ds3", 900), ("ragged", None)])
def test_fused_chain_full_size_vs_oracle(name, per_layer):
    """the real ds2 / ds3 layer structure (45 layers of 144 / 900 voxels) and an irregular one, 256 showers"""
    from vit4hep_b200.postprocess import FusedReverseTransforms
    dev = torch.device("cuda:0")
    if per_layer is None:
        sizes = [5, 160, 190, 7, 33] * 9
        bounds = [0]
        for s in sizes:
            bounds.append(bounds[-1] + s)
    else:
        bounds = list(range(0, 45 * per_layer + 1, per_layer))
    V = bounds[-1]
    chain = {k: dict(v) for k, v in CHAIN.items()}
    chain["AddFeaturesToCond"]["split_index"] = V
    chain["Reshape"]["shape"] = [1, V]
    g = torch.Generator().manual_seed(8)
    N = 256
    samples = torch.randn(N, V, generator=g) * 1.5
    cond = torch.cat([torch.randn(N, 45, generator=g) * 1.2 + 3.0, torch.rand(N, 1, generator=g)], dim=1)
    want, want_e = to.reverse_chain(samples, cond, bounds, mean=-7.5, std=2.25, **PARAMS)
    fused = FusedReverseTransforms(chain, bounds, -7.5, 2.25)
    got, got_e = fused(samples.to(dev), cond.to(dev))
    _close(got, want)
    assert vo.rel_l2(got_e, want_e) < 1e-6
    # size-independent property: every layer of a shower sums to its layer energy, the layers to E_inc * u_0
    assert torch.isfinite(got).all() and (got >= 0).all()
