"""The data feed: the oracle's forward transform chain against the golden output of the
reference's own transform objects (CPU), and the fused CUDA kernels + device-resident dataset against both (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_stubs
from oracle import transforms_oracle as to
from oracle import vit_oracle as vo
from tests.test_postprocess import CHAIN

PARAMS = dict(delta=1.0e-6, factor=0.35, e_min=6.907755, e_max=13.815510)
SHAPE = [1, 45, 4, 3]


def _golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "preprocess_ds2.npz"))
    return z, [int(b) for b in z["bounds"]]


def _raw(n, bounds, seed):
    """synthetic raw showers with the features the chain branches on: sparse hits, empty layers, early stops"""
    g = torch.Generator().manual_seed(seed)
    L, V = len(bounds) - 1, bounds[-1]
    raw = torch.exp(torch.randn(n, V, generator=g) * 2.0 + 3.0) * (torch.rand(n, V, generator=g) < 0.3)
    for l in range(L):
        raw[torch.rand(n, generator=g) < 0.1, bounds[l]:bounds[l + 1]] = 0.0
    e_inc = 10.0 ** (3.0 + 3.0 * torch.rand(n, 1, generator=g))
    raw = raw * (e_inc * (0.3 + 0.9 * torch.rand(n, 1, generator=g)) / raw.sum(1, keepdim=True))   # E_tot ~ E_inc
    return raw, e_inc


def test_oracle_forward_chain_matches_the_reference(golden_dir):
    z, bounds = _golden(golden_dir)
    x, cond, mean, std = to.forward_chain(torch.from_numpy(z["showers"]), torch.from_numpy(z["e_inc"]), bounds,
                                          shape=SHAPE, **PARAMS)
    assert abs(mean - float(z["mean"])) < 1e-6 and abs(std - float(z["std"])) < 1e-6
    assert tuple(x.shape) == z["x"].shape and tuple(cond.shape) == z["cond"].shape
    assert vo.rel_l2(x, torch.from_numpy(z["x"])) < 1e-6
    assert vo.rel_l2(cond, torch.from_numpy(z["cond"])) < 1e-6
    # with the statistics given (GlobalStandardizeFromFile after means.npy exists) the result is the same
    x2, cond2, _, _ = to.forward_chain(torch.from_numpy(z["showers"]), torch.from_numpy(z["e_inc"]), bounds, shape=SHAPE,
                                       mean=float(z["mean"]), std=float(z["std"]), **PARAMS)
    assert vo.rel_l2(x2, x) < 1e-6 and vo.rel_l2(cond2, cond) < 1e-6


def test_oracle_forward_then_reverse_restores_the_showers(golden_dir):
    z, bounds = _golden(golden_dir)
    raw, e_inc = torch.from_numpy(z["showers"]), torch.from_numpy(z["e_inc"])
    x, cond, mean, std = to.forward_chain(raw, e_inc, bounds, shape=SHAPE, **PARAMS)
    back, e = to.reverse_chain(x, cond, bounds, mean=mean, std=std, cut=1.0e-7, **PARAMS)
    assert vo.rel_l2(e, e_inc) < 1e-5
    # voxels below the normalised cut (1e-7 of their layer) are dropped by CutValues; the rest comes back
    kept = raw / (raw.reshape(len(raw), 45, -1).sum(-1).repeat_interleave(12, dim=1) + 1e-10) > 2e-7
    assert vo.rel_l2(back[kept], raw[kept]) < 1e-3


@pytest.mark.skipif(not ref_stubs.reference_available(), reason="live reference not present")
def test_oracle_chains_match_the_live_reference_at_ds2_size():
    """the reference's own transform objects (experiments/calochallenge/transforms.py) at the real ds2 geometry, both
    directions, against the restatement (the committed golden covers a small geometry)"""
    ref_stubs.install()
    import importlib
    tr = importlib.import_module("experiments.calochallenge.transforms")
    L, per = 45, 144
    V = L * per
    bounds = np.arange(0, V + 1, per)
    norm = object.__new__(tr.NormalizeByElayer)
    norm.eps, norm.cut, norm.layer_boundaries, norm.n_layers = 1.0e-10, 0.0, bounds, L
    gs = object.__new__(tr.GlobalStandardizeFromFile)
    gs.written, gs.exclude_zeros, gs.eps = False, True, torch.logit(torch.tensor(1.0e-6))
    gs.write = lambda: None
    chain = [norm, tr.ScaleTotalEnergy(factor=0.35, n_layers=L), tr.CutValues(cut=1.0e-7, n_layers=L),
             tr.ExclusiveLogitTransform(delta=1.0e-6, rescale=True), gs, tr.LogEnergy(),
             tr.ScaleEnergy(e_min=6.907755, e_max=13.815510), tr.AddFeaturesToCond(split_index=V),
             tr.Reshape(shape=[1, 45, 16, 9])]
    raw, e_inc = _raw(64, [int(b) for b in bounds], 13)
    x, c = raw.clone(), e_inc.clone()
    for fn in chain:
        x, c = fn(x, c, rank=1)
    xo, co, mean, std = to.forward_chain(raw, e_inc, [int(b) for b in bounds], shape=[1, 45, 16, 9], **PARAMS)
    assert abs(mean - float(gs.mean)) < 1e-6 and abs(std - float(gs.std)) < 1e-6
    assert vo.rel_l2(xo, x) < 1e-6 and vo.rel_l2(co, c) < 1e-6
    xb, cb = x.clone().squeeze(1), c.clone()
    for fn in chain[::-1]:
        xb, cb = fn(xb, cb, rev=True)
    bo, eo = to.reverse_chain(xo, co, [int(b) for b in bounds], mean=mean, std=std, cut=1.0e-7, **PARAMS)
    assert vo.rel_l2(bo, xb) < 1e-5 and vo.rel_l2(eo, cb) < 1e-6
    assert torch.equal(bo == 0, xb == 0)


def test_dataset_host_logic_on_cpu_tensors():
    """splits, indexing and batching of ShowerDataset do not depend on the device (no transform: nothing native runs)"""
    from vit4hep_b200.preprocess import ShowerDataset
    g = torch.Generator().manual_seed(0)
    showers, e = torch.rand(50, 12, generator=g), torch.rand(50, 1, generator=g)
    full = ShowerDataset.from_arrays(showers.numpy(), e.numpy(), split="full", device="cpu")
    trn = ShowerDataset.from_arrays(showers, e, train_val_frac=[0.6, 0.2], split="training", device="cpu")
    val = ShowerDataset.from_arrays(showers, e, train_val_frac=[0.6, 0.2], split="validation", device="cpu")
    assert (len(full), len(trn), len(val)) == (50, 30, 10)
    assert torch.equal(trn.layers, showers[:30]) and torch.equal(val.layers, showers[-10:]) and torch.equal(val.energy, e[-10:])
    assert float(full.max_bounds) == float(showers.max()) and float(full.min_bounds) == float(showers.min())
    x7, e7 = full[7]
    assert torch.equal(x7, showers[7]) and torch.equal(e7, e[7])
    assert [len(xb) for xb, _ in full.batches(16, shuffle=False)] == [16, 16, 16, 2]
    assert [len(xb) for xb, _ in full.batches(16, shuffle=False, drop_last=True)] == [16, 16, 16]
    gen = torch.Generator().manual_seed(1)
    seen = torch.cat([eb for _, eb in full.batches(16, shuffle=True, generator=gen)])
    assert not torch.equal(seen, e) and torch.equal(seen.sort(0).values, e.sort(0).values)
    # data-parallel sharding: same permutation on every rank, disjoint strided shares, equal step counts
    shares = []
    for r in range(3):
        gen = torch.Generator().manual_seed(5)
        shares.append(torch.cat([eb for _, eb in full.batches(4, shuffle=True, generator=gen, rank=r, world_size=3)]))
    assert [len(sh) for sh in shares] == [16, 16, 16]                       # 50 -> 48 = 3 x 16
    both = torch.cat(shares)
    assert len(torch.unique(both)) == 48 and set(both.flatten().tolist()) <= set(e.flatten().tolist())
    fixed = [torch.cat([eb for _, eb in full.batches(8, shuffle=False, rank=r, world_size=2)]) for r in range(2)]
    assert torch.equal(fixed[0], e[0:50:2]) and torch.equal(fixed[1], e[1:50:2])
    with pytest.raises(ValueError):
        next(full.batches(4, rank=2, world_size=2))
    with pytest.raises(ValueError):
        ShowerDataset.from_arrays(showers, e, split="test", device="cpu")
    empty = ShowerDataset.from_arrays(showers[:0], e[:0], device="cpu")
    assert len(empty) == 0 and empty.min_bounds is None and list(empty.batches(4)) == []


def test_layer_boundaries_from_xml(tmp_path):
    from vit4hep_b200.preprocess import layer_boundaries_from_xml
    xml = tmp_path / "binning.xml"
    xml.write_text('<Bins><Bin pid="22" name="photon">'
                   '<Layer id="0" r_edges="0,5,10,30" n_bin_alpha="1"/>'
                   '<Layer id="1" r_edges="0,2,4,6,8" n_bin_alpha="10"/>'
                   '<Layer id="2" r_edges="0" n_bin_alpha="1"/>'
                   '<Layer id="3" r_edges="0,1,2" n_bin_alpha="4"/></Bin>'
                   '<Bin pid="11" name="electron"><Layer id="0" r_edges="0,1,2" n_bin_alpha="16"/></Bin></Bins>')
    assert layer_boundaries_from_xml(str(xml), "photon").tolist() == [0, 3, 43, 51]     # the empty layer is dropped
    assert layer_boundaries_from_xml(str(xml), "electron").tolist() == [0, 32]
    with pytest.raises(ValueError):
        layer_boundaries_from_xml(str(xml), "pion")


def test_forward_transforms_validate_their_configuration(tmp_path):
    from vit4hep_b200.preprocess import FusedForwardTransforms, ShowerDataset
    bounds = list(range(0, 541, 12))
    f = FusedForwardTransforms(CHAIN, bounds)
    assert not f.written and f.shape == SHAPE
    with pytest.raises(RuntimeError):
        f.reverse()
    with pytest.raises(ValueError):
        FusedForwardTransforms(CHAIN, bounds, mean=1.0)
    bad = dict(CHAIN); bad["GlobalStandardizeFromFile"] = {"model_dir": None, "exclude_zeros": False}
    with pytest.raises(NotImplementedError):
        FusedForwardTransforms(bad, bounds)
    with pytest.raises(NotImplementedError):
        FusedForwardTransforms({k: CHAIN[k] for k in list(CHAIN)[:-1]}, bounds)
    # means.npy / stds.npy under model_dir are picked up like GlobalStandardizeFromFile does
    np.save(tmp_path / "means.npy", np.float32(-2.5)); np.save(tmp_path / "stds.npy", np.float32(3.0))
    cfg = dict(CHAIN); cfg["GlobalStandardizeFromFile"] = {"model_dir": str(tmp_path), "eps": 1.0e-6}
    f = FusedForwardTransforms(cfg, bounds)
    assert f.written and (f.mean, f.std) == (-2.5, 3.0)
    assert (f.reverse().mean, f.reverse().std) == (-2.5, 3.0)
    with pytest.raises(RuntimeError):   # no CPU fallback
        f(torch.zeros(2, 540), torch.ones(2, 1))
    with pytest.raises(ImportError):    # no h5py in this image: the HDF5 entry points say so
        ShowerDataset("/nonexistent.hdf5")
    from vit4hep_b200.preprocess import save_hdf5
    with pytest.raises(ImportError):
        save_hdf5(str(tmp_path / "samples.hdf5"), torch.zeros(2, 540), torch.ones(2, 1))


# --------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_fused_forward_matches_reference_golden_and_oracle(golden_dir):
    from vit4hep_b200.preprocess import FusedForwardTransforms
    z, bounds = _golden(golden_dir)
    raw, e_inc = torch.from_numpy(z["showers"]).cuda(), torch.from_numpy(z["e_inc"]).cuda()
    f = FusedForwardTransforms(CHAIN, bounds)                       # statistics computed on the device
    x, cond = f(raw, e_inc)
    assert tuple(x.shape) == z["x"].shape and tuple(cond.shape) == z["cond"].shape
    assert abs(f.mean - float(z["mean"])) < 2e-6 * abs(float(z["mean"])) + 1e-6
    assert abs(f.std - float(z["std"])) < 2e-6 * float(z["std"])
    assert vo.rel_l2(x.cpu(), torch.from_numpy(z["x"])) < 1e-6
    assert vo.rel_l2(cond.cpu(), torch.from_numpy(z["cond"])) < 1e-6
    # element-wise: the layer sums are added in another order than torch's, so voxel / layer energy is off by a few
    # ulp, which the logit amplifies by its slope 1 / (z (1 - z)) — 1e6 where one voxel holds all of its layer
    want = torch.from_numpy(z["x"])
    zz = torch.sigmoid(want.double() * float(z["std"]) + float(z["mean"]))
    tol = 1e-5 + 5e-7 / (zz * (1 - zz)).clamp_min(1e-7) / float(z["std"])
    err = (x.cpu().double() - want.double()).abs()
    worst = (err / tol).argmax()
    assert (err <= tol).all(), (float(err.flatten()[worst]), float(tol.flatten()[worst]), float(want.flatten()[worst]))
    g = FusedForwardTransforms(CHAIN, bounds, mean=float(z["mean"]), std=float(z["std"]))   # one-kernel path
    x2, cond2 = g(raw, e_inc)
    assert vo.rel_l2(x2.cpu(), torch.from_numpy(z["x"])) < 1e-6
    assert vo.rel_l2(cond2.cpu(), torch.from_numpy(z["cond"])) < 1e-6
    x0, c0 = g(raw[:0], e_inc[:0])
    assert tuple(x0.shape) == (0, *SHAPE) and tuple(c0.shape) == (0, 46)


@pytest.mark.gpu
def test_fused_forward_ds2_size_against_the_oracle_and_round_trip():
    """the real ds2 geometry (45 layers x 144 voxels), 2 000 showers: oracle parity, then forward -> reverse on the
    device restores the raw showers (the size-independent property of the chain)"""
    from vit4hep_b200.preprocess import FusedForwardTransforms
    bounds = list(range(0, 6481, 144))
    chain = dict(CHAIN); chain["AddFeaturesToCond"] = {"split_index": 6480}; chain["Reshape"] = {"shape": [1, 45, 16, 9]}
    raw, e_inc = _raw(2000, bounds, 5)
    f = FusedForwardTransforms(chain, bounds)
    x, cond = f(raw.cuda(), e_inc.cuda())
    xo, co, mean, std = to.forward_chain(raw, e_inc, bounds, shape=[1, 45, 16, 9], **PARAMS)
    assert abs(f.mean - mean) < 1e-5 and abs(f.std - std) < 1e-5
    assert vo.rel_l2(x.cpu(), xo) < 1e-5 and vo.rel_l2(cond.cpu(), co) < 1e-5
    back, e = f.reverse()(x, cond)
    assert vo.rel_l2(e.cpu(), e_inc) < 1e-5
    layer_e = raw.reshape(2000, 45, 144).sum(-1).repeat_interleave(144, dim=1)
    kept = raw / (layer_e + 1e-10) > 2e-7
    assert vo.rel_l2(back.cpu()[kept], raw[kept]) < 1e-3
    assert (back.cpu()[raw == 0] == 0).all()


@pytest.mark.gpu
def test_device_resident_dataset_splits_and_batches(golden_dir):
    from vit4hep_b200.preprocess import FusedForwardTransforms, ShowerDataset
    z, bounds = _golden(golden_dir)
    mk = lambda split: ShowerDataset.from_arrays(z["showers"], z["e_inc"], train_val_frac=[0.75, 0.25], split=split,
                                                 transform=FusedForwardTransforms(CHAIN, bounds))
    full, trn, val = mk("full"), mk("training"), mk("validation")
    assert (len(full), len(trn), len(val)) == (96, 72, 24)
    assert full.layers.is_cuda and tuple(full.layers.shape) == (96, *SHAPE) and tuple(full.energy.shape) == (96, 46)
    # the statistics are those of the whole file whatever the split (reference datasets.py:44-61)
    assert torch.equal(trn.layers, full.layers[:72]) and torch.equal(val.energy, full.energy[-24:])
    assert float(full.min_bounds) == float(full.layers.min()) and float(full.max_bounds) == float(full.layers.max())
    x0, c0 = full[5]
    assert torch.equal(x0, full.layers[5]) and torch.equal(c0, full.energy[5])
    # one shuffled epoch visits every shower once, batches stay paired, drop_last drops the ragged tail
    gen = torch.Generator(device="cuda").manual_seed(3)
    seen = []
    for xb, cb in full.batches(20, shuffle=True, generator=gen):
        assert xb.is_cuda and xb.shape[0] == cb.shape[0] <= 20
        seen.append(cb)
    seen = torch.cat(seen)
    assert seen.shape[0] == 96 and not torch.equal(seen, full.energy)
    assert torch.equal(seen.sort(dim=0).values, full.energy.sort(dim=0).values)
    assert sum(xb.shape[0] for xb, _ in full.batches(20, drop_last=True)) == 80
    assert [xb.shape[0] for xb, _ in full.batches(40, shuffle=False)] == [40, 40, 16]


@pytest.mark.gpu
@pytest.mark.parametrize("layers,per", [(5, 900), (3, 1500), (7, 33)])
def test_fused_chains_for_other_layer_sizes(layers, per):
    """layers of 900 voxels (ds3: the 32-registers-per-lane kernels), 1500 (generic two-sweep kernels) and a ragged 33:
    forward and reverse against the oracle, and the round trip"""
    from vit4hep_b200.preprocess import FusedForwardTransforms
    V = layers * per
    bounds = list(range(0, V + 1, per))
    chain = dict(CHAIN)
    chain["ScaleTotalEnergy"] = {"n_layers": layers, "factor": 0.35}
    chain["CutValues"] = {"cut": 1.0e-7, "n_layers": layers}
    chain["AddFeaturesToCond"] = {"split_index": V}
    chain["Reshape"] = {"shape": [1, layers, per]}
    raw, e_inc = _raw(300, bounds, 7)
    f = FusedForwardTransforms(chain, bounds)
    x, cond = f(raw.cuda(), e_inc.cuda())
    xo, co, mean, std = to.forward_chain(raw, e_inc, bounds, shape=[1, layers, per], **PARAMS)
    assert abs(f.mean - mean) < 1e-5 and abs(f.std - std) < 1e-5
    assert vo.rel_l2(x.cpu(), xo) < 1e-5 and vo.rel_l2(cond.cpu(), co) < 1e-5
    back, e = f.reverse()(x, cond)
    bo, eo = to.reverse_chain(xo, co, bounds, mean=mean, std=std, cut=1.0e-7, **PARAMS)
    assert vo.rel_l2(e.cpu(), eo) < 1e-5 and vo.rel_l2(back.cpu(), bo) < 1e-4
    layer_e = raw.reshape(300, layers, per).sum(-1).repeat_interleave(per, dim=1)
    kept = raw / (layer_e + 1e-10) > 2e-7
    assert vo.rel_l2(back.cpu()[kept], raw[kept]) < 1e-3


@pytest.mark.skipif(not ref_stubs.reference_available(), reason="live reference not present")
@pytest.mark.parametrize("cfg,voxels,shape", [("calochallenge_ds2.yaml", 6480, [1, 45, 16, 9]),
                                               ("calochallenge_ds3.yaml", 40500, [1, 45, 50, 18])])
def test_transform_objects_accept_the_reference_yaml_unchanged(cfg, voxels, shape):
    """the `data.transforms` mapping of the reference's shipped experiment configs builds both fused chains as it is"""
    import yaml
    from vit4hep_b200 import FusedForwardTransforms, FusedReverseTransforms
    with open(os.path.join(ref_stubs.REFERENCE_ROOT, "configs", "calochallenge", "cfm", cfg)) as fh:
        transforms = yaml.safe_load(fh)["data"]["transforms"]
    bounds = list(range(0, voxels + 1, voxels // 45))
    fwd = FusedForwardTransforms(transforms, bounds, mean=-2.0, std=3.0)
    rev = FusedReverseTransforms(transforms, bounds, -2.0, 3.0)
    assert fwd.shape == shape and fwd.voxels == voxels == rev.voxels and fwd.n_layers == 45
    assert (rev.factor, rev.delta, rev.cut) == (0.35, 1.0e-6, 1.0e-7) and rev.max_layer == voxels // 45
    assert (fwd.reverse().mean, fwd.reverse().std) == (-2.0, 3.0)
