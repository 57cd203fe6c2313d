"""Pin the CPU oracle (oracle/vit_oracle.py) against outputs of the unmodified reference.

Golden vectors: tests/golden/*.npz, written by oracle/make_golden.py from the live
reference in the build container.  When /root/reference is present the oracle is also
checked against the live reference at the full ds2 size.
"""
import os

import numpy as np
import pytest
import torch

from oracle import ref_stubs
from oracle import vit_oracle as vo

FP32_TOL = 1e-5  # north_star: rel-L2 <= 1e-5 in fp32


def _load(golden_dir, tag):
    z = np.load(os.path.join(golden_dir, f"net_{tag}.npz"))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    hidden, heads, depth, B = (int(v) for v in z["meta"][:4])
    return z, sd, heads


@pytest.mark.parametrize("name", list(vo.CONFIGS))
def test_patch_maps_bit_exact(golden_dir, name):
    z = np.load(os.path.join(golden_dir, "patch_maps.npz"))
    geom = vo.CONFIGS[name]["geom"]
    x = np.arange(geom.voxels * geom.in_channels, dtype=np.int32).reshape(1, *geom.sample_shape)
    tok = vo.to_patches(x, geom)
    assert tuple(tok.shape[1:]) == tuple(z[name + "_shape"])
    assert np.array_equal(tok[0], z[name])
    assert np.array_equal(vo.from_patches(tok, geom), x)


@pytest.mark.parametrize("tag,name", [("ds2_tiny", "ds2"), ("calogan_tiny", "calogan"), ("ds1_pions_tiny", "ds1_pions")])
def test_forward_loss_grads_sample(golden_dir, tag, name):
    z, sd, heads = _load(golden_dir, tag)
    geom = vo.CONFIGS[name]["geom"]
    x1, c, t = (torch.from_numpy(z[k]) for k in ("x1", "c", "t_fwd"))
    out = vo.vit_forward(sd, vo.to_patches(x1, geom), t, c, heads)
    assert vo.rel_l2(out, torch.from_numpy(z["net_out"])) < FP32_TOL
    w = vo.cfm_forward(sd, x1, t, c, geom, heads)
    assert vo.rel_l2(w, torch.from_numpy(z["wrapper_out"])) < FP32_TOL

    params = {k: v.clone().requires_grad_(k not in ("pos_x", "pos_y", "pos_z")) for k, v in sd.items()}
    loss = vo.cfm_loss(params, x1, c, torch.from_numpy(z["loss_x0"]),
                       torch.from_numpy(z["loss_t"]), geom, heads)
    assert abs(loss.item() - float(z["loss"])) / float(z["loss"]) < FP32_TOL
    loss.backward()
    for k in z.files:
        if k.startswith("grad/"):
            assert vo.rel_l2(params[k[5:]].grad, torch.from_numpy(z[k])) < 2e-5, k

    with torch.no_grad():
        s = vo.sample_batch(sd, c, torch.from_numpy(z["x_T"]), geom, heads)
    assert vo.rel_l2(s, torch.from_numpy(z["sample"])) < FP32_TOL


def test_lemurs_batch_loss(golden_dir):
    """LEMURSCFM._batch_loss permutes (B, R, A, L) batches before the common path (K = 53 conditions)."""
    z, sd, heads = _load(golden_dir, "lemurs_tiny")
    geom = vo.CONFIGS["lemurs"]["geom"]
    x1 = vo.lemurs_to_grid(torch.from_numpy(z["x"]))
    params = {k: v.clone().requires_grad_(k not in ("pos_x", "pos_y", "pos_z")) for k, v in sd.items()}
    loss = vo.cfm_loss(params, x1, torch.from_numpy(z["c"]), torch.from_numpy(z["loss_x0"]),
                       torch.from_numpy(z["loss_t"]), geom, heads)
    assert abs(loss.item() - float(z["loss"])) / float(z["loss"]) < FP32_TOL
    loss.backward()
    for k in z.files:
        if k.startswith("grad/"):
            assert vo.rel_l2(params[k[5:]].grad, torch.from_numpy(z[k])) < 2e-5, k


def test_finetuning_structures(golden_dir):
    """mapped x / c embedders + re-created final layer (reference experiment_finetuning.py:75-165)"""
    z, sd, heads = _load(golden_dir, "finetune_tiny")
    params = {k: v.clone().requires_grad_(k not in ("pos_x", "pos_y", "pos_z")) for k, v in sd.items()}
    y = vo.vit_forward(params, *(torch.from_numpy(z[k]) for k in ("x", "t", "c")), heads)
    assert vo.rel_l2(y, torch.from_numpy(z["net_out"])) < FP32_TOL
    (y * torch.from_numpy(z["wgt"])).sum().backward()
    grads = [k for k in z.files if k.startswith("grad/")]
    assert any(k.startswith("grad/x_embedder.0.") for k in grads) and any(k.startswith("grad/c_embedder.2.0.") for k in grads)
    for k in grads:
        assert vo.rel_l2(params[k[5:]].grad, torch.from_numpy(z[k])) < 2e-5, k


def test_fixed_positional_tables(golden_dir):
    """learn_pos_embed=False: both coordinate systems' tables, and a forward with the table as a buffer"""
    z, sd, heads = _load(golden_dir, "fixed_pos_tiny")
    hidden = int(z["meta"][0])
    num_patches = vo.CONFIGS["ds3"]["param"]["num_patches"]
    for coords in ("cylindrical", "cartesian"):
        got = vo.get_sincos_pos_embed(coords, num_patches, hidden)
        assert vo.rel_l2(got, torch.from_numpy(z["table/" + coords])) < 1e-6, coords
    y = vo.vit_forward(sd, *(torch.from_numpy(z[k]) for k in ("x", "t", "c")), heads)
    assert vo.rel_l2(y, torch.from_numpy(z["net_out"])) < FP32_TOL


def test_time_grid_is_21_points():
    grid = vo.rk4_38_grid(0.05)
    assert grid.numel() == 21 and grid[0] == 0 and grid[-1] == 1


@pytest.mark.skipif(not ref_stubs.reference_available(), reason="live reference not present")
def test_live_reference_full_ds2():
    ref = ref_stubs.load_reference()
    cfg = vo.CONFIGS["ds2"]
    torch.manual_seed(0)
    net = ref.ViT(cfg["param"])
    ref_stubs.rerandomise_zero_init(net)
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 135, 48, generator=g); t = torch.rand(2, 1, generator=g)
    c = torch.rand(2, 46, generator=g)
    with torch.no_grad():
        want = net(x, t, c)
        got = vo.vit_forward(sd, x, t, c, 6)
    assert vo.rel_l2(got, want) < FP32_TOL
    # the oracle's own initialiser must produce the reference's names and shapes
    mine = vo.init_state_dict(cfg["param"])
    assert {k: tuple(v.shape) for k, v in mine.items()} == {k: tuple(v.shape) for k, v in sd.items()}
    assert torch.equal(mine["pos_z"], sd["pos_z"]) and torch.equal(mine["pos_x"], sd["pos_x"])


@pytest.mark.skipif(not ref_stubs.reference_available(), reason="live reference not present")
@pytest.mark.parametrize("name", ["calogan", "calohad", "ds1_photons", "ds1_pions", "ds3", "lemurs"])
def test_live_reference_meshgrid(name):
    ref = ref_stubs.load_reference()
    p = dict(vo.CONFIGS[name]["param"]); p.update(hidden_dim=12, depth=1, num_heads=2)
    net = ref.ViT(p)
    z, y, x = vo.create_meshgrid(p["num_patches"])
    assert torch.equal(z, net.pos_z) and torch.equal(y, net.pos_y) and torch.equal(x, net.pos_x)
