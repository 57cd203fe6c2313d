"""CPU-side checks (no GPU): the C-ABI library loads and exports what include/vit4hep_b200.h declares,
the host-side index tables equal the oracle and the reference's golden vectors, the module mirrors the
reference's state_dict, and there is no CPU fallback."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import vit_oracle as vo
from tests.helpers import geometry_of

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "vit4hep_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(v4h_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from vit4hep_b200 import _cabi
    lib = _cabi.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _cabi.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_cabi.SIGNATURES) == set(names)
    assert lib.v4h_version() >= 100


@pytest.mark.parametrize("name", list(vo.CONFIGS))
def test_index_table_matches_oracle_and_golden(golden_dir, name):
    g = geometry_of(name)
    og = vo.CONFIGS[name]["geom"]
    table = g.index_table()
    assert table.dtype == np.int32 and table.shape == (og.voxels * og.in_channels,)
    assert np.array_equal(table.astype(np.int64), vo.patch_index_map(og))
    assert (g.tokens, g.patch_dim) == (og.tokens, og.patch_dim)
    z = np.load(os.path.join(golden_dir, "patch_maps.npz"))
    # the golden file is to_patches(arange) of the unmodified reference = the gather table itself
    assert np.array_equal(table.reshape(g.tokens, g.patch_dim), z[name])


def test_index_table_multichannel_matches_oracle():
    og = vo.Geometry((vo.Segment((6, 4, 6), (3, 2, 2)),), in_channels=3)
    from vit4hep_b200.cfm import PatchGeometry
    g = PatchGeometry([(6, 4, 6)], [(3, 2, 2)], 3, False)
    assert np.array_equal(g.index_table().astype(np.int64), vo.patch_index_map(og))


def test_geometry_rejects_indivisible_patch_like_the_reference():
    from vit4hep_b200.cfm import PatchGeometry
    with pytest.raises(AssertionError, match="should be divisible by patch size"):
        PatchGeometry([(45, 16, 9)], [(4, 16, 1)])


@pytest.mark.parametrize("name", ["ds2", "calogan"])
def test_state_dict_names_and_shapes_match_the_reference(name):
    import vit4hep_b200 as v4
    param = vo.CONFIGS[name]["param"]
    net = v4.ViT(param)
    mine = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    want = {k: tuple(v.shape) for k, v in vo.init_state_dict(param).items()}
    assert mine == want
    # same buffers as the reference (oracle's create_meshgrid is pinned against it)
    z, y, x = vo.create_meshgrid(param["num_patches"])
    assert torch.equal(net.pos_z, z) and torch.equal(net.pos_y, y) and torch.equal(net.pos_x, x)
    # adaLN-Zero init: modulation and output layers are zero, everything else is not
    assert net.final_layer.linear.weight.abs().max() == 0
    assert net.blocks[0].adaLN_modulation[1].weight.abs().max() == 0
    assert net.blocks[0].attn.qkv.weight.abs().max() > 0
    # the flat gradient order covers every parameter exactly once
    ordered = net.ordered_parameters()
    assert {id(p) for _, p in ordered} == {id(p) for p in net.parameters()}
    assert len(ordered) == len(list(net.parameters()))
    bounds = net.stage_boundaries()
    offs, total = net.flat_layout(ordered)
    nparams = sum(p.numel() for p in net.parameters())
    # gradients start 16-byte aligned in the flat buffer: at most 3 padding elements per parameter
    assert len(bounds) == len(net.blocks) + 2 and bounds[-1] == total and nparams <= total <= nparams + 3 * len(ordered)
    assert all(o % 4 == 0 for o in offs) and bounds == sorted(bounds)
    assert all(b in set(offs) | {total} for b in bounds)  # stage boundaries fall between parameters
    if name == "ds2":
        assert nparams == 26_042_528 == total  # parameter count of the ds2 network


def test_reference_state_dict_loads(golden_dir):
    import vit4hep_b200 as v4
    from tests.helpers import load_golden
    z, sd, over, B = load_golden(golden_dir, "ds2_tiny")
    p = dict(vo.CONFIGS["ds2"]["param"]); p.update(over)
    net = v4.ViT(p)
    missing, unexpected = net.load_state_dict(sd, strict=True)
    assert not missing and not unexpected


def test_no_cpu_fallback():
    import vit4hep_b200 as v4
    p = dict(vo.CONFIGS["ds2"]["param"]); p.update(hidden_dim=48, depth=1, num_heads=2)
    net = v4.ViT(p)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(1, 135, 48), torch.zeros(1, 1), torch.zeros(1, 46))
    with pytest.raises(RuntimeError, match="no fallback"):
        net.blocks[0](torch.zeros(1, 135, 48), torch.zeros(1, 48))
    g = geometry_of("ds2")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        g.to_patches(torch.zeros(1, 1, 45, 16, 9))


def test_unsupported_knobs_fail_loudly():
    import vit4hep_b200 as v4
    base = dict(vo.CONFIGS["ds2"]["param"]); base.update(hidden_dim=48, depth=1, num_heads=2)
    for bad in (dict(attn_drop=0.1), dict(causal_attn=True), dict(precision="fp8")):
        with pytest.raises((NotImplementedError, ValueError)):
            v4.ViT({**base, **bad})


def test_fixed_grid_is_torchdiffeq_grid():
    from vit4hep_b200.cfm import fixed_grid
    assert torch.equal(fixed_grid(0.05), vo.rk4_38_grid(0.05))
    assert fixed_grid(0.3).tolist() == pytest.approx([0.0, 0.3, 0.6, 0.9, 1.0])


def test_fixed_pos_embed_tables():
    from vit4hep_b200.vit import get_sincos_pos_embed
    pe = get_sincos_pos_embed("cylindrical", [[15, 1, 9]], 48, 3)
    assert pe.shape == (135, 48)
    # token 0 sits at the origin: sin = 0, cos = 1
    assert torch.allclose(pe[0, 0:8], torch.zeros(8)) and torch.allclose(pe[0, 8:16], torch.ones(8))
