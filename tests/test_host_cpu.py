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


# ----------------------------------------------------------------------------------------------
# Hydra-free instantiation from the reference's own YAML
# ----------------------------------------------------------------------------------------------
REF_CFG = "/root/reference/configs/model"


@pytest.mark.skipif(not os.path.isdir(REF_CFG), reason="reference configs not present (GPU box)")
@pytest.mark.parametrize("rel,name", [("cfm/cfm_ds2_electrons.yaml", "ds2"), ("cfm/cfm_ds3_electrons.yaml", "ds3"),
                                      ("cfm/cfm_ds1_photons.yaml", "ds1_photons"), ("cfm/cfm_ds1_pions.yaml", "ds1_pions"),
                                      ("cfm_calogan/cfm_eplus.yaml", "calogan"), ("cfm_calohad/cfm_calohad.yaml", "calohad"),
                                      ("cfm_lemurs/cfm_lemurs.yaml", "lemurs")])
def test_instantiate_from_the_reference_yaml(rel, name):
    """The reference's model YAML, unchanged except for its two ``_target_`` strings (here remapped on the fly),
    builds the B200 drop-ins; the tables in vit4hep_b200.configs restate exactly these files."""
    import yaml
    import vit4hep_b200 as v4
    from vit4hep_b200 import configs
    with open(os.path.join(REF_CFG, rel)) as fh:
        cfg = yaml.safe_load(fh)
    model = configs.instantiate(cfg, remap=True)
    assert type(model).__module__.startswith("vit4hep_b200") and isinstance(model.net, v4.ViT)
    og = vo.CONFIGS[name]["geom"]
    assert (model.geometry.tokens, model.geometry.patch_dim, model.geometry.voxels) == (og.tokens, og.patch_dim, og.voxels)
    assert model.net.pos_z.numel() == og.tokens
    # the restated table builds the same thing
    mine = configs.MODELS[name]
    want = dict(cfg); want["_target_"] = configs.TARGETS[cfg["_target_"]]
    want["net"] = dict(cfg["net"]); want["net"]["_target_"] = configs.TARGETS[cfg["net"]["_target_"]]
    want["net"]["param"] = {k: v for k, v in cfg["net"]["param"].items() if k != "use_rotary_emb"}
    assert {k: v for k, v in mine.items() if k != "net"} == {k: v for k, v in want.items() if k != "net"}
    assert mine["net"]["param"] == want["net"]["param"]
    # edited-by-hand variant of INTEGRATION.md: explicit drop-in targets, no remapping
    same = configs.instantiate(want)
    assert {k: tuple(v.shape) for k, v in same.net.state_dict().items()} == \
        {k: tuple(v.shape) for k, v in model.net.state_dict().items()}


def test_finetuning_structures_are_accepted_on_the_host():
    """x / c embedders behind mapper Linears and a swapped final layer (reference experiment_finetuning.py:75-165)
    change the parameter list and the plan key; anything else still raises NotImplementedError."""
    import torch.nn as nn
    import vit4hep_b200 as v4
    p = dict(vo.tiny_config("ds2", hidden_dim=48, depth=1, num_heads=2)["param"])
    net = v4.ViT(p)
    n0 = len(net.ordered_parameters())
    net.x_embedder = nn.Sequential(nn.Linear(90, p["patch_dim"]), nn.SiLU(), net.x_embedder)
    net.c_embedder = nn.Sequential(nn.Linear(40, p["condition_dim"]), nn.SiLU(), net.c_embedder)
    fields = [f for f, _ in net.ordered_parameters()]
    assert len(fields) == n0 + 4 and {"xm_w", "xm_b", "cm_w", "cm_b"} <= set(fields)
    assert {id(q) for _, q in net.ordered_parameters()} == {id(q) for q in net.parameters()}
    offs, total = net.flat_layout(net.ordered_parameters())
    assert net.stage_boundaries()[-1] == total
    net.c_embedder = nn.Sequential(nn.Linear(40, 48), nn.ReLU(), nn.Linear(48, 48))
    with pytest.raises(NotImplementedError):
        net.ordered_parameters()


def test_stage_boundaries_follow_backward_completion():
    """final layer (with its adaLN Linear), blocks depth-1..0 (each with its adaLN Linear), stage 0"""
    import vit4hep_b200 as v4
    p = dict(vo.tiny_config("ds2", hidden_dim=48, depth=3, num_heads=2)["param"])
    net = v4.ViT(p)
    ordered = net.ordered_parameters()
    offs, total = net.flat_layout(ordered)
    bounds = net.stage_boundaries()
    assert len(bounds) == 3 + 2 and bounds[-1] == total and bounds == sorted(bounds)
    names = [f for f, _ in ordered]
    ends = dict(zip(names, offs[1:] + [total]))
    assert bounds[0] == ends["final_ada_b"]
    assert bounds[1] == ends["blocks.2.ada_b"] and bounds[3] == ends["blocks.0.ada_b"]
    stage0 = names[names.index("blocks.0.ada_b") + 1:]
    assert stage0[0] == "pos_embed_freqs" and not any(n.startswith("blocks.") or "ada" in n for n in stage0)
    from vit4hep_b200 import dp
    full = v4.ViT(dict(vo.CONFIGS["ds2"]["param"]))
    b = dp.plan_buckets(full.stage_boundaries(), 4_000_000)
    assert (b[-1].stage_begin, b[-1].stage_end) == (0, 0) and b[-1].stop - b[-1].start < 1_000_000
    assert len(b) == 7 and all(x.stop - x.start >= 4_000_000 for x in b[:-1])


def test_ema_host_logic_matches_torch_ema_semantics():
    """ExponentialMovingAverage mirrors torch_ema's container behaviour (state_dict keys, store / restore /
    average_parameters, to()); the update itself is a CUDA kernel and raises on CPU tensors (no fallback)."""
    import vit4hep_b200 as v4
    params = [torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(5))]
    ema = v4.ExponentialMovingAverage(params, decay=0.99)
    sd = ema.state_dict()
    assert set(sd) == {"decay", "num_updates", "shadow_params", "collected_params"} and sd["num_updates"] == 0
    assert all(torch.equal(s, p) for s, p in zip(sd["shadow_params"], params))
    with torch.no_grad():
        params[0].add_(1.0)
    with ema.average_parameters():
        assert torch.equal(params[0], ema.shadow_params[0])
    assert not torch.equal(params[0], ema.shadow_params[0])
    other = v4.ExponentialMovingAverage(params, decay=0.5)
    other.load_state_dict(sd)
    assert other.decay == 0.99 and torch.equal(other.shadow_params[1], ema.shadow_params[1])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ema.update()
    with pytest.raises(ValueError):
        v4.ExponentialMovingAverage(params, decay=1.5)
