"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

TEST INFRASTRUCTURE.  Run once here (``python -m oracle.make_golden``); the vectors are
committed because ``/root/reference`` does not exist on the GPU box.  The reference files
are imported from /root/reference through the stubs in ``oracle/ref_stubs.py``; nothing
is copied.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_stubs  # noqa: E402
from oracle.vit_oracle import CONFIGS, tiny_config  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def build_wrapper(ref, name, param, geom):
    net = ref.ViT(param)
    ode = dict(method="rk4", options=dict(step_size=0.05))
    segs = geom.segments
    if name in ("ds2", "ds3"):
        m = ref.CaloChallengeCFM(net, list(segs[0].patch), 1, "uniform", "linear", ode,
                                 shape=list(segs[0].shape))
    elif name == "lemurs":
        m = ref.LEMURSCFM(net, list(segs[0].patch), 1, "uniform", "linear", ode, shape=list(segs[0].shape))
    elif name in ("ds1_photons", "ds1_pions"):
        m = ref.CaloChallengeCFM_DS1(net, [list(s.shape) for s in segs], [s.voxels for s in segs],
                                     list(segs[0].patch), 1, "uniform", "linear", ode,
                                     shape=[geom.voxels])
    elif name == "calogan":
        m = ref.CaloGANCFM(net, [list(s.shape) for s in segs], [s.voxels for s in segs],
                           [list(s.patch) for s in segs], 1, "uniform", "linear", ode,
                           shape=[geom.voxels])
    elif name == "calohad":
        m = ref.CaloHadCFM(net, [list(s.shape) for s in segs], [s.voxels for s in segs],
                           [list(s.patch) for s in segs], 1, "uniform", "linear", ode,
                           shape=[geom.voxels])
    else:
        raise KeyError(name)
    m.device, m.dtype = torch.device("cpu"), torch.float32
    return m


def golden_patch_maps(ref):
    """to_patches of an arange-valued shower for every shipped geometry (exact integers)."""
    out = {}
    for name, cfg in CONFIGS.items():
        geom = cfg["geom"]
        tiny = dict(cfg["param"]); tiny.update(hidden_dim=12, depth=1, num_heads=2)
        m = build_wrapper(ref, name, tiny, geom)
        x = torch.arange(2 * geom.voxels, dtype=torch.float32).reshape(2, *geom.sample_shape)
        tok = m.to_patches(x)
        back = m.from_patches(tok)
        assert torch.equal(back.reshape(2, -1), x.reshape(2, -1))
        out[name] = tok[0].to(torch.int32).numpy()
        out[name + "_shape"] = np.asarray(tok.shape[1:], dtype=np.int32)
    np.savez_compressed(os.path.join(OUT, "patch_maps.npz"), **out)


def golden_net(ref, name, hidden, heads, depth, B, tag):
    """Reference ViT forward, CFM loss + every parameter gradient, and RK4 sampling."""
    cfg = tiny_config(name, hidden_dim=hidden, depth=depth, num_heads=heads)
    geom, param = cfg["geom"], cfg["param"]
    torch.manual_seed(0)
    model = build_wrapper(ref, name, param, geom)
    ref_stubs.rerandomise_zero_init(model.net, seed=1)
    K = param["condition_dim"]
    g = torch.Generator().manual_seed(1234)
    x1 = torch.randn(B, *geom.sample_shape, generator=g)
    c = torch.rand(B, K, generator=g)
    t_fwd = torch.rand(B, 1, generator=g)

    out = {"sd/" + k: v.detach().numpy() for k, v in model.net.state_dict().items()}
    out.update(x1=x1.numpy(), c=c.numpy(), t_fwd=t_fwd.numpy())

    # (1) network forward on patchified input, and wrapper forward
    with torch.no_grad():
        tok = model.to_patches(x1)
        out["net_out"] = model.net(tok, t_fwd, c).numpy()
        out["wrapper_out"] = model.forward(x1, t_fwd, c).numpy()

    # (2) _batch_loss: replay its RNG draws (models/base_model.py:209-212) to record t, x0
    torch.manual_seed(77)
    t = model.time_distribution.sample([B] + [1] * (x1.dim() - 1))
    x0 = torch.randn_like(x1)
    torch.manual_seed(77)
    loss = model._batch_loss((x1, c))
    loss.backward()
    out.update(loss_t=t.numpy(), loss_x0=x0.numpy(), loss=np.asarray(loss.item(), dtype=np.float64))
    for k, p in model.net.named_parameters():
        out["grad/" + k] = p.grad.numpy()

    # (3) sample_batch: replay x_T (calochallenge_cfm/model.py:77)
    torch.manual_seed(99)
    x_T = torch.randn((B, 1, *model.shape))
    torch.manual_seed(99)
    sample = model.sample_batch(c)
    out.update(x_T=x_T.numpy(), sample=sample.numpy())
    out["meta"] = np.asarray([hidden, heads, depth, B], dtype=np.int32)
    np.savez_compressed(os.path.join(OUT, f"net_{tag}.npz"), **out)
    print(tag, "loss", loss.item(), "sample rms", sample.pow(2).mean().sqrt().item())


def golden_lemurs(ref, hidden=48, heads=2, depth=2, B=3):
    """LEMURSCFM._batch_loss on a (B, R, A, L) batch (reference experiments/lemurs/model.py:62-65): loss and
    every parameter gradient, plus sample_batch."""
    cfg = tiny_config("lemurs", hidden_dim=hidden, depth=depth, num_heads=heads)
    geom, param = cfg["geom"], cfg["param"]
    torch.manual_seed(0)
    model = build_wrapper(ref, "lemurs", param, geom)
    ref_stubs.rerandomise_zero_init(model.net, seed=1)
    g = torch.Generator().manual_seed(4321)
    L, A, R = geom.segments[0].shape
    x = torch.randn(B, R, A, L, generator=g)
    c = torch.rand(B, param["condition_dim"], generator=g)
    out = {"sd/" + k: v.detach().numpy() for k, v in model.net.state_dict().items()}
    out.update(x=x.numpy(), c=c.numpy())
    torch.manual_seed(55)
    t = model.time_distribution.sample([B, 1, 1, 1, 1])
    # randn_like of the PERMUTED view: the draw follows its (preserved) strides, not the logical order
    x0 = torch.randn_like(x.permute(0, 3, 2, 1).unsqueeze(1)).contiguous()
    torch.manual_seed(55)
    loss = model._batch_loss([x.clone(), c])
    loss.backward()
    out.update(loss_t=t.numpy(), loss_x0=x0.numpy(), loss=np.asarray(loss.item(), dtype=np.float64))
    for k, p in model.net.named_parameters():
        out["grad/" + k] = p.grad.numpy()
    out["meta"] = np.asarray([hidden, heads, depth, B], dtype=np.int32)
    np.savez_compressed(os.path.join(OUT, "net_lemurs_tiny.npz"), **out)
    print("lemurs_tiny loss", loss.item())


def golden_finetune(ref, hidden=48, heads=2, depth=2, B=2):
    """The module surgery of reference experiments/calochallenge/calochallenge_cfm/experiment_finetuning.py:75-165
    on a ds2 backbone that is finetuned to the ds3 grid: x_embedder / c_embedder wrapped behind mapper Linears
    (map_x_embedding / map_c_embedding), pos grid rebuilt for the new num_patches, final layer re-created for
    the new patch_dim.  Records the network output and every parameter gradient of sum(out * wgt)."""
    import torch.nn as nn
    back = tiny_config("ds2", hidden_dim=hidden, depth=depth, num_heads=heads)["param"]
    new = tiny_config("ds3", hidden_dim=hidden, depth=depth, num_heads=heads)
    geom_new, p_new = new["geom"], new["param"]
    torch.manual_seed(0)
    net = ref.ViT(back)
    # experiment_finetuning.py:79-91 (map_x_embedding) and :106-118 (map_c_embedding); the new task has 40 conditions
    K_new = 40
    net.x_embedder = nn.Sequential(nn.Linear(p_new["patch_dim"], back["patch_dim"]), nn.SiLU(), net.x_embedder)
    net.c_embedder = nn.Sequential(nn.Linear(K_new, back["condition_dim"]), nn.SiLU(), net.c_embedder)
    # :135-142 positional grid of the new geometry
    net.num_patches = p_new["num_patches"]
    pos_z, pos_y, pos_x = net.create_meshgrid()
    net.pos_z, net.pos_y, net.pos_x = pos_z, pos_y, pos_x
    # :160-165 new final layer
    net.final_layer = ref.vit.FinalLayer(hidden, p_new["patch_dim"], back["out_channels"])
    ref_stubs.rerandomise_zero_init(net, seed=2)
    g = torch.Generator().manual_seed(99)
    x = torch.randn(B, geom_new.tokens, p_new["patch_dim"], generator=g)
    c = torch.rand(B, K_new, generator=g)
    t = torch.rand(B, 1, generator=g)
    wgt = torch.randn(B, geom_new.tokens, p_new["patch_dim"], generator=g)
    out = {"sd/" + k: v.detach().numpy() for k, v in net.state_dict().items()}
    y = net(x, t, c)
    (y * wgt).sum().backward()
    out.update(x=x.numpy(), c=c.numpy(), t=t.numpy(), wgt=wgt.numpy(), net_out=y.detach().numpy())
    for k, p in net.named_parameters():
        out["grad/" + k] = p.grad.numpy()
    out["meta"] = np.asarray([hidden, heads, depth, B, K_new], dtype=np.int32)
    np.savez_compressed(os.path.join(OUT, "net_finetune_tiny.npz"), **out)
    print("finetune_tiny out rms", y.pow(2).mean().sqrt().item())


def golden_fixed_pos_embed(ref, hidden=48, heads=2, depth=1, B=2):
    """learn_pos_embed=False (reference nn/vit.py:92-103, :461-540): the fixed tables for both coordinate
    systems and one forward with the cylindrical one."""
    out = {}
    for coords in ("cylindrical", "cartesian"):
        cfg = tiny_config("ds3", hidden_dim=hidden, depth=depth, num_heads=heads)
        # the fixed-table branch unpacks num_patches as ONE (L, A, R) triple (reference nn/vit.py:497)
        param = dict(cfg["param"]); param.update(learn_pos_embed=False, pos_embedding_coords=coords,
                                                 num_patches=list(cfg["param"]["num_patches"][0]))
        torch.manual_seed(0)
        net = ref.ViT(param)
        out["table/" + coords] = net.pos_embed.numpy()
        if coords == "cylindrical":
            ref_stubs.rerandomise_zero_init(net, seed=3)
            g = torch.Generator().manual_seed(5)
            x = torch.randn(B, cfg["geom"].tokens, param["patch_dim"], generator=g)
            c = torch.rand(B, param["condition_dim"], generator=g)
            t = torch.rand(B, 1, generator=g)
            with torch.no_grad():
                y = net(x, t, c)
            out.update({"sd/" + k: v.detach().numpy() for k, v in net.state_dict().items()})
            out.update(x=x.numpy(), c=c.numpy(), t=t.numpy(), net_out=y.numpy())
    out["meta"] = np.asarray([hidden, heads, depth, B], dtype=np.int32)
    np.savez_compressed(os.path.join(OUT, "net_fixed_pos_tiny.npz"), **out)
    print("fixed_pos_tiny written")


def golden_postprocess():
    """The reference's transform objects of configs/calochallenge/cfm/calochallenge_ds2.yaml:15-28 run back to front
    on synthetic sampled showers, as experiments/calochallenge/experiment.py:286-289 does.  NormalizeByElayer and
    GlobalStandardizeFromFile read files in their constructors (binning XML, means.npy): the objects are created
    without __init__ and given the attributes those files would provide; their __call__ is the reference's."""
    import importlib
    tr = importlib.import_module("experiments.calochallenge.transforms")
    L, per = 45, 12                       # 45 layers of 12 voxels: the ds2 structure at a size that keeps the file small
    V = L * per
    bounds = np.arange(0, V + 1, per)
    mean, std = -7.5, 2.25
    norm = object.__new__(tr.NormalizeByElayer)
    norm.eps, norm.cut, norm.layer_boundaries, norm.n_layers = 1.0e-10, 0.0, bounds, L
    gs = object.__new__(tr.GlobalStandardizeFromFile)
    gs.mean, gs.std = torch.tensor(mean), torch.tensor(std)
    chain = [norm, tr.ScaleTotalEnergy(factor=0.35, n_layers=L), tr.CutValues(cut=1.0e-7, n_layers=L),
             tr.ExclusiveLogitTransform(delta=1.0e-6, rescale=True), gs, tr.LogEnergy(),
             tr.ScaleEnergy(e_min=6.907755, e_max=13.815510), tr.AddFeaturesToCond(split_index=V),
             tr.Reshape(shape=[1, L, 4, 3])]
    g = torch.Generator().manual_seed(2024)
    N = 64
    samples = torch.randn(N, 1, L, 4, 3, generator=g) * 1.5
    cond = torch.cat([torch.randn(N, L, generator=g) * 1.2 + 3.0, torch.rand(N, 1, generator=g)], dim=1)
    x, c = samples.clone().squeeze(1), cond.clone()
    for fn in chain[::-1]:
        x, c = fn(x, c, rev=True)
    np.savez_compressed(os.path.join(OUT, "postprocess_ds2.npz"), samples=samples.numpy(), cond=cond.numpy(),
                        bounds=bounds.astype(np.int32), mean=np.float32(mean), std=np.float32(std),
                        showers=x.numpy(), energies=c.numpy())
    print("postprocess golden: showers", tuple(x.shape), "nonzero frac", float((x > 0).float().mean()))


def golden_preprocess():
    """The same reference transform objects run FRONT TO BACK on synthetic raw showers, as the dataset constructor
    does (experiments/calochallenge/datasets.py:44-47), GlobalStandardizeFromFile in its compute-and-write state
    (transforms.py:55-63; write() replaced by a no-op, it only saves the two scalars)."""
    import importlib
    tr = importlib.import_module("experiments.calochallenge.transforms")
    L, per = 45, 12
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
             tr.Reshape(shape=[1, L, 4, 3])]
    g = torch.Generator().manual_seed(4048)
    N = 96
    raw = torch.exp(torch.randn(N, V, generator=g) * 2.0 + 3.0) * (torch.rand(N, V, generator=g) < 0.35)
    raw = raw.reshape(N, L, per)
    raw[torch.rand(N, L, generator=g) < 0.15] = 0.0             # empty layers
    raw[:, -3:][torch.rand(N, 3, generator=g) < 0.5] = 0.0      # and showers that stop early
    single = torch.rand(N, L, generator=g) < 0.05               # layers with a single hit (normalised value exactly 1)
    raw[single] = raw[single] * torch.nn.functional.one_hot(torch.tensor(3), per)
    raw = raw.reshape(N, V).contiguous()
    e_inc = 10.0 ** (3.0 + 3.0 * torch.rand(N, 1, generator=g))
    # a calorimeter sees about the incident energy: E_tot / E_inc in (0.3, 1.2), so that u_0 * factor stays below 1
    raw = raw * (e_inc * (0.3 + 0.9 * torch.rand(N, 1, generator=g)) / raw.sum(1, keepdim=True))
    x, c = raw.clone(), e_inc.clone()
    for fn in chain:
        x, c = fn(x, c, rank=1)
    np.savez_compressed(os.path.join(OUT, "preprocess_ds2.npz"), showers=raw.numpy(), e_inc=e_inc.numpy(),
                        bounds=bounds.astype(np.int32), mean=np.float32(gs.mean), std=np.float32(gs.std),
                        x=x.numpy(), cond=c.numpy())
    print("preprocess golden: x", tuple(x.shape), "cond", tuple(c.shape), "mean/std", float(gs.mean), float(gs.std))


def golden_energy():
    """The reference's ParallelTransformer (nn/cfm/transformer_cfm.py) at a small size, dims_c = 1 and 3: velocity
    for per-sample times, and a CFM.sample_batch solve of the base class (models/base_model.py:220-244)."""
    import importlib
    tc = importlib.import_module("nn.cfm.transformer_cfm")
    bm = importlib.import_module("models.base_model")
    for dims_c in (1, 3):
        param = dict(dims_in=7, dims_c=dims_c, dim_embedding=16, encode_t_dim=16, nhead=2, num_encoder_layers=2,
                     num_decoder_layers=2, dim_feedforward=64, dropout=0.0, activation="relu", embeds=True,
                     encode_t_scale=30)
        torch.manual_seed(3)
        net = tc.ParallelTransformer(param)
        with torch.no_grad():  # LayerNorm affine / biases away from their 1 / 0 initial values
            for name, p in net.named_parameters():
                if p.requires_grad and ("norm" in name or name.endswith("bias")):
                    p.add_(0.1 * torch.randn_like(p))
        model = bm.CFM(net, "uniform", "linear", dict(method="rk4", options=dict(step_size=0.05)), shape=[7])
        model.device, model.dtype = torch.device("cpu"), torch.float32
        g = torch.Generator().manual_seed(17)
        B = 5
        x = torch.randn(B, 7, generator=g); t = torch.rand(B, 1, generator=g); c = torch.rand(B, dims_c, generator=g)
        out = {"sd/" + k: v.detach().numpy() for k, v in net.state_dict().items()}
        with torch.no_grad():
            out["velocity"] = net(x, t, c).numpy()
        torch.manual_seed(99)
        x_T = torch.randn((B, 7))
        torch.manual_seed(99)
        sample = model.sample_batch(c)
        out.update(x=x.numpy(), t=t.numpy(), c=c.numpy(), x_T=x_T.numpy(), sample=sample.numpy())
        out["meta"] = np.asarray([param[k] for k in ("dims_in", "dims_c", "dim_embedding", "encode_t_dim", "nhead",
                                                      "num_encoder_layers", "num_decoder_layers", "dim_feedforward")],
                                 dtype=np.int32)
        np.savez_compressed(os.path.join(OUT, f"energy_tiny_c{dims_c}.npz"), **out)
        print("energy_tiny", dims_c, "velocity rms", float(np.sqrt((out["velocity"] ** 2).mean())))


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_stubs.load_reference()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    golden_patch_maps(ref)
    golden_net(ref, "ds2", hidden=48, heads=2, depth=2, B=3, tag="ds2_tiny")
    golden_net(ref, "calogan", hidden=48, heads=2, depth=2, B=2, tag="calogan_tiny")
    golden_net(ref, "ds1_pions", hidden=48, heads=2, depth=1, B=2, tag="ds1_pions_tiny")
    golden_lemurs(ref)
    golden_finetune(ref)
    golden_fixed_pos_embed(ref)
    golden_postprocess()
    golden_preprocess()
    golden_energy()


if __name__ == "__main__":
    main()
