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
    elif name == "ds1_photons":
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


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_stubs.load_reference()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    golden_patch_maps(ref)
    golden_net(ref, "ds2", hidden=48, heads=2, depth=2, B=3, tag="ds2_tiny")
    golden_net(ref, "calogan", hidden=48, heads=2, depth=2, B=2, tag="calogan_tiny")


if __name__ == "__main__":
    main()
