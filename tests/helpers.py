"""Shared builders for the parity tests: the B200 model classes configured like oracle.CONFIGS."""
import os

import numpy as np
import torch

from oracle import vit_oracle as vo

ODE = dict(method="rk4", options=dict(step_size=0.05))


def build_model(name, param, precision, device="cuda"):
    """vit4hep_b200 wrapper + net for the named geometry (same ctor arguments as the reference's)."""
    import vit4hep_b200 as v4
    geom = vo.CONFIGS[name]["geom"]
    p = dict(param); p["precision"] = precision
    net = v4.ViT(p)
    segs = geom.segments
    if name in ("ds2", "ds3"):
        m = v4.CaloChallengeCFM(net, list(segs[0].patch), 1, "uniform", "linear", ODE, shape=list(segs[0].shape))
    elif name == "lemurs":
        m = v4.LEMURSCFM(net, list(segs[0].patch), 1, "uniform", "linear", ODE, shape=list(segs[0].shape))
    elif name in ("ds1_photons", "ds1_pions"):
        m = v4.CaloChallengeCFM_DS1(net, [list(s.shape) for s in segs], [s.voxels for s in segs],
                                    list(segs[0].patch), 1, "uniform", "linear", ODE, shape=[geom.voxels])
    elif name == "calogan":
        m = v4.CaloGANCFM(net, [list(s.shape) for s in segs], [s.voxels for s in segs],
                          [list(s.patch) for s in segs], 1, "uniform", "linear", ODE, shape=[geom.voxels])
    elif name == "calohad":
        m = v4.CaloHadCFM(net, [list(s.shape) for s in segs], [s.voxels for s in segs],
                          [list(s.patch) for s in segs], 1, "uniform", "linear", ODE, shape=[geom.voxels])
    else:
        raise KeyError(name)
    m = m.to(device)
    m.device, m.dtype = torch.device(device), torch.float32
    return m


def load_golden(golden_dir, tag):
    z = np.load(os.path.join(golden_dir, f"net_{tag}.npz"))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    hidden, heads, depth, B = (int(v) for v in z["meta"][:4])
    return z, sd, dict(hidden_dim=hidden, num_heads=heads, depth=depth), B


def geometry_of(name):
    from vit4hep_b200.cfm import PatchGeometry
    g = vo.CONFIGS[name]["geom"]
    return PatchGeometry([s.shape for s in g.segments], [s.patch for s in g.segments], g.in_channels, g.flat_input)
