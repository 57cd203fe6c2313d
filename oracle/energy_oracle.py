"""CPU restatement of the energy-ratio velocity network.  TEST INFRASTRUCTURE ONLY.

``ParallelTransformer`` (reference nn/cfm/transformer_cfm.py:12-119, ``embeds=True`` as in every shipped energy
config) around ``torch.nn.Transformer`` (post-norm, ReLU, batch_first, final LayerNorms in encoder and decoder;
torch is present in the container, so its layer arithmetic is restated from the installed torch.nn sources and
pinned against the live module).  A pure function of a state dict with the reference's key names.
Pinned by tests/golden/energy_tiny*.npz (oracle/make_golden.py runs the unmodified reference class).
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F


def _ln(x, w, b, eps=1e-5):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def _mha(sd, prefix, q_in, kv_in, nhead):
    """nn.MultiheadAttention forward (no mask, no dropout): packed in_proj (3E, E) in q | k | v order"""
    E = q_in.shape[-1]
    W, b = sd[prefix + ".in_proj_weight"], sd[prefix + ".in_proj_bias"]
    q = F.linear(q_in, W[:E], b[:E])
    k = F.linear(kv_in, W[E:2 * E], b[E:2 * E])
    v = F.linear(kv_in, W[2 * E:], b[2 * E:])
    B, Tq, _ = q.shape
    Tk = k.shape[1]
    dh = E // nhead
    q = q.reshape(B, Tq, nhead, dh).transpose(1, 2)
    k = k.reshape(B, Tk, nhead, dh).transpose(1, 2)
    v = v.reshape(B, Tk, nhead, dh).transpose(1, 2)
    p = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh), dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, Tq, E)
    return F.linear(o, sd[prefix + ".out_proj.weight"], sd[prefix + ".out_proj.bias"])


def _ff(sd, prefix, x):
    return F.linear(F.relu(F.linear(x, sd[prefix + ".linear1.weight"], sd[prefix + ".linear1.bias"])),
                    sd[prefix + ".linear2.weight"], sd[prefix + ".linear2.bias"])


def time_embed(sd, t):
    """GaussianFourierProjection + Linear (reference transformer_cfm.py:39-42, :165-176): t (B, 1) -> (B, Dt)"""
    proj = t * sd["time_embed.0.W"] * 2 * torch.pi
    return F.linear(torch.cat([torch.sin(proj), torch.cos(proj)], dim=1), sd["time_embed.1.weight"], sd["time_embed.1.bias"])


def encode_condition(sd, c, nhead, n_enc):
    """memory = transformer.encoder(compute_embedding(condition)) (reference transformer_cfm.py:86-89, :111-113)"""
    src = F.linear(c.unsqueeze(-1), sd["c_embed.weight"], sd["c_embed.bias"]) + sd["pos_embed_c.weight"][None]
    x = src
    for i in range(n_enc):
        p = f"transformer.encoder.layers.{i}"
        x = _ln(x + _mha(sd, p + ".self_attn", x, x, nhead), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"])
        x = _ln(x + _ff(sd, p, x), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"])
    return _ln(x, sd["transformer.encoder.norm.weight"], sd["transformer.encoder.norm.bias"])


def energy_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, t: torch.Tensor, c: torch.Tensor, nhead: int):
    """ParallelTransformer.forward with a condition (reference transformer_cfm.py:99-119):
    x (B, dims_in), t (B, 1), c (B, dims_c) -> velocity (B, dims_in)"""
    n_enc = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("transformer.encoder.layers."))
    n_dec = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("transformer.decoder.layers."))
    mem = encode_condition(sd, c, nhead, n_enc)
    te = time_embed(sd, t)                                                   # (B, Dt)
    px = F.linear(x.unsqueeze(-1), sd["x_embed.weight"], sd["x_embed.bias"]) + sd["pos_embed_x.weight"][None]
    y = torch.cat([te.unsqueeze(1).expand(-1, px.shape[1], -1), px], dim=-1)  # (B, dims_in, Dt + De)
    for i in range(n_dec):
        p = f"transformer.decoder.layers.{i}"
        y = _ln(y + _mha(sd, p + ".self_attn", y, y, nhead), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"])
        y = _ln(y + _mha(sd, p + ".multihead_attn", y, mem, nhead), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"])
        y = _ln(y + _ff(sd, p, y), sd[p + ".norm3.weight"], sd[p + ".norm3.bias"])
    y = _ln(y, sd["transformer.decoder.norm.weight"], sd["transformer.decoder.norm.bias"])
    h = torch.cat([te.unsqueeze(1).expand(-1, y.shape[1], -1), y], dim=-1)
    h = F.silu(F.linear(h, sd["layers.0.weight"], sd["layers.0.bias"]))
    return F.linear(h, sd["layers.2.weight"], sd["layers.2.bias"]).squeeze(-1)


def init_state_dict(param: dict, seed: int = 0) -> Dict[str, torch.Tensor]:
    """random state dict with the reference's names and shapes (for the GPU box, where the reference is absent)"""
    g = torch.Generator().manual_seed(seed)
    De, Dt = param["dim_embedding"], param.get("encode_t_dim", 64)
    E, Fd = 2 * De, param["dim_feedforward"]
    assert De == Dt, "embeds=True concatenates time and x embeddings into d_model = 2 * dim_embedding"

    def r(*shape, s=None):
        s = s if s is not None else 1.0 / math.sqrt(shape[-1])
        return (torch.rand(*shape, generator=g) * 2 - 1) * s

    sd = {"time_embed.0.W": torch.randn(Dt // 2, generator=g) * param.get("encode_t_scale", 30),
          "time_embed.1.weight": r(Dt, Dt), "time_embed.1.bias": r(Dt, s=0.1),
          "x_embed.weight": r(De, 1), "x_embed.bias": r(De, s=0.5), "c_embed.weight": r(E, 1), "c_embed.bias": r(E, s=0.5),
          "pos_embed_x.weight": torch.randn(param["dims_in"], De, generator=g),
          "pos_embed_c.weight": torch.randn(param["dims_c"], E, generator=g)}

    def attn(p):
        sd[p + ".in_proj_weight"] = r(3 * E, E); sd[p + ".in_proj_bias"] = r(3 * E, s=0.1)
        sd[p + ".out_proj.weight"] = r(E, E); sd[p + ".out_proj.bias"] = r(E, s=0.1)

    def layer(p, norms):
        sd[p + ".linear1.weight"] = r(Fd, E); sd[p + ".linear1.bias"] = r(Fd, s=0.1)
        sd[p + ".linear2.weight"] = r(E, Fd); sd[p + ".linear2.bias"] = r(E, s=0.1)
        for n in norms:
            sd[f"{p}.{n}.weight"] = 1 + 0.1 * torch.randn(E, generator=g); sd[f"{p}.{n}.bias"] = 0.1 * torch.randn(E, generator=g)

    for i in range(param["num_encoder_layers"]):
        p = f"transformer.encoder.layers.{i}"
        attn(p + ".self_attn"); layer(p, ("norm1", "norm2"))
    for i in range(param["num_decoder_layers"]):
        p = f"transformer.decoder.layers.{i}"
        attn(p + ".self_attn"); attn(p + ".multihead_attn"); layer(p, ("norm1", "norm2", "norm3"))
    for p in ("transformer.encoder.norm", "transformer.decoder.norm"):
        sd[p + ".weight"] = 1 + 0.1 * torch.randn(E, generator=g); sd[p + ".bias"] = 0.1 * torch.randn(E, generator=g)
    sd["layers.0.weight"] = r(Fd, Dt + E); sd["layers.0.bias"] = r(Fd, s=0.1)
    sd["layer.weight"], sd["layer.bias"] = sd["layers.0.weight"], sd["layers.0.bias"]
    sd["layers.2.weight"] = r(1, Fd); sd["layers.2.bias"] = r(1, s=0.1)
    return sd


# reference configs/model/cfm/cfm_ds2_energy.yaml:13-26 (ds3 identical; LEMURS: dims_c 3; ds1: dims_in 5 / 7)
DS2_ENERGY = dict(dims_in=45, dims_c=1, dim_embedding=64, nhead=4, num_encoder_layers=4, num_decoder_layers=4,
                  dim_feedforward=512, dropout=0.0, activation="relu", embeds=True, encode_t_scale=30)


def sample_batch(sd, cond, x_T, nhead, step_size=0.05):
    """CFM.sample_batch of the base class (reference models/base_model.py:220-244) with torchdiffeq 'rk4' = 3/8 rule"""
    from oracle.vit_oracle import rk4_38_grid
    grid = rk4_38_grid(step_size, dtype=x_T.dtype)
    B = x_T.shape[0]
    f = lambda t, y: energy_forward(sd, y, t.repeat((B, 1)), cond, nhead)
    y = x_T
    for ta, tb in zip(grid[:-1], grid[1:]):
        dt = tb - ta
        k1 = f(ta, y)
        k2 = f(ta + dt / 3, y + dt * k1 / 3)
        k3 = f(ta + dt * 2 / 3, y + dt * (k2 - k1 / 3))
        k4 = f(tb, y + dt * (k1 - k2 + k3))
        y = y + (k1 + 3 * (k2 + k3) + k4) * dt * 0.125
    return y
