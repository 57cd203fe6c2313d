"""CPU restatement of the vit4hep CFM-ViT hot path.  TEST INFRASTRUCTURE ONLY.

Plain numpy (index maps) and plain torch-on-CPU fp32/fp64 (floating-point math),
written from the algorithm, not from the reference sources; each function cites the
reference file:line it restates.  It travels to the GPU box (``/root/reference`` does
not) and is the checker for every ``-m gpu`` parity test, ``smoke()`` and the
``cpu_baseline`` leg of ``bench.py``.

Pinning: ``tests/test_oracle_pinned.py`` compares every function here with
``tests/golden/*.npz`` (outputs of the unmodified reference run in the build container by
``oracle/make_golden.py``) and, when ``/root/reference`` is present, with the live
reference.  Parity is otherwise unpinned by the reference itself (no tests upstream).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------
# Geometry / patch index maps (integer work: must be bit-exact)
# ----------------------------------------------------------------------------------------


@dataclass(frozen=True)
class Segment:
    """One calorimeter segment with a regular (L, A, R) voxel grid and its patch shape."""
    shape: Tuple[int, int, int]
    patch: Tuple[int, int, int]

    @property
    def num_patches(self) -> Tuple[int, int, int]:
        return tuple(s // p for s, p in zip(self.shape, self.patch))

    @property
    def voxels(self) -> int:
        return self.shape[0] * self.shape[1] * self.shape[2]


@dataclass(frozen=True)
class Geometry:
    """Detector geometry as the reference's CFM wrappers see it.

    regular:  reference experiments/calochallenge/calochallenge_cfm/model.py:9-60 (one segment,
              input (B, C, L, A, R));
    segmented: reference calochallenge_cfm/model.py:97-173, experiments/calogan/model.py:8-87,
              experiments/calohadronic/model.py:8-86 (input flat (B, C, sum V_k), split at
              list_edges, one rearrange per segment, tokens concatenated).
    """
    segments: Tuple[Segment, ...]
    in_channels: int = 1
    flat_input: bool = False  # True for the segmented wrappers: input is (B, C, sum V)

    @property
    def voxels(self) -> int:
        return sum(s.voxels for s in self.segments)

    @property
    def tokens(self) -> int:
        return sum(math.prod(s.num_patches) for s in self.segments)

    @property
    def patch_dim(self) -> int:
        pd = {math.prod(s.patch) * self.in_channels for s in self.segments}
        assert len(pd) == 1, "patch volume must be equal across segments"
        return pd.pop()

    @property
    def sample_shape(self) -> Tuple[int, ...]:
        if self.flat_input:
            return (self.in_channels, self.voxels)
        return (self.in_channels, *self.segments[0].shape)


def patch_index_map(geom: Geometry) -> np.ndarray:
    """int64 array ``idx`` of length C*V with ``tokens.flat[j] = x.flat[idx[j]]`` per sample.

    Restates einops ``"b c (l p1) (a p2) (r p3) -> b (l a r) (p1 p2 p3 c)"`` (reference
    calochallenge_cfm/model.py:54-60) and, for segmented geometries, the split /
    per-segment rearrange / cat of calogan/model.py:77-87.  Pure-python index arithmetic.
    """
    C = geom.in_channels
    V = geom.voxels
    out: List[int] = []
    seg_off = 0
    for seg in geom.segments:
        L, A, R = seg.shape
        P1, P2, P3 = seg.patch
        nl, na, nr = seg.num_patches
        for l in range(nl):
            for a in range(na):
                for r in range(nr):
                    for p1 in range(P1):
                        for p2 in range(P2):
                            for p3 in range(P3):
                                for c in range(C):
                                    vox = ((l * P1 + p1) * A + (a * P2 + p2)) * R + (r * P3 + p3)
                                    if geom.flat_input:
                                        # (B, C, sumV): channel-major over the whole flat axis,
                                        # each segment reshaped to (C, L, A, R) after the split
                                        out.append(c * V + seg_off + vox)
                                    else:
                                        out.append(c * V + vox)
        seg_off += seg.voxels
    return np.asarray(out, dtype=np.int64)


def to_patches(x: np.ndarray | torch.Tensor, geom: Geometry):
    """(B, C, *grid) or (B, C, sumV) -> (B, T, P).  Exact copy semantics."""
    idx = patch_index_map(geom)
    B = x.shape[0]
    flat = x.reshape(B, -1)
    if isinstance(x, torch.Tensor):
        tok = flat[:, torch.from_numpy(idx)]
    else:
        tok = flat[:, idx]
    return tok.reshape(B, geom.tokens, geom.patch_dim)


def from_patches(tok: np.ndarray | torch.Tensor, geom: Geometry):
    """(B, T, P) -> (B, C, *grid) / (B, C, sumV): exact inverse of :func:`to_patches`."""
    idx = patch_index_map(geom)
    B = tok.shape[0]
    flat = tok.reshape(B, -1)
    if isinstance(tok, torch.Tensor):
        out = torch.empty_like(flat)
        out[:, torch.from_numpy(idx)] = flat
    else:
        out = np.empty_like(flat)
        out[:, idx] = flat
    return out.reshape(B, *geom.sample_shape)


# ----------------------------------------------------------------------------------------
# Network
# ----------------------------------------------------------------------------------------


def create_meshgrid(num_patches: Sequence[Sequence[int]]):
    """pos_z, pos_y, pos_x buffers (reference nn/vit.py:137-154): layer coordinate is
    cumulative over list entries and divided by sum L_k; y = a / A_k, x = r / R_k."""
    sum_l = sum(n[0] for n in num_patches)
    zs, ys, xs = [], [], []
    l0 = 0
    for (L, A, R) in num_patches:
        lgrid = torch.arange(sum_l, dtype=torch.float32)[l0:l0 + L] / sum_l
        agrid = torch.arange(A, dtype=torch.float32) / A
        rgrid = torch.arange(R, dtype=torch.float32) / R
        z = lgrid[:, None, None].expand(L, A, R)
        y = agrid[None, :, None].expand(L, A, R)
        x = rgrid[None, None, :].expand(L, A, R)
        zs.append(z.reshape(-1)); ys.append(y.reshape(-1)); xs.append(x.reshape(-1))
        l0 += L
    return torch.cat(zs), torch.cat(ys), torch.cat(xs)


def learnable_pos_embedding(sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """(T, D) table, D = 6 * len(freqs) (reference nn/vit.py:156-162)."""
    w = sd["pos_embed_freqs"] * 2 * math.pi
    z = sd["pos_z"][:, None] * w[None, :]
    y = sd["pos_y"][:, None] * w[None, :]
    x = sd["pos_x"][:, None] * w[None, :]
    return torch.cat((x.sin(), x.cos(), y.sin(), y.cos(), z.sin(), z.cos()), dim=1)


def timestep_embedding(t: torch.Tensor, dim: int = 256, max_period: float = 10000.0):
    """(B, 1) -> (B, dim): cat(cos(t f), sin(t f)), f_i = exp(-ln(max_period) i / half)
    (reference nn/vit.py:368-389; t is used raw in [0, 1], args forced to fp32 by t.float())."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period)
                      * torch.arange(0, half, dtype=t.dtype) / half)
    args = t.float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _ln(x: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """LayerNorm without affine, biased variance, eps 1e-6 (reference nn/vit.py:309-311)."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps)


def _modulate(x, shift, scale):
    """reference nn/vit.py:457-458"""
    return x * (1 + scale[:, None, :]) + shift[:, None, :]


def _mlp2(sd, prefix0, prefix2, x):
    h = F.linear(x, sd[prefix0 + ".weight"], sd[prefix0 + ".bias"])
    return F.linear(F.silu(h), sd[prefix2 + ".weight"], sd[prefix2 + ".bias"])


def attention(qkv: torch.Tensor, num_heads: int) -> torch.Tensor:
    """qkv (B, T, 3*D) with column order [3][H][dh] -> (B, T, D) heads merged as [H][dh]
    (reference nn/vit.py:425-451; softmax(q k^T / sqrt(dh)) v, no mask, no dropout)."""
    B, T, D3 = qkv.shape
    D = D3 // 3
    dh = D // num_heads
    q, k, v = qkv.reshape(B, T, 3, num_heads, dh).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * (dh ** -0.5)
    p = torch.softmax(s, dim=-1)
    return (p @ v).transpose(1, 2).reshape(B, T, D)


def get_sincos_pos_embed(pos_embedding_coords: str, num_patches, hidden_dim: int, dim: int = 3,
                         temperature: float = 10000.0) -> torch.Tensor:
    """Fixed (T, D) positional table of ``learn_pos_embed=False`` (reference nn/vit.py:461-540, used by
    nn/vit.py:92-103): frequencies temperature**(-i/(F-1)), F = D/6 per coordinate and sin/cos; coordinates
    (radial, angular, layer) for "cylindrical", (r cos a, r sin a, layer) with a in [0, 2 pi) for "cartesian"."""
    if len(num_patches) == 1 and isinstance(num_patches[0], (list, tuple)):
        num_patches = num_patches[0]
    L, A, R = num_patches
    F_ = hidden_dim // 6
    omega = torch.pow(torch.tensor(float(temperature)), -torch.arange(F_) / (F_ - 1))
    z = (torch.arange(L) / L)[:, None, None].expand(L, A, R).reshape(-1)
    if pos_embedding_coords == "cylindrical":
        a = (torch.arange(A) / A)[None, :, None].expand(L, A, R).reshape(-1)
        r = (torch.arange(R) / R)[None, None, :].expand(L, A, R).reshape(-1)
        coords = (r, a, z)
    elif pos_embedding_coords == "cartesian":
        a = (torch.arange(A) * (2 * math.pi / A))[None, :, None].expand(L, A, R).reshape(-1)
        r = (torch.arange(R) / R)[None, None, :].expand(L, A, R).reshape(-1)
        coords = (r * a.cos(), r * a.sin(), z)
    else:
        raise ValueError(pos_embedding_coords)
    parts = []
    for cvals in coords:
        arg = cvals[:, None] * omega[None, :]
        parts += [arg.sin(), arg.cos()]
    return torch.cat(parts, dim=1)


def vit_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, t: torch.Tensor,
                c: torch.Tensor, num_heads: int) -> torch.Tensor:
    """ViT.forward (reference nn/vit.py:185-206) as a pure function of a state dict with
    the reference's key names.  x (B,T,P), t (B,1), c (B,K) -> (B,T,P*out_channels).

    Also accepts the finetuning structures of reference
    experiments/calochallenge/calochallenge_cfm/experiment_finetuning.py:79-118: ``x_embedder`` =
    Sequential(mapper, SiLU, old Linear) (keys x_embedder.0 / .2) and ``c_embedder`` =
    Sequential(mapper, SiLU, old Sequential) (keys c_embedder.0, c_embedder.2.0, c_embedder.2.2)."""
    depth = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
    if "x_embedder.weight" in sd:
        h = F.linear(x, sd["x_embedder.weight"], sd["x_embedder.bias"])
    else:
        xm = F.silu(F.linear(x, sd["x_embedder.0.weight"], sd["x_embedder.0.bias"]))
        h = F.linear(xm, sd["x_embedder.2.weight"], sd["x_embedder.2.bias"])
    if "pos_embed_freqs" in sd:
        h = h + learnable_pos_embedding(sd)
    else:
        h = h + sd["pos_embed"]
    temb = _mlp2(sd, "t_embedder.mlp.0", "t_embedder.mlp.2", timestep_embedding(t).to(x.dtype))
    if "c_embedder.2.0.weight" in sd:
        cm = F.silu(F.linear(c, sd["c_embedder.0.weight"], sd["c_embedder.0.bias"]))
        cemb = _mlp2(sd, "c_embedder.2.0", "c_embedder.2.2", cm)
    else:
        cemb = _mlp2(sd, "c_embedder.0", "c_embedder.2", c)
    cond = F.silu(temb + cemb)
    for i in range(depth):
        p = f"blocks.{i}."
        mod = F.linear(cond, sd[p + "adaLN_modulation.1.weight"], sd[p + "adaLN_modulation.1.bias"])
        sh1, sc1, g1, sh2, sc2, g2 = mod.chunk(6, dim=1)
        a = _modulate(_ln(h), sh1, sc1)
        qkv = F.linear(a, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"])
        o = attention(qkv, num_heads)
        h = h + g1[:, None, :] * F.linear(o, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
        m = _modulate(_ln(h), sh2, sc2)
        u = F.linear(m, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
        g = F.gelu(u, approximate="tanh")
        h = h + g2[:, None, :] * F.linear(g, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    mod = F.linear(cond, sd["final_layer.adaLN_modulation.1.weight"],
                   sd["final_layer.adaLN_modulation.1.bias"])
    sh, sc = mod.chunk(2, dim=1)
    y = _modulate(_ln(h), sh, sc)
    return F.linear(y, sd["final_layer.linear.weight"], sd["final_layer.linear.bias"])


# ----------------------------------------------------------------------------------------
# Parameter construction (for the GPU box, where the reference cannot be instantiated)
# ----------------------------------------------------------------------------------------


def init_state_dict(param: dict, seed: int = 0, rerandomise: bool = True,
                    dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """State dict with the reference's names and shapes.

    Values: Xavier-uniform Linears like the reference (nn/vit.py:164-172), and - when
    ``rerandomise`` - N(0, 0.02) adaLN / final / bias tensors instead of the reference's
    zeros, so that outputs and gradients are non-trivial.  This is NOT bitwise the
    reference's RNG stream; tests that need reference-identical weights load a golden
    state dict instead.
    """
    g = torch.Generator().manual_seed(seed)
    D = param.get("hidden_dim", 180)
    P = param.get("patch_dim", 12)
    K = param.get("condition_dim", 46)
    depth = param.get("depth", 2)
    hid = int(D * param.get("mlp_ratio", 2.0))
    oc = param.get("out_channels", 1)
    num_patches = param.get("num_patches", [[15, 4, 9]])

    def xavier(o, i):
        a = math.sqrt(6.0 / (i + o))
        return (torch.rand(o, i, generator=g) * 2 - 1) * a

    def small(*shape):
        return torch.randn(*shape, generator=g) * 0.02 if rerandomise else torch.zeros(*shape)

    sd: Dict[str, torch.Tensor] = {}
    sd["pos_embed_freqs"] = torch.randn(D // 6, generator=g)
    sd["pos_z"], sd["pos_y"], sd["pos_x"] = create_meshgrid(num_patches)

    def lin(name, o, i, zero=False):
        sd[name + ".weight"] = small(o, i) if zero else xavier(o, i)
        sd[name + ".bias"] = small(o)

    lin("x_embedder", D, P)
    lin("c_embedder.0", D, K); lin("c_embedder.2", D, D)
    lin("t_embedder.mlp.0", D, 256); lin("t_embedder.mlp.2", D, D)
    for i in range(depth):
        p = f"blocks.{i}."
        lin(p + "attn.qkv", 3 * D, D); lin(p + "attn.proj", D, D)
        lin(p + "mlp.fc1", hid, D); lin(p + "mlp.fc2", D, hid)
        lin(p + "adaLN_modulation.1", 6 * D, D, zero=True)
    lin("final_layer.linear", oc * P, D, zero=True)
    lin("final_layer.adaLN_modulation.1", 2 * D, D, zero=True)
    return {k: v.to(dtype) for k, v in sd.items()}


# ----------------------------------------------------------------------------------------
# CFM loss and ODE sampling
# ----------------------------------------------------------------------------------------


def linear_trajectory(x0, x1, t):
    """reference models/trajectories.py:5-8"""
    return (1 - t) * x0 + t * x1, x1 - x0


def cfm_forward(sd, x, t, c, geom: Geometry, num_heads: int):
    """wrapper forward: from_patches(net(to_patches(x), t, c))
    (reference calochallenge_cfm/model.py:62-66)."""
    return from_patches(vit_forward(sd, to_patches(x, geom), t, c, num_heads), geom)


def cfm_loss(sd, x1, c, x0, t, geom: Geometry, num_heads: int) -> torch.Tensor:
    """CFM._batch_loss with the random draws (x0 ~ N(0,1), t ~ U(0,1) of shape (B,1,..,1))
    supplied by the caller (reference models/base_model.py:203-218)."""
    tt = t.reshape(-1, *([1] * (x1.dim() - 1)))
    x_t, x_t_dot = linear_trajectory(x0, x1, tt)
    v = cfm_forward(sd, x_t, t.reshape(-1, 1), c, geom, num_heads)
    return ((v - x_t_dot) ** 2).mean()


def rk4_38_grid(step_size: float, t0: float = 0.0, t1: float = 1.0, dtype=torch.float32):
    """torchdiffeq's fixed grid from ``step_size`` (published algorithm, see
    oracle/ref_stubs.py:_odeint): ceil((t1-t0)/step + 1) points, last forced to t1."""
    a = torch.tensor(t0, dtype=dtype); b = torch.tensor(t1, dtype=dtype)
    n = torch.ceil((b - a) / step_size + 1).item()
    grid = torch.arange(0, n, dtype=dtype) * step_size + a
    grid[-1] = b
    return grid


def sample_batch(sd, cond, x_T, geom: Geometry, num_heads: int, step_size: float = 0.05):
    """sample_batch with x_T supplied (reference calochallenge_cfm/model.py:68-94):
    integrate dx/dt = v(x, t, cond) from t=0 to t=1 with torchdiffeq 'rk4' = 3/8 rule,
    every sample of the batch sharing one t."""
    grid = rk4_38_grid(step_size, dtype=x_T.dtype)
    B = x_T.shape[0]

    def f(t, y):
        return cfm_forward(sd, y, t.repeat((B, 1)), cond, geom, num_heads)

    y = x_T
    third, two_thirds = 1.0 / 3.0, 2.0 / 3.0
    for ta, tb in zip(grid[:-1], grid[1:]):
        dt = tb - ta
        k1 = f(ta, y)
        k2 = f(ta + dt * third, y + dt * k1 * third)
        k3 = f(ta + dt * two_thirds, y + dt * (k2 - k1 * third))
        k4 = f(tb, y + dt * (k1 - k2 + k3))
        y = y + (k1 + 3 * (k2 + k3) + k4) * dt * 0.125
    return y


# ----------------------------------------------------------------------------------------
# Named geometries / model hyper-parameters of the shipped configs
# ----------------------------------------------------------------------------------------


def _vit_param(patch_dim, num_patches, condition_dim, **kw):
    p = dict(dim=3, condition_dim=condition_dim, hidden_dim=480, out_channels=1, depth=6,
             num_heads=6, mlp_ratio=4, attn_drop=0.0, proj_drop=0.0, learn_pos_embed=True,
             causal_attn=False, checkpoint_grads=False, num_patches=num_patches,
             patch_dim=patch_dim, use_torch_sdpa=False)
    p.update(kw)
    return p


CONFIGS = {
    # reference configs/model/cfm/cfm_ds2_electrons.yaml
    "ds2": dict(geom=Geometry((Segment((45, 16, 9), (3, 16, 1)),)),
                param=_vit_param(48, [[15, 1, 9]], 46)),
    # reference configs/model/cfm/cfm_ds3_electrons.yaml
    "ds3": dict(geom=Geometry((Segment((45, 50, 18), (3, 10, 3)),)),
                param=_vit_param(90, [[15, 5, 6]], 46)),
    # reference configs/model/cfm_calogan/cfm_eplus.yaml
    "calogan": dict(geom=Geometry((Segment((1, 96, 3), (1, 6, 1)), Segment((1, 12, 12), (1, 2, 3)),
                                   Segment((1, 6, 12), (1, 2, 3))), flat_input=True),
                    param=_vit_param(6, [[1, 16, 3], [1, 6, 4], [1, 3, 4]], 4)),
    # reference configs/model/cfm_calohad/cfm_calohad.yaml
    "calohad": dict(geom=Geometry((Segment((10, 15, 15), (5, 5, 3)), Segment((48, 30, 30), (3, 5, 5))),
                                  flat_input=True),
                    param=_vit_param(75, [[2, 3, 5], [16, 6, 6]], 59)),
    # reference configs/model/cfm/cfm_ds1_photons.yaml
    "ds1_photons": dict(geom=Geometry(tuple(Segment(s, (1, 1, 5)) for s in
                                            [(1, 8, 5), (1, 16, 10), (1, 19, 10), (1, 5, 5), (1, 5, 5)]),
                                      flat_input=True),
                        param=_vit_param(5, [[1, 8, 1], [1, 16, 2], [1, 19, 2], [1, 5, 1], [1, 5, 1]], 6)),
}


_DS1_PIONS_SHAPES = [(1, 8, 5), (1, 10, 10), (1, 10, 10), (1, 5, 5), (1, 15, 10), (1, 16, 10), (1, 10, 5)]
# reference configs/model/cfm/cfm_ds1_pions.yaml
CONFIGS["ds1_pions"] = dict(
    geom=Geometry(tuple(Segment(s, (1, 1, 5)) for s in _DS1_PIONS_SHAPES), flat_input=True),
    param=_vit_param(5, [[1, 8, 1], [1, 10, 2], [1, 10, 2], [1, 5, 1], [1, 15, 2], [1, 16, 2], [1, 10, 1]], 8))
# reference configs/model/cfm_lemurs/cfm_lemurs.yaml: the ds2 grid with 53 conditions; batches arrive as
# (B, R, A, L) and are permuted by LEMURSCFM._batch_loss (experiments/lemurs/model.py:62-65)
CONFIGS["lemurs"] = dict(geom=Geometry((Segment((45, 16, 9), (3, 16, 1)),)),
                         param=_vit_param(48, [[15, 1, 9]], 53))


def lemurs_to_grid(x: torch.Tensor) -> torch.Tensor:
    """(B, R, A, L) -> (B, 1, L, A, R) (reference experiments/lemurs/model.py:63-64)"""
    return x.permute(0, 3, 2, 1).unsqueeze(1)


def tiny_config(name: str = "ds2", hidden_dim: int = 96, depth: int = 2, num_heads: int = 2):
    """Same geometry, smaller network: for parity cases the oracle finishes in seconds."""
    cfg = CONFIGS[name]
    p = dict(cfg["param"]); p.update(hidden_dim=hidden_dim, depth=depth, num_heads=num_heads)
    return dict(geom=cfg["geom"], param=p)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu().reshape(-1); b = b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-300))
