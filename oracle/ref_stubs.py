"""sys.modules stubs that let the UNMODIFIED reference files import in this container.

TEST INFRASTRUCTURE.  Used only by ``oracle/make_golden.py`` and by the live-reference
pinning tests (skipped when ``/root/reference`` is absent, e.g. on the GPU box).

The reference imports three packages that are not installed here and cannot be fetched:

* ``timm.models.vision_transformer.Mlp``  (call site: reference nn/vit.py:7, :317-322)
* ``xformers.ops.memory_efficient_attention``  (reference nn/vit.py:9, :443-448)
* ``torchdiffeq.odeint``  (reference models/base_model.py:5, :235-242)

Each stub restates the published behaviour of the function the reference calls; none of
this is reference code.
"""
from __future__ import annotations

import math
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("VIT4HEP_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "nn", "vit.py"))


class _Mlp(nn.Module):
    """timm Mlp: fc1 -> act -> drop1 -> norm -> fc2 -> drop2 (state-dict names fc1.*, fc2.*)."""

    def __init__(self, in_features, hidden_features=None, out_features=None,
                 act_layer=nn.GELU, norm_layer=None, bias=True, drop=0.0, **_):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias)
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop)
        self.norm = norm_layer(hidden_features) if norm_layer is not None else nn.Identity()
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias)
        self.drop2 = nn.Dropout(drop)

    def forward(self, x):
        return self.drop2(self.fc2(self.norm(self.drop1(self.act(self.fc1(x))))))


def _memory_efficient_attention(q, k, v, attn_bias=None, p=0.0, scale=None):
    """xformers layout (B, T, H, dh); softmax(q k^T * scale) v with scale = dh**-0.5."""
    assert attn_bias is None and p == 0.0
    scale = q.shape[-1] ** -0.5 if scale is None else scale
    qh, kh, vh = (z.transpose(1, 2) for z in (q, k, v))
    att = torch.softmax((qh @ kh.transpose(-1, -2)) * scale, dim=-1)
    return (att @ vh).transpose(1, 2)


def _odeint(func, y0, t, *, method=None, options=None, **_):
    """torchdiffeq fixed-grid solver, method 'rk4' (= rk4_alt_step_func, the 3/8 rule).

    Restated from the published torchdiffeq algorithm (_impl/solvers.py FixedGridODESolver,
    _impl/fixed_grid.py RK4, _impl/rk_common.py rk4_alt_step_func): grid from step_size,
    last grid point forced to t[-1]; outputs only at the requested times (linear interp
    returns y1 exactly when t[j] == t1).
    """
    assert method == "rk4", "only the method the reference configs use is restated"
    step = options["step_size"]
    t0, t1 = t[0], t[-1]
    niters = torch.ceil((t1 - t0) / step + 1).item()
    grid = torch.arange(0, niters, dtype=t.dtype, device=t.device) * step + t0
    grid[-1] = t1
    sol = [y0]
    y = y0
    third, two_thirds = 1.0 / 3.0, 2.0 / 3.0
    j = 1
    for ta, tb in zip(grid[:-1], grid[1:]):
        dt = tb - ta
        k1 = func(ta, y)
        k2 = func(ta + dt * third, y + dt * k1 * third)
        k3 = func(ta + dt * two_thirds, y + dt * (k2 - k1 * third))
        k4 = func(tb, y + dt * (k1 - k2 + k3))
        y1 = y + (k1 + 3 * (k2 + k3) + k4) * dt * 0.125
        while j < len(t) and tb >= t[j]:
            if t[j] == tb:
                sol.append(y1)
            else:  # linear interpolation branch (not hit for t = [0, 1])
                sol.append(y + (t[j] - ta) / (tb - ta) * (y1 - y))
            j += 1
        y = y1
    return torch.stack(sol)


def install() -> None:
    """Insert the stubs and put the reference root on sys.path (idempotent)."""
    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        timm_models = types.ModuleType("timm.models")
        timm_vt = types.ModuleType("timm.models.vision_transformer")
        timm_vt.Mlp = _Mlp
        timm.models = timm_models
        timm_models.vision_transformer = timm_vt
        sys.modules.update({"timm": timm, "timm.models": timm_models,
                            "timm.models.vision_transformer": timm_vt})
    if "xformers" not in sys.modules:
        xf = types.ModuleType("xformers")
        xf_ops = types.ModuleType("xformers.ops")
        xf_ops.memory_efficient_attention = _memory_efficient_attention
        xf.ops = xf_ops
        sys.modules.update({"xformers": xf, "xformers.ops": xf_ops})
    if "torchdiffeq" not in sys.modules:
        td = types.ModuleType("torchdiffeq")
        td.odeint = _odeint
        sys.modules["torchdiffeq"] = td
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load_reference():
    """Return (ViT, wrappers) imported from the unmodified reference tree."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    install()
    # the reference's package names ("nn", "models", "experiments") are generic: import
    # them fresh from REFERENCE_ROOT and make sure we did not pick up something else
    import importlib
    vit = importlib.import_module("nn.vit")
    assert vit.__file__.startswith(REFERENCE_ROOT), vit.__file__
    cc = importlib.import_module("experiments.calochallenge.calochallenge_cfm.model")
    cg = importlib.import_module("experiments.calogan.model")
    ch = importlib.import_module("experiments.calohadronic.model")
    lm = importlib.import_module("experiments.lemurs.model")
    return types.SimpleNamespace(
        ViT=vit.ViT, vit=vit, LEMURSCFM=lm.LEMURSCFM,
        CaloChallengeCFM=cc.CaloChallengeCFM, CaloChallengeCFM_DS1=cc.CaloChallengeCFM_DS1,
        CaloGANCFM=cg.CaloGANCFM, CaloHadCFM=ch.CaloHadCFM,
    )


def rerandomise_zero_init(net: nn.Module, seed: int = 1, std: float = 0.02) -> None:
    """The reference zero-initialises every adaLN Linear and the final Linear
    (reference nn/vit.py:174-183), which makes the network output identically 0.  Parity
    tests re-draw those layers N(0, std) in both paths."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in net.named_parameters():
            if "adaLN_modulation" in name or name.startswith("final_layer.linear"):
                p.copy_(torch.randn(p.shape, generator=g) * std)
            elif name.endswith(".bias"):
                # biases are zero at init too; give them signal so bias paths are checked
                p.copy_(torch.randn(p.shape, generator=g) * std)
