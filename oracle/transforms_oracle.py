"""CPU restatement of the REVERSE and FORWARD passes of the reference's CaloChallenge ds2 / ds3 shape-model
transforms.  TEST INFRASTRUCTURE ONLY (checker of vit4hep_b200.postprocess / v4h_postprocess_showers and of
vit4hep_b200.preprocess / v4h_preprocess_showers).

Plain torch-on-CPU fp32, written from the algorithm; every step cites the reference lines it restates
(/root/reference/experiments/calochallenge/transforms.py).  Pinned by tests/golden/postprocess_ds2.npz, which
oracle/make_golden.py generates by running the UNMODIFIED reference classes back to front like
experiments/calochallenge/experiment.py:286-289 does, and by tests/golden/preprocess_ds2.npz (the same classes run
front to back like experiments/calochallenge/datasets.py:44-47).
"""
from __future__ import annotations

from typing import Sequence

import torch


def reverse_chain(samples: torch.Tensor, cond: torch.Tensor, bounds: Sequence[int], *, mean: float, std: float,
                  delta: float, cut: float, factor: float, e_min: float, e_max: float, alpha: float = 0.0,
                  eps: float = 1.0e-10, norm_cut: float = 0.0):
    """samples (N, *grid), cond (N, L + 1) = [u_0..u_{L-1}, scaled log E] -> (showers (N, V), E_inc (N, 1))."""
    L = len(bounds) - 1
    N = samples.shape[0]
    x = samples.reshape(N, -1)                                   # Reshape rev            transforms.py:323-326
    e, us = cond[:, -1:], cond[:, :-1]                           # AddFeaturesToCond rev  :139-142
    x = torch.cat([x, us], dim=1)
    e = e * (e_max - e_min)                                      # ScaleEnergy rev        :217-220
    e = e + e_min
    e = torch.exp(e) - alpha                                     # LogEnergy rev          :159-161
    x = x * std + mean                                           # GlobalStandardize rev  :50-52
    z = torch.sigmoid(x)                                         # ExclusiveLogit rev     :240-243, logit(inv) :11-14
    x = (z - delta) / (1 - 2 * delta)
    if cut:                                                      # CutValues rev          :302-308
        mask = x <= cut
        mask[:, -L:] = False
        x = x.masked_fill(mask, 0.0)
    x = x.clone()
    x[..., -L] = x[..., -L] / factor                             # ScaleTotalEnergy rev   :197-199
    us = x[:, -L:].clone()                                       # NormalizeByElayer rev  :344-378
    us[:, 1:] = us[:, 1:].clamp(0.0, 1.0)
    vox = x[:, :-L]
    total = e.flatten() * us[:, 0]
    cum = torch.zeros_like(total)
    layer_es = []
    for i in range(L - 1):
        le = (total - cum) * us[:, i + 1]
        layer_es.append(le)
        cum = cum + le
    layer_es.append(total - cum)
    out = torch.zeros_like(vox)
    for l, (a, b) in enumerate(zip(bounds[:-1], bounds[1:])):
        layer = vox[:, a:b] / (vox[:, a:b].sum(-1, keepdim=True) + eps)
        layer = layer.masked_fill(layer <= norm_cut, 0.0)
        out[:, a:b] = layer * layer_es[l][:, None]
    return out, e


def forward_chain(showers: torch.Tensor, e_inc: torch.Tensor, bounds: Sequence[int], *, delta: float, factor: float,
                  e_min: float, e_max: float, alpha: float = 0.0, eps: float = 1.0e-10, mean=None, std=None,
                  shape=None, std_eps: float = 1.0e-6):
    """showers (N, V) raw energies, e_inc (N, 1) -> (x (N, *shape), cond (N, L + 1), mean, std).  mean / std None:
    computed like GlobalStandardizeFromFile's first call."""
    L = len(bounds) - 1
    x = showers.clone()
    layer_es = []                                                # NormalizeByElayer fwd      transforms.py:380-398
    for a, b in zip(bounds[:-1], bounds[1:]):
        le = x[:, a:b].sum(dim=1, keepdim=True)
        x[:, a:b] = x[:, a:b] / (le + eps)
        layer_es.append(le)
    layer_es = torch.cat(layer_es, dim=1)
    us = [layer_es.sum(dim=1, keepdim=True) / e_inc]
    for l in range(L - 1):
        us.append(layer_es[:, [l]] / (layer_es[:, l:].sum(dim=1, keepdim=True) + eps))
    x = torch.cat([x] + us, dim=1)
    x[..., -L] = x[..., -L] * factor                             # ScaleTotalEnergy fwd       :200-201
    x = torch.logit(x * (1 - 2 * delta) + delta)                 # ExclusiveLogit fwd         :244-247, logit :15-17
    if mean is None:                                             # GlobalStandardize fwd      :54-64
        sat = torch.logit(torch.tensor(std_eps))
        keep = (x > sat) & (x < -sat)
        mean, std = x[keep].mean(), x[keep].std()
    x = (x - mean) / std
    e = torch.log(e_inc + alpha)                                 # LogEnergy fwd              :162-163
    e = (e - e_min) / (e_max - e_min)                            # ScaleEnergy fwd            :222-223
    V = bounds[-1]
    cond = torch.cat([x[:, V:], e], dim=1)                       # AddFeaturesToCond fwd      :143-145
    x = x[:, :V]
    if shape is not None:                                        # Reshape fwd                :327-328
        x = x.reshape(-1, *shape)
    return x, cond, float(mean), float(std)
