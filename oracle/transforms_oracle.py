"""CPU restatement of the REVERSE pass of the reference's CaloChallenge ds2 / ds3 shape-model transforms.
TEST INFRASTRUCTURE ONLY (checker of vit4hep_b200.postprocess / v4h_postprocess_showers).

Plain torch-on-CPU fp32, written from the algorithm; every step cites the reference lines it restates
(/root/reference/experiments/calochallenge/transforms.py).  Pinned by tests/golden/postprocess_ds2.npz, which
oracle/make_golden.py generates by running the UNMODIFIED reference classes back to front like
experiments/calochallenge/experiment.py:286-289 does.
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
