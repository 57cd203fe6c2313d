"""Post-processing of sampled showers on the GPU.

The reference maps sampled showers back to detector energies by running its pre-processing transforms in
reverse, one after the other, on the CPU after ``.cpu()`` (experiments/calochallenge/experiment.py:223, :286-289;
classes in experiments/calochallenge/transforms.py).  :class:`FusedReverseTransforms` takes the same ordered
transform configuration (the ``data.transforms`` mapping of e.g. reference
configs/calochallenge/cfm/calochallenge_ds2.yaml:15-28), checks that it is the chain the fused kernel implements
(v4h_postprocess_showers) and applies all of it in one launch on the device the showers already live on:

    samples, conditions = fused(samples.squeeze(1), conditions)      # instead of: for fn in transforms[::-1]: ...

Unsupported chains raise ``NotImplementedError`` (there is no CPU fallback).
"""
from __future__ import annotations

from typing import Mapping, Optional, Sequence

import torch

from . import _cabi

__all__ = ["FusedReverseTransforms"]

# the CaloChallenge ds2 / ds3 shape-model chain, in the (forward) order of the reference's YAML
_CHAIN = ["NormalizeByElayer", "ScaleTotalEnergy", "CutValues", "ExclusiveLogitTransform", "GlobalStandardizeFromFile",
          "LogEnergy", "ScaleEnergy", "AddFeaturesToCond", "Reshape"]


class FusedReverseTransforms:
    """Reverse pass of NormalizeByElayer, ScaleTotalEnergy, CutValues, ExclusiveLogitTransform(rescale=True),
    GlobalStandardizeFromFile, LogEnergy, ScaleEnergy, AddFeaturesToCond, Reshape in one kernel.

    ``transforms``: ordered mapping name -> kwargs exactly as in the reference's YAML; ``layer_boundaries``: the voxel
    offsets of the calorimeter layers (``np.unique(XMLHandler.GetBinEdges())`` in the reference's NormalizeByElayer,
    transforms.py:339-341; for the regular ds2 / ds3 grids ``range(0, V + 1, V // 45)``); ``mean`` / ``std``: the
    scalars GlobalStandardizeFromFile loads from ``means.npy`` / ``stds.npy``."""

    def __init__(self, transforms: Mapping[str, Optional[Mapping]], layer_boundaries: Sequence[int], mean: float,
                 std: float):
        names = list(transforms)
        if names != _CHAIN:
            raise NotImplementedError(f"the fused post-processing implements the chain {_CHAIN}; got {names}")
        kw = {k: dict(v or {}) for k, v in transforms.items()}
        self.bounds = [int(b) for b in layer_boundaries]
        self.n_layers = len(self.bounds) - 1
        self.voxels = self.bounds[-1]
        self.max_layer = max(b - a for a, b in zip(self.bounds, self.bounds[1:])) if self.n_layers else 0
        if self.bounds[0] != 0 or any(b <= a for a, b in zip(self.bounds, self.bounds[1:])):
            raise ValueError("layer_boundaries must start at 0 and increase")
        nb = kw["NormalizeByElayer"]
        self.eps = float(nb.get("eps", 1.0e-10))
        self.norm_cut = float(nb.get("cut", 0.0))
        st = kw["ScaleTotalEnergy"]
        self.factor = float(st["factor"])
        cv = kw["CutValues"]
        self.cut = float(cv.get("cut", 0.0))
        el = kw["ExclusiveLogitTransform"]
        if not el.get("rescale", False) or el.get("exclusions") is not None:
            raise NotImplementedError("ExclusiveLogitTransform: only rescale=True without exclusions is implemented")
        self.delta = float(el["delta"])
        for name, n in (("ScaleTotalEnergy", st.get("n_layers", 45)), ("CutValues", cv.get("n_layers", 45))):
            if int(n) != self.n_layers:
                raise ValueError(f"{name}.n_layers = {n} but layer_boundaries describe {self.n_layers} layers")
        self.alpha = float(kw["LogEnergy"].get("alpha", 0.0))
        se = kw["ScaleEnergy"]
        self.e_min, self.e_max = float(se["e_min"]), float(se["e_max"])
        if int(kw["AddFeaturesToCond"]["split_index"]) != self.voxels:
            raise ValueError("AddFeaturesToCond.split_index must equal the number of voxels")
        shape = list(kw["Reshape"]["shape"])
        n = 1
        for s in shape:
            n *= int(s)
        if n != self.voxels:
            raise ValueError(f"Reshape.shape {shape} does not hold {self.voxels} voxels")
        self.mean, self.std = float(mean), float(std)
        self._bounds_dev = {}

    def __call__(self, samples: torch.Tensor, conditions: torch.Tensor):
        """samples (N, *grid) as returned by the shape model (channel axis squeezed or not), conditions
        (N, n_layers + 1) = [u_0 .. u_{L-1}, scaled log E_inc] -> (showers (N, V), incident energies (N, 1))"""
        if not (samples.is_cuda and conditions.is_cuda):
            raise RuntimeError("FusedReverseTransforms runs on a B200 GPU only (no CPU fallback)")
        if samples.dtype != torch.float32 or conditions.dtype != torch.float32:
            raise TypeError("samples and conditions must be float32")
        N = samples.shape[0]
        c = conditions.contiguous()
        per_sample = 1
        for d in samples.shape[1:]:
            per_sample *= int(d)
        x = samples.reshape(N, per_sample).contiguous()
        if per_sample != self.voxels or tuple(c.shape) != (N, self.n_layers + 1):
            raise ValueError(f"expected samples with {self.voxels} voxels and conditions (N, {self.n_layers + 1}), got "
                             f"{tuple(samples.shape)} and {tuple(conditions.shape)}")
        dev = x.device
        _cabi.require_device(dev.index if dev.index is not None else torch.cuda.current_device())
        out = torch.empty_like(x)
        e = torch.empty((N, 1), dtype=torch.float32, device=dev)
        if N == 0:
            return out, e
        if dev not in self._bounds_dev:
            self._bounds_dev[dev] = torch.tensor(self.bounds, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _cabi.check(_cabi.load().v4h_postprocess_showers(
                x.data_ptr(), c.data_ptr(), N, self.voxels, self.n_layers, self._bounds_dev[dev].data_ptr(), self.max_layer,
                self.mean,
                self.std, self.delta, self.cut, self.factor, self.e_min, self.e_max, self.alpha, self.eps, self.norm_cut,
                out.data_ptr(), e.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        return out, e
