"""The data feed on the GPU: pre-processing of raw showers and a device-resident dataset.

The reference's ``CaloChallengeDataset`` (experiments/calochallenge/datasets.py:9-75) reads the showers from HDF5,
runs its transform objects forwards one after the other on the CPU (:44-47; ~10 elementwise passes and two
45-iteration Python loops over a (100 000, 6 480) tensor for ds2), splits, and hands tensors to a ``DataLoader``.
Here :class:`FusedForwardTransforms` takes the same ordered transform configuration (the ``data.transforms`` mapping
of e.g. reference configs/calochallenge/cfm/calochallenge_ds2.yaml:15-28) and applies all of it in one launch
(two when GlobalStandardizeFromFile has to compute its mean / std first) on the device the training step reads
from; :class:`ShowerDataset` keeps the result resident in HBM and serves shuffled batches with no host round trip.

Unsupported chains raise ``NotImplementedError`` (there is no CPU fallback).  Reading HDF5 needs ``h5py``, which is
imported only inside :func:`load_hdf5`; arrays from any other source go through :meth:`ShowerDataset.from_arrays`.
"""
from __future__ import annotations

import os
from typing import Iterator, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi
from .postprocess import FusedReverseTransforms

__all__ = ["FusedForwardTransforms", "ShowerDataset", "layer_boundaries_from_xml", "load_hdf5", "save_hdf5"]


def layer_boundaries_from_xml(xml_filename: str, particle_type: str) -> np.ndarray:
    """Voxel offsets of the calorimeter layers from a CaloChallenge binning XML: per layer (number of radial bins)
    x ``n_bin_alpha`` voxels, cumulated, empty layers dropped — what ``np.unique(XMLHandler(...).GetBinEdges())``
    gives in the reference (experiments/calo_utils/ugr_evaluation/XMLHandler.py:56-71,113-121,
    experiments/calochallenge/utils.py:12-14)."""
    import xml.etree.ElementTree as ET
    root = ET.parse(xml_filename).getroot()
    for particle in root:
        if particle.attrib.get("name") == particle_type:
            edges = [0]
            for layer in particle:
                r_bins = len(layer.attrib["r_edges"].split(",")) - 1
                edges.append(edges[-1] + r_bins * int(layer.attrib["n_bin_alpha"]))
            return np.unique(np.asarray(edges, dtype=np.int64))
    raise ValueError(f"Particle {particle_type} not found in {xml_filename}")


def load_hdf5(hdf5_file: str) -> Tuple[np.ndarray, np.ndarray]:
    """(showers (N, V), incident energies (N, 1)) of a CaloChallenge file (reference
    experiments/calochallenge/utils.py:21-29 slices the same two datasets layer by layer and concatenates them back,
    :34-55)."""
    try:
        import h5py
    except ImportError as e:  # pragma: no cover - h5py is not part of this image
        raise ImportError("reading CaloChallenge HDF5 files needs h5py; pass arrays to ShowerDataset.from_arrays "
                          "instead") from e
    with h5py.File(hdf5_file, "r") as f:
        return f["showers"][:], f["incident_energies"][:].reshape(-1, 1)


def save_hdf5(hdf5_file: str, showers, incident_energies) -> None:
    """Write generated showers in the CaloChallenge layout (`showers`, `incident_energies`, gzip), what the reference's
    ``save_sample`` does (experiments/calochallenge/experiment.py:305-310).  Device tensors are copied to the host."""
    try:
        import h5py
    except ImportError as e:  # pragma: no cover - h5py is not part of this image
        raise ImportError("writing CaloChallenge HDF5 files needs h5py") from e
    to_np = lambda a: a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    with h5py.File(hdf5_file, "w") as f:
        f.create_dataset("incident_energies", data=to_np(incident_energies), compression="gzip")
        f.create_dataset("showers", data=to_np(showers), compression="gzip")


class FusedForwardTransforms:
    """Forward pass of NormalizeByElayer, ScaleTotalEnergy, CutValues, ExclusiveLogitTransform(rescale=True),
    GlobalStandardizeFromFile, LogEnergy, ScaleEnergy, AddFeaturesToCond, Reshape on the GPU.

    ``transforms`` / ``layer_boundaries`` as for :class:`FusedReverseTransforms`.  ``mean`` / ``std``: the scalars of
    GlobalStandardizeFromFile; when omitted they are loaded from ``means.npy`` / ``stds.npy`` under the transform's
    ``model_dir`` if present, and otherwise computed from the first batch of showers that is transformed (and saved
    there by rank 0) — the reference's behaviour (transforms.py:29-63)."""

    def __init__(self, transforms: Mapping[str, Optional[Mapping]], layer_boundaries: Sequence[int],
                 mean: Optional[float] = None, std: Optional[float] = None):
        self._transforms = {k: dict(v or {}) for k, v in transforms.items()}
        self._cfg = FusedReverseTransforms(transforms, layer_boundaries, 0.0, 1.0)   # validates the chain
        gs = self._transforms["GlobalStandardizeFromFile"]
        if not gs.get("exclude_zeros", True) or float(gs.get("eps", 1.0e-6)) != 1.0e-6:
            raise NotImplementedError("GlobalStandardizeFromFile: only exclude_zeros=True, eps=1e-6 is implemented")
        self.model_dir = gs.get("model_dir")
        self.shape = [int(s) for s in self._transforms["Reshape"]["shape"]]
        self.n_layers, self.voxels = self._cfg.n_layers, self._cfg.voxels
        if (mean is None) != (std is None):
            raise ValueError("give both mean and std, or neither")
        if mean is None and self.model_dir is not None:
            try:
                mean = float(np.load(os.path.join(self.model_dir, "means.npy")))
                std = float(np.load(os.path.join(self.model_dir, "stds.npy")))
            except FileNotFoundError:
                pass
        self.mean, self.std = mean, std
        self._bounds_dev, self._ms_dev = {}, {}

    @property
    def written(self) -> bool:
        return self.mean is not None

    def reverse(self) -> FusedReverseTransforms:
        """The matching post-processing (needs mean / std: given, loaded, or computed by a forward call)."""
        if not self.written:
            raise RuntimeError("mean / std are not known yet: transform the training showers first")
        return FusedReverseTransforms(self._transforms, self._cfg.bounds, self.mean, self.std)

    def __call__(self, showers: torch.Tensor, energy: torch.Tensor, rank: int = 0):
        """showers (N, V) raw energies per voxel, energy (N, 1) or (N,) incident energies (both on the GPU, float32)
        -> (x (N, *Reshape.shape), cond (N, n_layers + 1))"""
        if not (showers.is_cuda and energy.is_cuda):
            raise RuntimeError("FusedForwardTransforms runs on a B200 GPU only (no CPU fallback)")
        if showers.dtype != torch.float32 or energy.dtype != torch.float32:
            raise TypeError("showers and energy must be float32")
        N = showers.shape[0]
        if showers.dim() != 2 or showers.shape[1] != self.voxels or energy.numel() != N:
            raise ValueError(f"expected showers (N, {self.voxels}) and N incident energies, got {tuple(showers.shape)} "
                             f"and {tuple(energy.shape)}")
        dev = showers.device
        _cabi.require_device(dev.index if dev.index is not None else torch.cuda.current_device())
        raw, e = showers.contiguous(), energy.reshape(N).contiguous()
        x = torch.empty((N, self.voxels), dtype=torch.float32, device=dev)
        cond = torch.empty((N, self.n_layers + 1), dtype=torch.float32, device=dev)
        if N == 0:
            return x.reshape(0, *self.shape), cond
        if dev not in self._bounds_dev:
            self._bounds_dev[dev] = torch.tensor(self._cfg.bounds, dtype=torch.int32, device=dev)
        compute = not self.written
        if compute or self._ms_dev.get(dev, (None,))[0] != (self.mean, self.std):
            self._ms_dev[dev] = ((self.mean, self.std), torch.tensor([0.0, 1.0] if compute else [self.mean, self.std],
                                                                     dtype=torch.float32, device=dev),
                                 torch.zeros(3, dtype=torch.float64, device=dev))
        _, mean_std, stats = self._ms_dev[dev]
        c = self._cfg
        with torch.cuda.device(dev):
            _cabi.check(_cabi.load().v4h_preprocess_showers(
                raw.data_ptr(), e.data_ptr(), N, self.voxels, self.n_layers, self._bounds_dev[dev].data_ptr(), c.max_layer, c.eps,
                c.factor, c.delta, c.alpha, c.e_min, c.e_max, mean_std.data_ptr(), int(compute), stats.data_ptr(),
                x.data_ptr(), cond.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        if compute:
            self.mean, self.std = (float(v) for v in mean_std.tolist())
            if rank == 0 and self.model_dir is not None:
                np.save(os.path.join(self.model_dir, "means.npy"), np.float32(self.mean))
                np.save(os.path.join(self.model_dir, "stds.npy"), np.float32(self.std))
        return x.reshape(N, *self.shape), cond


class ShowerDataset:
    """Device-resident drop-in for the reference's ``CaloChallengeDataset`` (experiments/calochallenge/datasets.py):
    same constructor arguments and attributes (``layers``, ``energy``, ``min_bounds``, ``max_bounds``, ``len``,
    indexing -> ``(layers[i], energy[i])``), with ``transform`` a :class:`FusedForwardTransforms` instead of a list of
    CPU transform objects, plus :meth:`batches` in place of a ``DataLoader``."""

    def __init__(self, hdf5_file, particle_type=None, xml_filename=None, train_val_frac=(0.7, 0.3), transform=None,
                 split="full", device="cuda", dtype=torch.float32, rank=0):
        showers, energy = load_hdf5(hdf5_file)
        self._setup(showers, energy, train_val_frac, transform, split, device, dtype, rank)

    @classmethod
    def from_arrays(cls, showers, incident_energies, train_val_frac=(0.7, 0.3), transform=None, split="full",
                    device="cuda", dtype=torch.float32, rank=0) -> "ShowerDataset":
        """The same dataset from in-memory arrays (numpy or torch; (N, V) and (N, 1) or (N,))."""
        self = cls.__new__(cls)
        self._setup(showers, incident_energies, train_val_frac, transform, split, device, dtype, rank)
        return self

    def _setup(self, showers, energy, train_val_frac, transform, split, device, dtype, rank):
        if split not in ("full", "training", "validation"):
            raise ValueError(f"unknown split {split!r}")
        assert split == "full" or train_val_frac[0] + train_val_frac[1] <= 1.0
        self.transform, self.device, self.dtype = transform, torch.device(device), dtype
        layers = torch.as_tensor(showers, dtype=torch.float32).to(self.device, non_blocking=True)
        energy = torch.as_tensor(energy, dtype=torch.float32)
        width = energy.numel() // len(layers) if len(layers) else (energy.shape[-1] if energy.dim() > 1 else 1)
        energy = energy.reshape(len(layers), width).to(self.device, non_blocking=True)
        # pre-process ALL showers before the split, like the reference (the statistics of GlobalStandardizeFromFile
        # are those of the whole file)                                                          datasets.py:44-61
        if transform is not None:
            layers, energy = transform(layers, energy, rank=rank)
        n = len(energy)
        val_size, trn_size = int(n * train_val_frac[1]), int(n * train_val_frac[0])
        if split == "training":
            layers, energy = layers[:trn_size], energy[:trn_size]
        elif split == "validation":
            layers, energy = layers[n - val_size:], energy[n - val_size:]
        self.layers = layers.to(dtype=dtype)
        self.energy = energy.to(dtype=dtype)
        self.min_bounds = self.layers.min() if len(self.layers) else None
        self.max_bounds = self.layers.max() if len(self.layers) else None

    def __len__(self) -> int:
        return len(self.energy)

    def __getitem__(self, idx):
        return self.layers[idx], self.energy[idx]

    def batches(self, batch_size: int, shuffle: bool = True, drop_last: bool = False,
                generator: Optional[torch.Generator] = None, rank: int = 0,
                world_size: int = 1) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        """One epoch of ``(x, cond)`` batches gathered on the device (the permutation is drawn on the device too;
        ``generator`` must then be a generator of that device).

        Data-parallel training (``world_size`` > 1): every rank draws the SAME permutation — pass generators seeded
        alike, as ``DistributedSampler`` does with its epoch seed — and takes the indices ``rank, rank + world_size,
        ...`` of it; the tail that does not divide by ``world_size`` is dropped so that all ranks run the same
        number of steps."""
        n = len(self)
        if not 0 <= rank < world_size:
            raise ValueError(f"rank {rank} outside world_size {world_size}")
        if world_size > 1 or shuffle:
            order = (torch.randperm(n, device=self.device, generator=generator) if shuffle
                     else torch.arange(n, device=self.device))
            if world_size > 1:
                order = order[: n - n % world_size][rank::world_size]
        else:
            order = None
        n = n if order is None else len(order)
        stop = n - (n % batch_size) if drop_last else n
        for lo in range(0, stop, batch_size):
            hi = min(lo + batch_size, n)
            if order is None:
                yield self.layers[lo:hi], self.energy[lo:hi]
            else:
                idx = order[lo:hi]
                yield self.layers.index_select(0, idx), self.energy.index_select(0, idx)
