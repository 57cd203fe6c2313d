"""Drop-ins for the reference's CFM wrappers on B200.

Mirrors (same class names, constructor arguments, methods and attributes):

* ``models.base_model.CFM``                                   reference models/base_model.py:159-247
* ``...calochallenge_cfm.model.CaloChallengeCFM`` / ``_DS1``   reference experiments/calochallenge/calochallenge_cfm/model.py:8-173
* ``experiments.calogan.model.CaloGANCFM``                    reference experiments/calogan/model.py:8-121
* ``experiments.calohadronic.model.CaloHadCFM``               reference experiments/calohadronic/model.py:8-120
* ``experiments.lemurs.model.LEMURSCFM``                      reference experiments/lemurs/model.py:8-99

``to_patches`` / ``from_patches`` are bit-exact permutation kernels (v4h_to_patches / v4h_from_patches),
``_batch_loss`` fuses the linear trajectory with patchification (v4h_cfm_prepare) and the MSE with its
gradient (v4h_cfm_loss), ``sample_batch`` integrates the ODE with torchdiffeq's fixed-grid schemes
('rk4' = 3/8 rule, 'euler', 'midpoint') entirely in token layout with one fused kernel per stage
(v4h_axpy4).  Random draws are made exactly where and how the reference makes them, so identical
seeds give identical ``t``, ``x_0`` and ``x_T``.
"""
from __future__ import annotations

import contextlib
import ctypes
import math
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import _cabi

__all__ = ["PatchGeometry", "GraphedTrainStep", "CFM", "CaloChallengeCFM", "CaloChallengeCFM_DS1", "CaloGANCFM", "CaloHadCFM",
           "LEMURSCFM", "fixed_grid", "linear_trajectory"]


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda_f32(name: str, v: torch.Tensor) -> None:
    if not v.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: vit4hep_b200 has no CPU fallback")
    if v.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {v.dtype}")


class PatchGeometry:
    """(layer, angular, radial) voxel grids <-> (tokens, patch_dim), regular or segmented.

    Host side of ``v4h_geometry``: ``shapes`` / ``patches`` are lists of (L, A, R) / (P1, P2, P3), one per
    calorimeter segment; ``flat_input`` says the sample is (C, sum V) split at the segment edges
    instead of (C, L, A, R)."""

    def __init__(self, shapes: Sequence[Sequence[int]], patches: Sequence[Sequence[int]], in_channels: int = 1,
                 flat_input: bool = False):
        self.shapes = [tuple(int(v) for v in s) for s in shapes]
        self.patches = [tuple(int(v) for v in p) for p in patches]
        if len(self.shapes) != len(self.patches):
            raise AssertionError("list_shape and list_patch_shape must have the same length")
        self.in_channels = int(in_channels)
        self.flat_input = bool(flat_input)
        n = len(self.shapes)
        sh = (ctypes.c_int32 * (3 * n))(*[v for s in self.shapes for v in s])
        pa = (ctypes.c_int32 * (3 * n))(*[v for p in self.patches for v in p])
        handle = ctypes.c_void_p()
        lib = _cabi.load()
        rc = lib.v4h_geometry_create(sh, pa, n, self.in_channels, int(self.flat_input), ctypes.byref(handle))
        if rc != _cabi.V4H_OK:
            msg = _cabi.last_error()
            if "should be divisible" in msg:  # the reference raises AssertionError for this
                raise AssertionError(msg)
            raise RuntimeError(f"v4h_geometry_create failed: {msg}")
        self._handle = handle
        self.tokens = lib.v4h_geometry_tokens(handle)
        self.patch_dim = lib.v4h_geometry_patch_dim(handle)
        self.voxels = lib.v4h_geometry_voxels(handle)  # per sample, channels included
        if self.flat_input:
            self.sample_shape = (self.in_channels, self.voxels // self.in_channels)
        else:
            self.sample_shape = (self.in_channels, *self.shapes[0])

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                _cabi.load().v4h_geometry_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    def index_table(self) -> np.ndarray:
        """int32 gather table: tokens.flat[j] = x.flat[table[j]] per sample (host copy)."""
        buf = (ctypes.c_int32 * self.voxels)()
        _cabi.check(_cabi.load().v4h_geometry_table_host(self._handle, buf, self.voxels))
        return np.ctypeslib.as_array(buf).copy()

    def _check(self, name, v, per_sample):
        _require_cuda_f32(name, v)
        got = math.prod(v.shape[1:]) if v.dim() >= 2 else -1
        if got != per_sample:
            raise ValueError(f"{name} has {got} values per sample, expected {per_sample}")

    def to_patches(self, x: torch.Tensor) -> torch.Tensor:
        self._check("x", x, self.voxels)
        x = x.contiguous()
        B = x.shape[0]
        out = torch.empty((B, self.tokens, self.patch_dim), dtype=torch.float32, device=x.device)
        if B == 0:
            return out
        with torch.cuda.device(x.device):
            _cabi.check(_cabi.load().v4h_to_patches(self._handle, x.data_ptr(), out.data_ptr(), B, _stream(x.device)))
        return out

    def from_patches(self, tok: torch.Tensor) -> torch.Tensor:
        self._check("tokens", tok, self.voxels)
        tok = tok.contiguous()
        B = tok.shape[0]
        out = torch.empty((B, *self.sample_shape), dtype=torch.float32, device=tok.device)
        if B == 0:
            return out
        with torch.cuda.device(tok.device):
            _cabi.check(_cabi.load().v4h_from_patches(self._handle, tok.data_ptr(), out.data_ptr(), B,
                                                      _stream(tok.device)))
        return out

    def cfm_prepare(self, x1: torch.Tensor, x0: torch.Tensor, t: torch.Tensor):
        """x_t = (1 - t) x0 + t x1 and x1 - x0, both directly in token layout."""
        self._check("x1", x1, self.voxels)
        self._check("x0", x0, self.voxels)
        _require_cuda_f32("t", t)
        B = x1.shape[0]
        if t.numel() != B:
            raise ValueError("t must hold one time per sample")
        x1, x0, t = x1.contiguous(), x0.contiguous(), t.contiguous()
        xt = torch.empty((B, self.tokens, self.patch_dim), dtype=torch.float32, device=x1.device)
        target = torch.empty_like(xt)
        with torch.cuda.device(x1.device):
            _cabi.check(_cabi.load().v4h_cfm_prepare(self._handle, x1.data_ptr(), x0.data_ptr(), t.data_ptr(),
                                                     xt.data_ptr(), target.data_ptr(), B, _stream(x1.device)))
        return xt, target


class _Permute(torch.autograd.Function):
    """to_patches / from_patches with the inverse permutation as backward."""

    @staticmethod
    def forward(ctx, geom: PatchGeometry, x, to_tokens: bool):
        ctx.geom, ctx.to_tokens = geom, to_tokens
        return geom.to_patches(x) if to_tokens else geom.from_patches(x)

    @staticmethod
    def backward(ctx, g):
        geom = ctx.geom
        return None, (geom.from_patches(g) if ctx.to_tokens else geom.to_patches(g)), None


class _MSELoss(torch.autograd.Function):
    """mean((v - target)^2) with d loss / d v produced by the same kernel (reference models/base_model.py:217-218)."""

    @staticmethod
    def forward(ctx, v, target):
        _require_cuda_f32("velocity", v)
        v, target = v.contiguous(), target.contiguous()
        loss = torch.empty((), dtype=torch.float32, device=v.device)
        need = ctx.needs_input_grad[0]
        dv = torch.empty_like(v) if need else None
        with torch.cuda.device(v.device):
            _cabi.check(_cabi.load().v4h_cfm_loss(v.data_ptr(), target.data_ptr(), v.numel(), 1.0, loss.data_ptr(),
                                                  None if dv is None else dv.data_ptr(), _stream(v.device)))
        ctx.dv = dv
        return loss

    @staticmethod
    def backward(ctx, g):
        dv, ctx.dv = ctx.dv, None
        return dv * g, None


def linear_trajectory(x_0, x_1, t):
    """reference models/trajectories.py:5-8 (host-visible helper; the training path uses the fused kernel)."""
    return (1 - t) * x_0 + t * x_1, x_1 - x_0


def fixed_grid(step_size: float, t0: float = 0.0, t1: float = 1.0) -> torch.Tensor:
    """torchdiffeq's fixed integration grid from ``options.step_size``: ceil((t1 - t0) / step + 1) points in
    fp32, the last forced onto t1."""
    a, b = torch.tensor(t0, dtype=torch.float32), torch.tensor(t1, dtype=torch.float32)
    n = int(torch.ceil((b - a) / step_size + 1).item())
    grid = torch.arange(0, n, dtype=torch.float32) * step_size + a
    grid[-1] = b
    return grid


# Butcher-style stage tables of torchdiffeq's fixed-grid solvers: per stage the time offset (fraction of dt)
# and the coefficients of k1.. in the stage state; `final` combines the stages.
_SCHEMES = {
    "euler": dict(c=[0.0], a=[[]], final=[1.0]),
    "midpoint": dict(c=[0.0, 0.5], a=[[], [0.5]], final=[0.0, 1.0]),
    # 'rk4' in torchdiffeq is rk4_alt_step_func, the 3/8 rule
    "rk4": dict(c=[0.0, 1.0 / 3.0, 2.0 / 3.0, 1.0],
                a=[[], [1.0 / 3.0], [-1.0 / 3.0, 1.0], [1.0, -1.0, 1.0]],
                final=[0.125, 0.375, 0.375, 0.125]),
}


class CFM(nn.Module):
    """Conditional-flow-matching model around a velocity network (reference models/base_model.py:159-247)."""

    def __init__(self, net, time_distribution="uniform", trajectory="linear", odeint_kwargs=None, *args,
                 shape=None, **kwargs):
        super().__init__()
        if args:
            shape = args[0]
        self.shape = shape
        self.time_distribution = self.get_time_distribution(time_distribution)
        self.trajectory = self.get_trajectory(trajectory)
        self.odeint_kwargs = odeint_kwargs
        self.net = net
        self._geometry: Optional[PatchGeometry] = None
        # extension: replay the whole ODE solve of a sample batch as one CUDA graph (launch-bound loop of
        # 80 network evaluations); off by default because it pins the batch size's buffers
        self.graph_sampling = False
        self._sample_graphs = {}

    # -- reference API ---------------------------------------------------------------
    def get_trajectory(self, trajectory):
        if trajectory == "linear":
            return linear_trajectory
        raise ValueError

    def get_time_distribution(self, time_distribution):
        if time_distribution == "uniform":
            return torch.distributions.uniform.Uniform(low=0.0, high=1.0)
        raise ValueError

    def build_net(self):
        raise NotImplementedError

    # -- geometry ----------------------------------------------------------------------
    def _make_geometry(self) -> Optional[PatchGeometry]:
        """Base CFM: no patch geometry, the network takes the sample as it is (reference CFM.forward calls the net
        directly, models/base_model.py:199-201; used with the energy-ratio network, shape [45])."""
        return None

    @property
    def geometry(self) -> Optional[PatchGeometry]:
        if self._geometry is None:
            self._geometry = self._make_geometry()
        return self._geometry

    def to_patches(self, x):
        return x if self.geometry is None else _Permute.apply(self.geometry, x, True)

    def from_patches(self, x):
        return x if self.geometry is None else _Permute.apply(self.geometry, x, False)

    def _net(self):
        """the ViT itself when `net` was wrapped (e.g. by DistributedDataParallel)"""
        return self.net

    # -- forward / loss ----------------------------------------------------------------
    def forward(self, x, t, c):
        """from_patches(net(to_patches(x), t, c)) (reference calochallenge_cfm/model.py:62-66)"""
        return self.from_patches(self.net(self.to_patches(x), t, c))

    def _batch_loss(self, x, device_rng: bool = False):
        """CFM loss of one batch (x, c) (reference models/base_model.py:203-218): t ~ U(0,1) drawn on the host
        with shape (B, 1, ..), x_0 = randn_like(x) on the device, linear trajectory, MSE on the velocity.
        ``device_rng`` (extension, used under CUDA-graph capture) draws t from the device generator instead."""
        x, c = x[0], x[1]
        device = getattr(self, "device", None) or x.device
        dtype = getattr(self, "dtype", None) or torch.float32
        if dtype != torch.float32:
            raise TypeError("vit4hep_b200 computes with fp32 inputs; model.dtype must be torch.float32")
        x = x.to(dtype=dtype, device=device, non_blocking=True)
        c = c.to(dtype=dtype, device=device, non_blocking=True)
        if device_rng:
            if self.time_distribution.low != 0.0 or self.time_distribution.high != 1.0:
                raise NotImplementedError("device_rng supports the uniform(0, 1) time distribution")
            t = torch.rand([x.shape[0]] + [1] * (x.dim() - 1), device=device, dtype=dtype)
        else:
            t = self.time_distribution.sample([x.shape[0]] + [1] * (x.dim() - 1))
            t = t.to(device, dtype, non_blocking=True)
        x_0 = torch.randn_like(x)
        if self.geometry is None:
            raise NotImplementedError("CFM._batch_loss without a patch geometry (the energy-ratio network) is not "
                                      "implemented: vit4hep_b200.ParallelTransformer is forward-only")
        x_t, target = self.geometry.cfm_prepare(x, x_0, t.view(-1))
        velocity = self.net(x_t, t.view(-1, 1), c)
        return _MSELoss.apply(velocity, target)

    # -- sampling ----------------------------------------------------------------------
    def _stage_times(self, times, device):
        """stage times of the fixed-grid solve as a cached device tensor (no host copy inside a graph capture)"""
        cache = self.__dict__.setdefault("_time_cache", {})
        key = (times, device)
        if key not in cache:
            cache[key] = torch.tensor(times, dtype=torch.float32).to(device)
        return cache[key]

    def _sample_noise(self, batch):
        return torch.randn((batch.shape[0], *self.shape), dtype=batch.dtype, device=batch.device)

    @torch.inference_mode()
    def sample_batch(self, batch):
        """Solve dx/dt = v(x, t, cond) from Gaussian noise at t=0 to t=1 (reference
        calochallenge_cfm/model.py:68-94 with torchdiffeq's fixed-grid solver)."""
        x_T = self._sample_noise(batch)
        return self.integrate(x_T, batch)

    @torch.inference_mode()
    def integrate(self, x_T: torch.Tensor, cond: torch.Tensor) -> torch.Tensor:
        if self.graph_sampling and x_T.is_cuda:
            return self._integrate_graphed(x_T, cond)
        return self._integrate(x_T, cond)

    def _integrate_graphed(self, x_T, cond):
        """one CUDA graph per (batch shape, device): static input / output buffers, the eager solve captured once"""
        key = (tuple(x_T.shape), tuple(cond.shape), x_T.device)
        entry = self._sample_graphs.get(key)
        if entry is None:
            xs, cs = torch.empty_like(x_T), torch.empty_like(cond)
            xs.copy_(x_T); cs.copy_(cond)
            side = torch.cuda.Stream(device=x_T.device)
            side.wait_stream(torch.cuda.current_stream(x_T.device))
            with torch.cuda.stream(side):  # warm-up off the capture: plans, tensor maps, weight arena
                self._integrate(xs, cs, max_steps=1)
            torch.cuda.current_stream(x_T.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="relaxed"):
                out = self._integrate(xs, cs)
            entry = (graph, xs, cs, out)
            self._sample_graphs[key] = entry
        graph, xs, cs, out = entry
        xs.copy_(x_T); cs.copy_(cond)
        graph.replay()
        return out.clone()

    def _integrate(self, x_T: torch.Tensor, cond: torch.Tensor, max_steps: Optional[int] = None) -> torch.Tensor:
        kw = dict(self.odeint_kwargs or {})
        method = kw.get("method", None)
        if method not in _SCHEMES:
            raise NotImplementedError(f"odeint method {method!r}: only torchdiffeq's fixed-grid "
                                      f"{sorted(_SCHEMES)} are implemented")
        step = (kw.get("options") or {}).get("step_size")
        if step is None:
            raise NotImplementedError("odeint_kwargs.options.step_size is required (fixed grid)")
        _require_cuda_f32("conditions", cond)
        scheme = _SCHEMES[method]
        grid = fixed_grid(float(step))
        # every stage time of the whole solve, computed like torchdiffeq does (fp32 tensor arithmetic)
        times = []
        for ta, tb in zip(grid[:-1], grid[1:]):
            dt = tb - ta
            for frac in scheme["c"]:
                times.append(tb if frac == 1.0 else ta + dt * frac)
        t_dev = self._stage_times(tuple(float(v) for v in times), cond.device)
        geom = self.geometry
        lib = _cabi.load()
        net = self._net()
        _require_cuda_f32("x_T", x_T)
        y = x_T.contiguous().clone() if geom is None else geom.to_patches(x_T)
        n = y.numel()
        stage_y = torch.empty_like(y)
        s = _stream(y.device)
        nstage = len(scheme["c"])

        def axpy(out, base, ks, coefs):
            ptrs = [k.data_ptr() for k in ks] + [None] * (4 - len(ks))
            cf = list(coefs) + [0.0] * (4 - len(coefs))
            _cabi.check(lib.v4h_axpy4(out.data_ptr(), base.data_ptr(), ptrs[0], cf[0], ptrs[1], cf[1], ptrs[2], cf[2],
                                      ptrs[3], cf[3], n, s))

        i = 0
        # a network whose condition side is independent of (x, t) (the energy network's encoder) computes it once
        # per solve inside this scope -- also inside a captured solve, whose replays then re-encode new conditions
        scope = getattr(net, "condition_scope", None)
        with torch.cuda.device(y.device), (scope() if scope is not None else contextlib.nullcontext()):
            for step, (ta, tb) in enumerate(zip(grid[:-1], grid[1:])):
                if max_steps is not None and step >= max_steps:
                    break
                dt = float(tb - ta)
                ks: List[torch.Tensor] = []
                for st in range(nstage):
                    if st == 0:
                        src = y
                    else:
                        axpy(stage_y, y, ks, [dt * a for a in scheme["a"][st]])
                        src = stage_y
                    ks.append(net(src, t_dev[i:i + 1], cond, shared_t=True))
                    i += 1
                axpy(y, y, ks, [dt * b for b in scheme["final"]])
        return y if geom is None else geom.from_patches(y)


class GraphedTrainStep:
    """One training step -- ``_batch_loss`` (device RNG), ``zero_grad``, backward (with the data-parallel
    all-reduce when enabled), optimizer step -- captured once as a CUDA graph and replayed per batch.

    Extension for launch-bound regimes (the reference's ``BaseExperiment._step`` issues the same work
    eagerly, experiments/base_experiment.py:555-597).  ``step(x, c)`` copies the batch into the static input
    buffers (host tensors should be pinned) and returns the static 0-dim loss tensor of the replay.

    Construct it before any eager step of the same model on the default stream, or drop every reference to the
    losses of such steps first: a live autograd graph keeps its gradient-accumulation nodes bound to the stream it was
    built on, torch would make that (legacy) stream wait on the capture and the capture fails with
    cudaErrorStreamCaptureImplicit (the usual rule for whole-step CUDA graphs in torch)."""

    def __init__(self, model, optimizer, x_example: torch.Tensor, c_example: torch.Tensor, warmup: int = 3):
        dev = getattr(model, "device", None) or next(model.parameters()).device
        self.model, self.optimizer = model, optimizer
        self.x = torch.empty(x_example.shape, dtype=torch.float32, device=dev)
        self.c = torch.empty(c_example.shape, dtype=torch.float32, device=dev)
        self.x.copy_(x_example); self.c.copy_(c_example)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(0, warmup)):  # 0: the caller has already run this step eagerly (lazy allocations)
                self._eager()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        # relaxed: the optimizer may allocate a pinned job table, the autograd thread runs the backward
        try:
            with torch.cuda.graph(self.graph, capture_error_mode="relaxed"):
                self.loss = self._eager()
        except RuntimeError as e:  # torch.AcceleratorError is a RuntimeError
            if "capture" in str(e).lower():
                raise RuntimeError(
                    "GraphedTrainStep: the CUDA-graph capture of the training step failed. The usual cause is an autograd "
                    "graph of an earlier EAGER step of this model that is still alive (e.g. a `loss` tensor still "
                    "referenced): its gradient-accumulation nodes are bound to the stream that step ran on. Drop those "
                    "references (del loss) or build the GraphedTrainStep before the first eager step.") from e
            raise

    def _eager(self):
        self.optimizer.zero_grad(set_to_none=True)
        loss = self.model._batch_loss((self.x, self.c), device_rng=True)
        loss.backward()
        self.optimizer.step()
        ema = getattr(self.optimizer, "ema", None)
        if ema is not None:  # the fused pass has updated it; nobody calls ema.update() inside the graph
            ema._fused_pending = False
        return loss.detach()

    def step(self, x: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
        self.x.copy_(x, non_blocking=True)
        self.c.copy_(c, non_blocking=True)
        if hasattr(self.optimizer, "sync_lr"):
            self.optimizer.sync_lr()
        self.graph.replay()
        return self.loss

    # ---- input prefetch: the host -> device copy of the NEXT batch runs on its own stream under the current
    # step (what a DataLoader with pin_memory + non_blocking copies does for an eager loop)
    def stage(self, x: torch.Tensor, c: torch.Tensor) -> None:
        """start copying a (pinned) host batch into the staging buffers; returns immediately"""
        if not hasattr(self, "_stage_stream"):
            self._stage_stream = torch.cuda.Stream(device=self.x.device)
            self._stage_x, self._stage_c = torch.empty_like(self.x), torch.empty_like(self.c)
            self._staged = self._consumed = None
        if self._consumed is not None:  # the previous staged batch must have been copied out
            self._stage_stream.wait_event(self._consumed)
        with torch.cuda.stream(self._stage_stream):
            self._stage_x.copy_(x, non_blocking=True)
            self._stage_c.copy_(c, non_blocking=True)
            self._staged = self._stage_stream.record_event()

    def step_staged(self) -> torch.Tensor:
        """one training step on the batch passed to the last ``stage`` call"""
        cur = torch.cuda.current_stream(self.x.device)
        cur.wait_event(self._staged)
        self.x.copy_(self._stage_x, non_blocking=True)
        self.c.copy_(self._stage_c, non_blocking=True)
        self._consumed = cur.record_event()
        if hasattr(self.optimizer, "sync_lr"):
            self.optimizer.sync_lr()
        self.graph.replay()
        return self.loss


class CaloChallengeCFM(CFM):
    """Regular (L, A, R) grid: CaloChallenge ds2 / ds3 (reference calochallenge_cfm/model.py:8-94)."""

    def __init__(self, net, patch_shape, in_channels=1, time_distribution="uniform", trajectory="linear",
                 odeint_kwargs=None, *args, **kwargs):
        super().__init__(None, time_distribution, trajectory, odeint_kwargs, *args, **kwargs)
        self.patch_shape = patch_shape
        self.num_patches = [s // p for s, p in zip(self.shape, self.patch_shape)]
        self.in_channels = in_channels
        for i, (s, p) in enumerate(zip(self.shape, self.patch_shape)):
            assert s % p == 0, f"Input size ({s}) should be divisible by patch size ({p}) in axis {i}."
        self.net = net

    def _make_geometry(self):
        return PatchGeometry([self.shape], [self.patch_shape], self.in_channels, flat_input=False)

    def _sample_noise(self, batch):
        return torch.randn((batch.shape[0], self.in_channels, *self.shape), dtype=batch.dtype, device=batch.device)


class _SegmentedCFM(CFM):
    """Flat (B, C, sum V) input split at ``list_edges`` into per-layer grids with their own patch shapes."""

    def _init_segments(self, net, list_shape, list_edges, list_patch_shape, in_channels):
        self.list_shape = list(list_shape)
        self.list_edges = list(list_edges)
        self.list_patch_shape = [list(p) for p in list_patch_shape]
        self.in_channels = in_channels
        if len(self.list_shape) != len(self.list_patch_shape):
            raise AssertionError("list_shape and list_patch_shape must have the same length")
        self.num_patches_per_dim = []
        self.num_patches_per_layer = []
        for i, (shape, patch) in enumerate(zip(self.list_shape, self.list_patch_shape)):
            for L, m in zip(shape, patch):
                assert L % m == 0, f"Input size ({L}) should be divisible by patch size ({m}) in axis {i}."
            dims = tuple(s // p for s, p in zip(shape, patch))
            self.num_patches_per_dim.append(dims)
            self.num_patches_per_layer.append(math.prod(dims))
        for shape, edge in zip(self.list_shape, self.list_edges):
            if math.prod(shape) != edge:
                raise AssertionError(f"list_edges entry {edge} does not match layer shape {list(shape)}")
        self.net = net
        self.net.num_patches = self.num_patches_per_dim

    def _make_geometry(self):
        return PatchGeometry(self.list_shape, self.list_patch_shape, self.in_channels, flat_input=True)

    def _sample_noise(self, batch):
        return torch.randn((batch.shape[0], self.in_channels, *self.shape), dtype=batch.dtype, device=batch.device)


class CaloChallengeCFM_DS1(_SegmentedCFM):
    """CaloChallenge ds1: irregular layers, one shared patch shape (reference calochallenge_cfm/model.py:97-173)."""

    def __init__(self, net, list_shape, list_edges, patch_shape, in_channels=1, time_distribution="uniform",
                 trajectory="linear", odeint_kwargs=None, *args, **kwargs):
        super().__init__(None, time_distribution, trajectory, odeint_kwargs, *args, **kwargs)
        self.patch_shape = patch_shape
        self._init_segments(net, list_shape, list_edges, [patch_shape] * len(list(list_shape)), in_channels)


class CaloGANCFM(_SegmentedCFM):
    """CaloGAN 3-layer geometry with per-layer patch shapes (reference experiments/calogan/model.py:8-121)."""

    def __init__(self, net, list_shape, list_edges, list_patch_shape, in_channels=1, time_distribution="uniform",
                 trajectory="linear", odeint_kwargs=None, *args, **kwargs):
        super().__init__(None, time_distribution, trajectory, odeint_kwargs, *args, **kwargs)
        self._init_segments(net, list_shape, list_edges, list_patch_shape, in_channels)


class CaloHadCFM(CaloGANCFM):
    """CaloHadronic ECal + HCal geometry (reference experiments/calohadronic/model.py:8-120)."""


class LEMURSCFM(CaloChallengeCFM):
    """LEMURS: ds2-like grid whose batches arrive as (B, R, A, L) (reference experiments/lemurs/model.py:8-99)."""

    def _batch_loss(self, x, device_rng: bool = False):
        x = list(x)
        x[0] = x[0].permute(0, 3, 2, 1).unsqueeze(1)  # layers first, then add the channel axis
        return super()._batch_loss(x, device_rng=device_rng)
