"""Fused tail of the training step for vit4hep_b200.ViT.

Replaces, for the parameters of one ViT, the reference's

    grad_norm = clip_grad_norm_(model.parameters(), max_grad_norm)     # experiments/base_experiment.py:573-585
    optimizer.step()                                                   # AdamW, :592
    ema.update()                                                       # torch_ema, :594 (optional)

by two native launches: the squared global gradient norm (v4h_grad_norm_sq) and ONE multi-tensor pass
(v4h_adamw_step) that applies the clip coefficient, the AdamW update (torch.optim.AdamW arithmetic), the
exponential moving average of the parameters (torch_ema arithmetic) and rewrites the bf16 tensor-core
operand copy of every GEMM weight, so the next forward does not recast.
No host synchronisation: the norm stays on the device (``last_grad_norm`` is a 0-dim CUDA tensor).
"""
from __future__ import annotations

import contextlib
import ctypes
import weakref
from typing import Iterable, List, Optional

import torch

from . import _cabi

__all__ = ["FusedAdamW", "ExponentialMovingAverage"]


def _job_table(entries, dev):
    """(pinned host table, device copy) of v4h_adamw_job entries; ``entries`` = dicts of addresses."""
    n = len(entries)
    nbytes = n * ctypes.sizeof(_cabi.AdamWJob)
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    table = (_cabi.AdamWJob * n).from_address(host.data_ptr())
    for i, e in enumerate(entries):
        table[i].p, table[i].g = e.get("p"), e.get("g")
        table[i].m, table[i].v = e.get("m"), e.get("v")
        table[i].bf16_dst, table[i].f32_dst = e.get("bf16_dst"), e.get("f32_dst")
        table[i].ema = e.get("ema")
        table[i].n = e["n"]
    devt = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    devt.copy_(host, non_blocking=True)
    return host, devt


class ExponentialMovingAverage:
    """Exponential moving average of a set of parameters with torch_ema's interface and arithmetic (the
    reference builds ``torch_ema.ExponentialMovingAverage(model.parameters(), decay)``,
    experiments/base_experiment.py:127-134, calls ``update()`` after every optimizer step, :594, validates
    under ``average_parameters()``, :630-632, and checkpoints ``state_dict()``, :674; torch_ema is a third-party
    package that is absent here, its published update rule is restated):

        num_updates += 1;  d = min(decay, (1 + num_updates) / (10 + num_updates))
        shadow -= (1 - d) * (shadow - param)

    ``update()`` is one native multi-tensor launch (v4h_ema_update); when the object is handed to
    ``FusedAdamW(..., ema=...)`` the optimizer pass performs the update itself and the ``update()`` call
    that follows ``optimizer.step()`` in the reference's loop is recognised and skipped.
    """

    def __init__(self, parameters: Iterable[torch.nn.Parameter], decay: float, use_num_updates: bool = True):
        if decay < 0.0 or decay > 1.0:
            raise ValueError("Decay must be between 0 and 1")
        self.decay = decay
        self.num_updates = 0 if use_num_updates else None
        parameters = list(parameters)
        self.shadow_params: List[torch.Tensor] = [p.clone().detach() for p in parameters]
        self.collected_params: Optional[List[torch.Tensor]] = None
        self._params_refs = [weakref.ref(p) for p in parameters]
        self._updates_dev: Optional[torch.Tensor] = None
        self._tables = {}
        self._fused_pending = False  # set by FusedAdamW.step(): the next update() call is already done

    # ---- torch_ema API
    def _get_parameters(self, parameters):
        if parameters is None:
            parameters = [p() for p in self._params_refs]
            if any(p is None for p in parameters):
                raise ValueError("(One of) the parameters with which this ExponentialMovingAverage was "
                                 "initialized no longer exists (was garbage collected); please provide `parameters`")
            return parameters
        parameters = list(parameters)
        if len(parameters) != len(self.shadow_params):
            raise ValueError("Number of parameters passed as argument is different from number of shadow "
                             "parameters maintained by this ExponentialMovingAverage")
        return parameters

    def _device_counter(self, dev) -> torch.Tensor:
        if self._updates_dev is None or self._updates_dev.device != dev:
            self._updates_dev = torch.full((1,), int(self.num_updates or 0), dtype=torch.int32, device=dev)
        return self._updates_dev

    def sync_num_updates(self) -> None:
        """host counter <- device counter (CUDA-graph replays only advance the latter)"""
        if self._updates_dev is not None and self.num_updates is not None:
            self.num_updates = int(self._updates_dev.item())

    @torch.no_grad()
    def update(self, parameters=None) -> None:
        if self._fused_pending:  # FusedAdamW.step() of this iteration has applied it in its own pass
            self._fused_pending = False
            return
        params = [p for p in self._get_parameters(parameters)]
        pairs = [(p, s) for p, s in zip(params, self.shadow_params) if p.requires_grad]
        if not pairs:
            return
        dev = pairs[0][0].device
        if dev.type != "cuda":
            raise RuntimeError("ExponentialMovingAverage.update runs on a B200 GPU only (no CPU fallback)")
        for p, s in pairs:
            if p.dtype != torch.float32 or s.dtype != torch.float32 or not p.is_contiguous() or s.device != dev:
                raise TypeError("ExponentialMovingAverage needs contiguous float32 parameters and shadows on one device")
        lib = _cabi.load()
        stream = torch.cuda.current_stream(dev).cuda_stream
        key = tuple((p.data_ptr(), s.data_ptr()) for p, s in pairs)
        if key not in self._tables:
            if len(self._tables) >= 16:
                self._tables.pop(next(iter(self._tables)))
            self._tables[key] = _job_table([dict(p=p.data_ptr(), ema=s.data_ptr(), n=p.numel()) for p, s in pairs], dev)
        _, devt = self._tables[key]
        decay = float(self.decay)
        with torch.cuda.device(dev):
            if self.num_updates is not None:
                cnt = self._device_counter(dev)  # created from the host count BEFORE this update is booked
                self.num_updates += 1
                _cabi.check(lib.v4h_counter_increment(cnt.data_ptr(), stream))
                cnt_ptr = cnt.data_ptr()
            else:  # fixed decay: a huge update count makes (1 + n) / (10 + n) irrelevant
                cnt_ptr = None
            _cabi.check(lib.v4h_ema_update(devt.data_ptr(), len(pairs), max(p.numel() for p, _ in pairs), decay,
                                           int(self.num_updates) if self.num_updates is not None else (1 << 30),
                                           cnt_ptr, stream))

    @torch.no_grad()
    def copy_to(self, parameters=None) -> None:
        for s, p in zip(self.shadow_params, self._get_parameters(parameters)):
            p.data.copy_(s.data)

    def store(self, parameters=None) -> None:
        self.collected_params = [p.clone() for p in self._get_parameters(parameters)]

    @torch.no_grad()
    def restore(self, parameters=None) -> None:
        if self.collected_params is None:
            raise RuntimeError("This ExponentialMovingAverage has no `store()`ed weights to `restore()`")
        for c, p in zip(self.collected_params, self._get_parameters(parameters)):
            p.data.copy_(c.data)

    @contextlib.contextmanager
    def average_parameters(self, parameters=None):
        parameters = self._get_parameters(parameters)
        self.store(parameters)
        self.copy_to(parameters)
        try:
            yield
        finally:
            self.restore(parameters)

    def to(self, device=None, dtype=None) -> None:
        def move(t):
            return t.to(device=device, dtype=dtype) if t.is_floating_point() else t.to(device=device)
        self.shadow_params = [move(s) for s in self.shadow_params]
        if self.collected_params is not None:
            self.collected_params = [move(c) for c in self.collected_params]
        self._tables.clear()
        self._updates_dev = None

    def state_dict(self) -> dict:
        self.sync_num_updates()
        return {"decay": self.decay, "num_updates": self.num_updates, "shadow_params": self.shadow_params,
                "collected_params": self.collected_params}

    def load_state_dict(self, state_dict: dict) -> None:
        import copy
        state_dict = copy.deepcopy(state_dict)
        self.decay = state_dict["decay"]
        if self.decay < 0.0 or self.decay > 1.0:
            raise ValueError("Decay must be between 0 and 1")
        self.num_updates = state_dict["num_updates"]
        assert self.num_updates is None or isinstance(self.num_updates, int), "Invalid num_updates"
        shadow = state_dict["shadow_params"]
        assert isinstance(shadow, list) and all(isinstance(p, torch.Tensor) for p in shadow), \
            "shadow_params must all be Tensors"
        if len(shadow) != len(self.shadow_params):
            raise ValueError("Tried to `load_state_dict()` with the wrong number of parameters in the saved state.")
        # keep the placement of the current shadows (the reference loads on the CPU and moves later)
        self.shadow_params = [s.to(device=old.device, dtype=old.dtype) for s, old in zip(shadow, self.shadow_params)]
        self.collected_params = state_dict["collected_params"]
        self._tables.clear()
        self._updates_dev = None


class FusedAdamW(torch.optim.Optimizer):
    """AdamW + gradient-norm clipping (+ parameter EMA) for the parameters of a ``vit4hep_b200.ViT``.

    ``lr`` may be changed through ``param_groups[0]["lr"]`` (torch LR schedulers work unchanged).
    ``state_dict()`` / ``load_state_dict()`` carry exp_avg / exp_avg_sq / step per parameter like
    torch.optim.AdamW, so the reference's warm start (experiments/base_experiment.py:377-388) resumes the
    moments and the bias correction.
    """

    def __init__(self, net, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2,
                 max_grad_norm: Optional[float] = None, ema: Optional[ExponentialMovingAverage] = None):
        params = [p for p in net.parameters() if p.requires_grad]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.net = net
        self.max_grad_norm = max_grad_norm
        self.ema = ema
        self.last_grad_norm: Optional[torch.Tensor] = None
        self._jobs_host = None
        self._jobs_dev = None
        self._norm_sq = None
        self._lr_dev = None
        self._step_dev = None
        self._key = None
        self._tables = {}
        self._step = 0

    def sync_lr(self):
        """push param_groups[0]["lr"] to the device scalar the kernel reads (call before a graph replay when a
        scheduler changed it; step() does it itself)"""
        lr = float(self.param_groups[0]["lr"])
        if getattr(self, "_lr_dev", None) is not None and lr != self._lr_host:
            self._lr_dev.fill_(lr)
            self._lr_host = lr

    def _state_for(self, p):
        st = self.state[p]
        if not st:
            st["step"] = torch.tensor(float(self._step))
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    # ---- checkpointing (a loaded state must reach the kernel's job table and step counters)
    def sync_step(self) -> int:
        """host step count <- device step counter (CUDA-graph replays only advance the latter)"""
        if self._step_dev is not None:
            self._step = int(self._step_dev.item())
            for st in self.state.values():
                if "step" in st:
                    st["step"] = torch.tensor(float(self._step))
        return self._step

    def state_dict(self):
        self.sync_step()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        # the moments are new tensors: every cached job table points at the old ones
        self._tables.clear()
        self._key = None
        self._jobs_host = self._jobs_dev = None
        steps = set()
        for p, st in self.state.items():
            if "step" in st:
                steps.add(int(float(st["step"])))
                st["step"] = torch.tensor(float(st["step"]))  # host scalar like torch.optim.AdamW (capturable=False)
            for k in ("exp_avg", "exp_avg_sq"):
                if k in st:
                    st[k] = st[k].to(device=p.device, dtype=torch.float32).contiguous()
        if len(steps) > 1:
            raise ValueError(f"FusedAdamW keeps ONE step count for all parameters, the loaded state has {sorted(steps)}")
        self._step = steps.pop() if steps else 0
        if self._step_dev is not None:
            self._step_dev.fill_(self._step)

    def _arena_targets(self):
        """parameter -> (device address of its copy in the ViT's weight arena, is_fp32) (bf16 precision only)"""
        net = self.net
        nat = getattr(net, "_native", None)
        if nat is None or nat.plan is None or nat.arena is None:
            return {}
        lib = _cabi.load()
        out = {}
        for field, p in net.ordered_parameters():
            off = lib.v4h_vit_arena_offset(nat.plan, field.encode())
            if off >= 0:
                out[id(p)] = (nat.arena.data_ptr() + off, field.endswith("_b"))
        return out

    def _ema_shadows(self, params, dev):
        """parameter id -> device address of its EMA shadow (empty without an attached EMA)"""
        ema = self.ema
        if ema is None:
            return {}
        owners = [r() for r in ema._params_refs]
        out = {}
        for q, s in zip(owners, ema.shadow_params):
            if q is None:
                continue
            if s.device != dev or s.dtype != torch.float32 or not s.is_contiguous():
                raise TypeError("the attached ExponentialMovingAverage must live on the parameters' device in float32 "
                                "(call ema.to(device) like the reference does)")
            out[id(q)] = s.data_ptr()
        missing = [p for p in params if id(p) not in out]
        if missing:
            raise ValueError("the attached ExponentialMovingAverage does not track every optimised parameter")
        return out

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise NotImplementedError("FusedAdamW does not take a closure")
        group = self.param_groups[0]
        params = [p for p in group["params"] if p.grad is not None]
        if not params:
            return None
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdamW runs on a B200 GPU only (no CPU fallback)")
        lib = _cabi.load()
        stream = torch.cuda.current_stream(dev).cuda_stream
        for p in params:
            if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_contiguous() \
                    or not p.grad.is_contiguous():
                raise TypeError("FusedAdamW needs contiguous float32 parameters and gradients")
        targets = self._arena_targets()
        shadows = self._ema_shadows(params, dev)
        states = [self._state_for(p) for p in params]
        key = tuple((p.data_ptr(), p.grad.data_ptr(), targets.get(id(p), (0, False))[0], st["exp_avg"].data_ptr(),
                     st["exp_avg_sq"].data_ptr(), shadows.get(id(p), 0)) for p, st in zip(params, states))
        n = len(params)
        if self._norm_sq is None:
            self._norm_sq = torch.zeros(1, dtype=torch.float32, device=dev)
            self._step_dev = torch.full((1,), self._step, dtype=torch.int32, device=dev)
            self._lr_dev = torch.full((1,), float(group["lr"]), dtype=torch.float32, device=dev)
            self._lr_host = float(group["lr"])
        if key != self._key:
            # one immutable (pinned, device) table per distinct set of addresses (gradient buffers move
            # rarely and between few places): no reuse hazard, so no stream synchronisation -- which also
            # keeps step() legal under CUDA-graph capture
            if key not in self._tables:
                entries = []
                for p, st in zip(params, states):
                    addr, is_f32 = targets.get(id(p), (0, False))
                    entries.append(dict(p=p.data_ptr(), g=p.grad.data_ptr(), m=st["exp_avg"].data_ptr(),
                                        v=st["exp_avg_sq"].data_ptr(),
                                        bf16_dst=None if (is_f32 or not addr) else addr,
                                        f32_dst=addr if (is_f32 and addr) else None,
                                        ema=shadows.get(id(p)) or None, n=p.numel()))
                if len(self._tables) >= 64:
                    self._tables.pop(next(iter(self._tables)))
                self._tables[key] = _job_table(entries, dev)
            self._jobs_host, self._jobs_dev = self._tables[key]
            self._key = key
        self.sync_lr()
        norm_ptr = None
        if self.max_grad_norm is not None:
            # gradients of a vit4hep_b200.ViT are views of one flat buffer: one pass over it
            lo = min(p.grad.data_ptr() for p in params)
            hi = max(p.grad.data_ptr() + 4 * p.numel() for p in params)
            store = params[0].grad.untyped_storage()
            one_buffer = all(p.grad.untyped_storage().data_ptr() == store.data_ptr() for p in params) and \
                hi - lo <= store.nbytes() and sum(p.numel() for p in params) * 4 >= (hi - lo) - 16 * len(params)
            with torch.cuda.device(dev):
                if one_buffer:  # alignment padding between the views is zero: it does not change the norm
                    _cabi.check(lib.v4h_grad_norm_sq(lo, (hi - lo) // 4, self._norm_sq.data_ptr(), stream))
                else:  # gradients live in separate allocations (not produced by the fused backward)
                    self._norm_sq.copy_(torch.stack([p.grad.square().sum() for p in params]).sum().reshape(1))
            norm_ptr = self._norm_sq.data_ptr()
            self.last_grad_norm = self._norm_sq.sqrt().squeeze(0)
        self._step += 1
        for st in states:
            st["step"] += 1
        b1, b2 = group["betas"]
        ema_decay, ema_n, ema_ptr = 0.0, 0, None
        with torch.cuda.device(dev):
            # step count and learning rate travel through device scalars so that a CUDA graph of this call
            # (GraphedTrainStep) keeps the bias correction and the schedule right on every replay
            _cabi.check(lib.v4h_counter_increment(self._step_dev.data_ptr(), stream))
            if self.ema is not None:
                ema = self.ema
                ema_decay = float(ema.decay)
                if ema.num_updates is not None:
                    cnt = ema._device_counter(dev)  # created from the host count BEFORE this update is booked
                    ema.num_updates += 1
                    _cabi.check(lib.v4h_counter_increment(cnt.data_ptr(), stream))
                    ema_n, ema_ptr = int(ema.num_updates), cnt.data_ptr()
                else:
                    ema_n = 1 << 30
                ema._fused_pending = True
            _cabi.check(lib.v4h_adamw_step(self._jobs_dev.data_ptr(), n, max(p.numel() for p in params), norm_ptr,
                                           float(self.max_grad_norm or 0.0), float(group["lr"]), float(b1), float(b2),
                                           float(group["eps"]), float(group["weight_decay"]), self._step,
                                           self._step_dev.data_ptr(), self._lr_dev.data_ptr(), ema_decay, ema_n, ema_ptr,
                                           stream))
        # the bf16 arena now matches the parameters: spare the next forward its recast
        nat = getattr(self.net, "_native", None)
        if targets and nat is not None:
            ordered = self.net.ordered_parameters()
            if all(id(p) in {id(q) for q in params} or id(p) not in targets for _, p in ordered):
                nat.arena_key = tuple((p.data_ptr(), p._version) for _, p in ordered)
        return None
