"""Fused tail of the training step for vit4hep_b200.ViT.

Replaces, for the parameters of one ViT, the reference's

    grad_norm = clip_grad_norm_(model.parameters(), max_grad_norm)     # experiments/base_experiment.py:573-585
    optimizer.step()                                                   # AdamW, :592

by two native launches: the squared global gradient norm (v4h_grad_norm_sq) and ONE multi-tensor pass
(v4h_adamw_step) that applies the clip coefficient, the AdamW update (torch.optim.AdamW arithmetic) and
rewrites the bf16 tensor-core operand copy of every GEMM weight, so the next forward does not recast.
No host synchronisation: the norm stays on the device (``last_grad_norm`` is a 0-dim CUDA tensor).
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional

import torch

from . import _cabi

__all__ = ["FusedAdamW"]


class FusedAdamW(torch.optim.Optimizer):
    """AdamW + gradient-norm clipping for the parameters of a ``vit4hep_b200.ViT``.

    ``lr`` may be changed through ``param_groups[0]["lr"]`` (torch LR schedulers work unchanged).
    ``state_dict()`` carries exp_avg / exp_avg_sq / step per parameter like torch.optim.AdamW.
    """

    def __init__(self, net, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2,
                 max_grad_norm: Optional[float] = None):
        params = [p for p in net.parameters() if p.requires_grad]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.net = net
        self.max_grad_norm = max_grad_norm
        self.last_grad_norm: Optional[torch.Tensor] = None
        self._jobs_host = None
        self._jobs_dev = None
        self._norm_sq = None
        self._lr_dev = None
        self._key = None
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
            st["step"] = torch.tensor(0.0)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

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
        states = [self._state_for(p) for p in params]
        key = tuple((p.data_ptr(), p.grad.data_ptr(), targets.get(id(p), (0, False))[0]) for p in params)
        n = len(params)
        if self._norm_sq is None:
            self._norm_sq = torch.zeros(1, dtype=torch.float32, device=dev)
            self._step_dev = torch.full((1,), self._step, dtype=torch.int32, device=dev)
            self._lr_dev = torch.full((1,), float(group["lr"]), dtype=torch.float32, device=dev)
            self._lr_host = float(group["lr"])
            self._tables = {}
        if key != self._key:
            # one immutable (pinned, device) table per distinct set of addresses (gradient buffers move
            # rarely and between few places): no reuse hazard, so no stream synchronisation -- which also
            # keeps step() legal under CUDA-graph capture
            if key not in self._tables:
                nbytes = n * ctypes.sizeof(_cabi.AdamWJob)
                host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
                table = (_cabi.AdamWJob * n).from_address(host.data_ptr())
                for i, (p, st) in enumerate(zip(params, states)):
                    table[i].p, table[i].g = p.data_ptr(), p.grad.data_ptr()
                    table[i].m, table[i].v = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
                    addr, is_f32 = targets.get(id(p), (0, False))
                    table[i].bf16_dst = None if (is_f32 or not addr) else addr
                    table[i].f32_dst = addr if (is_f32 and addr) else None
                    table[i].n = p.numel()
                devt = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                devt.copy_(host, non_blocking=True)
                if len(self._tables) >= 64:
                    self._tables.pop(next(iter(self._tables)))
                self._tables[key] = (host, devt)
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
        with torch.cuda.device(dev):
            # step count and learning rate travel through device scalars so that a CUDA graph of this call
            # (GraphedTrainStep) keeps the bias correction and the schedule right on every replay
            _cabi.check(lib.v4h_counter_increment(self._step_dev.data_ptr(), stream))
            _cabi.check(lib.v4h_adamw_step(self._jobs_dev.data_ptr(), n, max(p.numel() for p in params), norm_ptr,
                                           float(self.max_grad_norm or 0.0), float(group["lr"]), float(b1), float(b2),
                                           float(group["eps"]), float(group["weight_decay"]), self._step,
                                           self._step_dev.data_ptr(), self._lr_dev.data_ptr(), stream))
        # the bf16 arena now matches the parameters: spare the next forward its recast
        nat = getattr(self.net, "_native", None)
        if targets and nat is not None:
            ordered = self.net.ordered_parameters()
            if all(id(p) in {id(q) for q in params} or id(p) not in targets for _, p in ordered):
                nat.arena_key = tuple((p.data_ptr(), p._version) for _, p in ordered)
        return None
