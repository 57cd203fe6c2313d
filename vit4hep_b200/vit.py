"""Drop-in for the reference's ``nn.vit.ViT`` (reference nn/vit.py:12-206) on B200.

Same constructor (one ``param`` mapping with the reference's keys and defaults), same
sub-module / parameter / buffer names (so ``state_dict()`` is interchangeable with reference
checkpoints), same ``forward(x, t, c)`` signature.  The arithmetic runs in
the hand-written sm_100a kernels of ``libvit4hep_b200.so`` through the C ABI
(include/vit4hep_b200.h); the sub-modules below only *hold* the parameters.  There is no
PyTorch/CPU fallback: tensors that are not on a B200, or knobs the kernels do not implement,
raise.

One extra, optional key: ``param["precision"]`` in {"bf16", "fp32"} - arithmetic of the GEMM
operands and saved activations (the reference only has fp32; "bf16" is the B200 production mode:
bf16 tensor-core operands, fp32 accumulation, fp32 residual stream / LayerNorm / softmax).
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _cabi

__all__ = ["ViT", "DiTBlock", "FinalLayer", "TimestepEmbedder", "Attention", "Mlp", "get_sincos_pos_embed"]


def _no_eager(name: str):
    raise RuntimeError(
        f"{name}.forward is not an eager PyTorch module: it is evaluated by the fused B200 kernels through "
        "vit4hep_b200.ViT.forward (there is no fallback path)")


class TimestepEmbedder(nn.Module):
    """Parameter holder for reference nn/vit.py:354-394 (Linear -> SiLU -> Linear on a sinusoidal embedding)."""

    def __init__(self, hidden_size: int, frequency_embedding_size: int = 256):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(frequency_embedding_size, hidden_size), nn.SiLU(),
                                 nn.Linear(hidden_size, hidden_size))
        self.frequency_embedding_size = frequency_embedding_size

    def forward(self, t):
        _no_eager("TimestepEmbedder")


class Attention(nn.Module):
    """Parameter holder for reference nn/vit.py:397-454 (qkv Linear with bias, proj Linear)."""

    def __init__(self, dim: int, num_heads: int):
        super().__init__()
        if dim % num_heads != 0:
            raise AssertionError("dim should be divisible by num_heads")
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.q_norm = nn.Identity()
        self.k_norm = nn.Identity()
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        _no_eager("Attention")


class Mlp(nn.Module):
    """Parameter holder with timm's Mlp state-dict names (fc1, fc2); reference call site nn/vit.py:317-322."""

    def __init__(self, in_features: int, hidden_features: int):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = nn.GELU(approximate="tanh")
        self.fc2 = nn.Linear(hidden_features, in_features)

    def forward(self, x):
        _no_eager("Mlp")


class DiTBlock(nn.Module):
    """Parameter holder for reference nn/vit.py:302-333."""

    def __init__(self, hidden_size: int, num_heads: int, mlp_ratio: float = 4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.attn = Attention(hidden_size, num_heads)
        self.norm2 = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.mlp = Mlp(hidden_size, int(hidden_size * mlp_ratio))
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, 6 * hidden_size, bias=True))

    def forward(self, x, c):
        _no_eager("DiTBlock")


class FinalLayer(nn.Module):
    """Parameter holder for reference nn/vit.py:336-351."""

    def __init__(self, hidden_dim: int, patch_dim: int, out_channels: int = 1, x_out: int = 1):
        super().__init__()
        self.norm_final = nn.LayerNorm(hidden_dim, elementwise_affine=False, eps=1e-6)
        self.linear = nn.Linear(hidden_dim, out_channels * x_out * patch_dim)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_dim, 2 * hidden_dim))

    def forward(self, x, c):
        _no_eager("FinalLayer")


def get_sincos_pos_embed(pos_embedding_coords, num_patches, hidden_dim, dim, temperature=10000):
    """Fixed positional table for ``learn_pos_embed=False`` (reference nn/vit.py:459-540): a constant
    (T, D) buffer the x_embedder epilogue adds.  Geometric frequency ladder temperature**(-i/(F-1));
    3d tables concatenate sin/cos of (x, y, z) with x = radial, y = angular, z = layer ("cylindrical")
    or of the cartesian image of (r, alpha) ("cartesian"); the 1d table spans prod(num_patches) / 2
    positions like upstream."""
    if len(num_patches) == 1 and isinstance(num_patches[0], (list, tuple)):
        num_patches = num_patches[0]

    def ladder(n):
        return torch.pow(torch.tensor(float(temperature)), -torch.arange(n) / (n - 1))

    def sincos(coords, n):
        w = ladder(n)
        waves = [c.reshape(-1, 1) * w.reshape(1, -1) for c in coords]
        return torch.cat([f(a) for a in waves for f in (torch.sin, torch.cos)], dim=1)

    if dim == 1:
        n = int(math.prod(num_patches) / 2)
        return sincos([torch.arange(n) / n], hidden_dim // 2)
    if dim != 3 or pos_embedding_coords not in ("cylindrical", "cartesian"):
        raise ValueError(f"unknown positional embedding {pos_embedding_coords!r} in {dim}d")
    L, A, R = num_patches
    ang_step = 1.0 / A if pos_embedding_coords == "cylindrical" else 2 * math.pi / A
    z, ang, rad = torch.meshgrid(torch.arange(L) / L, torch.arange(A) * ang_step, torch.arange(R) / R,
                                 indexing="ij")
    if pos_embedding_coords == "cylindrical":
        coords = [rad, ang, z]
    else:
        coords = [rad * ang.cos(), rad * ang.sin(), z]
    return sincos(coords, hidden_dim // 6)


def _dev_ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class _Native:
    """Per-module native state: plan handles (one per shape key), bf16 weight arena, cached pointer tables."""

    def __init__(self):
        self.plans = {}       # key -> handle, kept until the module dies: a pending autograd graph may hold one
        self.plan = None      # the plan of the most recent call
        self.plan_key = None
        self.arena = None
        self.arena_key = None
        self.arena_layout = None  # the part of the plan key the arena layout depends on

    def close(self):
        lib = _cabi.load() if self.plans else None
        for handle in self.plans.values():
            lib.v4h_plan_destroy(handle)
        self.plans = {}
        self.plan = None
        self.plan_key = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ViT(nn.Module):
    """3D Vision-Transformer velocity network (DiT-style adaLN-Zero), reference nn/vit.py:12-206."""

    DEFAULTS = {
        "dim": 3, "condition_dim": 46, "hidden_dim": 180, "out_channels": 1, "depth": 2, "num_heads": 4,
        "mlp_ratio": 2.0, "attn_drop": 0.0, "proj_drop": 0.0, "pos_embedding_coords": "cartesian",
        "temperature": 10000, "learn_pos_embed": True, "causal_attn": False, "checkpoint_grads": False,
        "patch_dim": 12, "num_patches": [[15, 4, 9]], "use_torch_sdpa": True,
        # extension (not in the reference): arithmetic of GEMM operands / saved activations
        "precision": os.environ.get("V4H_PRECISION", "bf16"),
    }

    def __init__(self, param):
        super().__init__()
        for k, p in self.DEFAULTS.items():
            setattr(self, k, param[k] if k in param else p)
        if self.precision not in ("bf16", "fp32"):
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {self.precision!r}")
        if self.attn_drop != 0.0 or self.proj_drop != 0.0:
            raise NotImplementedError("attn_drop / proj_drop != 0 are not implemented (every shipped config uses 0)")
        if self.causal_attn:
            raise NotImplementedError("causal_attn is not implemented (no shipped config enables it)")
        if self.hidden_dim % self.num_heads != 0:
            raise AssertionError("dim should be divisible by num_heads")
        if self.checkpoint_grads:
            import warnings
            warnings.warn("vit4hep_b200.ViT: checkpoint_grads=True has no effect (the fused backward always uses the "
                          "saved activation workspace; results are identical)")

        self.x_embedder = nn.Linear(self.patch_dim, self.hidden_dim)
        self.c_embedder = nn.Sequential(nn.Linear(self.condition_dim, self.hidden_dim), nn.SiLU(),
                                        nn.Linear(self.hidden_dim, self.hidden_dim))
        self.t_embedder = TimestepEmbedder(self.hidden_dim)
        if self.learn_pos_embed:
            self.pos_embed_freqs = nn.Parameter(torch.randn(self.hidden_dim // 6))
            pos_z, pos_y, pos_x = self.create_meshgrid()
            self.register_buffer("pos_z", pos_z)
            self.register_buffer("pos_y", pos_y)
            self.register_buffer("pos_x", pos_x)
        else:
            self.register_buffer("pos_embed", get_sincos_pos_embed(
                self.pos_embedding_coords, self.num_patches, self.hidden_dim, self.dim, self.temperature))
        self.blocks = nn.ModuleList([DiTBlock(self.hidden_dim, self.num_heads, mlp_ratio=self.mlp_ratio)
                                     for _ in range(self.depth)])
        self.final_layer = FinalLayer(self.hidden_dim, self.patch_dim, self.out_channels, x_out=1)
        self.initialize_weights()
        self._native = _Native()
        self._dp = None  # set by vit4hep_b200.dp.enable_data_parallel
        self._aux_streams = {}

    # ------------------------------------------------------------------ reference-visible helpers
    def create_meshgrid(self):
        """pos_z / pos_y / pos_x per token (reference nn/vit.py:137-154): cumulative layer index over the
        segment list divided by the total layer count; angular / radial index over its segment's extent."""
        segs = [tuple(n) for n in self.num_patches]
        total_l = sum(s[0] for s in segs)
        zs, ys, xs = [], [], []
        start = 0
        for (L, A, R) in segs:
            z = (torch.arange(total_l) / total_l)[start:start + L]
            grid = torch.meshgrid(z, torch.arange(A) / A, torch.arange(R) / R, indexing="ij")
            zs.append(grid[0].flatten()); ys.append(grid[1].flatten()); xs.append(grid[2].flatten())
            start += L
        return torch.cat(zs), torch.cat(ys), torch.cat(xs)

    def learnable_pos_embedding(self):
        raise RuntimeError("the positional table is computed inside the fused forward (v4h_vit_forward)")

    def initialize_weights(self):
        """Xavier-uniform Linears with zero bias; zero adaLN modulation and output Linears
        (adaLN-Zero, reference nn/vit.py:164-183)."""
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        zero = [b.adaLN_modulation[-1] for b in self.blocks]
        zero += [self.final_layer.adaLN_modulation[-1], self.final_layer.linear]
        for lin in zero:
            nn.init.zeros_(lin.weight)
            nn.init.zeros_(lin.bias)

    # ------------------------------------------------------------------ parameter plumbing
    @staticmethod
    def _is_mlp2(seq) -> bool:
        return (isinstance(seq, nn.Sequential) and len(seq) == 3 and isinstance(seq[0], nn.Linear)
                and isinstance(seq[1], nn.SiLU) and isinstance(seq[2], nn.Linear))

    @classmethod
    def _mlp2(cls, seq, what: str) -> Tuple[nn.Linear, nn.Linear]:
        if not cls._is_mlp2(seq):
            raise NotImplementedError(f"{what} must be Sequential(Linear, SiLU, Linear) for the fused path (got {seq})")
        return seq[0], seq[2]

    def _x_parts(self) -> Tuple[Optional[nn.Linear], nn.Linear]:
        """(mapper or None, embedding Linear) of ``x_embedder``: a Linear, or the finetuning structure
        Sequential(mapper Linear, SiLU, old Linear) (reference experiment_finetuning.py:79-91)."""
        xe = self.x_embedder
        if isinstance(xe, nn.Linear):
            return None, xe
        if self._is_mlp2(xe):
            return xe[0], xe[2]
        raise NotImplementedError(f"x_embedder must be a Linear or Sequential(Linear, SiLU, Linear) for the fused "
                                  f"path (got {xe})")

    def _c_parts(self) -> Tuple[Optional[nn.Linear], nn.Linear, nn.Linear]:
        """(mapper or None, first, second Linear) of ``c_embedder``: Sequential(Linear, SiLU, Linear), or the
        finetuning structure Sequential(mapper Linear, SiLU, old Sequential) (reference experiment_finetuning.py:106-118)."""
        ce = self.c_embedder
        if self._is_mlp2(ce):
            return None, ce[0], ce[2]
        if (isinstance(ce, nn.Sequential) and len(ce) == 3 and isinstance(ce[0], nn.Linear)
                and isinstance(ce[1], nn.SiLU) and self._is_mlp2(ce[2])):
            return ce[0], ce[2][0], ce[2][2]
        raise NotImplementedError(f"c_embedder must be Sequential(Linear, SiLU, Linear), optionally behind a mapper "
                                  f"Linear + SiLU, for the fused path (got {ce})")

    def ordered_parameters(self) -> List[Tuple[str, nn.Parameter]]:
        """Parameters as (struct field, tensor), in the order the backward chain completes their
        gradients: final layer (with its adaLN Linear), blocks depth-1..0 (each with its adaLN Linear: a block's
        modulation gradients are final when its backward stage ends), then embeddings / conditioning (stage 0).
        The flat gradient buffer and the data-parallel buckets use this order."""
        out: List[Tuple[str, nn.Parameter]] = []
        fl = self.final_layer
        out += [("final_w", fl.linear.weight), ("final_b", fl.linear.bias),
                ("final_ada_w", fl.adaLN_modulation[-1].weight), ("final_ada_b", fl.adaLN_modulation[-1].bias)]
        for i in reversed(range(len(self.blocks))):
            b = self.blocks[i]
            ada = b.adaLN_modulation[-1]
            out += [(f"blocks.{i}.qkv_w", b.attn.qkv.weight), (f"blocks.{i}.qkv_b", b.attn.qkv.bias),
                    (f"blocks.{i}.proj_w", b.attn.proj.weight), (f"blocks.{i}.proj_b", b.attn.proj.bias),
                    (f"blocks.{i}.fc1_w", b.mlp.fc1.weight), (f"blocks.{i}.fc1_b", b.mlp.fc1.bias),
                    (f"blocks.{i}.fc2_w", b.mlp.fc2.weight), (f"blocks.{i}.fc2_b", b.mlp.fc2.bias),
                    (f"blocks.{i}.ada_w", ada.weight), (f"blocks.{i}.ada_b", ada.bias)]
        if self.learn_pos_embed:
            out.append(("pos_embed_freqs", self.pos_embed_freqs))
        xm, xe = self._x_parts()
        cm, c0, c2 = self._c_parts()
        t0, t2 = self._mlp2(self.t_embedder.mlp, "t_embedder.mlp")
        out += [("x_w", xe.weight), ("x_b", xe.bias)]
        if xm is not None:
            out += [("xm_w", xm.weight), ("xm_b", xm.bias)]
        out += [("c0_w", c0.weight), ("c0_b", c0.bias), ("c2_w", c2.weight), ("c2_b", c2.bias)]
        if cm is not None:
            out += [("cm_w", cm.weight), ("cm_b", cm.bias)]
        out += [("t0_w", t0.weight), ("t0_b", t0.bias), ("t2_w", t2.weight), ("t2_b", t2.bias)]
        return out

    def _aux_stream(self, device) -> "torch.cuda.Stream":
        """side stream for host-issued work that is independent of the kernel chain (gradient buffer clear)"""
        key = torch.device(device).index
        st = self._aux_streams.get(key)
        if st is None:
            st = self._aux_streams[key] = torch.cuda.Stream(device=device)
        return st

    FLAT_ALIGN = 4  # elements: every gradient starts on a 16-byte boundary of the flat buffer

    @classmethod
    def flat_layout(cls, ordered) -> Tuple[List[int], int]:
        """Offsets (in elements) of every parameter's gradient inside the flat gradient buffer, and its
        total length.  Each gradient starts 16-byte aligned so the native kernels (vector atomics of the
        LayerNorm backward, the fused optimizer) keep their 128-bit paths for any patch_dim / cond_dim."""
        offs, off = [], 0
        a = cls.FLAT_ALIGN
        for _, p in ordered:
            offs.append(off)
            off += (p.numel() + a - 1) // a * a
        return offs, off

    def stage_boundaries(self) -> List[int]:
        """Element offsets into the flat gradient buffer at which each backward stage's parameters
        end: [after final layer, after block depth-1, ..., after block 0, after stage 0]."""
        names = self.ordered_parameters()
        offs, total = self.flat_layout(names)
        ends = offs[1:] + [total]
        bounds = [ends[3]]                       # final_w, final_b, final_ada_w, final_ada_b
        idx = 4
        for _ in range(len(self.blocks)):
            idx += 10
            bounds.append(ends[idx - 1])
        bounds.append(total)
        return bounds

    @staticmethod
    def _fill(struct: _cabi.VitParams, field: str, ptr: int) -> None:
        if field.startswith("blocks."):
            _, i, name = field.split(".")
            setattr(struct.blocks[int(i)], name, ptr)
        else:
            setattr(struct, field, ptr)

    def _weights_struct(self, ordered) -> _cabi.VitParams:
        w = _cabi.VitParams()
        for field, p in ordered:
            self._fill(w, field, p.data_ptr())
        if self.learn_pos_embed:
            w.pos_z, w.pos_y, w.pos_x = self.pos_z.data_ptr(), self.pos_y.data_ptr(), self.pos_x.data_ptr()
        else:
            w.pos_embed = self.pos_embed.data_ptr()
        return w

    def _plan(self, tokens: int):
        xm, xe = self._x_parts()
        cm, c0, _ = self._c_parts()
        out_dim = self.final_layer.linear.out_features
        key = (self.hidden_dim, len(self.blocks), self.num_heads, self.blocks[0].mlp.fc1.out_features,
               xe.in_features, out_dim, c0.in_features, tokens,
               self.t_embedder.frequency_embedding_size, bool(self.learn_pos_embed), self.precision,
               xm.in_features if xm is not None else 0, cm.in_features if cm is not None else 0)
        nat = self._native
        if nat.plan_key != key:
            handle = nat.plans.get(key)
            if handle is None:
                dims = _cabi.VitDims(*[int(v) for v in key[:9]], int(key[9]),
                                     _cabi.V4H_BF16 if self.precision == "bf16" else _cabi.V4H_FP32,
                                     int(key[11]), int(key[12]))
                handle = ctypes.c_void_p()
                _cabi.check(_cabi.load().v4h_plan_create(ctypes.byref(dims), ctypes.byref(handle)))
                nat.plans[key] = handle
            nat.plan, nat.plan_key = handle, key
            layout = key[:7] + key[8:]  # the weight arena does not depend on the token count
            if nat.arena_layout != layout:
                nat.arena, nat.arena_key, nat.arena_layout = None, None, layout
            else:
                nat.arena_key = None  # same layout, another plan's cast-job table: refresh through this plan once
        return nat.plan

    def invalidate_weights(self) -> None:
        """Force the bf16 operand copies to be rebuilt by the next forward.  Writers that change parameter
        storage WITHOUT bumping ``Tensor._version`` (native kernels, collectives into raw pointers) must call
        this; in-place torch ops, optimizers and ``load_state_dict`` are detected through ``_version``."""
        self._native.arena_key = None

    def _prepare_arena(self, plan, ordered, w, stream: int):
        """bf16 operand copies of the GEMM weights, rebuilt whenever a parameter was written
        (optimizer step, EMA swap, load_state_dict: SURVEY.md appendix B 'live weights')."""
        if self.precision != "bf16":
            return None
        nat = self._native
        lib = _cabi.load()
        if nat.arena is None:
            nbytes = lib.v4h_vit_weight_arena_bytes(plan)
            nat.arena = torch.empty(nbytes, dtype=torch.uint8, device=ordered[0][1].device)
            nat.arena_key = None
        key = tuple((p.data_ptr(), p._version) for _, p in ordered)
        if key != nat.arena_key:
            _cabi.check(lib.v4h_vit_prepare_weights(plan, ctypes.byref(w), nat.arena.data_ptr(), stream))
            nat.arena_key = key
        return nat.arena.data_ptr()

    def _check_inputs(self, x, t, c, shared_t: bool):
        if not (x.is_cuda and t.is_cuda and c.is_cuda):
            raise RuntimeError("vit4hep_b200.ViT runs on a B200 GPU only: inputs must be CUDA tensors "
                               "(there is no CPU fallback)")
        _cabi.require_device(x.device.index if x.device.index is not None else torch.cuda.current_device())
        for name, v in (("x", x), ("t", t), ("c", c)):
            if v.dtype != torch.float32:
                raise TypeError(f"{name} must be float32 (the reference supports fp32/fp64; fp64 is not implemented)")
        xm, xe = self._x_parts()
        in_features = (xm if xm is not None else xe).in_features
        if x.dim() != 3 or x.shape[2] != in_features:
            raise ValueError(f"x must be (B, T, {in_features}), got {tuple(x.shape)}")
        B = x.shape[0]
        tokens = self.pos_z.numel() if self.learn_pos_embed else self.pos_embed.shape[0]
        if x.shape[1] != tokens:
            raise ValueError(f"x has {x.shape[1]} tokens but the positional grid has {tokens}")
        if shared_t:
            if t.numel() != 1:
                raise ValueError("shared_t expects a single time value")
        elif t.numel() != B:
            raise ValueError(f"t must hold one time per sample ({B}), got {tuple(t.shape)}")
        cm, c0, _ = self._c_parts()
        k_in = (cm if cm is not None else c0).in_features
        if c.dim() != 2 or c.shape[0] != B or c.shape[1] != k_in:
            raise ValueError(f"c must be (B, {k_in}), got {tuple(c.shape)}")
        if x.requires_grad or t.requires_grad or c.requires_grad:
            raise NotImplementedError("gradients w.r.t. x, t, c are not implemented (training feeds leaf inputs)")

    # ------------------------------------------------------------------ forward
    def forward(self, x, t, c, shared_t: bool = False):
        """x (B, T, P), t (B, 1) [or one value when ``shared_t``], c (B, K) -> (B, T, P * out_channels)
        (reference nn/vit.py:185-206)."""
        self._check_inputs(x, t, c, shared_t)
        ordered = self.ordered_parameters()
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for _, p in ordered)
        if needs_grad:
            return _ViTFunction.apply(self, x.contiguous(), t.contiguous(), c.contiguous(), shared_t,
                                      *[p for _, p in ordered])
        out, _, _ = self._run_forward(x.contiguous(), t.contiguous(), c.contiguous(), shared_t, False, ordered)
        return out

    def _run_forward(self, x, t, c, shared_t: bool, save: bool, ordered):
        lib = _cabi.load()
        B, T = x.shape[0], x.shape[1]
        plan = self._plan(T)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        w = self._weights_struct(ordered)
        arena = self._prepare_arena(plan, ordered, w, stream)
        nbytes = lib.v4h_vit_workspace_bytes(plan, B, int(save))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        out = torch.empty((B, T, self.final_layer.linear.out_features), dtype=torch.float32, device=x.device)
        _cabi.check(lib.v4h_vit_forward(plan, ctypes.byref(w), arena, x.data_ptr(), t.data_ptr(), c.data_ptr(),
                                        out.data_ptr(), B, int(shared_t), int(save), ws.data_ptr(), nbytes, stream))
        return out, ws, plan

    def _run_backward(self, plan, x, c, dout, ws, ordered, flat_grad, stage_begin: int, stage_end: int):
        lib = _cabi.load()
        stream = torch.cuda.current_stream(x.device).cuda_stream
        w = self._weights_struct(ordered)
        g = _cabi.VitParams()
        offs, _ = self.flat_layout(ordered)
        for (field, p), off in zip(ordered, offs):
            self._fill(g, field, flat_grad.data_ptr() + 4 * off)
        arena = None if self._native.arena is None else self._native.arena.data_ptr()
        _cabi.check(lib.v4h_vit_backward(plan, ctypes.byref(w), arena, ctypes.byref(g), x.data_ptr(), c.data_ptr(),
                                         dout.data_ptr(), x.shape[0], stage_begin, stage_end, ws.data_ptr(),
                                         ws.numel(), stream))


class _ViTFunction(torch.autograd.Function):
    """Autograd node of the whole network: forward saves the activation workspace, backward runs the
    native backward chain into one flat gradient buffer (optionally all-reducing it bucket by bucket
    while the chain is still running, see vit4hep_b200.dp)."""

    @staticmethod
    def forward(ctx, module: ViT, x, t, c, shared_t, *params):
        ordered = module.ordered_parameters()
        # the flat gradient buffer (split-K weight gradients accumulate into it) is allocated and cleared on a
        # side stream, next to the forward kernels, instead of at the head of the backward chain
        _, total = module.flat_layout(ordered)
        cur = torch.cuda.current_stream(x.device)
        side = module._aux_stream(x.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            flat = torch.zeros(total, dtype=torch.float32, device=x.device)
        ctx.flat, ctx.flat_ready = flat, side.record_event()
        out, ws, plan = module._run_forward(x, t, c, shared_t, True, ordered)
        ctx.plan = plan  # the backward must lay the workspace out like this forward did, whatever ran in between
        cur.wait_event(ctx.flat_ready)  # also rejoins the side stream when the step is being captured
        flat.record_stream(cur)
        ctx.module = module
        ctx.ws = ws
        ctx.save_for_backward(x, c)
        ctx.mark_non_differentiable()
        return out

    @staticmethod
    def backward(ctx, dout):
        module: ViT = ctx.module
        x, c = ctx.saved_tensors
        ordered = module.ordered_parameters()
        offs, total = module.flat_layout(ordered)
        if ctx.flat is None:
            raise RuntimeError("vit4hep_b200.ViT: the saved activations of this forward were already consumed by a "
                               "backward pass (retain_graph / a second backward are not supported: run the forward again)")
        flat, ctx.flat = ctx.flat, None
        dout = dout.contiguous()
        depth = len(module.blocks)
        if module._dp is None:
            module._run_backward(ctx.plan, x, c, dout, ctx.ws, ordered, flat, depth + 1, 0)
        else:
            module._dp.backward(module, ctx.plan, x, c, dout, ctx.ws, ordered, flat)
        ctx.ws = None
        grads = []
        for (_, p), off in zip(ordered, offs):
            grads.append(flat[off:off + p.numel()].view(p.shape) if p.requires_grad else None)
        return (None, None, None, None, None, *grads)
