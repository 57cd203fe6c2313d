"""Drop-in for the reference's energy-ratio velocity network ``nn.cfm.transformer_cfm.ParallelTransformer``
(reference nn/cfm/transformer_cfm.py:12-119) on B200, forward only.

The reference samples this small CFM (45 energy ratios conditioned on the incident energy) right before every
shape-sampling job, for the same conditions (experiments/calochallenge/experiment.py:225-247).  Same constructor
(one ``param`` mapping), same sub-module / parameter names -- the module really holds a ``torch.nn.Transformer``,
so ``state_dict()`` is interchangeable with reference checkpoints -- and the same ``forward(x, t, condition)``.
The arithmetic runs in the native library (v4h_energy_encode once per condition batch, v4h_energy_forward per
velocity evaluation); training this network is not implemented here (use the reference's module: the state
dicts are interchangeable) and raises.
"""
from __future__ import annotations

import contextlib
import ctypes
import os

import torch
import torch.nn as nn

from . import _cabi

__all__ = ["ParallelTransformer", "GaussianFourierProjection"]


class GaussianFourierProjection(nn.Module):
    """Parameter holder of reference transformer_cfm.py:165-176 (fixed random frequencies)."""

    def __init__(self, embed_dim, scale=30.0):
        super().__init__()
        self.W = nn.Parameter(torch.randn(embed_dim // 2) * scale, requires_grad=False)

    def forward(self, x):
        raise RuntimeError("GaussianFourierProjection is evaluated inside the fused energy-network forward")


class ParallelTransformer(nn.Module):
    """Velocity field of the whole energy-ratio vector in one pass (reference transformer_cfm.py:12-119)."""

    DEFAULTS = {
        "dims_in": 46, "dims_c": 1, "dim_embedding": 180, "nhead": 4, "num_encoder_layers": 2,
        "num_decoder_layers": 4, "dim_feedforward": 256, "dropout": 0.0, "activation": "relu", "embeds": False,
        "encode_t_scale": 30, "encode_t_dim": 64,
        "precision": os.environ.get("V4H_PRECISION", "bf16"),  # extension, like vit4hep_b200.ViT
    }

    def __init__(self, param):
        super().__init__()
        for k, p in self.DEFAULTS.items():
            setattr(self, k, param[k] if k in param else p)
        if not self.embeds:
            raise NotImplementedError("ParallelTransformer: only embeds=True (every shipped energy config) is implemented")
        if self.dropout != 0.0 or self.activation != "relu":
            raise NotImplementedError("ParallelTransformer: dropout != 0 / activation != 'relu' are not implemented")
        if self.precision not in ("bf16", "fp32"):
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {self.precision!r}")
        self.time_embed = nn.Sequential(GaussianFourierProjection(embed_dim=self.encode_t_dim, scale=self.encode_t_scale),
                                        nn.Linear(self.encode_t_dim, self.encode_t_dim))
        self.d_model = 2 * self.dim_embedding
        if self.encode_t_dim + self.dim_embedding != self.d_model:
            raise ValueError("embeds=True concatenates the time and x embeddings into d_model = 2 * dim_embedding: "
                             "encode_t_dim must equal dim_embedding")
        self.x_embed = nn.Linear(1, self.dim_embedding)
        self.c_embed = nn.Linear(1, 2 * self.dim_embedding)
        self.pos_embed_x = nn.Embedding(self.dims_in, self.dim_embedding)
        self.pos_embed_c = nn.Embedding(self.dims_c, 2 * self.dim_embedding)
        self.layer = nn.Linear(3 * self.dim_embedding, self.dim_feedforward)
        self.transformer = nn.Transformer(d_model=self.d_model, nhead=self.nhead,
                                          num_encoder_layers=self.num_encoder_layers,
                                          num_decoder_layers=self.num_decoder_layers,
                                          dim_feedforward=self.dim_feedforward, dropout=self.dropout,
                                          activation=self.activation, batch_first=True)
        self.layers = nn.Sequential(self.layer, nn.SiLU(), nn.Linear(self.dim_feedforward, 1))
        self._plan = None
        self._arena = None
        self._arena_key = None
        self._ws = {}        # batch size -> persistent workspace (holds the encoded condition between evaluations)
        self._encoded = {}   # batch size -> (condition data_ptr, version) the workspace was encoded for
        self._in_scope = False

    def __del__(self):
        try:
            if self._plan is not None:
                _cabi.load().v4h_energy_plan_destroy(self._plan)
                self._plan = None
        except Exception:
            pass

    # ------------------------------------------------------------------ native plumbing
    def _params_struct(self) -> _cabi.EnergyParams:
        w = _cabi.EnergyParams()
        ptr = lambda t: t.data_ptr()
        w.gfp_w = ptr(self.time_embed[0].W)
        w.time_w, w.time_b = ptr(self.time_embed[1].weight), ptr(self.time_embed[1].bias)
        w.x_embed_w, w.x_embed_b = ptr(self.x_embed.weight), ptr(self.x_embed.bias)
        w.c_embed_w, w.c_embed_b = ptr(self.c_embed.weight), ptr(self.c_embed.bias)
        w.pos_x, w.pos_c = ptr(self.pos_embed_x.weight), ptr(self.pos_embed_c.weight)
        enc, dec = self.transformer.encoder, self.transformer.decoder
        w.enc_norm_w, w.enc_norm_b = ptr(enc.norm.weight), ptr(enc.norm.bias)
        w.dec_norm_w, w.dec_norm_b = ptr(dec.norm.weight), ptr(dec.norm.bias)
        w.head0_w, w.head0_b = ptr(self.layers[0].weight), ptr(self.layers[0].bias)
        w.head2_w, w.head2_b = ptr(self.layers[2].weight), ptr(self.layers[2].bias)
        for i, L in enumerate(enc.layers):
            e = w.enc[i]
            e.in_w, e.in_b = ptr(L.self_attn.in_proj_weight), ptr(L.self_attn.in_proj_bias)
            e.out_w, e.out_b = ptr(L.self_attn.out_proj.weight), ptr(L.self_attn.out_proj.bias)
            e.l1_w, e.l1_b, e.l2_w, e.l2_b = ptr(L.linear1.weight), ptr(L.linear1.bias), ptr(L.linear2.weight), ptr(L.linear2.bias)
            e.n1_w, e.n1_b, e.n2_w, e.n2_b = ptr(L.norm1.weight), ptr(L.norm1.bias), ptr(L.norm2.weight), ptr(L.norm2.bias)
        for i, L in enumerate(dec.layers):
            d = w.dec[i]
            d.sa_in_w, d.sa_in_b = ptr(L.self_attn.in_proj_weight), ptr(L.self_attn.in_proj_bias)
            d.sa_out_w, d.sa_out_b = ptr(L.self_attn.out_proj.weight), ptr(L.self_attn.out_proj.bias)
            d.ca_in_w, d.ca_in_b = ptr(L.multihead_attn.in_proj_weight), ptr(L.multihead_attn.in_proj_bias)
            d.ca_out_w, d.ca_out_b = ptr(L.multihead_attn.out_proj.weight), ptr(L.multihead_attn.out_proj.bias)
            d.l1_w, d.l1_b, d.l2_w, d.l2_b = ptr(L.linear1.weight), ptr(L.linear1.bias), ptr(L.linear2.weight), ptr(L.linear2.bias)
            d.n1_w, d.n1_b, d.n2_w, d.n2_b = ptr(L.norm1.weight), ptr(L.norm1.bias), ptr(L.norm2.weight), ptr(L.norm2.bias)
            d.n3_w, d.n3_b = ptr(L.norm3.weight), ptr(L.norm3.bias)
        return w

    def _get_plan(self):
        if self._plan is None:
            if max(self.num_encoder_layers, self.num_decoder_layers) > _cabi.V4H_ENERGY_MAX_LAYERS:
                raise NotImplementedError(f"more than {_cabi.V4H_ENERGY_MAX_LAYERS} encoder / decoder layers")
            dims = _cabi.EnergyDims(self.dims_in, self.dims_c, self.dim_embedding, self.encode_t_dim, self.nhead,
                                    self.num_encoder_layers, self.num_decoder_layers, self.dim_feedforward,
                                    _cabi.V4H_BF16 if self.precision == "bf16" else _cabi.V4H_FP32)
            handle = ctypes.c_void_p()
            _cabi.check(_cabi.load().v4h_energy_plan_create(ctypes.byref(dims), ctypes.byref(handle)))
            self._plan = handle
        return self._plan

    def invalidate_weights(self) -> None:
        """force the bf16 operand copies and the cached condition encoding to be rebuilt (see ViT.invalidate_weights)"""
        self._arena_key = None
        self._encoded.clear()

    @contextlib.contextmanager
    def condition_scope(self):
        """Inside this scope (CFM wraps every ODE solve in it) the condition side -- c_embed, the encoder and the
        cross-attention K / V of every decoder layer -- is computed once per condition tensor and reused by the
        evaluations that follow; outside it every forward encodes its condition (a tensor's address and version are
        not a safe cache key across allocations)."""
        self._encoded.clear()
        self._in_scope = True
        try:
            yield self
        finally:
            self._in_scope = False
            self._encoded.clear()

    def _prepare(self, dev, w, stream):
        lib = _cabi.load()
        plan = self._get_plan()
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self.precision == "bf16":
            if self._arena is None or self._arena.device != dev:
                self._arena = torch.empty(max(16, lib.v4h_energy_weight_arena_bytes(plan)), dtype=torch.uint8, device=dev)
                self._arena_key = None
            if key != self._arena_key:
                _cabi.check(lib.v4h_energy_prepare_weights(plan, ctypes.byref(w), self._arena.data_ptr(), stream))
        if key != self._arena_key:
            self._arena_key = key
            self._encoded.clear()  # the encoded conditions were computed with the old weights
        return plan, (None if self._arena is None else self._arena.data_ptr())

    # ------------------------------------------------------------------ forward
    def forward(self, x, t, condition=None, shared_t: bool = False):
        """x (B, dims_in), t (B, 1) [one value when ``shared_t``], condition (B, dims_c) -> (B, dims_in)
        (reference transformer_cfm.py:99-119)."""
        if condition is None:
            raise NotImplementedError("ParallelTransformer without a condition (decoder-only branch) is not implemented")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("vit4hep_b200.ParallelTransformer is forward-only (sampling): wrap the call in "
                                      "torch.no_grad() / inference_mode(); train the energy network with the reference "
                                      "module (the state dicts are interchangeable)")
        for name, v in (("x", x), ("t", t), ("condition", condition)):
            if not v.is_cuda:
                raise RuntimeError("vit4hep_b200.ParallelTransformer runs on a B200 GPU only (no CPU fallback)")
            if v.dtype != torch.float32:
                raise TypeError(f"{name} must be float32")
        B = x.shape[0]
        if tuple(x.shape) != (B, self.dims_in) or tuple(condition.shape) != (B, self.dims_c):
            raise ValueError(f"expected x (B, {self.dims_in}) and condition (B, {self.dims_c}), got {tuple(x.shape)} "
                             f"and {tuple(condition.shape)}")
        if t.numel() != (1 if shared_t else B):
            raise ValueError("t must hold one time per sample (or a single value with shared_t)")
        dev = x.device
        _cabi.require_device(dev.index if dev.index is not None else torch.cuda.current_device())
        lib = _cabi.load()
        stream = torch.cuda.current_stream(dev).cuda_stream
        x, t, condition = x.contiguous(), t.contiguous(), condition.contiguous()
        w = self._params_struct()
        with torch.cuda.device(dev):
            plan, arena = self._prepare(dev, w, stream)
            ws = self._ws.get((B, dev))
            if ws is None:
                ws = torch.empty(lib.v4h_energy_workspace_bytes(plan, B), dtype=torch.uint8, device=dev)
                self._ws[(B, dev)] = ws
                self._encoded.pop((B, dev), None)
            # the condition side only depends on the condition: encode once per condition batch (the 80 evaluations of
            # an ODE solve pass the same tensor), again when it changes
            # (inference tensors do not track a version; within a scope the condition of a solve is not rewritten)
            ckey = (condition.data_ptr(), 0 if condition.is_inference() else condition._version)
            if not self._in_scope or self._encoded.get((B, dev)) != ckey:
                _cabi.check(lib.v4h_energy_encode(plan, ctypes.byref(w), arena, condition.data_ptr(), B, ws.data_ptr(),
                                                  ws.numel(), stream))
                self._encoded[(B, dev)] = ckey
            out = torch.empty((B, self.dims_in), dtype=torch.float32, device=dev)
            _cabi.check(lib.v4h_energy_forward(plan, ctypes.byref(w), arena, x.data_ptr(), t.data_ptr(), int(shared_t),
                                               out.data_ptr(), B, ws.data_ptr(), ws.numel(), stream))
        return out
