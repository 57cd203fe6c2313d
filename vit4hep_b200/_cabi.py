"""ctypes binding of libvit4hep_b200.so (include/vit4hep_b200.h).

This is the only place the Python host touches native code.  There is no CPU fallback: if the
library is missing it is built with nvcc; if that is impossible, or a call fails, a
RuntimeError carrying ``v4h_last_error()`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

V4H_OK = 0
V4H_FP32, V4H_BF16 = 0, 1
V4H_MAX_DEPTH = 32
V4H_MAX_SEGMENTS = 16

_f32p = C.c_void_p  # device pointers travel as integers


class VitDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "hidden_dim", "depth", "num_heads", "mlp_hidden", "patch_dim", "out_dim", "cond_dim",
        "tokens", "freq_dim", "learn_pos_embed", "precision", "x_map_dim", "c_map_dim")]


class BlockParams(C.Structure):
    _fields_ = [(n, _f32p) for n in (
        "qkv_w", "qkv_b", "proj_w", "proj_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b", "ada_w", "ada_b")]


class VitParams(C.Structure):
    _fields_ = [(n, _f32p) for n in (
        "pos_embed_freqs", "pos_z", "pos_y", "pos_x", "pos_embed", "x_w", "x_b",
        "c0_w", "c0_b", "c2_w", "c2_b", "t0_w", "t0_b", "t2_w", "t2_b",
        "final_w", "final_b", "final_ada_w", "final_ada_b", "xm_w", "xm_b", "cm_w", "cm_b")] + [("blocks", BlockParams * V4H_MAX_DEPTH)]


V4H_ENERGY_MAX_LAYERS = 16


class EnergyDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "dims_in", "dims_c", "dim_embedding", "encode_t_dim", "nhead", "n_enc", "n_dec", "dim_feedforward", "precision")]


class EnergyEncLayer(C.Structure):
    _fields_ = [(n, _f32p) for n in (
        "in_w", "in_b", "out_w", "out_b", "l1_w", "l1_b", "l2_w", "l2_b", "n1_w", "n1_b", "n2_w", "n2_b")]


class EnergyDecLayer(C.Structure):
    _fields_ = [(n, _f32p) for n in (
        "sa_in_w", "sa_in_b", "sa_out_w", "sa_out_b", "ca_in_w", "ca_in_b", "ca_out_w", "ca_out_b",
        "l1_w", "l1_b", "l2_w", "l2_b", "n1_w", "n1_b", "n2_w", "n2_b", "n3_w", "n3_b")]


class EnergyParams(C.Structure):
    _fields_ = [(n, _f32p) for n in (
        "gfp_w", "time_w", "time_b", "x_embed_w", "x_embed_b", "c_embed_w", "c_embed_b", "pos_x", "pos_c",
        "enc_norm_w", "enc_norm_b", "dec_norm_w", "dec_norm_b", "head0_w", "head0_b", "head2_w", "head2_b")] + [
        ("enc", EnergyEncLayer * V4H_ENERGY_MAX_LAYERS), ("dec", EnergyDecLayer * V4H_ENERGY_MAX_LAYERS)]


class AdamWJob(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("bf16_dst", C.c_void_p),
                ("f32_dst", C.c_void_p), ("ema", C.c_void_p), ("n", C.c_int64)]


class ProfileEntry(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("launches", C.c_int64), ("ms", C.c_double), ("flops", C.c_double),
                ("bytes", C.c_double)]


# name -> (restype, argtypes); every symbol include/vit4hep_b200.h declares
_i32, _i64, _sz, _vp, _fl = C.c_int32, C.c_int64, C.c_size_t, C.c_void_p, C.c_float
SIGNATURES = {
    "v4h_last_error": (C.c_char_p, []),
    "v4h_version": (C.c_int, []),
    "v4h_check_device": (C.c_int, [C.c_int]),
    "v4h_geometry_create": (C.c_int, [C.POINTER(_i32), C.POINTER(_i32), _i32, _i32, _i32, C.POINTER(_vp)]),
    "v4h_geometry_destroy": (None, [_vp]),
    "v4h_geometry_tokens": (_i32, [_vp]),
    "v4h_geometry_patch_dim": (_i32, [_vp]),
    "v4h_geometry_voxels": (_i32, [_vp]),
    "v4h_geometry_table_host": (C.c_int, [_vp, C.POINTER(_i32), _i32]),
    "v4h_to_patches": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "v4h_from_patches": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "v4h_plan_create": (C.c_int, [C.POINTER(VitDims), C.POINTER(_vp)]),
    "v4h_plan_destroy": (None, [_vp]),
    "v4h_vit_workspace_bytes": (_sz, [_vp, _i64, _i32]),
    "v4h_vit_weight_arena_bytes": (_sz, [_vp]),
    "v4h_vit_prepare_weights": (C.c_int, [_vp, C.POINTER(VitParams), _vp, _vp]),
    "v4h_vit_forward": (C.c_int, [_vp, C.POINTER(VitParams), _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32,
                                  _vp, _sz, _vp]),
    "v4h_vit_backward": (C.c_int, [_vp, C.POINTER(VitParams), _vp, C.POINTER(VitParams), _vp, _vp, _vp,
                                   _i64, _i32, _i32, _vp, _sz, _vp]),
    "v4h_cfm_prepare": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "v4h_cfm_loss": (C.c_int, [_vp, _vp, _i64, _fl, _vp, _vp, _vp]),
    "v4h_axpy4": (C.c_int, [_vp, _vp, _vp, _fl, _vp, _fl, _vp, _fl, _vp, _fl, _i64, _vp]),
    "v4h_energy_plan_create": (C.c_int, [C.POINTER(EnergyDims), C.POINTER(_vp)]),
    "v4h_energy_plan_destroy": (None, [_vp]),
    "v4h_energy_workspace_bytes": (_sz, [_vp, _i64]),
    "v4h_energy_weight_arena_bytes": (_sz, [_vp]),
    "v4h_energy_prepare_weights": (C.c_int, [_vp, C.POINTER(EnergyParams), _vp, _vp]),
    "v4h_energy_encode": (C.c_int, [_vp, C.POINTER(EnergyParams), _vp, _vp, _i64, _vp, _sz, _vp]),
    "v4h_energy_forward": (C.c_int, [_vp, C.POINTER(EnergyParams), _vp, _vp, _vp, _i32, _vp, _i64, _vp, _sz, _vp]),
    "v4h_postprocess_showers": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _i32, _fl, _fl, _fl, _fl, _fl, _fl, _fl, _fl, _fl,
                                          _fl, _vp, _vp, _vp]),
    "v4h_preprocess_showers": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _i32, _fl, _fl, _fl, _fl, _fl, _fl, _vp, _i32, _vp,
                                         _vp, _vp, _vp]),
    "v4h_grad_norm_sq": (C.c_int, [_vp, _i64, _vp, _vp]),
    "v4h_adamw_step": (C.c_int, [_vp, _i32, _i64, _vp, _fl, _fl, _fl, _fl, _fl, _fl, _i32, _vp, _vp, _fl, _i32, _vp, _vp]),
    "v4h_ema_update": (C.c_int, [_vp, _i32, _i64, _fl, _i32, _vp, _vp]),
    "v4h_counter_increment": (C.c_int, [_vp, _vp]),
    "v4h_vit_arena_offset": (C.c_int64, [_vp, C.c_char_p]),
    "v4h_launch_count": (C.c_int64, []),
    "v4h_profile_begin": (C.c_int, []),
    "v4h_profile_end": (C.c_int, [C.POINTER(ProfileEntry), _i32, C.POINTER(_i32)]),
    "v4h_test_gemm": (C.c_int, [_i32, _i32, _vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "v4h_debug_gemm": (C.c_int, [_i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "v4h_debug_gemm_ln": (C.c_int, [_i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "v4h_debug_tma_probe": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "v4h_debug_attention_counters": (C.c_int, [_vp]),
    "v4h_test_attention_fwd": (C.c_int, [_i32, _i32, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "v4h_test_attention_bwd": (C.c_int, [_i32, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
}

_lock = threading.Lock()
_lib = None


def lib_path() -> str:
    from . import build as _build
    return _build.LIB_PATH


def load(build_if_missing: bool = True):
    """Return the loaded library (ctypes.CDLL) with typed entry points."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = lib_path()
        if not os.path.isfile(path):
            if not build_if_missing:
                raise RuntimeError(f"{path} is missing: run `python -m vit4hep_b200.build` (no CPU fallback)")
            from . import build as _build
            _build.build()
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header and library out of sync
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return load().v4h_last_error().decode(errors="replace")


def check(rc: int) -> None:
    if rc != V4H_OK:
        raise RuntimeError(f"vit4hep_b200 native call failed (code {rc}): {last_error()}")


_device_ok = set()


def require_device(index: int) -> None:
    """Raise unless CUDA device `index` is a B200-class (sm_100) GPU."""
    if index in _device_ok:
        return
    check(load().v4h_check_device(int(index)))
    _device_ok.add(index)


def profile_begin() -> None:
    check(load().v4h_profile_begin())


def profile_end(max_entries: int = 64):
    """[{name, launches, ms, flops, bytes}] per kernel class since profile_begin()."""
    buf = (ProfileEntry * max_entries)()
    n = C.c_int32(0)
    check(load().v4h_profile_end(buf, max_entries, C.byref(n)))
    return [dict(name=buf[i].name.decode(), launches=buf[i].launches, ms=buf[i].ms, flops=buf[i].flops,
                 bytes=buf[i].bytes) for i in range(n.value)]
