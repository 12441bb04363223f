"""ctypes binding of libslam_b200.so (C ABI declared in include/slam_b200.h).

The product path has NO fallback: if the shared library is missing or a symbol cannot be
resolved, loading raises.  (INTEGRATION.md shows the same stub a reference maintainer would add.)
"""
from __future__ import annotations

import ctypes as C
import os
import re
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libslam_b200.so"
HEADER = PKG.parent / "include" / "slam_b200.h"

_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", _vp), ("lda", _i64),
        ("b", _vp), ("ldb", _i64),
        ("k1", _i32), ("k2", _i32),
        ("a2", _vp), ("lda2", _i64),
        ("b2", _vp), ("ldb2", _i64),
        ("out", _vp), ("ldo", _i64),
        ("out_f32", _i32), ("act", _i32),
        ("bias", _vp),
        ("residual", _vp), ("ldr", _i64),
        ("alpha", _f32),
        ("m", _i32), ("n", _i32),
        ("block_n", _i32),
        ("split_k", _i32),
        ("workspace", _vp), ("workspace_bytes", _i64),
        ("tail_split", _i32), ("transpose_out", _i32),
        ("aux", _vp), ("ld_aux", _i64),
        ("static_operands", _i32), ("reserved0", _i32),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("q", _vp), ("ldq", _i64),
        ("k", _vp), ("ldk", _i64),
        ("v", _vp), ("ldv", _i64),
        ("out", _vp), ("ldo", _i64),
        ("lse", _vp),
        ("key_mask", _vp),
        ("batch", _i32), ("sq", _i32), ("sk", _i32), ("hq", _i32), ("hkv", _i32), ("dh", _i32),
        ("causal", _i32),
        ("scale", _f32),
        ("dout", _vp), ("lddo", _i64),
        ("dq", _vp), ("lddq", _i64),
        ("dk", _vp), ("lddk", _i64),
        ("dv", _vp), ("lddv", _i64),
        ("delta", _vp),
        ("dq_accum", _vp),
        ("dkv_part", _vp),
        ("rope_cos", _vp),
        ("rope_sin", _vp),
    ]


# name -> argtypes (return type is int unless listed in _RESTYPES)
_SIGS = {
    "slam_abi_version": [],
    "slam_last_error": [],
    "slam_launch_count": [],
    "slam_gemm_bf16": [C.POINTER(GemmArgs), _vp],
    "slam_gemm_workspace_bytes": [],
    "slam_wgrad_thin": [_vp, _i64, _i32, _vp, _i64, _i32, _i32, _f32, _vp, _i64, _i32, _vp],
    "slam_logmel": [_vp, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _vp],
    "slam_conv_im2col": [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _i64, _vp],
    "slam_add_pos": [_vp, _vp, _i32, _i32, _i32, _vp],
    "slam_layernorm": [_vp, _vp, _vp, _vp, _i32, _i32, _f32, _vp],
    "slam_attn_fwd": [C.POINTER(AttnArgs), _vp],
    "slam_attn_bwd": [C.POINTER(AttnArgs), _vp],
    "slam_embed_merge": [_vp, _vp, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _vp],
    "slam_embed_merge_bwd": [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp],
    "slam_rmsnorm_fwd": [_vp, _vp, _vp, _vp, _i32, _i32, _f32, _vp],
    "slam_rmsnorm_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp],
    "slam_rope": [_vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _vp],
    "slam_swiglu_fwd": [_vp, _vp, _i32, _i32, _i32, _vp],
    "slam_swiglu_bwd": [_vp, _vp, _vp, _i32, _i32, _i32, _vp],
    "slam_dropout": [_vp, _vp, _i64, _f32, C.c_uint64, _vp],
    "slam_dropout_bwd_add": [_vp, _vp, _vp, _i64, _f32, C.c_uint64, _vp],
    "slam_cross_entropy": [_vp, _i64, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _i64, _vp, _vp],
    "slam_adamw": [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _f32, _i32, _f32, _vp],
    "slam_cast_f32_to_bf16": [_vp, _vp, _i64, _f32, _vp],
    "slam_cast_bf16_to_f32": [_vp, _vp, _i64, _vp],
    "slam_transpose_bf16": [_vp, _i64, _vp, _i64, _i32, _i32, _vp],
    "slam_transpose_f32_batched": [_vp, _vp, _i32, _i32, _i32, _vp],
    "slam_gather_rows": [_vp, _vp, _vp, _i32, _i32, _vp],
    "slam_scatter_rows": [_vp, _vp, _vp, _i32, _i32, _vp],
    "slam_relu_bwd": [_vp, _vp, _vp, _i64, _vp],
    "slam_colsum": [_vp, _i64, _i32, _i32, _vp, _vp],
    "slam_rmsnorm_wgrad": [_vp, _vp, _vp, _i32, _i32, _vp, _vp],
    "slam_embed_grad": [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp],
    "slam_pack2d": [_vp, _i64, _i64, _vp, _i64, _i64, _i32, _i32, _i32, _f32, _i32, _vp],
    "slam_add_bf16": [_vp, _vp, _vp, _i64, _vp],
}
_RESTYPES = {"slam_last_error": C.c_char_p, "slam_launch_count": _i64, "slam_gemm_workspace_bytes": _i64}

_lib = None


def header_symbols() -> list[str]:
    """Every function name declared in include/slam_b200.h (used by the ABI export test)."""
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(slam_[a-z0-9_]+)\s*\(", text)))


def load() -> C.CDLL:
    """Load the shared library and declare every entry point.  Raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("SLAM_B200_LIB", str(LIB_PATH)))
    if not path.exists():
        raise RuntimeError(
            f"{path} not found: build it with `python -m slam_llm_b200.build` "
            "(there is no CPU or PyTorch fallback for the slam_b200 kernels)")
    lib = C.CDLL(str(path))
    for name, argtypes in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    if lib.slam_abi_version() != 6:
        raise RuntimeError(f"libslam_b200 ABI version {lib.slam_abi_version()} != 6")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().slam_last_error()
        raise RuntimeError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")
