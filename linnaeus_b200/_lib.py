"""ctypes binding of ``liblinnaeus_b200.so`` (the C-ABI declared in include/linnaeus_b200.h).

There is NO CPU or PyTorch fallback: if the shared library is missing, or a kernel
returns an error, the call raises.  PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LNX_LIB_PATH") or os.path.join(_PKG, "liblinnaeus_b200.so")  # override: profiling builds (tools/build_variant.sh)

F32, BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_RELU, ACT_GELU_DG, ACT_MUL, ACT_SWISH = 0, 1, 2, 3, 4, 5
LOSS_CE, LOSS_LS, LOSS_TAXONOMY = 0, 1, 2

P = c_void_p
I = c_int
L = c_int64
F = c_float

# name -> argtypes (restype is int unless listed in _RESTYPE)
SIGNATURES = {
    "lnx_version": [],
    "lnx_strerror": [I],
    "lnx_patchify_nchw": [P, P, I, I, I, I, I, I, I, P],
    "lnx_space_to_depth": [P, P, I, I, I, I, I, I, P],
    "lnx_tokens_assemble": [P, L, P, P, P, I, I, I, I, I, P],
    "lnx_tokens_split": [P, P, P, P, I, I, I, I, I, P],
    "lnx_colsum": [P, P, L, I, I, P],
    "lnx_cast_f32_to_bf16": [P, P, L, P],
    "lnx_act_bwd": [P, P, P, L, I, I, P],
    "lnx_rowscale_cast_bf16": [P, P, P, I, I, P],
    "lnx_layerscale_bwd": [P, P, P, P, P, P, P, P, I, I, P],
    "lnx_layernorm_fwd": [P, P, P, P, P, P, P, L, I, F, I, P],
    "lnx_layernorm_bwd": [P, P, P, P, P, P, P, P, P, L, I, I, P],
    "lnx_dwconv7_fwd": [P, P, I, P, P, P, I, I, I, I, I, P],
    "lnx_dwconv7_wgrad": [P, P, P, I, P, I, I, I, I, I, P],
    "lnx_dwconv7_set_impl": [I],
    "lnx_gemm": [I, P, L, I, P, L, I, P, I, I, I, I, P, I, P, P, P, P, P, I, P, I, I, P],
    "lnx_wgrad": [P, L, P, L, P, P, L, I, I, I, P],
    "lnx_mlp_fused_fwd": [P, P, P, P, P, P, P, I, P, P, L, I, I, P],
    "lnx_mlp_fused_bwd": [P, P, P, P, P, P, P, P, L, I, I, P],
    "lnx_mlp_fused_wgrad": [P, P, P, P, P, P, P, P, L, I, I, P],
    "lnx_im2col3x3": [P, I, P, I, I, I, I, I, I, I, I, I, P],
    "lnx_conv3x3_s1": [P, P, P, P, I, I, I, I, I, I, I, P],
    "lnx_maxpool3s2": [P, P, I, I, I, I, I, P],
    "lnx_dwconv3_fwd": [P, P, P, P, P, P, I, I, I, I, I, I, I, I, I, I, I, P],
    "lnx_se_scale": [P, P, P, I, I, I, I, P],
    "lnx_attn_bias_fwd": [P, P, P, I, I, I, I, F, I, P],
    "lnx_hier_metrics": [P, I, L, I, I, P, P, I, P, P, P],
    "lnx_hier_topk": [P, I, L, I, I, P, I, P, P, P],
    "lnx_hier_consistency": [P, P, P, P, P, P, I, I, I, P],
    "lnx_mix_pairs": [P, P, P, P, I, L, P],
    "lnx_mix_meta_chunks": [P, P, P, P, P, I, P, P, I, I, P],
    "lnx_cutmix_paste": [P, P, P, P, I, I, I, I, I, I, I, I, P],
    "lnx_mix_pairs_valid": [P, P, P, F, F, P, I, L, P],
    "lnx_rowscale": [P, P, P, L, I, I, I, P],
    "lnx_rope_table": [P, P, P, I, I, I, I, P],
    "lnx_rope_qk_fwd": [P, P, P, P, P, I, I, I, I, I, F, I, P],
    "lnx_rope_qk_bwd": [P, P, P, P, P, P, P, P, I, I, I, I, I, F, I, P],
    "lnx_rope_qk_bwd_scaled": [P, P, P, P, P, P, P, P, I, I, I, I, I, F, I, P],
    "lnx_rope_freq_grad": [P, P, I, I, I, I, P],
    "lnx_gemm_pair": [P, P, P, P, L, I, I, P],
    "lnx_qkv_rope_gemm": [P, P, P, P, P, L, I, I, I, I, I, F, P],
    "lnx_attn_qkv_fwd": [P, P, P, I, I, I, I, I, P],
    "lnx_attn_qkv_bwd": [P, P, P, P, P, P, P, I, I, I, I, I, P],
    "lnx_attn_fwd": [P, P, P, P, P, I, I, I, I, I, I, P],
    "lnx_attn_bwd": [P, P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, P],
    "lnx_aggregate2_fwd": [P, P, P, P, P, L, I, I, P],
    "lnx_aggregate2_bwd": [P, P, P, P, P, P, P, P, L, I, I, P],
    "lnx_loss_fwd": [P, I, I, I, P, P, P, P, I, F, P, I, P, P, P, P, P],
    "lnx_loss_reduce": [P, P, I, I, I, P, P, P, P, P],
    "lnx_loss_bwd": [P, I, I, I, P, P, P, I, F, P, P, P, P, P, P],
    "lnx_sumsq": [P, L, P, P, P],
    "lnx_clip_coef": [P, F, F, P, P, P],
    "lnx_adamw": [P, P, P, P, L, F, F, F, F, F, F, F, F, P, P, P, P],
}
_RESTYPE = {"lnx_strerror": c_char_p}

_lib = None
launch_count = 0  # number of kernel-launching C-ABI calls made by this process


class LnxError(RuntimeError):
    pass


def load(build_if_missing: bool = False) -> ctypes.CDLL:
    """Load the shared library and bind every declared symbol (raises if one is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if build_if_missing:
            from . import _build

            _build.build()
        else:
            raise LnxError(
                f"{LIB_PATH} not found: build it with `python -m linnaeus_b200._build` "
                "(there is no CPU / PyTorch fallback for the linnaeus_b200 hot path)"
            )
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = _RESTYPE.get(name, c_int)
    _lib = lib
    return lib


def strerror(code: int) -> str:
    return load().lnx_strerror(code).decode()


# LNX_PROFILE=1: every C-ABI call is bracketed by CUDA events on its stream (tools/profile_step.py
# aggregates them per entry point and shape).  Diagnostic only; never set inside a timed region.
PROFILE = bool(int(os.environ.get("LNX_PROFILE", "0")))
profile_log: list = []  # (name, key, start_event, end_event)


def call(name: str, *args) -> None:
    """Invoke a C-ABI entry point on the current CUDA stream; raise on a non-zero code."""
    global launch_count
    lib = _lib if _lib is not None else load()
    if PROFILE:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        code = getattr(lib, name)(*args, torch.cuda.current_stream().cuda_stream)
        e1.record()
        key = tuple(a for a in args if isinstance(a, (int, bool)) and not isinstance(a, bool) and abs(a) < (1 << 32)) if args else ()
        profile_log.append((name, key, e0, e1))
    else:
        code = getattr(lib, name)(*args, torch.cuda.current_stream().cuda_stream)
    launch_count += 1
    if code != 0:
        raise LnxError(f"{name} failed: {strerror(code)} (code {code})")


def dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise LnxError(f"unsupported dtype {t.dtype}; the kernels take float32 or bfloat16")


def ptr(t: torch.Tensor | None):
    return None if t is None else t.data_ptr()


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise LnxError("linnaeus_b200 kernels need CUDA tensors (no CPU fallback exists)")
