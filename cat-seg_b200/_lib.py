"""ctypes binding of libcatseg_b200.so (declared in include/catseg_b200.h).

There is no fallback: if the shared library is missing and cannot be built, importing the product
path raises.  PyTorch is only used by callers for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_LIB = None


class CatsegConfig(C.Structure):
    _fields_ = [
        ("text_guidance_dim", C.c_int32), ("text_guidance_proj_dim", C.c_int32),
        ("appearance_guidance_dim", C.c_int32), ("appearance_guidance_proj_dim", C.c_int32),
        ("decoder_dims", C.c_int32 * 2), ("decoder_guidance_dims", C.c_int32 * 2),
        ("decoder_guidance_proj_dims", C.c_int32 * 2),
        ("num_layers", C.c_int32), ("nheads", C.c_int32), ("hidden_dim", C.c_int32),
        ("pooling_size", C.c_int32 * 2), ("feature_resolution", C.c_int32 * 2),
        ("window_size", C.c_int32), ("attention_type", C.c_int32), ("prompt_channel", C.c_int32),
        ("pad_len", C.c_int32), ("precision", C.c_int32),
    ]


class CatsegTaps(C.Structure):
    _fields_ = [
        ("corr", C.c_void_p), ("classes", C.c_void_p), ("embed", C.c_void_p),
        ("app_guidance", C.c_void_p), ("text_guidance", C.c_void_p),
        ("dec_guidance0", C.c_void_p), ("dec_guidance1", C.c_void_p),
        ("swin_b1", C.c_void_p * 4), ("swin_b2", C.c_void_p * 4), ("class_out", C.c_void_p * 4),
        ("up1", C.c_void_p), ("up2", C.c_void_p),
    ]


class ClipDenseWeights(C.Structure):
    _fields_ = [("width", C.c_int32), ("out_dim", C.c_int32)] + [(n, C.c_void_p) for n in (
        "ln_1_weight", "ln_1_bias", "v_proj_weight", "v_proj_bias", "out_proj_weight", "out_proj_bias",
        "ln_2_weight", "ln_2_bias", "c_fc_weight", "c_fc_bias", "c_proj_weight", "c_proj_bias",
        "ln_post_weight", "ln_post_bias", "proj")]


STAGES = ("prep", "embed", "swin", "class", "decoder", "swin_mlp", "exchange")
FAST_BITS = {"swin_mlp": 1, "swin_attn": 2, "class": 4, "decoder": 8, "prep": 16}


PRECISE_SPLIT = 0x100


def precision_mask(spec: str) -> int:
    """'exact' -> 0; 'fast' / 'precise' -> every stage with a tcgen05 kernel (single fp16 operands / hi+lo fp16
    pairs on the value path); 'fast:swin_mlp,decoder' / 'precise:class' -> those stages, the others EXACT."""
    if spec == "exact":
        return 0
    for mode, extra in (("fast", 0), ("precise", PRECISE_SPLIT)):
        if spec == mode:
            return 0x1F | extra
        if spec.startswith(mode + ":"):
            m = extra
            for part in spec[len(mode) + 1:].split(","):
                m |= FAST_BITS[part.strip()]
            return m
    raise ValueError(f"bad precision {spec!r}: use 'exact', 'fast', 'precise' or '<mode>:<stage>[,<stage>]' with stages "
                     f"{sorted(FAST_BITS)}")

# every symbol include/catseg_b200.h declares: (name, restype, argtypes)
_SIGS = [
    ("catseg_version", C.c_char_p, []),
    ("catseg_create", C.c_int, [C.POINTER(CatsegConfig), C.POINTER(C.c_void_p)]),
    ("catseg_destroy", C.c_int, [C.c_void_p]),
    ("catseg_last_error", C.c_char_p, [C.c_void_p]),
    ("catseg_num_params", C.c_int, [C.c_void_p]),
    ("catseg_param_name", C.c_char_p, [C.c_void_p, C.c_int]),
    ("catseg_param_numel", C.c_int64, [C.c_void_p, C.c_int]),
    ("catseg_set_param", C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_int]),
    ("catseg_finalize_params", C.c_int, [C.c_void_p, C.c_void_p]),
    ("catseg_set_vocabulary", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    ("catseg_kept_classes", C.c_int, [C.c_void_p, C.c_int]),
    ("catseg_workspace_bytes", C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    ("catseg_forward", C.c_int, [C.c_void_p] + [C.c_void_p] * 7 + [C.c_size_t, C.c_int, C.c_int, C.c_void_p]),
    ("catseg_forward_taps", C.c_int, [C.c_void_p] + [C.c_void_p] * 7 +
     [C.c_size_t, C.c_int, C.c_int, C.POINTER(CatsegTaps), C.c_void_p]),
    ("catseg_forward_class_sharded", C.c_int, [C.c_void_p] + [C.c_void_p] * 8 +
     [C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("catseg_exchange_buffer_bytes", C.c_size_t, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    ("catseg_forward_class_sharded_a2a", C.c_int, [C.c_void_p] + [C.c_void_p] * 8 +
     [C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_size_t,
      C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("catseg_exchange_guidance_bytes", C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    ("catseg_exchange_timed_out", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    ("catseg_exchange_logits_bytes", C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    ("catseg_assemble_class_sharded", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                                C.c_int64, C.c_void_p]),
    ("catseg_peer_alloc", C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    ("catseg_peer_free", C.c_int, [C.c_void_p]),
    ("catseg_peer_export", C.c_int, [C.c_void_p, C.c_char_p]),
    ("catseg_peer_open", C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    ("catseg_peer_close", C.c_int, [C.c_void_p]),
    ("catseg_set_profiling", C.c_int, [C.c_void_p, C.c_int]),
    ("catseg_stage_times", C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_int]),
    ("catseg_last_launch_count", C.c_int, [C.c_void_p]),
    ("catseg_stitch_argmax", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p]),
    ("catseg_stitch_scratch_bytes", C.c_size_t, [C.c_int]),
    ("catseg_stitch_argmax_ws", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    ("catseg_argmax", C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    ("catseg_argmax_batched", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    ("catseg_guidance_upsample", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.c_void_p]),
    ("catseg_strip_cls_nchw", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    ("catseg_clip_dense_workspace_bytes", C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    ("catseg_clip_dense_last_block", C.c_int, [C.POINTER(ClipDenseWeights), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
]
# int (*catseg_allreduce_fn)(void* ctx, float* buf, size_t count, catseg_stream stream)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)
# int (*catseg_barrier_fn)(void* ctx, catseg_stream stream)
BARRIER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p)
EXPORTED_SYMBOLS = tuple(s[0] for s in _SIGS)


def lib_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True):
    """Loads (building first if needed) the CUDA library.  Raises if it cannot be had."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB
    if not os.path.exists(path) or (build_if_missing and _build.needs_build() and _can_build()):
        if not build_if_missing:
            raise RuntimeError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        _build.build()
    lib = C.CDLL(path)
    for name, res, args in _SIGS:
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype, fn.argtypes = res, args
    _LIB = lib
    return lib


def _can_build() -> bool:
    try:
        _build._nvcc()
        return True
    except RuntimeError:
        return False
