"""Sliding-window tiling and stitch around the boundary operator.

Mirrors the sliding-window branch of ``CATSeg.forward`` (cat_seg/cat_seg_model.py:156-176 tiling,
:204-218 stitch) and the evaluator's ``argmax(dim=0)`` (train_net.py:58).  Tiling is index
plumbing done with torch views on the device; the stitch + postprocess resize + argmax is one fused
CUDA kernel behind ``catseg_stitch_argmax`` (no ``[T,384,384]`` / ``[T,640,640]`` intermediates).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import _lib

KERNEL = 384            # cat_seg_model.py:158
OVERLAP = 0.333         # :159
OUT_RES = 640           # :160


def window_stride(kernel: int = KERNEL, overlap: float = OVERLAP) -> int:
    return int(kernel * (1 - overlap))            # :162  -> 256


def tile_origins(kernel: int = KERNEL, stride: Optional[int] = None, out_res: int = OUT_RES) -> List[Tuple[int, int]]:
    """Top-left corners of the nn.Unfold tiles in L order (row-major): (0,0),(0,256),(256,0),(256,256)."""
    stride = window_stride(kernel) if stride is None else stride
    n = (out_res - kernel) // stride + 1
    return [(ty * stride, tx * stride) for ty in range(n) for tx in range(n)]


def count_map(kernel: int = KERNEL, stride: Optional[int] = None, out_res: int = OUT_RES) -> torch.Tensor:
    """fold(unfold(ones)) (:212): how many tiles cover each pixel — values in {1,2,4}. int32 [out_res,out_res]."""
    stride = window_stride(kernel) if stride is None else stride
    cm = torch.zeros(out_res, out_res, dtype=torch.int32)
    for y, x in tile_origins(kernel, stride, out_res):
        cm[y:y + kernel, x:x + kernel] += 1
    return cm


def make_windows(image: torch.Tensor, kernel: int = KERNEL, overlap: float = OVERLAP, out_res: int = OUT_RES) -> torch.Tensor:
    """image [3,H,W] float -> [ntile+1, 3, kernel, kernel]: the tiles of the out_res-resized image in
    row-major order followed by the global view (:166-169)."""
    stride = window_stride(kernel, overlap)
    img = image.to(torch.float32).unsqueeze(0)
    big = F.interpolate(img, size=(out_res, out_res), mode="bilinear", align_corners=False)[0]
    tiles = [big[:, y:y + kernel, x:x + kernel] for y, x in tile_origins(kernel, stride, out_res)]
    glob = F.interpolate(img, size=(kernel, kernel), mode="bilinear", align_corners=False)[0]
    return torch.stack(tiles + [glob], dim=0)


def stitch(win_logits: torch.Tensor, height: int = OUT_RES, width: int = OUT_RES, kernel: int = KERNEL,
           overlap: float = OVERLAP, out_res: int = OUT_RES, want_probs: bool = False, want_labels: bool = True,
           skip_dropped: bool = True):
    """win_logits [ntile+1, T, S, S] (CUDA fp32) -> (probs [T,height,width] or None, labels int32 [height,width] or None).
    skip_dropped: planes that hold -100 everywhere (classes dropped by the top-256 truncation; their sigmoid is exactly 0)
    are found by a pre-pass and skipped -- the result is bit-identical, only faster."""
    if not win_logits.is_cuda:
        raise RuntimeError("catseg_b200.stitch runs on CUDA tensors only (no CPU fallback)")
    lib = _lib.load()
    stride = window_stride(kernel, overlap)
    ntile = ((out_res - kernel) // stride + 1) ** 2
    nwin, T, S, S2 = win_logits.shape
    if nwin != ntile + 1 or S != S2:
        raise ValueError(f"expected {ntile + 1} windows of square logits, got {tuple(win_logits.shape)}")
    x = win_logits.detach().to(torch.float32).contiguous()
    dev = x.device
    probs = torch.empty(T, height, width, dtype=torch.float32, device=dev) if want_probs else None
    labels = torch.empty(height, width, dtype=torch.int32, device=dev) if want_labels else None
    with torch.cuda.device(dev):
        args = [C.c_void_p(x.data_ptr()), T, S, kernel, stride, out_res, height, width,
                C.c_void_p(probs.data_ptr()) if want_probs else None, C.c_void_p(labels.data_ptr()) if want_labels else None]
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        if skip_dropped:
            scratch = torch.empty(lib.catseg_stitch_scratch_bytes(T), dtype=torch.uint8, device=dev)
            rc = lib.catseg_stitch_argmax_ws(*args, C.c_void_p(scratch.data_ptr()), scratch.numel(), stream)
        else:
            rc = lib.catseg_stitch_argmax(*args, stream)
    if rc != 0:
        raise RuntimeError(f"catseg_stitch_argmax failed ({rc}): {lib.catseg_last_error(None).decode()}")
    return probs, labels


def argmax(scores: torch.Tensor) -> torch.Tensor:
    """scores [T, ...] CUDA fp32 -> int32 labels [...], first maximum wins (train_net.py:58)."""
    if not scores.is_cuda:
        raise RuntimeError("catseg_b200.argmax runs on CUDA tensors only (no CPU fallback)")
    lib = _lib.load()
    x = scores.detach().to(torch.float32).contiguous()
    T = x.shape[0]
    npix = x[0].numel()
    out = torch.empty(x.shape[1:], dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.catseg_argmax(C.c_void_p(x.data_ptr()), T, npix, C.c_void_p(out.data_ptr()),
                               C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
    if rc != 0:
        raise RuntimeError(f"catseg_argmax failed ({rc})")
    return out


def argmax_batched(scores: torch.Tensor) -> torch.Tensor:
    """scores [B, T, ...] CUDA fp32 -> int32 labels [B, ...] in one launch (per image: train_net.py:58)."""
    if not scores.is_cuda:
        raise RuntimeError("catseg_b200.argmax_batched runs on CUDA tensors only (no CPU fallback)")
    lib = _lib.load()
    x = scores.detach().to(torch.float32).contiguous()
    B, T = x.shape[0], x.shape[1]
    npix = x[0, 0].numel()
    out = torch.empty((B,) + tuple(x.shape[2:]), dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.catseg_argmax_batched(C.c_void_p(x.data_ptr()), B, T, npix, C.c_void_p(out.data_ptr()),
                                       C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
    if rc != 0:
        raise RuntimeError(f"catseg_argmax_batched failed ({rc})")
    return out
