"""Guidance pyramid producers: the step immediately before the boundary operator.

Mirrors the part of ``CATSeg.forward`` that turns CLIP outputs into the ``Aggregator`` arguments
(cat_seg/cat_seg_model.py:176-185, modules built at :80-82; CLS strip also at cat_seg_head.py:2009):

    res3 = rearrange(clip_features[:, 1:, :], "B (H W) C -> B C H W")          -> img_feats and guidance[0]
    res4 = upsample1(rearrange(layers[0][1:], "(H W) B C -> B C H W"))          -> guidance[1]   [B,256,48,48]
    res5 = upsample2(rearrange(layers[1][1:], "(H W) B C -> B C H W"))          -> guidance[2]   [B,128,96,96]

``GuidancePyramid`` keeps the reference's parameter names (``upsample1.weight`` ... ``upsample2.bias``, the keys of
the CATSeg checkpoint) by holding two ``nn.ConvTranspose2d`` modules as parameter containers; the arithmetic runs in
the CUDA library (``catseg_guidance_upsample``: a token-row GEMM with a pixel-shuffle store that reads the hooked
``[1 + HW, B, width]`` layer output in place, ``catseg_strip_cls_nchw``: CLS strip + transpose).  CUDA tensors only.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch
from torch import nn

from . import _lib


def _stream(dev) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def strip_cls_nchw(clip_features: torch.Tensor, grid: int = 24) -> torch.Tensor:
    """clip_features [B, 1 + grid^2, C] -> [B, C, grid, grid] (cat_seg_model.py:179,182)."""
    if not clip_features.is_cuda:
        raise RuntimeError("cat_seg_b200.guidance runs on CUDA tensors only (no CPU fallback)")
    B, L, Cc = clip_features.shape
    if L != 1 + grid * grid:
        raise ValueError(f"expected {1 + grid * grid} tokens (CLS + {grid}x{grid}), got {L}")
    x = clip_features.detach().to(torch.float32).contiguous()
    out = torch.empty(B, Cc, grid, grid, dtype=torch.float32, device=x.device)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        rc = lib.catseg_strip_cls_nchw(C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), B, Cc, grid, _stream(x.device))
    if rc != 0:
        raise RuntimeError(f"catseg_strip_cls_nchw failed ({rc}): {lib.catseg_last_error(None).decode()}")
    return out


def upsample_tokens(tokens: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, grid: int = 24) -> torch.Tensor:
    """ConvTranspose2d(kernel == stride) of a hooked CLIP layer output [1 + grid^2, B, width] -> [B, cout, grid*k, grid*k]."""
    if not tokens.is_cuda:
        raise RuntimeError("cat_seg_b200.guidance runs on CUDA tensors only (no CPU fallback)")
    L, B, width = tokens.shape
    cin, cout, k, k2 = weight.shape
    if L != 1 + grid * grid or cin != width or k != k2 or bias.shape != (cout,):
        raise ValueError(f"shape mismatch: tokens {tuple(tokens.shape)}, weight {tuple(weight.shape)}, bias {tuple(bias.shape)}")
    x = tokens.detach().to(torch.float32).contiguous()
    w = weight.detach().to(device=x.device, dtype=torch.float32).contiguous()
    b = bias.detach().to(device=x.device, dtype=torch.float32).contiguous()
    out = torch.empty(B, cout, grid * k, grid * k, dtype=torch.float32, device=x.device)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        rc = lib.catseg_guidance_upsample(C.c_void_p(x.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()),
                                          C.c_void_p(out.data_ptr()), B, width, cout, k, grid, _stream(x.device))
    if rc != 0:
        raise RuntimeError(f"catseg_guidance_upsample failed ({rc}): {lib.catseg_last_error(None).decode()}")
    return out


class GuidancePyramid(nn.Module):
    """``upsample1`` / ``upsample2`` of CATSeg (cat_seg_model.py:80-82) plus the rearranges of :179-185."""

    def __init__(self, proj_dim: int = 1024, grid: int = 24):
        super().__init__()
        self.grid = grid
        self.upsample1 = nn.ConvTranspose2d(proj_dim, 256, kernel_size=2, stride=2)
        self.upsample2 = nn.ConvTranspose2d(proj_dim, 128, kernel_size=4, stride=4)

    @torch.no_grad()
    def forward(self, clip_features: torch.Tensor, layer_a: torch.Tensor, layer_b: torch.Tensor) -> Dict[str, torch.Tensor]:
        """clip_features [B, 1+HW, C]; layer_a / layer_b: hooked resblock outputs [1+HW, B, width] (:84-87).
        Returns the reference's ``features`` dict; the Aggregator takes ``img_feats = res3`` and
        ``appearance_guidance = [res3, res4, res5]`` (cat_seg_predictor.py:151-161 reverses the dict order)."""
        return {
            "res5": upsample_tokens(layer_b, self.upsample2.weight, self.upsample2.bias, self.grid),
            "res4": upsample_tokens(layer_a, self.upsample1.weight, self.upsample1.bias, self.grid),
            "res3": strip_cls_nchw(clip_features, self.grid),
        }
