"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the guidance pyramid producers (never imported by the product).

Follows cat_seg/cat_seg_model.py:176-185 with the modules of :80-82:
    res3 = rearrange(clip_features[:, 1:, :], "B (H W) C -> B C H W", H=24)
    res4 = upsample1(rearrange(layers[0][1:, :, :], "(H W) B C -> B C H W", H=24))   # ConvTranspose2d(width, 256, 2, 2)
    res5 = upsample2(rearrange(layers[1][1:, :, :], "(H W) B C -> B C H W", H=24))   # ConvTranspose2d(width, 128, 4, 4)
The transposed convolution is written out as the contraction it is when stride == kernel (no overlapping taps):
    out[b, co, k y + a, k x + c] = bias[co] + sum_ci in[b, ci, y, x] W[ci, co, a, c]
Pinned against the reference's own op, nn.ConvTranspose2d, by tests/golden/guidance_pyramid.npz
(tests/golden/make_guidance_golden.py) and live in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import torch


def tokens_to_nchw(tokens: torch.Tensor, grid: int = 24) -> torch.Tensor:
    """[1 + grid^2, B, C] -> [B, C, grid, grid]  ("(H W) B C -> B C H W" after dropping the CLS row)."""
    L, B, Cc = tokens.shape
    assert L == 1 + grid * grid
    return tokens[1:].reshape(grid, grid, B, Cc).permute(2, 3, 0, 1).contiguous()


def strip_cls_nchw(clip_features: torch.Tensor, grid: int = 24) -> torch.Tensor:
    """[B, 1 + grid^2, C] -> [B, C, grid, grid]  (:179, :182)."""
    B, L, Cc = clip_features.shape
    assert L == 1 + grid * grid
    return clip_features[:, 1:, :].reshape(B, grid, grid, Cc).permute(0, 3, 1, 2).contiguous()


def conv_transpose_stride_eq_kernel(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """x [B, Cin, H, W], weight [Cin, Cout, k, k], bias [Cout] -> [B, Cout, kH, kW] in float64 accumulate."""
    B, Cin, H, W = x.shape
    _, Cout, k, _ = weight.shape
    y = torch.einsum("bihw,ioac->bohawc", x.double(), weight.double())          # [B, Cout, H, k, W, k]
    y = y.reshape(B, Cout, H * k, W * k) + bias.double().view(1, Cout, 1, 1)
    return y.to(torch.float32)


def guidance_pyramid(clip_features, layer_a, layer_b, w1, b1, w2, b2, grid: int = 24):
    return {
        "res5": conv_transpose_stride_eq_kernel(tokens_to_nchw(layer_b, grid), w2, b2),
        "res4": conv_transpose_stride_eq_kernel(tokens_to_nchw(layer_a, grid), w1, b1),
        "res3": strip_cls_nchw(clip_features, grid),
    }
