"""Synthetic weights and inputs (checkpoints and datasets are unavailable offline).

``param_shapes`` is the checkpoint contract: the exact ``state_dict`` keys and shapes of the
reference ``Aggregator`` (cat_seg/modeling/transformer/model.py:602-636), so a real
``sem_seg_head.predictor.transformer.*`` checkpoint loads into the B200 module unchanged.

Weight recipe (documented because parity numbers depend on it): every tensor is drawn from its own
``torch.Generator`` seeded with ``seed`` and a stable hash of the parameter name, so tensors do
not depend on creation order.
  * Linear / Conv / ConvTranspose weights and biases: U(-b, b), b = 1/sqrt(fan_in) (the PyTorch
    default init bound, so magnitudes match a freshly constructed reference);
  * LayerNorm / GroupNorm: weight = 1 + 0.1 N(0,1), bias = 0.1 N(0,1)  (the default 1/0 would hide
    gamma/beta indexing bugs);
  * padding_tokens / padding_guidance: 0.5 N(0,1)  (the default zeros would hide padding bugs,
    model.py:372-373).
Inputs are N(0,1) from ``torch.Generator().manual_seed(1234 + seed)`` on the CPU;
guidance[0] is the same tensor as img_feats, as in the real model.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict
from typing import Dict, List, Tuple

import torch

from .config import AggregatorConfig


def param_shapes(cfg: AggregatorConfig) -> "OrderedDict[str, Tuple[int, ...]]":
    hid, P = cfg.hidden_dim, cfg.prompt_channel
    ag, tg = cfg.appearance_guidance_proj_dim, cfg.text_guidance_proj_dim
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    for l in range(cfg.num_layers):
        p = f"layers.{l}.swin_block"
        for b in ("block_1", "block_2"):
            q = f"{p}.{b}"
            s[f"{q}.norm1.weight"] = (hid,); s[f"{q}.norm1.bias"] = (hid,)
            s[f"{q}.attn.q.weight"] = (hid, hid + ag); s[f"{q}.attn.q.bias"] = (hid,)
            s[f"{q}.attn.k.weight"] = (hid, hid + ag); s[f"{q}.attn.k.bias"] = (hid,)
            s[f"{q}.attn.v.weight"] = (hid, hid); s[f"{q}.attn.v.bias"] = (hid,)
            s[f"{q}.attn.proj.weight"] = (hid, hid); s[f"{q}.attn.proj.bias"] = (hid,)
            s[f"{q}.norm2.weight"] = (hid,); s[f"{q}.norm2.bias"] = (hid,)
            s[f"{q}.mlp.fc1.weight"] = (4 * hid, hid); s[f"{q}.mlp.fc1.bias"] = (4 * hid,)
            s[f"{q}.mlp.fc2.weight"] = (hid, 4 * hid); s[f"{q}.mlp.fc2.bias"] = (hid,)
        s[f"{p}.guidance_norm.weight"] = (ag,); s[f"{p}.guidance_norm.bias"] = (ag,)
        a = f"layers.{l}.attention"
        if cfg.pad_len > 0:                      # None (absent from the state_dict) in the reference when pad_len == 0
            s[f"{a}.padding_tokens"] = (1, 1, hid)
            s[f"{a}.padding_guidance"] = (1, 1, tg)
        s[f"{a}.attention.q.weight"] = (hid, hid + tg); s[f"{a}.attention.q.bias"] = (hid,)
        s[f"{a}.attention.k.weight"] = (hid, hid + tg); s[f"{a}.attention.k.bias"] = (hid,)
        s[f"{a}.attention.v.weight"] = (hid, hid); s[f"{a}.attention.v.bias"] = (hid,)
        s[f"{a}.MLP.0.weight"] = (4 * hid, hid); s[f"{a}.MLP.0.bias"] = (4 * hid,)
        s[f"{a}.MLP.2.weight"] = (hid, 4 * hid); s[f"{a}.MLP.2.bias"] = (hid,)
        s[f"{a}.norm1.weight"] = (hid,); s[f"{a}.norm1.bias"] = (hid,)
        s[f"{a}.norm2.weight"] = (hid,); s[f"{a}.norm2.bias"] = (hid,)
    s["conv1.weight"] = (hid, P, 7, 7); s["conv1.bias"] = (hid,)
    s["guidance_projection.0.weight"] = (ag, cfg.appearance_guidance_dim, 3, 3)
    s["guidance_projection.0.bias"] = (ag,)
    s["text_guidance_projection.0.weight"] = (tg, cfg.text_guidance_dim)
    s["text_guidance_projection.0.bias"] = (tg,)
    for i, (d, dp) in enumerate(zip(cfg.decoder_guidance_dims, cfg.decoder_guidance_proj_dims)):
        s[f"decoder_guidance_projection.{i}.0.weight"] = (dp, d, 3, 3)
        s[f"decoder_guidance_projection.{i}.0.bias"] = (dp,)
    cin = hid
    for i, (cout, gp) in enumerate(zip(cfg.decoder_dims, cfg.decoder_guidance_proj_dims)):
        d = f"decoder{i + 1}"
        s[f"{d}.up.weight"] = (cin, cin - gp, 2, 2); s[f"{d}.up.bias"] = (cin - gp,)
        s[f"{d}.conv.double_conv.0.weight"] = (cout, cin, 3, 3)
        s[f"{d}.conv.double_conv.1.weight"] = (cout,); s[f"{d}.conv.double_conv.1.bias"] = (cout,)
        s[f"{d}.conv.double_conv.3.weight"] = (cout, cout, 3, 3)
        s[f"{d}.conv.double_conv.4.weight"] = (cout,); s[f"{d}.conv.double_conv.4.bias"] = (cout,)
        cin = cout
    s["head.weight"] = (1, cin, 3, 3); s["head.bias"] = (1,)
    return s


def _fan_in(name: str, shape: Tuple[int, ...]) -> int:
    if ".up.weight" in name:                    # ConvTranspose2d weight [Cin, Cout, kh, kw]: fan_in = Cout*kh*kw
        return shape[1] * shape[2] * shape[3]
    f = 1
    for d in shape[1:]:
        f *= d
    return f


def _is_norm(name: str) -> bool:
    return (".norm1." in name or ".norm2." in name or "guidance_norm" in name
            or ".double_conv.1." in name or ".double_conv.4." in name)


def make_state_dict(cfg: AggregatorConfig, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    shapes = param_shapes(cfg)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape in shapes.items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
        if "padding_" in name:
            t = 0.5 * torch.randn(shape, generator=g)
        elif _is_norm(name):
            t = 0.1 * torch.randn(shape, generator=g)
            if name.endswith(".weight"):
                t = t + 1.0
        else:
            wname = name[:-4] + "weight" if name.endswith(".bias") else name
            bound = 1.0 / math.sqrt(_fan_in(wname, shapes[wname]))
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        sd[name] = t.float().contiguous()
    return sd


def make_inputs(cfg: AggregatorConfig, B: int, T: int, seed: int = 0, same_text: bool = True):
    """Returns (img_feats [B,C,H,W], text_feats [B,T,P,C], guidance list of 3), fp32 CPU tensors."""
    g = torch.Generator().manual_seed(1234 + seed)
    H, W = cfg.feature_resolution
    C, P = cfg.appearance_guidance_dim, cfg.prompt_channel
    img = torch.randn(B, C, H, W, generator=g)
    if same_text:
        text = torch.randn(1, T, P, cfg.text_guidance_dim, generator=g).repeat(B, 1, 1, 1)
    else:
        text = torch.randn(B, T, P, cfg.text_guidance_dim, generator=g)
    g1 = torch.randn(B, cfg.decoder_guidance_dims[0], 2 * H, 2 * W, generator=g)
    g2 = torch.randn(B, cfg.decoder_guidance_dims[1], 4 * H, 4 * W, generator=g)
    return img, text.contiguous(), [img, g1, g2]


def make_pyramid_inputs(width: int, B: int, seed: int = 0, grid: int = 24, feat_dim: int = 0):
    """Seeded synthetic inputs of the guidance pyramid producers (cat_seg_model.py:176-185): clip_features
    [B, 1+grid^2, feat_dim or width], two hooked layer outputs [1+grid^2, B, width], and ConvTranspose2d parameters
    in the module's own layout / default-init scale (upsample1: width->256 k2, upsample2: width->128 k4)."""
    g = torch.Generator().manual_seed(4321 + seed)
    L = 1 + grid * grid
    fd = feat_dim or width
    clip = torch.randn(B, L, fd, generator=g)
    la = torch.randn(L, B, width, generator=g)
    lb = torch.randn(L, B, width, generator=g)
    def uni(shape, bound):
        return (torch.rand(shape, generator=g) * 2 - 1) * bound
    w1 = uni((width, 256, 2, 2), (1.0 / (256 * 4)) ** 0.5)
    b1 = uni((256,), (1.0 / (256 * 4)) ** 0.5)
    w2 = uni((width, 128, 4, 4), (1.0 / (128 * 16)) ** 0.5)
    b2 = uni((128,), (1.0 / (128 * 16)) ** 0.5)
    return clip, la, lb, w1, b1, w2, b2


CLIP_DENSE_KEYS = ("ln_1.weight", "ln_1.bias", "attn.q_proj_weight", "attn.k_proj_weight", "attn.v_proj_weight", "attn.in_proj_bias", "attn.out_proj.weight",
                   "attn.out_proj.bias", "ln_2.weight", "ln_2.bias", "mlp.c_fc.weight", "mlp.c_fc.bias",
                   "mlp.c_proj.weight", "mlp.c_proj.bias")


def make_clip_dense_inputs(width: int, L: int, N: int, out_dim: int, seed: int = 0):
    """Seeded synthetic input and parameters of the CLIP dense last block (model_vpt.py:186-240, 283-284): x [L, N, width]
    (the input of the last resblock, LND) and a dict with the resblock's own state_dict keys (the fork splits in_proj_weight into q/k/v_proj_weight, :169-178) plus ``ln_post.*`` and ``proj``.
    Magnitudes follow CLIP's initialisation (model_vpt.py:399-419): attn std width^-0.5, fc std (2 width)^-0.5, proj std
    width^-0.5; LayerNorm affine is randomised (the defaults 1 / 0 would hide a swapped or missing affine)."""
    g = torch.Generator().manual_seed(9876 + seed)
    def nrm(shape, std):
        return torch.randn(shape, generator=g) * std
    sd = {
        "ln_1.weight": 1.0 + 0.1 * torch.randn(width, generator=g), "ln_1.bias": 0.1 * torch.randn(width, generator=g),
        "attn.q_proj_weight": nrm((width, width), width ** -0.5), "attn.k_proj_weight": nrm((width, width), width ** -0.5),
        "attn.v_proj_weight": nrm((width, width), width ** -0.5), "attn.in_proj_bias": nrm((3 * width,), 0.02),
        "attn.out_proj.weight": nrm((width, width), width ** -0.5), "attn.out_proj.bias": nrm((width,), 0.02),
        "ln_2.weight": 1.0 + 0.1 * torch.randn(width, generator=g), "ln_2.bias": 0.1 * torch.randn(width, generator=g),
        "mlp.c_fc.weight": nrm((4 * width, width), (2 * width) ** -0.5), "mlp.c_fc.bias": nrm((4 * width,), 0.02),
        "mlp.c_proj.weight": nrm((width, 4 * width), width ** -0.5), "mlp.c_proj.bias": nrm((width,), 0.02),
        "ln_post.weight": 1.0 + 0.1 * torch.randn(width, generator=g), "ln_post.bias": 0.1 * torch.randn(width, generator=g),
        "proj": nrm((width, out_dim), width ** -0.5),
    }
    x = torch.randn(L, N, width, generator=g)
    return x, sd
