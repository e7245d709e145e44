"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the CLIP dense last block (never imported by the product).

Follows cat_seg/third_party/model_vpt.py:
  ResidualAttentionBlock.forward_dense :219-240
      y = ln_1(x); y = F.linear(y, cat(q_proj_weight, k_proj_weight, v_proj_weight), in_proj_bias)   (:220-225)
      y.reshape(L, N, 3, D).permute(2, 1, 0, 3).reshape(3 N, L, D); out_proj           (:227-230)
      q, k, v = y.tensor_split(3, dim=0); v = v.transpose(1, 0) + x[:1]                (:231-235; q, k discarded)
      v = v + mlp(ln_2(v)),  mlp = c_fc -> QuickGELU (x * sigmoid(1.702 x), :165-167) -> c_proj    (:237)
      if prompt is not None: v = cat(v[0:1], v[prompt + 1:])                           (:238-239)
  VisualTransformer.forward, dense branch :300-312
      x = x.permute(1, 0, 2); x = ln_post(x); x = x @ proj
LayerNorm: eps 1e-5, biased variance (:156-162, computed in fp32).  Written per token with the v third of in_proj only,
in float64 accumulation where `double=True` (the error bar of the fp32 reference itself).
Pinned against the reference's own classes by tests/golden/clip_dense_*.npz (tests/golden/make_clip_dense_golden.py) and
live in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import torch


def _ln(x, w, b):
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    return (x - mean) / torch.sqrt(var + 1e-5) * w + b


def dense_last_block(sd, x: torch.Tensor, prompt: int = 0, double: bool = False):
    """x [L, N, D] -> (block_out [L - prompt, N, D], clip_features [N, L - prompt, out_dim])."""
    dt = torch.float64 if double else torch.float32
    p = {k: v.to(dt) for k, v in sd.items()}
    x = x.to(dt)
    D = x.shape[-1]
    rows = torch.cat((x[0:1], x[prompt + 1:]), dim=0)                       # token-wise ops: drop the prompt rows up front
    y = _ln(rows, p["ln_1.weight"], p["ln_1.bias"])
    wv = p["attn.v_proj_weight"] if "attn.v_proj_weight" in p else p["attn.in_proj_weight"][2 * D:]
    v = y @ wv.T + p["attn.in_proj_bias"][2 * D:]
    v = v @ p["attn.out_proj.weight"].T + p["attn.out_proj.bias"]
    v = v + x[:1]                                                           # the CLS row of each image, broadcast over L
    z = _ln(v, p["ln_2.weight"], p["ln_2.bias"])
    h = z @ p["mlp.c_fc.weight"].T + p["mlp.c_fc.bias"]
    h = h * torch.sigmoid(1.702 * h)
    v = v + (h @ p["mlp.c_proj.weight"].T + p["mlp.c_proj.bias"])
    f = _ln(v.permute(1, 0, 2), p["ln_post.weight"], p["ln_post.bias"]) @ p["proj"]
    return v.to(torch.float32), f.to(torch.float32)
