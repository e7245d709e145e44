"""CLIP dense last block: the producer of ``img_feats`` on the other side of the boundary.

Mirrors ``ResidualAttentionBlock.forward_dense`` (cat_seg/third_party/model_vpt.py:219-240) and the dense tail of
``VisualTransformer.forward`` (:300-312).  ``DenseLastBlock`` keeps the reference's parameter names, so the relevant slice of
a CLIP visual state_dict loads with ``strict=True``:

    ln_1.{weight,bias}  attn.{q,k,v}_proj_weight  attn.in_proj_bias  attn.out_proj.{weight,bias}  ln_2.{weight,bias}
    mlp.c_fc.{weight,bias}  mlp.c_proj.{weight,bias}          <- visual.transformer.resblocks.<last>.*
    ln_post.{weight,bias}  proj                               <- visual.ln_post.*, visual.proj

(``attn.q_proj_weight`` / ``attn.k_proj_weight`` are kept only as checkpoint keys: the reference computes q and k in this
block and discards them, :231-233.)  The arithmetic runs in the CUDA library (``catseg_clip_dense_last_block``: five
fp32-accurate tcgen05 GEMMs with fused bias / QuickGELU / residual epilogues, three LayerNorm kernels that also do the
prompt drop and the LND -> NLD transpose).  CUDA tensors only: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Tuple

import torch
from torch import nn

from . import _lib


class _Attn(nn.Module):
    def __init__(self, width: int):
        super().__init__()
        self.q_proj_weight = nn.Parameter(torch.empty(width, width))
        self.k_proj_weight = nn.Parameter(torch.empty(width, width))
        self.v_proj_weight = nn.Parameter(torch.empty(width, width))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * width))
        self.out_proj = nn.Linear(width, width)
        for p in (self.q_proj_weight, self.k_proj_weight, self.v_proj_weight):
            nn.init.normal_(p, std=width ** -0.5)


class DenseLastBlock(nn.Module):
    """forward(x [L, N, width]) -> (block_out [L', N, width], clip_features [N, L', out_dim]) with L' = L - prompt_length."""

    def __init__(self, width: int, output_dim: int, prompt_length: int = 0):
        super().__init__()
        self.width, self.output_dim, self.prompt_length = width, output_dim, prompt_length
        self.attn = _Attn(width)
        self.ln_1 = nn.LayerNorm(width)
        self.mlp = nn.Sequential()
        self.mlp.add_module("c_fc", nn.Linear(width, 4 * width))
        self.mlp.add_module("c_proj", nn.Linear(4 * width, width))
        self.ln_2 = nn.LayerNorm(width)
        self.ln_post = nn.LayerNorm(width)
        self.proj = nn.Parameter(width ** -0.5 * torch.randn(width, output_dim))
        self._ws = None

    @staticmethod
    def from_clip_visual(visual_sd: Dict[str, torch.Tensor], prompt_length: int = 0) -> "DenseLastBlock":
        """Builds the module from a CLIP ``visual.*`` state_dict (keys without the ``visual.`` prefix): the last resblock,
        ``ln_post`` and ``proj``.  A stock ``attn.in_proj_weight`` is split as the fork does at load time (model_vpt.py:523)."""
        last = max(int(k.split(".")[2]) for k in visual_sd if k.startswith("transformer.resblocks."))
        pre = f"transformer.resblocks.{last}."
        sd = {k[len(pre):]: v for k, v in visual_sd.items() if k.startswith(pre)}
        if "attn.in_proj_weight" in sd:
            q, k, v = sd.pop("attn.in_proj_weight").chunk(3, dim=0)
            sd.update({"attn.q_proj_weight": q, "attn.k_proj_weight": k, "attn.v_proj_weight": v})
        sd.update({"ln_post.weight": visual_sd["ln_post.weight"], "ln_post.bias": visual_sd["ln_post.bias"], "proj": visual_sd["proj"]})
        width, out_dim = sd["proj"].shape
        m = DenseLastBlock(width, out_dim, prompt_length)
        m.load_state_dict(sd, strict=True)
        return m

    @torch.no_grad()
    def forward(self, x: torch.Tensor, want_block_out: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
        if not x.is_cuda:
            raise RuntimeError("cat_seg_b200.clip_dense runs on CUDA tensors only (no CPU fallback)")
        L, N, D = x.shape
        if D != self.width:
            raise ValueError(f"expected width {self.width}, got {D}")
        dev = x.device
        x = x.detach().to(torch.float32).contiguous()
        prm = {n: p.detach().to(device=dev, dtype=torch.float32).contiguous() for n, p in self.named_parameters()}
        vb = prm["attn.in_proj_bias"][2 * D:].contiguous()
        w = _lib.ClipDenseWeights()
        w.width, w.out_dim = D, self.output_dim
        for field, t in (("ln_1_weight", prm["ln_1.weight"]), ("ln_1_bias", prm["ln_1.bias"]),
                         ("v_proj_weight", prm["attn.v_proj_weight"]), ("v_proj_bias", vb),
                         ("out_proj_weight", prm["attn.out_proj.weight"]), ("out_proj_bias", prm["attn.out_proj.bias"]),
                         ("ln_2_weight", prm["ln_2.weight"]), ("ln_2_bias", prm["ln_2.bias"]),
                         ("c_fc_weight", prm["mlp.c_fc.weight"]), ("c_fc_bias", prm["mlp.c_fc.bias"]),
                         ("c_proj_weight", prm["mlp.c_proj.weight"]), ("c_proj_bias", prm["mlp.c_proj.bias"]),
                         ("ln_post_weight", prm["ln_post.weight"]), ("ln_post_bias", prm["ln_post.bias"]), ("proj", prm["proj"])):
            setattr(w, field, t.data_ptr())
        Lp = L - self.prompt_length
        lib = _lib.load()
        need = lib.catseg_clip_dense_workspace_bytes(L, N, D, self.prompt_length)
        if need == 0:
            raise ValueError(f"bad shape L={L} N={N} width={D} prompt={self.prompt_length}")
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        block_out = torch.empty(Lp, N, D, dtype=torch.float32, device=dev) if want_block_out else None
        feats = torch.empty(N, Lp, self.output_dim, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = lib.catseg_clip_dense_last_block(C.byref(w), C.c_void_p(x.data_ptr()), L, N, self.prompt_length,
                                                  C.c_void_p(block_out.data_ptr() if want_block_out else None),
                                                  C.c_void_p(feats.data_ptr()), C.c_void_p(self._ws.data_ptr()), need,
                                                  C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"catseg_clip_dense_last_block failed ({rc}): {lib.catseg_last_error(None).decode()}")
        return block_out, feats
