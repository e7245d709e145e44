"""Aggregator configuration: a plain mirror of the reference ctor kwargs.

Mirrors ``Aggregator.__init__`` (cat_seg/modeling/transformer/model.py:559-576) as wired by
``CATSegPredictor`` (cat_seg/modeling/transformer/cat_seg_predictor.py:97-113) from
``configs/vitb_384.yaml:16-31`` / ``configs/vitl_336.yaml:16-31``.  No Detectron2 CfgNode.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict, field
from typing import Tuple


@dataclass
class AggregatorConfig:
    text_guidance_dim: int = 512
    text_guidance_proj_dim: int = 128
    appearance_guidance_dim: int = 512
    appearance_guidance_proj_dim: int = 128
    decoder_dims: Tuple[int, int] = (64, 32)
    decoder_guidance_dims: Tuple[int, int] = (256, 128)
    decoder_guidance_proj_dims: Tuple[int, int] = (32, 16)
    num_layers: int = 2
    nheads: int = 4
    hidden_dim: int = 128
    pooling_size: Tuple[int, int] = (1, 1)      # eval protocol (eval.sh:37); yaml trains with (2, 2)
    feature_resolution: Tuple[int, int] = (24, 24)
    window_size: int = 12
    attention_type: str = "linear"
    prompt_channel: int = 1
    pad_len: int = 256

    def ctor_kwargs(self) -> dict:
        d = asdict(self)
        for k in ("decoder_dims", "decoder_guidance_dims", "decoder_guidance_proj_dims",
                  "pooling_size", "feature_resolution"):
            d[k] = list(d[k])
        return d

    def oracle_cfg(self) -> dict:
        return dict(num_layers=self.num_layers, nheads=self.nheads, hidden_dim=self.hidden_dim,
                    pooling_size=tuple(self.pooling_size), feature_resolution=tuple(self.feature_resolution),
                    window_size=self.window_size, pad_len=self.pad_len)


def vitb(**over) -> AggregatorConfig:
    """CAT-Seg (B): CLIP ViT-B/16, 384x384, C=512 (configs/vitb_384.yaml)."""
    return AggregatorConfig(**over)


def vitl(**over) -> AggregatorConfig:
    """CAT-Seg (L): CLIP ViT-L/14@336, C=768 (configs/vitl_336.yaml)."""
    kw = dict(text_guidance_dim=768, appearance_guidance_dim=768)
    kw.update(over)
    return AggregatorConfig(**kw)


# BASELINE.json configs restated.  (name, cfg factory, B, T)
BENCH_CONFIGS = {
    "cfg1": dict(model="vitb", B=1, T=20),      # VOC-20, the reference's CPU-runnable case
    "cfg2": dict(model="vitb", B=8, T=150),     # A-150, single B200
    "cfg3": dict(model="vitl", B=8, T=459),     # PC-459
    "cfg4": dict(model="vitl", B=16, T=847),    # A-847
    "cfg5": dict(model="vitl", B=5, T=847),     # sliding window: 5 crops of one 640x640 image
}
