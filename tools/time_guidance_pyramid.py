"""GPU: time the guidance pyramid producers at the cfg4 batch (B=16, ViT-L width 1024) with CUDA events.

    python tools/time_guidance_pyramid.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cat_seg_b200.guidance import GuidancePyramid  # noqa: E402
from cat_seg_b200.synth import make_pyramid_inputs  # noqa: E402


def main():
    B, width = 16, 1024
    clip, la, lb, w1, b1, w2, b2 = make_pyramid_inputs(width, B, 0, feat_dim=768)
    m = GuidancePyramid(width)
    m.load_state_dict({"upsample1.weight": w1, "upsample1.bias": b1, "upsample2.weight": w2, "upsample2.bias": b2})
    m = m.cuda()
    args = (clip.cuda(), la.cuda(), lb.cuda())
    for _ in range(3):
        m(*args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        m(*args)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    gmac = B * 576 * width * (256 * 4 + 128 * 16) / 1e9
    print(f"guidance pyramid, B={B}, width={width}: {ms:.3f} ms per batch ({gmac:.1f} GMAC fp32 -> {2 * gmac / ms:.1f} TFLOP/s)")


if __name__ == "__main__":
    main()
