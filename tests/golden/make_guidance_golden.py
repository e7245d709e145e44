"""Generate tests/golden/guidance_pyramid.npz from the reference's own op (run in the build container).

CATSeg itself cannot be imported here (detectron2 / CLIP weights), so the five lines of
cat_seg/cat_seg_model.py:179-185 are executed verbatim on nn.ConvTranspose2d modules built as at :81-82 and loaded with
the seeded parameters of cat_seg_b200.synth.make_pyramid_inputs; outputs are stored sub-sampled with float64 checksums.

    python tests/golden/make_guidance_golden.py
"""
import os
import sys

import numpy as np
import torch
from einops import rearrange
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cat_seg_b200.synth import make_pyramid_inputs  # noqa: E402

CASES = {"w64_B2": (64, 2, 0), "w1024_B1": (1024, 1, 1)}   # name: (width, B, seed)
SUB = 97


def fingerprint(t):
    d = t.double()
    return np.array([d.sum().item(), d.abs().sum().item(), (d * d).sum().item()], dtype=np.float64)


def main():
    out = {}
    for name, (width, B, seed) in CASES.items():
        clip, la, lb, w1, b1, w2, b2 = make_pyramid_inputs(width, B, seed)
        upsample1 = nn.ConvTranspose2d(width, 256, kernel_size=2, stride=2)     # cat_seg_model.py:81
        upsample2 = nn.ConvTranspose2d(width, 128, kernel_size=4, stride=4)     # :82
        with torch.no_grad():
            upsample1.weight.copy_(w1); upsample1.bias.copy_(b1); upsample2.weight.copy_(w2); upsample2.bias.copy_(b2)
            image_features = clip[:, 1:, :]                                                   # :179
            res3 = rearrange(image_features, "B (H W) C -> B C H W", H=24)                    # :182
            res4 = rearrange(la[1:, :, :], "(H W) B C -> B C H W", H=24)                      # :183
            res5 = rearrange(lb[1:, :, :], "(H W) B C -> B C H W", H=24)                      # :184
            res4 = upsample1(res4)                                                            # :185
            res5 = upsample2(res5)                                                            # :186
        for k, v in (("res3", res3), ("res4", res4), ("res5", res5)):
            out[f"{name}/{k}/shape"] = np.array(v.shape, dtype=np.int64)
            out[f"{name}/{k}/sub"] = v.contiguous().flatten()[::SUB].numpy().copy()
            out[f"{name}/{k}/fp"] = fingerprint(v)
    path = os.path.join(ROOT, "tests", "golden", "guidance_pyramid.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
