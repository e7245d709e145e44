"""CPU oracle for the sliding-window tiling + stitch + argmax.  TEST INFRASTRUCTURE ONLY.

Restates ``CATSeg.forward``'s sliding-window branch (cat_seg/cat_seg_model.py:156-176, 204-218)
with the same torch ops the reference uses (nn.Unfold / nn.Fold / F.interpolate), Detectron2's
``sem_seg_postprocess`` (not vendored in the reference: crop to ``img_size``, bilinear
``align_corners=False`` resize to ``(height, width)``) and the evaluator's ``argmax(dim=0)``
(train_net.py:58).  ``CATSeg`` itself cannot be imported here (needs detectron2 + CLIP weights,
SURVEY.md §8c).  Pinning: tests/golden/make_stitch_golden.py extracts the reference's own statements
(cat_seg_model.py:158-176, 206-218) with ``ast``, executes them with a stub ``self`` and writes
tests/golden/stitch_*.npz; tests/test_stitch_golden.py requires this oracle to reproduce those fixtures BIT
EXACTLY (windows, stitched probabilities, labels), and the CUDA kernel to match them within 2e-6.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F
from einops import rearrange


def make_windows(image: torch.Tensor, kernel: int = 384, overlap: float = 0.333, out_res=(640, 640)) -> torch.Tensor:
    """cat_seg_model.py:162-169.  image [3,H,W] -> [L+1, 3, kernel, kernel]."""
    stride = int(kernel * (1 - overlap))
    unfold = nn.Unfold(kernel_size=kernel, stride=stride)
    img = F.interpolate(image.unsqueeze(0).float(), size=list(out_res), mode="bilinear", align_corners=False).squeeze()
    tiles = rearrange(unfold(img), "(C H W) L-> L C H W", C=3, H=kernel)
    glob = F.interpolate(image.unsqueeze(0).float(), size=(kernel, kernel), mode="bilinear", align_corners=False)
    return torch.cat((tiles, glob), dim=0)


def stitch(win_logits: torch.Tensor, height: int, width: int, kernel: int = 384, overlap: float = 0.333,
           out_res=(640, 640)):
    """cat_seg_model.py:204-218.  win_logits [L+1,T,S,S] -> (sem_seg [T,height,width], labels [height,width])."""
    stride = int(kernel * (1 - overlap))
    unfold = nn.Unfold(kernel_size=kernel, stride=stride)
    fold = nn.Fold(list(out_res), kernel_size=kernel, stride=stride)
    outputs = F.interpolate(win_logits, size=kernel, mode="bilinear", align_corners=False)
    outputs = outputs.sigmoid()
    global_output = outputs[-1:]
    global_output = F.interpolate(global_output, size=list(out_res), mode="bilinear", align_corners=False)
    outputs = outputs[:-1]
    outputs = fold(outputs.flatten(1).T) / fold(unfold(torch.ones([1] + list(out_res))))
    outputs = (outputs + global_output) / 2.0
    # sem_seg_postprocess(outputs[0], out_res, height, width)
    result = outputs[0][:, : out_res[0], : out_res[1]].expand(1, -1, -1, -1)
    result = F.interpolate(result, size=(height, width), mode="bilinear", align_corners=False)[0]
    return result, result.argmax(dim=0)


def count_map(kernel: int = 384, overlap: float = 0.333, out_res=(640, 640)) -> torch.Tensor:
    stride = int(kernel * (1 - overlap))
    unfold = nn.Unfold(kernel_size=kernel, stride=stride)
    fold = nn.Fold(list(out_res), kernel_size=kernel, stride=stride)
    return fold(unfold(torch.ones([1] + list(out_res))))[0]
