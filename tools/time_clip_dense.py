"""Times catseg_clip_dense_last_block on the GPU (ViT-L/14: width 1024, 577 tokens, out 768; ViT-B/16: 768 / 577 / 512).

    python tools/time_clip_dense.py [N images]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cat_seg_b200.clip_dense import DenseLastBlock  # noqa: E402
from cat_seg_b200.synth import make_clip_dense_inputs  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
for width, od in ((1024, 768), (768, 512)):
    x, sd = make_clip_dense_inputs(width, 577, N, od, 0)
    m = DenseLastBlock(width, od)
    m.load_state_dict(sd)
    m = m.cuda()
    xd = x.cuda()
    for _ in range(3):
        m(xd, want_block_out=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        m(xd, want_block_out=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    macs = 577 * N * (10 * width * width + width * od)
    print(f"width {width} N {N}: {ms:.3f} ms per call, {2 * macs / ms / 1e9:.1f} TFLOP/s algorithmic "
          f"({6 * macs / ms / 1e9:.1f} issued: 3 MMAs per product)")
