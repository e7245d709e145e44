"""Run-to-run determinism of each precision / stage mask (debug aid)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cat_seg_b200.aggregator import Aggregator
from cat_seg_b200.config import vitb, vitl
from cat_seg_b200.synth import make_inputs, make_state_dict
cfg = vitb()
sd = make_state_dict(cfg, 11)
img, text, g = make_inputs(cfg, 2, 5, 11, same_text=True)
cu = (img.cuda(), text.cuda(), [x.cuda() for x in g])
for prec in sys.argv[1:]:
    m = Aggregator(**cfg.ctor_kwargs(), precision=prec)
    m.load_state_dict(sd, strict=False)
    m = m.cuda()
    ys = [m(*cu).clone() for _ in range(4)]
    d = max((ys[0] - y).abs().max().item() for y in ys[1:])
    m.set_vocabulary(text[0])
    yv = m(cu[0], None, cu[2])
    print(f"{prec:22s} run-to-run max diff {d:.3e}   vocab vs call {(yv - ys[0]).abs().max().item():.3e}")
