"""One boundary call (for ncu captures): python tools/one_forward.py <precision> <B> <T> [model]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cat_seg_b200.aggregator import Aggregator
from cat_seg_b200.config import vitb, vitl
from cat_seg_b200.synth import make_inputs, make_state_dict
prec, B, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
cfg = vitb() if (len(sys.argv) > 4 and sys.argv[4] == "vitb") else vitl()
sd = make_state_dict(cfg, 0)
img, text, g = make_inputs(cfg, B, T, 0)
m = Aggregator(**cfg.ctor_kwargs(), precision=prec)
m.load_state_dict(sd, strict=False)
m = m.cuda()
a = (img.cuda(), text.cuda(), [x.cuda() for x in g])
for _ in range(int(os.environ.get("NFWD", "1"))):
    y = m(*a)
torch.cuda.synchronize()
print("done", tuple(y.shape))
