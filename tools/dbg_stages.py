"""Stage-wise comparison of a precision mode with the CPU oracle (debug aid; test infrastructure)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cat_seg_b200.aggregator import Aggregator
from cat_seg_b200.config import vitb, vitl
from cat_seg_b200.synth import make_inputs, make_state_dict
from oracle.aggregator_oracle import aggregator_forward

prec = sys.argv[1] if len(sys.argv) > 1 else "precise"
B, T = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (2, 5)
cfg = vitb()
sd = make_state_dict(cfg, 21)
img, text, g = make_inputs(cfg, B, T, 21, same_text=False)
ref, st = aggregator_forward(sd, cfg.oracle_cfg(), img, text, g, return_stages=True)
m = Aggregator(**cfg.ctor_kwargs(), precision=prec)
m.load_state_dict(sd, strict=False)
names = ["embed", "swin_l0_b1", "swin_l0_b2", "class_l0", "swin_l1_b1", "swin_l1_b2", "class_l1"]
y, taps = m.cuda()(img.cuda(), text.cuda(), [x.cuda() for x in g], taps=names)
for n in names:
    d = (taps[n].cpu() - st[n]).abs()
    print(f"{prec} {n:12s} max-abs {d.max().item():.3e}  (ref max {st[n].abs().max().item():.2f})  worst slice {d.flatten(2).max(2)[0].flatten().argmax().item()}")
torch.save({n: taps[n].cpu() for n in names}, "/tmp/taps_%s.pt" % os.environ.get("TAG", "x"))
print(f"{prec} logits max-abs {(y.cpu() - ref).abs().max().item():.3e}")
