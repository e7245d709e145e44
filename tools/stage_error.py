"""Which PRECISE stage contributes how much error?  Each tensor-core stage alone (the others on the fp32 EXACT kernels) against
the CPU oracle, on one image of a workload: max-abs, rel-L2 and raw argmax agreement of the logits.

    python tools/stage_error.py [cfg2|cfg4|...]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cat_seg_b200.aggregator import Aggregator  # noqa: E402
from cat_seg_b200.config import BENCH_CONFIGS, vitb, vitl  # noqa: E402
from cat_seg_b200.synth import make_inputs, make_state_dict  # noqa: E402
from oracle.aggregator_oracle import aggregator_forward  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
w = BENCH_CONFIGS[wl]
cfg = vitb() if w["model"] == "vitb" else vitl()
T = w["T"]
sd = make_state_dict(cfg, 0)
img, text, g = make_inputs(cfg, 1, T, 0)
torch.set_num_threads(os.cpu_count() or 1)
ref = aggregator_forward(sd, cfg.oracle_cfg(), img, text, g)
kept = ref != -100.0
for spec in ("exact", "precise:prep", "precise:swin_attn", "precise:swin_mlp", "precise:class", "precise:decoder", "precise",
             "fast:swin_attn", "fast"):
    m = Aggregator(**cfg.ctor_kwargs(), precision=spec)
    m.load_state_dict(sd, strict=False)
    y = m.cuda()(img.cuda(), text.cuda(), [x.cuda() for x in g]).cpu()
    err = (y[kept] - ref[kept]).abs().max().item()
    rl2 = ((y[kept].double() - ref[kept].double()).norm() / ref[kept].double().norm()).item()
    agree = (y.argmax(1) == ref.argmax(1)).float().mean().item()
    print(f"{wl} {spec:20s} max-abs {err:.3e}  rel-L2 {rl2:.3e}  argmax {agree:.5f}")
    del m
