"""GPU: per-stage error of a precision spec against the CPU oracle (and timing per stage).

    python tests/tools/check_fast.py --precision fast:swin_mlp [--model vitb --B 1 --T 5 --pool 1]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from cat_seg_b200.aggregator import Aggregator  # noqa: E402
from cat_seg_b200.config import vitb, vitl  # noqa: E402
from cat_seg_b200.synth import make_inputs, make_state_dict  # noqa: E402
from helpers import argmax_agreement, rel_l2  # noqa: E402
from oracle.aggregator_oracle import aggregator_forward  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="fast")
    ap.add_argument("--model", default="vitb")
    ap.add_argument("--B", type=int, default=1)
    ap.add_argument("--T", type=int, default=5)
    ap.add_argument("--pool", type=int, default=1)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--time", type=int, default=0, help="timed repetitions")
    a = ap.parse_args()
    cfg = (vitb if a.model == "vitb" else vitl)(pooling_size=(a.pool, a.pool))
    sd = make_state_dict(cfg, a.seed)
    img, text, g = make_inputs(cfg, a.B, a.T, a.seed, same_text=False)
    ref, st = aggregator_forward(sd, cfg.oracle_cfg(), img, text, g, return_stages=True)
    names = ["app_guidance", "dec_guidance0", "dec_guidance1", "embed"] + [f"{k}{l}{s}" for l in range(cfg.num_layers) for k, s in
                         (("swin_l", "_b1"), ("swin_l", "_b2"), ("class_l", ""))] + ["up1", "up2"]
    for prec in ["exact", a.precision]:
        if prec == "fast" or "decoder" in prec:
            names = [n for n in names if n not in ("up1", "up2")]
        m = Aggregator(**cfg.ctor_kwargs(), precision=prec)
        m.load_state_dict(sd, strict=False)
        m = m.cuda()
        cu = (img.cuda(), text.cuda(), [x.cuda() for x in g])
        y, taps = m(*cu, taps=names)
        torch.cuda.synchronize()
        print(f"== precision {prec}")
        for n in names:
            got = taps[n].cpu()
            if n.startswith("dec_guidance"):
                st[n] = st[n].permute(0, 2, 3, 1).reshape(got.shape) if st[n].shape != got.shape else st[n]
            print(f"  {n:12s} max-abs {float((got - st[n]).abs().max()):.3e}  rel-L2 {rel_l2(got, st[n]):.3e}  "
                  f"(ref rms {float(st[n].pow(2).mean().sqrt()):.3e})")
        yc = y.cpu()
        raw, filt, frac, err = argmax_agreement(yc, ref)
        print(f"  logits       max-abs {err:.3e}  rel-L2 {rel_l2(yc, ref):.3e}  argmax raw {raw:.4f} "
              f"margin-filtered {filt:.4f} (on {frac:.3f} of pixels)  mask-equal {bool(((yc == -100) == (ref == -100)).all())}")
        if a.time:
            m.set_profiling(True)
            m.stage_times(reset=True)
            for _ in range(a.time):
                m(*cu)
            ms, calls = m.stage_times()
            print("  stage ms/forward:", {k: round(v / calls, 3) for k, v in ms.items()})


if __name__ == "__main__":
    main()
