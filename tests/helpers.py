"""Shared helpers for the parity tests (test infrastructure)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from cat_seg_b200.config import vitb, vitl  # noqa: E402
from cat_seg_b200.synth import make_inputs, make_state_dict  # noqa: E402

# must stay in sync with tests/golden/make_golden.py
GOLDEN_CASES = {
    "vitb_T5_B1": (vitb(), 1, 5, 0, True),
    "vitb_T3_B2_pool2": (vitb(pooling_size=(2, 2)), 2, 3, 1, False),
    "vitb_T20_B1_cfg1": (vitb(), 1, 20, 2, True),
    "vitb_T9_B2_pad4": (vitb(pad_len=4), 2, 9, 3, False),
    "vitl_T300_B1": (vitl(), 1, 300, 4, True),
}
SUB_TOK, SUB_CH, SUB_PIX = 7, 5, 4
STAGES = ["embed", "swin_l0_b1", "swin_l0_b2", "class_l0", "swin_l1_b1", "swin_l1_b2", "class_l1", "up1", "up2"]


def fingerprint(t: torch.Tensor) -> np.ndarray:
    d = t.double()
    return np.array([d.sum().item(), d.abs().sum().item(), (d * d).sum().item()], dtype=np.float64)


def load_case(name):
    cfg, B, T, seed, same_text = GOLDEN_CASES[name]
    sd = make_state_dict(cfg, seed)
    img, text, g = make_inputs(cfg, B, T, seed, same_text)
    gold = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    return cfg, B, T, sd, (img, text, g), gold


def check_inputs_match_golden(sd, inputs, gold):
    """The goldens were made from regenerated seeded tensors; make sure this machine regenerates the same bits."""
    img, text, g = inputs
    w = torch.cat([v.flatten() for v in sd.values()])
    for key, t in (("fp_weights", w), ("fp_img", img), ("fp_text", text), ("fp_g1", g[1]), ("fp_g2", g[2])):
        np.testing.assert_allclose(fingerprint(t), gold[key], rtol=1e-12, atol=0, err_msg=key)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def argmax_agreement(logits: torch.Tensor, ref: torch.Tensor):
    """(raw agreement, margin-filtered agreement, filtered fraction, max abs err) over pixels.

    `raw` is the number north_star gates on.  The margin-filtered value (pixels whose reference top1-top2
    margin exceeds 2x the max absolute logit error) is a SANITY check only: such pixels cannot flip by
    construction, so it can only fail on an indexing bug; `frac` says how many pixels it covers."""
    err = (logits - ref).abs().max().item()
    a, r = logits.argmax(dim=1), ref.argmax(dim=1)
    raw = (a == r).float().mean().item()
    if ref.shape[1] < 2:
        return raw, 1.0, 1.0, err
    top2 = ref.topk(2, dim=1)[0]
    safe = (top2[:, 0] - top2[:, 1]) > 2 * err
    filt = (a == r)[safe].float().mean().item() if safe.any() else 1.0
    return raw, filt, safe.float().mean().item(), err

