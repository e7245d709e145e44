"""Generate tests/golden/*.npz from the UNMODIFIED reference Aggregator (run in the build container).

The reference ships no tests or golden vectors, and /root/reference does not exist
on the GPU box, so the pin is made here: the reference module is imported from
/root/reference/cat_seg/modeling/transformer/model.py (oracle/ref_loader.py), fed the seeded
synthetic weights/inputs of cat_seg_b200.synth, and its logits plus hook-captured intermediates
are stored (sub-sampled, with float64 checksums) in the kernels' token-major layout.

    python tests/golden/make_golden.py            # rewrites tests/golden/
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from cat_seg_b200.config import vitb, vitl  # noqa: E402
from cat_seg_b200.synth import make_inputs, make_state_dict  # noqa: E402
from oracle.ref_loader import build_reference_aggregator  # noqa: E402

CASES = {
    # name: (cfg, B, T, seed, same_text)
    "vitb_T5_B1": (vitb(), 1, 5, 0, True),
    "vitb_T3_B2_pool2": (vitb(pooling_size=(2, 2)), 2, 3, 1, False),
    "vitb_T20_B1_cfg1": (vitb(), 1, 20, 2, True),
    "vitb_T9_B2_pad4": (vitb(pad_len=4), 2, 9, 3, False),          # cheap truncation case
    "vitl_T300_B1": (vitl(), 1, 300, 4, True),                      # real pad_len=256 truncation
}

SUB_TOK, SUB_CH, SUB_PIX = 7, 5, 4     # sub-sampling strides (tokens, channels, logit pixels)


def fingerprint(t: torch.Tensor) -> np.ndarray:
    d = t.double()
    return np.array([d.sum().item(), d.abs().sum().item(), (d * d).sum().item()], dtype=np.float64)


def case_fingerprints(cfg, B, T, seed, same_text):
    sd = make_state_dict(cfg, seed)
    img, text, g = make_inputs(cfg, B, T, seed, same_text)
    w = torch.cat([v.flatten() for v in sd.values()])
    return sd, (img, text, g), dict(fp_weights=fingerprint(w), fp_img=fingerprint(img), fp_text=fingerprint(text),
                                    fp_g1=fingerprint(g[1]), fp_g2=fingerprint(g[2]))


def sub_tokens(x: torch.Tensor) -> np.ndarray:          # [B,Te,HW,C]
    return x[:, :, ::SUB_TOK, ::SUB_CH].contiguous().numpy()


def run_reference(cfg, B, T, seed, same_text):
    sd, (img, text, g), fps = case_fingerprints(cfg, B, T, seed, same_text)
    ref = build_reference_aggregator(cfg.ctor_kwargs(), sd)
    Te = min(T, cfg.pad_len) if cfg.pad_len > 0 else T
    H, W = cfg.feature_resolution
    cap = {}

    def tok_from_bthwc(o):          # [(B T), HW, C]
        return o.detach().reshape(B, Te, H * W, -1)

    def tok_from_bcthw(o):          # [B, C, T, H, W]
        return o.detach().permute(0, 2, 3, 4, 1).reshape(B, Te, H * W, -1)

    def tok_from_nchw(o):           # [(B T), C, h, w]
        n, c, h, w = o.shape
        return o.detach().permute(0, 2, 3, 1).reshape(B, Te, h * w, c)

    hooks = [ref.conv1.register_forward_hook(lambda m, i, o: cap.__setitem__("embed", tok_from_nchw(o)))]
    for l, layer in enumerate(ref.layers):
        hooks.append(layer.swin_block.block_1.register_forward_hook(
            lambda m, i, o, l=l: cap.__setitem__(f"swin_l{l}_b1", tok_from_bthwc(o))))
        hooks.append(layer.swin_block.block_2.register_forward_hook(
            lambda m, i, o, l=l: cap.__setitem__(f"swin_l{l}_b2", tok_from_bthwc(o))))
        hooks.append(layer.register_forward_hook(
            lambda m, i, o, l=l: cap.__setitem__(f"class_l{l}", tok_from_bcthw(o))))
    hooks.append(ref.decoder1.register_forward_hook(lambda m, i, o: cap.__setitem__("up1", tok_from_nchw(o))))
    hooks.append(ref.decoder2.register_forward_hook(lambda m, i, o: cap.__setitem__("up2", tok_from_nchw(o))))
    with torch.no_grad():
        logits = ref(img, text, g)
    for h in hooks:
        h.remove()
    out = dict(fps)
    out["meta"] = np.array([B, T, Te, seed, int(same_text)], dtype=np.int64)
    out["logits_sub"] = logits[:, :, ::SUB_PIX, ::SUB_PIX].contiguous().numpy()
    kept = (logits[:, :, 0, 0] != -100.0)
    out["kept_mask"] = kept.numpy()
    out["logits_fp"] = fingerprint(logits[kept])
    if T <= 20:
        out["logits_full"] = logits.numpy()
    truncated = Te < T
    for k, v in cap.items():
        # sums are invariant to the (unspecified) order of the kept set, sub-samples are not
        out[k + "_fp"] = fingerprint(v)
        if not truncated and T <= 9:
            out[k + "_sub"] = sub_tokens(v)
    return out


def main():
    gdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gdir, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    for name, (cfg, B, T, seed, same_text) in CASES.items():
        out = run_reference(cfg, B, T, seed, same_text)
        path = os.path.join(gdir, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, {k: v.shape for k, v in out.items() if k.endswith("_sub")}, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
