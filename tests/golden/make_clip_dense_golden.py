"""Generate tests/golden/clip_dense.npz from the reference's own classes (run in the build container).

cat_seg/third_party/model_vpt.py imports only torch, so it is loaded by path, UNMODIFIED; a ResidualAttentionBlock and the
ln_post / proj of VisualTransformer are built exactly as the reference builds them (:186-200, :283-284), loaded with the
seeded parameters of cat_seg_b200.synth.make_clip_dense_inputs, and run through forward_dense (:219-240) and the dense tail of
VisualTransformer.forward (:300-312, executed statement by statement).

    python tests/golden/make_clip_dense_golden.py
"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cat_seg_b200.synth import CLIP_DENSE_KEYS, make_clip_dense_inputs  # noqa: E402

REF = "/root/reference/cat_seg/third_party/model_vpt.py"
# name: (width, heads, L, N, out_dim, prompt, seed)
CASES = {"w64_L10_N2": (64, 4, 10, 2, 32, 0, 0), "w64_L12_N3_prompt2": (64, 4, 12, 3, 40, 2, 1),
         "w768_L577_N1": (768, 12, 577, 1, 512, 0, 2)}
SUB = 53


def load_ref():
    spec = importlib.util.spec_from_file_location("ref_model_vpt", REF)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def reference_dense(ref, width, heads, sd, x, prompt):
    blk = ref.ResidualAttentionBlock(width, heads).eval()                      # model_vpt.py:186-200
    blk.load_state_dict({k: sd[k] for k in CLIP_DENSE_KEYS}, strict=True)
    ln_post = ref.LayerNorm(width)                                              # :283
    ln_post.load_state_dict({"weight": sd["ln_post.weight"], "bias": sd["ln_post.bias"]})
    proj = sd["proj"]                                                           # :284
    with torch.no_grad():
        v = blk.forward_dense(x, prompt)                                        # Transformer.forward :262
        y = v.permute(1, 0, 2)                                                  # :302
        y = ln_post(y[:, :, :])                                                 # :305
        y = y @ proj                                                            # :310
    return v, y


def main():
    ref = load_ref()
    out = {}
    for name, (width, heads, L, N, od, prompt, seed) in CASES.items():
        x, sd = make_clip_dense_inputs(width, L, N, od, seed)
        v, y = reference_dense(ref, width, heads, sd, x, prompt)
        for k, t in (("block_out", v), ("feats", y)):
            out[f"{name}/{k}/shape"] = np.array(t.shape, dtype=np.int64)
            out[f"{name}/{k}/sub"] = t.contiguous().flatten()[::SUB].numpy().copy()
            d = t.double()
            out[f"{name}/{k}/fp"] = np.array([d.sum().item(), d.abs().sum().item(), (d * d).sum().item()], dtype=np.float64)
        out[f"{name}/cfg"] = np.array([width, heads, L, N, od, prompt, seed], dtype=np.int64)
    path = os.path.join(ROOT, "tests", "golden", "clip_dense.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
