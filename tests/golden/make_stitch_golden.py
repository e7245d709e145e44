"""Golden vectors for the sliding-window tiling / stitch path, produced by the REFERENCE's own statements.

``CATSeg`` cannot be imported here (detectron2, CLIP weights), so the two sliding-window blocks of
``CATSeg.forward`` -- cat_seg/cat_seg_model.py:156-176 (tiling) and :204-218 (stitch) -- are located with ``ast`` in
the unmodified file under /root/reference, their source text is taken verbatim from that file at run time (nothing is
copied into this repository) and executed with a stub ``self`` (pixel statistics, a no-op ``encode_image``) and
Detectron2's ``sem_seg_postprocess`` semantics (crop to ``img_size``, bilinear ``align_corners=False`` resize; the
function is third-party, detectron2/modeling/postprocessing.py, and not vendored by the reference).

    python tests/golden/make_stitch_golden.py      # writes tests/golden/stitch_*.npz (build container only)
"""
import ast
import os
import textwrap
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from einops import rearrange

REF = "/root/reference/cat_seg/cat_seg_model.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def sem_seg_postprocess(result, img_size, output_height, output_width):
    result = result[:, : img_size[0], : img_size[1]].expand(1, -1, -1, -1)
    return F.interpolate(result, size=(output_height, output_width), mode="bilinear", align_corners=False)[0]


def reference_blocks():
    """Source text of the bodies of the two `elif self.sliding_window:` branches of CATSeg.forward."""
    src = open(REF).read()
    tree = ast.parse(src)
    fwd = next(n for c in tree.body if isinstance(c, ast.ClassDef) and c.name == "CATSeg"
               for n in c.body if isinstance(n, ast.FunctionDef) and n.name == "forward")
    found = []

    def visit(node):
        if isinstance(node, ast.If):
            t = node.test
            if isinstance(t, ast.Attribute) and t.attr == "sliding_window" and not isinstance(t.ctx, ast.Store):
                found.append(node)
        for ch in ast.iter_child_nodes(node):
            visit(ch)

    visit(fwd)
    found.sort(key=lambda n: n.lineno)
    assert len(found) == 2, [n.lineno for n in found]
    lines = src.splitlines()
    out = []
    for n in found:
        w = n.body[0]
        assert isinstance(w, ast.With)
        body = "\n".join(lines[w.body[0].lineno - 1: w.body[-1].end_lineno])
        out.append((w.body[0].lineno, w.body[-1].end_lineno, textwrap.dedent(body)))
    return out


def run_reference(image_u8, win_logits_fn, height, width):
    (a0, a1, tiling), (b0, b1, stitching) = reference_blocks()
    stub = types.SimpleNamespace(
        device="cpu", sliding_window=True, clip_resolution=(384, 384),
        pixel_mean=torch.tensor([123.675, 116.280, 103.530]).view(-1, 1, 1), pixel_std=torch.tensor([58.395, 57.120, 57.375]).view(-1, 1, 1),
        clip_pixel_mean=torch.tensor([122.7709383, 116.7460125, 104.09373615]).view(-1, 1, 1),
        clip_pixel_std=torch.tensor([68.5005327, 66.6321579, 70.3231630]).view(-1, 1, 1))
    stub.sem_seg_head = types.SimpleNamespace(predictor=types.SimpleNamespace(
        clip_model=types.SimpleNamespace(encode_image=lambda x, dense=True: None)))
    ns = dict(self=stub, torch=torch, nn=nn, F=F, rearrange=rearrange, sem_seg_postprocess=sem_seg_postprocess,
              batched_inputs=[{"image": image_u8, "height": height, "width": width}])
    exec(compile(tiling, f"{REF}:{a0}-{a1}", "exec"), ns)
    windows = ns["image"].clone()                           # [5, 3, 384, 384]: 4 tiles + the global view (:164-168)
    ns["outputs"] = win_logits_fn(windows.shape[0])
    # the stitch block ends with `return [...]`: run it as a function body
    fn_src = "def _stitch(self, outputs, kernel, out_res, fold, unfold, batched_inputs):\n" + textwrap.indent(stitching, "    ")
    exec(compile(fn_src, f"{REF}:{b0}-{b1}", "exec"), ns)
    res = ns["_stitch"](stub, ns["outputs"], ns["kernel"], ns["out_res"], ns["fold"], ns["unfold"], ns["batched_inputs"])
    return windows, ns["outputs"], res[0]["sem_seg"], (a0, a1, b0, b1)


def main():
    cases = {"stitch_T3": (3, 120, 168, 5, None), "stitch_T6_dropped": (6, 97, 133, 7, [1, 4])}
    for name, (T, h, w, seed, dropped) in cases.items():
        g = torch.Generator().manual_seed(seed)
        image = torch.randint(0, 256, (3, 211, 307), generator=g).float()

        def logits(nwin, T=T, g=g, dropped=dropped):
            y = torch.randn(nwin, T, 96, 96, generator=g)
            if dropped:
                y[:, dropped] = -100.0                       # classes removed by the top-k truncation (model.py:721-724)
            return y

        windows, win_logits, sem, lines = run_reference(image, logits, h, w)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), image=image.numpy().astype(np.uint8),
                            win_logits=win_logits.numpy(), sem_seg=sem.numpy(), labels=sem.argmax(dim=0).numpy().astype(np.int32),
                            windows_sub=windows[:, :, ::16, ::16].numpy(),
                            windows_fp=np.array([windows.double().sum().item(), (windows.double() ** 2).sum().item()]),
                            height=h, width=w, ref_lines=np.array(lines))
        print(name, tuple(sem.shape), "reference lines", lines)


if __name__ == "__main__":
    main()
