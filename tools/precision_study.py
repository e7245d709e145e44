"""Operand-precision study on the CPU oracle (test infrastructure; not product code).

Question (VERDICT r1 item 1 / SURVEY.md §0.9, §7.3): which tensor-core operand scheme keeps the
raw argmax agreement with the fp32 reference above 99.9 % under random-init weights?  The oracle's
contractions inside the Swin blocks, class layers and decoder are re-run with emulated operand
rounding; accumulation stays fp32 (what tcgen05 does).  Everything in front of the aggregation
layers (cost volume, 7x7 embedding, guidance projections) stays fp32: the product computes those
fp32-accurately in every mode.

Schemes: <act>/<weight> with each side one of
  b1 = one bf16 term, b2 = hi+lo bf16 terms, h1 = one fp16 term, h2 = hi+lo fp16 terms, f = fp32.
A product of an n-term and an m-term operand is emulated with all cross terms except lo*lo.

    python tools/precision_study.py [case ...]
"""
from __future__ import annotations

import os
import sys
import time

import torch
import torch.nn.functional as F
from torch.overrides import TorchFunctionMode

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cat_seg_b200.config import vitb, vitl  # noqa: E402
from cat_seg_b200.synth import make_inputs, make_state_dict  # noqa: E402
from oracle import aggregator_oracle as orc  # noqa: E402


def split(x: torch.Tensor, kind: str):
    if kind == "f":
        return [x]
    dt = torch.bfloat16 if kind[0] == "b" else torch.float16
    hi = x.to(dt).float()
    if kind[1] == "1":
        return [hi]
    lo = (x - hi).to(dt).float()
    return [hi, lo]


class Emu(TorchFunctionMode):
    """Rounds the operands of every contraction; `wk` applies to parameters, `ak` to activations."""

    def __init__(self, ak: str, wk: str, aak: str | None = None):
        super().__init__()
        self.ak, self.wk, self.aak = ak, wk, aak or ak

    def _prod(self, fn, a, b, ka, kb):
        if ka == "n1":     # single fp16 attention weights, renormalised by the sum of the ROUNDED weights
            hi = a.to(torch.float16).float()
            return self._prod(fn, hi, b, "f", kb) / hi.sum(-1, keepdim=True) * a.sum(-1, keepdim=True)
        A, Bs = split(a, ka), split(b, kb)
        out = None
        for i, ai in enumerate(A):
            for j, bj in enumerate(Bs):
                if i == 1 and j == 1:
                    continue
                t = fn(ai, bj)
                out = t if out is None else out + t
        return out

    conv_idx = 0
    cur_lin = None      # set by the patched orc._lin: name of the linear layer being evaluated

    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func is F.linear:
            x, w = args[0], args[1]
            b = args[2] if len(args) > 2 else kwargs.get("bias")
            ka, kb = self.ak, self.wk
            if isinstance(self.aak, dict) and Emu.cur_lin is not None:
                for key, v in self.aak.items():
                    if key.startswith("L:") and Emu.cur_lin.endswith(key[2:]):
                        ka, kb = v if isinstance(v, tuple) else (v, v)
            y = self._prod(lambda p, q: func(p, q), x, w, ka, kb)
            return y if b is None else y + b
        if func in (F.conv2d, F.conv_transpose2d):
            x, w = args[0], args[1]
            b = args[2] if len(args) > 2 else kwargs.get("bias")
            kw = {k: v for k, v in kwargs.items() if k != "bias"}
            ka, kb = self.ak, self.wk
            key = "C:%d" % Emu.conv_idx
            Emu.conv_idx += 1
            if isinstance(self.aak, dict) and key in self.aak:
                v = self.aak[key]
                ka, kb = v if isinstance(v, tuple) else (v, v)
            y = self._prod(lambda p, q: func(p, q, None, *args[3:], **kw), x, w, ka, kb)
            return y if b is None else y + b.reshape(1, -1, 1, 1)
        if func in (torch.matmul, torch.Tensor.matmul, torch.Tensor.__matmul__):
            # swin_block: first matmul of a block is q k^T, second is att v
            site = "qk" if args[1].shape[-2] == 32 else "pv"
            ka, kb = self.site(site)
            return self._prod(lambda p, q: func(p, q), args[0], args[1], ka, kb)
        if func is torch.einsum and len(args) == 3:
            site = {"nshd,nshv->nhdv": "kv", "nlhd,nhd->nlh": "qz"}[args[0]]
            ka, kb = self.site(site)
            return self._prod(lambda p, q: func(args[0], p, q), args[1], args[2], ka, kb)
        if func is torch.einsum and len(args) == 4:      # "nlhd,nhdv,nlh->nlhv": apply, then the normaliser
            ka, kb = self.site("ap")
            y = self._prod(lambda p, q: func("nlhd,nhdv->nlhv", p, q), args[1], args[2], ka, kb)
            return y * args[3][..., None]
        return func(*args, **kwargs)

    def site(self, name):
        if isinstance(self.aak, dict):
            v = self.aak.get(name, self.aak.get("*", self.ak))
        else:
            v = self.aak
        return (v, v) if isinstance(v, str) else v


def run(case: str, schemes):
    cfg, B, T, seed = {
        "vitb_T20": (vitb(), 1, 20, 2),
        "vitb_T150": (vitb(), 1, 150, 5),
        "vitl_T300": (vitl(), 1, 300, 4),
    }[case]
    sd = make_state_dict(cfg, seed)
    img, text, g = make_inputs(cfg, B, T, seed, True)
    ref = orc.aggregator_forward(sd, cfg.oracle_cfg(), img, text, g)
    keep = ref[0, :, 0, 0] != -100.0
    top2 = ref[:, keep].topk(2, dim=1)[0]
    margin = top2[:, 0] - top2[:, 1]
    print(f"== {case}: logits std {ref[:, keep].std():.3f}, margin quantiles 0.1%/1%/10% = "
          f"{margin.flatten().quantile(0.001):.2e} {margin.flatten().quantile(0.01):.2e} {margin.flatten().quantile(0.1):.2e}")
    names = ("swin_block", "class_layer", "decoder")
    orig = {n: getattr(orc, n) for n in names}
    lin0 = orc._lin

    def lin_named(x, sd_, prefix):
        Emu.cur_lin = prefix
        try:
            return lin0(x, sd_, prefix)
        finally:
            Emu.cur_lin = None
    orc._lin = lin_named
    for sch in schemes:
        parts = sch.split("/")
        ak, wk = parts[0], parts[1]
        aak = parts[2] if len(parts) > 2 else None
        if aak and "=" in aak:       # e.g. *=h2,qk=h1,pv=h2:h1  (site=kindA[:kindB])
            d = {}
            for it in aak.split(","):
                k, v = it.split("=")
                d[k] = tuple(v.split(":")) if ":" in v else v
            aak = d

        def wrap(fn):
            def inner(*a, **k):
                with Emu(ak, wk, aak):
                    return fn(*a, **k)
            return inner

        for n in names:
            setattr(orc, n, wrap(orig[n]))
        t0 = time.time()
        Emu.conv_idx = 0
        out = orc.aggregator_forward(sd, cfg.oracle_cfg(), img, text, g)
        for n in names:
            setattr(orc, n, orig[n])
        d = (out[:, keep] - ref[:, keep])
        agree = (out[:, keep].argmax(1) == ref[:, keep].argmax(1)).float().mean().item()
        print(f"  {sch:10s} max-abs {d.abs().max():.2e} rms {d.pow(2).mean().sqrt():.2e} rel-L2 "
              f"{(d.norm() / ref[:, keep].norm()):.2e} argmax {100 * agree:.3f} %   ({time.time() - t0:.0f}s)", flush=True)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    cases = [a for a in sys.argv[1:] if "/" not in a] or ["vitb_T20"]
    schemes = [a for a in sys.argv[1:] if "/" in a] or ["b1/b1", "h1/h1", "b2/b1", "h2/h1", "b2/b2", "h2/h2", "h2/h1/h1"]
    for c in cases:
        run(c, schemes)
