"""CPU oracle for the CAT-Seg cost-aggregation hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain torch-fp32 (CPU) restatement of the reference algorithm in
``cat_seg/modeling/transformer/model.py`` (``Aggregator.forward`` :683-725 and everything it
calls).  It is written functionally against a reference-named ``state_dict`` and keeps every
activation in the token-major layout the CUDA kernels use (``[B, Te, H*W, hidden]``), so that the
stage-wise parity tests can compare intermediates directly.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline / ``--impl reference``
legs may import it.  The product path (``cat-seg_b200/``) never does.

Parity pinning: the reference ships no tests or golden vectors, so the oracle is
pinned against the *reference itself*, imported unchanged from ``/root/reference`` in the build
container (``oracle/ref_loader.py``, ``tests/test_oracle_vs_reference.py``) and against fixtures
generated from the reference by ``tests/golden/make_golden.py`` (``tests/golden/*.npz``).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------
def _ln(x: Tensor, sd: Dict[str, Tensor], prefix: str) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], 1e-5)


def _lin(x: Tensor, sd: Dict[str, Tensor], prefix: str) -> Tensor:
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def shift_region_ids(H: int, W: int, window: int, shift: int) -> Tensor:
    """Region id of every *shifted* grid position, as built at model.py:161-176.

    Three bands per axis: [0, H-window), [H-window, H-shift), [H-shift, H).
    """
    def band(n: int, size: int) -> Tensor:
        b = torch.zeros(size, dtype=torch.long)
        b[size - window: size - shift] = 1
        b[size - shift:] = 2
        return b

    return band(H, H)[:, None] * 3 + band(W, W)[None, :]


# --------------------------------------------------------------------------------------
# a2/a3: cost volume and top-k class selection  (model.py:648-652, 694-702)
# --------------------------------------------------------------------------------------
def cost_volume(img_feats: Tensor, text_feats: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """Returns (corr [B,T,P,HW], img_n [B,C,HW], text_n [B,T,P,C])."""
    B, C, H, W = img_feats.shape
    img_n = F.normalize(img_feats, dim=1).reshape(B, C, H * W)
    text_n = F.normalize(text_feats, dim=-1)
    corr = torch.einsum("bcx,btpc->btpx", img_n, text_n)
    return corr, img_n, text_n


def select_classes(corr: Tensor, pad_len: int) -> Optional[Tensor]:
    """Top-`pad_len` classes per image by max-over-(P,HW) similarity (model.py:694-696).

    The reference uses topk(sorted=False); its output is invariant to the order of the kept
    set, so the oracle (and the kernels) use ascending class id.  Ties at the
    cut are broken towards the lower class id.
    """
    B, T = corr.shape[:2]
    if pad_len <= 0 or T <= pad_len:
        return None
    score = corr.reshape(B, T, -1).max(dim=-1)[0]
    # rank = number of classes that beat this one (higher score, or equal score and lower id)
    gt = score[:, None, :] > score[:, :, None]
    eq = (score[:, None, :] == score[:, :, None]) & (
        torch.arange(T)[None, None, :] < torch.arange(T)[None, :, None])
    rank = (gt | eq).sum(dim=-1)
    keep = rank < pad_len
    idx = torch.arange(T)[None, :].expand(B, T)[keep].reshape(B, pad_len)
    return idx


# --------------------------------------------------------------------------------------
# a4: cost embedding (model.py:654-659)
# --------------------------------------------------------------------------------------
def cost_embed(corr_kept: Tensor, sd: Dict[str, Tensor], H: int, W: int) -> Tensor:
    """corr_kept [B,Te,P,HW] -> [B,Te,HW,hidden]; 7x7 conv with P input channels."""
    B, Te, P, HW = corr_kept.shape
    x = corr_kept.reshape(B * Te, P, H, W)
    y = F.conv2d(x, sd["conv1.weight"], sd["conv1.bias"], padding=3)
    return y.reshape(B, Te, -1, HW).permute(0, 1, 3, 2).contiguous()


# --------------------------------------------------------------------------------------
# a5: guidance projections (model.py:707-715)
# --------------------------------------------------------------------------------------
def project_guidance(sd: Dict[str, Tensor], guidance: Sequence[Tensor]) -> Tuple[Tensor, List[Tensor]]:
    """Returns appearance guidance [B,HW,128] (token-major) and decoder guidance NCHW list."""
    g0 = F.relu(F.conv2d(guidance[0], sd["guidance_projection.0.weight"],
                         sd["guidance_projection.0.bias"], padding=1))
    B, Cg = g0.shape[:2]
    g0 = g0.reshape(B, Cg, -1).permute(0, 2, 1).contiguous()
    dec = [F.relu(F.conv2d(g, sd[f"decoder_guidance_projection.{i}.0.weight"],
                           sd[f"decoder_guidance_projection.{i}.0.bias"], padding=1))
           for i, g in enumerate(guidance[1:])]
    return g0, dec


def project_text(sd: Dict[str, Tensor], text_used: Tensor) -> Tensor:
    """text_used [B,Te,P,C] (raw, or normalised+gathered when truncated) -> [B,Te,128]."""
    t = text_used.mean(dim=-2)
    t = t / t.norm(dim=-1, keepdim=True)
    return F.relu(_lin(t, sd, "text_guidance_projection.0"))


# --------------------------------------------------------------------------------------
# a7-a9: spatial aggregation (Swin blocks with appearance guidance), model.py:51-253
# --------------------------------------------------------------------------------------
def swin_block(x: Tensor, g_norm: Tensor, sd: Dict[str, Tensor], prefix: str, *, H: int, W: int,
               window: int, shift: int, nheads: int) -> Tensor:
    """x [N, HW, C] (N = B*Te slices), g_norm [N or B-broadcastable, HW, Cg] already LayerNorm'd."""
    N, HW, C = x.shape
    hd = C // nheads
    nwh, nww = H // window, W // window
    xn = _ln(x, sd, prefix + ".norm1")
    cat = torch.cat([xn, g_norm.expand(N, HW, -1)], dim=-1).reshape(N, H, W, -1)
    if shift > 0:
        cat = torch.roll(cat, shifts=(-shift, -shift), dims=(1, 2))
    # window partition: [N, nwh, window, nww, window, C'] -> [N*nw, window*window, C']
    win = cat.reshape(N, nwh, window, nww, window, -1).permute(0, 1, 3, 2, 4, 5)
    win = win.reshape(N * nwh * nww, window * window, -1)
    q = _lin(win, sd, prefix + ".attn.q")
    k = _lin(win, sd, prefix + ".attn.k")
    v = _lin(win[..., :C], sd, prefix + ".attn.v")
    L = window * window

    def heads(t: Tensor) -> Tensor:
        return t.reshape(-1, L, nheads, hd).permute(0, 2, 1, 3)

    q, k, v = heads(q) * hd ** -0.5, heads(k), heads(v)
    att = q @ k.transpose(-2, -1)                              # [N*nw, heads, L, L]
    if shift > 0:
        ids = shift_region_ids(H, W, window, shift)
        ids = ids.reshape(nwh, window, nww, window).permute(0, 2, 1, 3).reshape(nwh * nww, L)
        mask = torch.where(ids[:, :, None] != ids[:, None, :], -100.0, 0.0).to(x.dtype)
        att = (att.reshape(N, nwh * nww, nheads, L, L) + mask[None, :, None]).reshape(-1, nheads, L, L)
    att = att.softmax(dim=-1)
    o = (att @ v).transpose(1, 2).reshape(-1, L, C)
    o = _lin(o, sd, prefix + ".attn.proj")
    o = o.reshape(N, nwh, nww, window, window, C).permute(0, 1, 3, 2, 4, 5).reshape(N, H, W, C)
    if shift > 0:
        o = torch.roll(o, shifts=(shift, shift), dims=(1, 2))
    x = x + o.reshape(N, HW, C)
    h = F.gelu(_lin(_ln(x, sd, prefix + ".norm2"), sd, prefix + ".mlp.fc1"))
    return x + _lin(h, sd, prefix + ".mlp.fc2")


# --------------------------------------------------------------------------------------
# a10-a12: class aggregation (linear attention over classes), model.py:256-286, 323-424
# --------------------------------------------------------------------------------------
def class_layer(x: Tensor, text_g: Tensor, sd: Dict[str, Tensor], prefix: str, *, H: int, W: int,
                pool: Tuple[int, int], nheads: int, pad_len: int) -> Tensor:
    """x [B,Te,HW,C], text_g [B,Te,Cg] -> [B,Te,HW,C]."""
    B, Te, HW, C = x.shape
    hd = C // nheads
    xp = x.reshape(B * Te, H, W, C).permute(0, 3, 1, 2)
    xp = F.avg_pool2d(xp, pool)
    Hp, Wp = xp.shape[-2:]
    xp = xp.reshape(B, Te, C, Hp * Wp).permute(0, 3, 1, 2)       # [B, HpWp, Te, C]
    g = text_g
    n_pad = pad_len - Te if (pad_len > 0 and Te < pad_len) else 0
    if n_pad > 0:
        xp = torch.cat([xp, sd[prefix + ".padding_tokens"].reshape(1, 1, 1, C).expand(B, Hp * Wp, n_pad, C)], dim=2)
        g = torch.cat([g, sd[prefix + ".padding_guidance"].reshape(1, 1, -1).expand(B, n_pad, -1)], dim=1)
    S = xp.shape[2]
    tok = xp.reshape(B * Hp * Wp, S, C)
    gg = g[:, None].expand(B, Hp * Wp, S, g.shape[-1]).reshape(B * Hp * Wp, S, -1)

    xn = _ln(tok, sd, prefix + ".norm1")
    cat = torch.cat([xn, gg], dim=-1)
    q = _lin(cat, sd, prefix + ".attention.q").reshape(-1, S, nheads, hd)
    k = _lin(cat, sd, prefix + ".attention.k").reshape(-1, S, nheads, hd)
    v = _lin(xn, sd, prefix + ".attention.v").reshape(-1, S, nheads, hd)
    Q = F.elu(q) + 1
    K = F.elu(k) + 1
    v = v / S
    KV = torch.einsum("nshd,nshv->nhdv", K, v)
    Z = 1 / (torch.einsum("nlhd,nhd->nlh", Q, K.sum(dim=1)) + 1e-6)
    att = (torch.einsum("nlhd,nhdv,nlh->nlhv", Q, KV, Z) * S).reshape(-1, S, C)
    tok = tok + att
    h = F.relu(_lin(_ln(tok, sd, prefix + ".norm2"), sd, prefix + ".MLP.0"))
    tok = tok + _lin(h, sd, prefix + ".MLP.2")

    tok = tok.reshape(B, Hp * Wp, S, C)[:, :, :Te]              # drop padding tokens
    up = tok.permute(0, 2, 3, 1).reshape(B * Te, C, Hp, Wp)
    up = F.interpolate(up, size=(H, W), mode="bilinear", align_corners=True)
    up = up.reshape(B, Te, C, HW).permute(0, 1, 3, 2)
    return x + up


# --------------------------------------------------------------------------------------
# a13-a15: decoder (model.py:520-555, 674-681)
# --------------------------------------------------------------------------------------
def _up(x: Tensor, g: Tensor, sd: Dict[str, Tensor], prefix: str, Te: int) -> Tensor:
    x = F.conv_transpose2d(x, sd[prefix + ".up.weight"], sd[prefix + ".up.bias"], stride=2)
    gg = g[:, None].expand(-1, Te, -1, -1, -1).reshape(-1, *g.shape[1:])
    x = torch.cat([x, gg], dim=1)
    for i in (0, 3):
        w = sd[f"{prefix}.conv.double_conv.{i}.weight"]
        x = F.conv2d(x, w, None, padding=1)
        x = F.group_norm(x, w.shape[0] // 16, sd[f"{prefix}.conv.double_conv.{i + 1}.weight"],
                         sd[f"{prefix}.conv.double_conv.{i + 1}.bias"], 1e-5)
        x = F.relu(x)
    return x


def decoder(x: Tensor, dec_g: Sequence[Tensor], sd: Dict[str, Tensor], H: int, W: int,
            return_stages: bool = False):
    """x [B,Te,HW,C] -> logits [B,Te,4H,4W]."""
    B, Te, HW, C = x.shape
    y = x.reshape(B * Te, H, W, C).permute(0, 3, 1, 2)
    u1 = _up(y, dec_g[0], sd, "decoder1", Te)
    u2 = _up(u1, dec_g[1], sd, "decoder2", Te)
    out = F.conv2d(u2, sd["head.weight"], sd["head.bias"], padding=1)
    out = out.reshape(B, Te, 4 * H, 4 * W)
    if return_stages:
        return out, {"up1": u1.permute(0, 2, 3, 1).reshape(B, Te, 4 * HW, -1),
                     "up2": u2.permute(0, 2, 3, 1).reshape(B, Te, 16 * HW, -1)}
    return out


# --------------------------------------------------------------------------------------
# a1: the boundary function (model.py:683-725)
# --------------------------------------------------------------------------------------
@torch.no_grad()
def aggregator_forward(sd: Dict[str, Tensor], cfg: dict, img_feats: Tensor, text_feats: Tensor,
                       guidance: Sequence[Tensor], return_stages: bool = False):
    """cfg keys: num_layers, nheads, hidden_dim, pooling_size, feature_resolution, window_size,
    pad_len.  Returns logits [B,T,4H,4W] (and a dict of stage outputs)."""
    H, W = cfg["feature_resolution"]
    L, nheads, pad_len = cfg["num_layers"], cfg["nheads"], cfg["pad_len"]
    window, pool = cfg["window_size"], tuple(cfg["pooling_size"])
    B, T = text_feats.shape[:2]
    stages: Dict[str, Tensor] = {}

    corr, img_n, text_n = cost_volume(img_feats, text_feats)
    classes = select_classes(corr, pad_len)
    if classes is not None:
        gi = classes[:, :, None, None]
        corr_kept = torch.gather(corr, 1, gi.expand(-1, -1, *corr.shape[2:]))
        text_used = torch.gather(text_n, 1, gi.expand(-1, -1, *text_n.shape[2:]))
    else:
        corr_kept, text_used = corr, text_feats
    Te = corr_kept.shape[1]
    x = cost_embed(corr_kept, sd, H, W)
    app_g, dec_g = project_guidance(sd, guidance)
    text_g = project_text(sd, text_used)
    if return_stages:
        stages.update(corr=corr, embed=x, app_guidance=app_g, text_guidance=text_g,
                      dec_guidance0=dec_g[0], dec_guidance1=dec_g[1])
        if classes is not None:
            stages["classes"] = classes

    for l in range(L):
        p = f"layers.{l}"
        g_norm = _ln(app_g, sd, p + ".swin_block.guidance_norm")            # [B,HW,128]
        g_rep = g_norm[:, None].expand(B, Te, *g_norm.shape[1:]).reshape(B * Te, *g_norm.shape[1:])
        xs = x.reshape(B * Te, H * W, -1)
        xs = swin_block(xs, g_rep, sd, p + ".swin_block.block_1", H=H, W=W, window=window, shift=0, nheads=nheads)
        if return_stages:
            stages[f"swin_l{l}_b1"] = xs.reshape(x.shape)
        xs = swin_block(xs, g_rep, sd, p + ".swin_block.block_2", H=H, W=W, window=window,
                        shift=window // 2, nheads=nheads)
        x = xs.reshape(x.shape)
        if return_stages:
            stages[f"swin_l{l}_b2"] = x
        x = class_layer(x, text_g, sd, p + ".attention", H=H, W=W, pool=pool, nheads=nheads, pad_len=pad_len)
        if return_stages:
            stages[f"class_l{l}"] = x

    if return_stages:
        logit, dstages = decoder(x, dec_g, sd, H, W, return_stages=True)
        stages.update(dstages)
    else:
        logit = decoder(x, dec_g, sd, H, W)
    if classes is not None:
        out = torch.full((B, T, logit.shape[2], logit.shape[3]), -100.0, dtype=logit.dtype)
        out.scatter_(1, classes[:, :, None, None].expand(-1, -1, *logit.shape[2:]), logit)
        logit = out
    if return_stages:
        stages["logits_kept"] = logit if classes is None else torch.gather(
            logit, 1, classes[:, :, None, None].expand(-1, -1, *logit.shape[2:]))
        return logit, stages
    return logit
