"""GPU (-m gpu): the CUDA path, called through the C ABI, against the oracle and the committed goldens.

Tolerances (stated per north_star): EXACT precision is fp32 CUDA-core arithmetic, so it differs from the
CPU oracle only by fp32 re-association and libm (expf/erff) rounding: logits max-abs <= 2e-4, rel-L2 <= 2e-5,
margin-filtered argmax agreement == 100 %, raw agreement >= 99.9 %.
"""
import numpy as np
import pytest
import torch

from helpers import (GOLDEN_CASES, STAGES, SUB_PIX, SUB_TOK, SUB_CH, argmax_agreement, check_inputs_match_golden,
                     fingerprint, load_case, rel_l2)
from cat_seg_b200.aggregator import Aggregator
from cat_seg_b200.config import vitb, vitl
from cat_seg_b200.synth import make_inputs, make_state_dict
from cat_seg_b200 import sliding_window as sw
from oracle.aggregator_oracle import aggregator_forward
from oracle import stitch_oracle

pytestmark = pytest.mark.gpu

EXACT_MAXABS, EXACT_RELL2 = 2e-4, 2e-5


def _module(cfg, sd, precision="exact"):
    m = Aggregator(**cfg.ctor_kwargs(), precision=precision)
    m.load_state_dict(sd, strict=False)
    return m.cuda()


def _cuda(inputs):
    img, text, g = inputs
    return img.cuda(), text.cuda(), [x.cuda() for x in g]


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_exact_logits_match_reference_golden(name):
    cfg, B, T, sd, inputs, gold = load_case(name)
    check_inputs_match_golden(sd, inputs, gold)
    y = _module(cfg, sd)(*_cuda(inputs)).cpu()
    assert tuple(y.shape) == (B, T, 96, 96)
    kept = (y[:, :, 0, 0] != -100.0).numpy()
    assert (kept == gold["kept_mask"]).all()                       # exact equality of the -100 mask
    assert bool(((y == -100.0).reshape(B, T, -1).all(-1).numpy() == ~gold["kept_mask"]).all())
    np.testing.assert_allclose(y[:, :, ::SUB_PIX, ::SUB_PIX].numpy(), gold["logits_sub"], rtol=0, atol=EXACT_MAXABS)
    if "logits_full" in gold:
        ref = torch.from_numpy(gold["logits_full"])
        assert rel_l2(y, ref) <= EXACT_RELL2
        raw, filt, frac, err = argmax_agreement(y, ref)
        assert filt == 1.0 and raw >= 0.999, (raw, filt, frac, err)


@pytest.mark.parametrize("name", ["vitb_T5_B1", "vitb_T3_B2_pool2", "vitb_T9_B2_pad4"])
def test_exact_stages_match_oracle(name):
    cfg, B, T, sd, inputs, gold = load_case(name)
    ref_logits, st = aggregator_forward(sd, cfg.oracle_cfg(), *inputs, return_stages=True)
    names = ["corr", "embed", "app_guidance", "text_guidance", "dec_guidance0", "dec_guidance1"] + STAGES
    if "classes" in st:
        names.append("classes")
    y, taps = _module(cfg, sd)(*_cuda(inputs), taps=names)
    for n in names:
        got, ref = taps[n].cpu(), st[n]
        if n == "classes":
            assert torch.equal(got.long(), ref), "kept class ids must be bit-exact (ascending order)"
            continue
        if n.startswith("dec_guidance"):
            ref = ref.permute(0, 2, 3, 1).reshape(got.shape)       # oracle keeps NCHW, kernels NHWC
        assert got.shape == ref.shape, n
        err = (got - ref).abs().max().item()
        assert err <= 2e-4 and rel_l2(got, ref) <= 2e-5, (n, err, rel_l2(got, ref))
        if n + "_sub" in gold:                                      # and against the reference's own tensors
            np.testing.assert_allclose(got[:, :, ::SUB_TOK, ::SUB_CH].numpy(), gold[n + "_sub"], rtol=0, atol=2e-4,
                                       err_msg=n)
    assert (y.cpu() - ref_logits).abs().max().item() <= EXACT_MAXABS


def test_exact_edge_cases():
    # T == 1 (single class), T == pad_len (no padding, no truncation), P == 3 prompt templates, text differing per image
    for cfg, B, T, same in [(vitb(), 1, 1, True), (vitb(pad_len=6), 2, 6, False), (vitb(prompt_channel=3), 1, 4, False),
                            (vitb(pad_len=0), 1, 3, True), (vitb(pooling_size=(4, 4), pad_len=5), 1, 7, True)]:
        sd = make_state_dict(cfg, 7)
        inputs = make_inputs(cfg, B, T, 7, same_text=same)
        ref = aggregator_forward(sd, cfg.oracle_cfg(), *inputs)
        y = _module(cfg, sd)(*_cuda(inputs)).cpu()
        assert bool(((y == -100.0) == (ref == -100.0)).all())
        assert (y - ref).abs().max().item() <= EXACT_MAXABS, (cfg, B, T)


def test_batch_and_class_order_invariance():
    """Size-independent properties: images are independent; permuting classes permutes the logits."""
    cfg = vitb()
    sd = make_state_dict(cfg, 3)
    img, text, g = make_inputs(cfg, 2, 6, 3, same_text=False)
    m = _module(cfg, sd)
    y = m(*_cuda((img, text, g)))
    y0 = m(*_cuda((img[:1], text[:1], [x[:1] for x in g])))
    assert torch.equal(y[:1], y0)                                   # same kernels, same order: bit-exact
    perm = torch.tensor([3, 0, 5, 1, 4, 2])
    yp = m(*_cuda((img, text[:, perm].contiguous(), g)))
    assert (yp - y[:, perm]).abs().max().item() <= 1e-5             # class sum is re-associated


def test_weight_update_is_picked_up():
    cfg = vitb()
    sd = make_state_dict(cfg, 5)
    inputs = _cuda(make_inputs(cfg, 1, 2, 5))
    m = _module(cfg, sd)
    y1 = m(*inputs).clone()
    with torch.no_grad():
        m.head.bias.add_(1.0)
    y2 = m(*inputs)
    assert torch.allclose(y2, y1 + 1.0, atol=1e-6)


def test_stitch_matches_oracle():
    g = torch.Generator().manual_seed(11)
    T = 7
    logits = torch.randn(5, T, 96, 96, generator=g) * 3
    logits[:, 2] = -100.0                                           # a class dropped by the top-k truncation
    for (h, w) in [(640, 640), (500, 375)]:
        ref_p, ref_l = stitch_oracle.stitch(logits, h, w)
        p, l = sw.stitch(logits.cuda(), h, w, want_probs=True, want_labels=True)
        p, l = p.cpu(), l.cpu().long()
        assert (p - ref_p).abs().max().item() <= 2e-6               # same arithmetic order; expf rounding only
        assert float(p[2].abs().max()) == 0.0
        top2 = ref_p.topk(2, dim=0)[0]
        safe = (top2[0] - top2[1]) > 4e-6
        assert bool((l == ref_l)[safe].all()) and (l == ref_l).float().mean().item() > 0.9999
        # argmax kernel itself is bit-exact on identical probabilities
        assert torch.equal(sw.argmax(ref_p.cuda()).cpu().long(), ref_l)


def test_stitch_skipping_dropped_planes_is_bit_identical():
    """The pre-pass that finds planes holding -100 everywhere (dropped classes) and skips them must not change a single bit:
    classes dropped in every window, in some windows only, in the global view only, and a live plane that merely STARTS with
    -100 values."""
    g = torch.Generator().manual_seed(17)
    T = 23
    logits = torch.randn(5, T, 96, 96, generator=g) * 2
    logits[:, 3] = -100.0                       # dropped everywhere
    logits[0, 5] = -100.0                       # dropped in one tile
    logits[1:4, 6] = -100.0                     # dropped in three tiles
    logits[4, 7] = -100.0                       # dropped in the global view only
    logits[:4, 8] = -100.0                      # live in the global view only
    logits[2, 9, :40] = -100.0                  # NOT a dropped plane: only its first rows are -100
    logits[:, 0] = -100.0                       # class 0 dropped: the first-maximum rule must still see its zeros
    x = logits.cuda()
    for (h, w) in [(640, 640), (500, 375)]:
        p0, l0 = sw.stitch(x, h, w, want_probs=True, want_labels=True, skip_dropped=False)
        p1, l1 = sw.stitch(x, h, w, want_probs=True, want_labels=True, skip_dropped=True)
        assert torch.equal(p0, p1) and torch.equal(l0, l1)
        l2 = sw.stitch(x, h, w, want_probs=False, want_labels=True, skip_dropped=True)[1]
        assert torch.equal(l2, l0)


def test_argmax_batched_matches_torch():
    g = torch.Generator().manual_seed(3)
    s = torch.randn(3, 37, 50, 7, generator=g)
    assert torch.equal(sw.argmax_batched(s.cuda()).cpu().long(), s.argmax(dim=1))


def test_argmax_first_max_wins():
    s = torch.zeros(4, 10, 10)
    s[1] = 1.0
    s[3] = 1.0
    assert torch.equal(sw.argmax(s.cuda()).cpu().long(), s.argmax(dim=0))


def test_class_sharded_path_emulated_on_one_gpu():
    """The class-sharded entry point on ONE GPU: the two 'ranks' of a shard group run one after the other and
    the all-reduce callback is emulated by record/replay -- pass 1 records the layer-0 partial states, pass 2 injects their
    sum and records layer 1, pass 3 injects both -- so every rank ends with exactly the sums a real all-reduce would give.
    The assembled logits must match the unsharded forward (fp32 summation order of the state only)."""
    import ctypes as C
    from cat_seg_b200 import _lib
    from cat_seg_b200.aggregator import assemble_class_sharded
    cfg = vitb(pad_len=4)                                           # T = 9 > pad_len: 4 kept classes, 2 per rank
    B, T, world = 2, 9, 2
    sd = make_state_dict(cfg, 13)
    inputs = _cuda(make_inputs(cfg, B, T, 13, same_text=False))
    m = _module(cfg, sd)
    ref = m(*inputs)
    lib = _lib.load()
    Te = m.kept_classes(T)
    assert Te % world == 0
    img, text, g = inputs
    ws = torch.empty(lib.catseg_workspace_bytes(m._handle, B, T), dtype=torch.uint8, device="cuda")
    recorded = {}                                                   # (pass, rank, layer) -> partial state
    sums = {}

    def run(rank, npass):
        calls = {"n": 0}

        def cb(_ctx, buf, count, _stream):
            layer = calls["n"]
            calls["n"] += 1
            off = int(buf) - ws.data_ptr()
            view = ws[off:off + 4 * count].view(torch.float32)
            recorded[(npass, rank, layer)] = view.clone()
            if layer in sums:
                view.copy_(sums[layer])
            return 0

        fn = _lib.ALLREDUCE_FN(cb)
        local = torch.empty(B, Te // world, 96, 96, device="cuda")
        kept = torch.empty(B, Te, dtype=torch.int32, device="cuda")
        rc = lib.catseg_forward_class_sharded(
            m._handle, *[C.c_void_p(t.data_ptr()) for t in (img, text, g[0], g[1], g[2], local, kept)],
            C.c_void_p(ws.data_ptr()), ws.numel(), B, T, rank, world, C.cast(fn, C.c_void_p), None,
            C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, lib.catseg_last_error(m._handle)
        torch.cuda.synchronize()
        return local, kept

    out = None
    for npass in range(cfg.num_layers + 1):
        locals_ = [run(r, npass) for r in range(world)]
        if npass < cfg.num_layers:                                  # layer `npass` partials are now computed from correct inputs
            sums[npass] = sum(recorded[(npass, r, npass)] for r in range(world))
        out = locals_
    gathered = torch.stack([o[0] for o in out])
    assert torch.equal(out[0][1], out[1][1])                        # every rank reports the same kept-class list
    y = assemble_class_sharded(gathered, out[0][1], T)
    assert bool(((y == -100.0) == (ref == -100.0)).all())
    assert (y - ref).abs().max().item() <= 1e-5


def test_cuda_graph_replay_matches_eager():
    from cat_seg_b200.host_pipeline import GraphRunner
    cfg = vitb()
    sd = make_state_dict(cfg, 2)
    a = _cuda(make_inputs(cfg, 1, 5, 2))
    b = _cuda(make_inputs(cfg, 1, 5, 7))
    for prec in ("exact", "fast"):
        m = Aggregator(**cfg.ctor_kwargs(), precision=prec)
        m.load_state_dict(sd, strict=False)
        m = m.cuda()
        eager_b = m(*b).clone()
        run = GraphRunner(m, *a)
        yb = run(*b).clone()
        ya = run(*a).clone()
        assert torch.equal(yb, eager_b), prec                    # same kernels, same order: bit-exact
        assert torch.equal(ya, m(*a)), prec


# ---- guidance pyramid producers: C-ABI kernels vs the oracle
@pytest.mark.parametrize("width,B,seed", [(64, 2, 0), (1024, 1, 1), (768, 3, 2)])
def test_guidance_pyramid_matches_oracle(width, B, seed):
    from cat_seg_b200.guidance import GuidancePyramid
    from cat_seg_b200.synth import make_pyramid_inputs
    from oracle.guidance_oracle import guidance_pyramid
    clip, la, lb, w1, b1, w2, b2 = make_pyramid_inputs(width, B, seed, feat_dim=width // 2 + 3)
    ref = guidance_pyramid(clip, la, lb, w1, b1, w2, b2)
    m = GuidancePyramid(proj_dim=width)
    missing = m.load_state_dict({"upsample1.weight": w1, "upsample1.bias": b1, "upsample2.weight": w2, "upsample2.bias": b2})
    assert not missing.missing_keys and not missing.unexpected_keys
    out = m.cuda()(clip.cuda(), la.cuda(), lb.cuda())
    assert list(out) == ["res5", "res4", "res3"]                      # the reference's dict order (cat_seg_model.py:186)
    assert torch.equal(out["res3"].cpu(), ref["res3"])                # pure re-layout: bit exact
    for k in ("res4", "res5"):
        got = out[k].cpu()
        assert got.shape == ref[k].shape
        # fp32 FFMA accumulation over `width` terms vs the float64-accumulated oracle: |err| <= ~width * 2^-24 * |terms|
        assert (got - ref[k]).abs().max().item() <= 2e-5, (k, (got - ref[k]).abs().max().item())


def test_guidance_pyramid_rejects_bad_arguments():
    from cat_seg_b200.guidance import strip_cls_nchw, upsample_tokens
    with pytest.raises(RuntimeError):
        strip_cls_nchw(torch.zeros(1, 577, 8))                        # CPU tensor: no fallback
    with pytest.raises(ValueError):
        strip_cls_nchw(torch.zeros(1, 500, 8, device="cuda"))
    with pytest.raises(ValueError):
        upsample_tokens(torch.zeros(577, 1, 8, device="cuda"), torch.zeros(9, 4, 2, 2), torch.zeros(4))


# ---- CLIP dense last block: C-ABI kernels vs the oracle and the reference's golden
@pytest.mark.parametrize("name", ["w64_L10_N2", "w64_L12_N3_prompt2", "w768_L577_N1"])
def test_clip_dense_matches_reference_golden(name):
    import os
    from cat_seg_b200.clip_dense import DenseLastBlock
    from cat_seg_b200.synth import make_clip_dense_inputs
    from oracle.clip_dense_oracle import dense_last_block
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "clip_dense.npz"))
    width, heads, L, N, od, prompt, seed = (int(v) for v in gold[f"{name}/cfg"])
    x, sd = make_clip_dense_inputs(width, L, N, od, seed)
    m = DenseLastBlock(width, od, prompt_length=prompt)
    m.load_state_dict(sd, strict=True)                                # the reference's own parameter names
    block_out, feats = m.cuda()(x.cuda())
    ref_v, ref_f = dense_last_block(sd, x, prompt, double=True)       # float64-accumulated oracle
    # hi+lo fp16 operand pairs (~2^-21 per product) accumulated by the tensor core in fp32 over up to 3 x 4 width terms: the
    # stated tolerance is rel-L2 <= 2e-5, max-abs <= 5e-5 of the output scale (measured at width 768: 7.9e-6 rel-L2; the
    # reference's own fp32-vs-fp64 difference is ~5e-7)
    for got, ref, key in ((block_out.cpu(), ref_v, "block_out"), (feats.cpu(), ref_f, "feats")):
        assert got.shape == ref.shape
        err = (got - ref).abs().max().item()
        assert err <= 5e-5 * max(1.0, ref.abs().max().item()), (key, err)
        assert rel_l2(got, ref) <= 2e-5, (key, rel_l2(got, ref))
        np.testing.assert_allclose(got.contiguous().flatten()[::53].numpy(), gold[f"{name}/{key}/sub"], rtol=0,
                                   atol=6e-5 * max(1.0, ref.abs().max().item()))


def test_clip_dense_vitl_batch_properties():
    """ViT-L/14 size (width 1024, 577 tokens, 4 images): token-wise independence (any subset of images gives the same rows,
    bit for bit) and agreement with the oracle on a sub-sampled set of tokens."""
    from cat_seg_b200.clip_dense import DenseLastBlock
    from cat_seg_b200.synth import make_clip_dense_inputs
    from oracle.clip_dense_oracle import dense_last_block
    x, sd = make_clip_dense_inputs(1024, 577, 4, 768, 3)
    m = DenseLastBlock(1024, 768)
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    _, feats = m(x.cuda(), want_block_out=False)
    _, feats1 = m(x[:, 1:2].contiguous().cuda(), want_block_out=False)
    assert torch.equal(feats[1:2], feats1)
    _, ref_f = dense_last_block(sd, x[:, 2:3], 0)
    assert (feats[2:3].cpu() - ref_f).abs().max().item() <= 1e-4 * max(1.0, ref_f.abs().max().item())


def test_clip_dense_rejects_bad_arguments():
    from cat_seg_b200.clip_dense import DenseLastBlock
    m = DenseLastBlock(64, 32)
    with pytest.raises(RuntimeError):
        m(torch.zeros(10, 2, 64))                                     # CPU tensor: no fallback
    with pytest.raises(ValueError):
        m.cuda()(torch.zeros(10, 2, 48, device="cuda"))
