"""Drop-in replacement for the reference ``Aggregator`` (the boundary operator).

Reference: ``cat_seg/modeling/transformer/model.py:558-725`` — built at
``cat_seg/modeling/transformer/cat_seg_predictor.py:97-113`` and called at ``:161`` as
``self.transformer(img_feats, text_feats, appearance_guidance)``.

Same constructor kwargs, same ``state_dict`` keys/shapes (a reference checkpoint under
``sem_seg_head.predictor.transformer.*`` loads with ``strict=True``), same forward signature and
output (``logits [B, T, 4H, 4W]`` fp32, ``-100`` for classes dropped by the top-``pad_len``
truncation).  All arithmetic happens in ``libcatseg_b200.so`` (hand-written sm_100a CUDA behind a C
ABI); PyTorch only owns the tensors and the stream.  CPU tensors are rejected: there is no
fallback path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib
from .config import AggregatorConfig
from .synth import param_shapes


def _shift_mask(H: int, W: int, window: int, shift: int) -> torch.Tensor:
    """The ``attn_mask`` buffer of the shifted block (model.py:161-183); kept only so that the
    state_dict matches the reference — the kernels derive the mask from index arithmetic."""
    band_h = torch.zeros(H, dtype=torch.long)
    band_h[H - window: H - shift] = 1
    band_h[H - shift:] = 2
    band_w = torch.zeros(W, dtype=torch.long)
    band_w[W - window: W - shift] = 1
    band_w[W - shift:] = 2
    ids = band_h[:, None] * 3 + band_w[None, :]
    ids = ids.reshape(H // window, window, W // window, window).permute(0, 2, 1, 3).reshape(-1, window * window)
    diff = ids[:, None, :] - ids[:, :, None]
    return torch.where(diff != 0, torch.tensor(-100.0), torch.tensor(0.0))


class _Node(nn.Module):
    """Anonymous container used to reproduce the reference's nested parameter names."""


class Aggregator(nn.Module):
    def __init__(self, text_guidance_dim=512, text_guidance_proj_dim=128, appearance_guidance_dim=512,
                 appearance_guidance_proj_dim=128, decoder_dims=(64, 32), decoder_guidance_dims=(256, 128),
                 decoder_guidance_proj_dims=(32, 16), num_layers=4, nheads=4, hidden_dim=128,
                 pooling_size=(6, 6), feature_resolution=(24, 24), window_size=12, attention_type="linear",
                 prompt_channel=1, pad_len=256, precision: str = "exact") -> None:
        super().__init__()
        if attention_type != "linear":
            raise NotImplementedError("attention_type='full' is dead code in every shipped config (SURVEY.md §2.1)")
        _lib.precision_mask(precision)      # validates
        self.cfg = AggregatorConfig(
            text_guidance_dim=text_guidance_dim, text_guidance_proj_dim=text_guidance_proj_dim,
            appearance_guidance_dim=appearance_guidance_dim, appearance_guidance_proj_dim=appearance_guidance_proj_dim,
            decoder_dims=tuple(decoder_dims), decoder_guidance_dims=tuple(decoder_guidance_dims),
            decoder_guidance_proj_dims=tuple(decoder_guidance_proj_dims), num_layers=num_layers, nheads=nheads,
            hidden_dim=hidden_dim, pooling_size=tuple(pooling_size), feature_resolution=tuple(feature_resolution),
            window_size=window_size, attention_type=attention_type, prompt_channel=prompt_channel, pad_len=pad_len)
        self.num_layers, self.hidden_dim, self.pad_len = num_layers, hidden_dim, pad_len
        self.precision = precision
        self._names: List[str] = []
        for name, shape in param_shapes(self.cfg).items():
            self._register(name, nn.Parameter(torch.zeros(shape), requires_grad=False))
            self._names.append(name)
        H, W = self.cfg.feature_resolution
        for l in range(num_layers):
            self._register(f"layers.{l}.swin_block.block_2.attn_mask",
                           _shift_mask(H, W, window_size, window_size // 2), buffer=True)
        self._handle: Optional[C.c_void_p] = None
        self._handle_device: Optional[torch.device] = None
        self._synced: Dict[str, tuple] = {}
        self._workspace: Optional[torch.Tensor] = None
        self._peer = None          # PeerExchange of the all-to-all class split
        self._vocab: Optional[torch.Tensor] = None        # [T,P,C] class embeddings of the current vocabulary
        self._vocab_pushed = False

    # ------------------------------------------------------------------ parameter tree
    def _register(self, dotted: str, value, buffer: bool = False) -> None:
        parts = dotted.split(".")
        node: nn.Module = self
        for p in parts[:-1]:
            if p not in node._modules:
                node.add_module(p, _Node())
            node = node._modules[p]
        if buffer:
            node.register_buffer(parts[-1], value)
        else:
            node.register_parameter(parts[-1], value)

    def _param(self, dotted: str) -> torch.Tensor:
        node = self
        for p in dotted.split("."):
            node = getattr(node, p)
        return node

    # ------------------------------------------------------------------ library handle
    def _lib_error(self, code: int) -> RuntimeError:
        lib = _lib.load()
        msg = lib.catseg_last_error(self._handle).decode() if self._handle else lib.catseg_last_error(None).decode()
        return RuntimeError(f"catseg_b200 error {code}: {msg}")

    def _ensure_handle(self, device: torch.device) -> None:
        lib = _lib.load()
        if self._handle is not None and self._handle_device == device:
            return
        if self._handle is not None:
            lib.catseg_destroy(self._handle)
            self._handle = None
        c = self.cfg
        cc = _lib.CatsegConfig(
            c.text_guidance_dim, c.text_guidance_proj_dim, c.appearance_guidance_dim, c.appearance_guidance_proj_dim,
            (C.c_int32 * 2)(*c.decoder_dims), (C.c_int32 * 2)(*c.decoder_guidance_dims),
            (C.c_int32 * 2)(*c.decoder_guidance_proj_dims), c.num_layers, c.nheads, c.hidden_dim,
            (C.c_int32 * 2)(*c.pooling_size), (C.c_int32 * 2)(*c.feature_resolution), c.window_size, 0,
            c.prompt_channel, c.pad_len, _lib.precision_mask(self.precision))
        h = C.c_void_p()
        with torch.cuda.device(device):
            rc = lib.catseg_create(C.byref(cc), C.byref(h))
        if rc != 0:
            raise RuntimeError(f"catseg_create failed ({rc}): {lib.catseg_last_error(None).decode()}")
        self._handle, self._handle_device = h, device
        self._synced = {}
        self._vocab_pushed = False
        n = lib.catseg_num_params(h)
        names = [lib.catseg_param_name(h, i).decode() for i in range(n)]
        if names != self._names:
            raise RuntimeError("parameter table of libcatseg_b200.so does not match the module")

    def sync_weights(self, force: bool = False) -> None:
        """Pushes changed parameters to the library and re-packs them into kernel layouts."""
        lib = _lib.load()
        dirty = False
        for name in self._names:
            p = self._param(name)
            key = (p.data_ptr(), p._version, p.device)
            if not force and self._synced.get(name) == key:
                continue
            src = p.detach().to(dtype=torch.float32).contiguous()
            rc = lib.catseg_set_param(self._handle, name.encode(), C.c_void_p(src.data_ptr()), src.numel(),
                                      1 if src.is_cuda else 0)
            if rc != 0:
                raise self._lib_error(rc)
            self._synced[name] = key
            dirty = True
        if dirty:
            rc = lib.catseg_finalize_params(self._handle, C.c_void_p(torch.cuda.current_stream().cuda_stream))
            if rc != 0:
                raise self._lib_error(rc)

    def set_vocabulary(self, text_feats: Optional[torch.Tensor]) -> None:
        """Registers the class embeddings of a vocabulary once (``[T,P,C]`` or the predictor's cached ``[T,1,C]``,
        cat_seg_predictor.py:190-224); ``forward(img_feats, None, guidance)`` then skips the per-call text work
        (normalisation, text-guidance projection, guidance half of the class-attention q/k).  ``None`` forgets it."""
        if text_feats is None:
            self._vocab = None
        else:
            t = text_feats.detach()
            if t.dim() == 4:
                t = t[0]
            if t.dim() != 3 or t.shape[1] != self.cfg.prompt_channel or t.shape[2] != self.cfg.text_guidance_dim:
                raise ValueError(f"vocabulary must be [T,{self.cfg.prompt_channel},{self.cfg.text_guidance_dim}], got {tuple(text_feats.shape)}")
            self._vocab = t.to(dtype=torch.float32).contiguous()
        self._vocab_pushed = False

    def _push_vocabulary(self, dev) -> int:
        lib = _lib.load()
        if self._vocab is None:
            raise RuntimeError("text_feats is None and no vocabulary is set (Aggregator.set_vocabulary)")
        if not self._vocab_pushed:
            v = self._vocab.to(dev)
            rc = lib.catseg_set_vocabulary(self._handle, C.c_void_p(v.data_ptr()), v.shape[0],
                                           C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
            if rc != 0:
                raise self._lib_error(rc)
            self._vocab, self._vocab_pushed = v, True         # keeps the source alive until the copy has run
        return self._vocab.shape[0]

    def kept_classes(self, T: int) -> int:
        return self.pad_len if (self.pad_len > 0 and T > self.pad_len) else T

    # ------------------------------------------------------------------ forward
    def _check(self, img_feats, text_feats, guidance):
        if not (torch.is_tensor(img_feats) and img_feats.is_cuda):
            raise RuntimeError("catseg_b200.Aggregator runs on CUDA tensors only (no CPU fallback)")
        c = self.cfg
        B, Cc, H, W = img_feats.shape
        if (H, W) != tuple(c.feature_resolution):
            raise ValueError(f"img_feats grid {(H, W)} != feature_resolution {c.feature_resolution}")
        if text_feats is None:
            if self._vocab is None:
                raise RuntimeError("text_feats is None and no vocabulary is set (Aggregator.set_vocabulary)")
            if Cc != c.text_guidance_dim:
                raise ValueError(f"img_feats channels {Cc} != text_guidance_dim {c.text_guidance_dim}")
        elif text_feats.dim() != 4 or text_feats.shape[0] != B or text_feats.shape[2] != c.prompt_channel \
                or text_feats.shape[3] != c.text_guidance_dim or Cc != c.text_guidance_dim:
            raise ValueError(f"text_feats {tuple(text_feats.shape)} / img_feats {tuple(img_feats.shape)} mismatch")
        if len(guidance) != 3:
            raise ValueError("appearance_guidance must hold 3 tensors [res3, res4, res5]")
        exp = [(B, c.appearance_guidance_dim, H, W), (B, c.decoder_guidance_dims[0], 2 * H, 2 * W),
               (B, c.decoder_guidance_dims[1], 4 * H, 4 * W)]
        for g, e in zip(guidance, exp):
            if tuple(g.shape) != e:
                raise ValueError(f"appearance guidance shape {tuple(g.shape)} != {e}")
        return B, (text_feats.shape[1] if text_feats is not None else self._vocab.shape[0]), H, W

    @staticmethod
    def _f32c(t: torch.Tensor) -> torch.Tensor:
        return t.detach().to(dtype=torch.float32).contiguous()

    @torch.no_grad()
    def forward(self, img_feats: torch.Tensor, text_feats: torch.Tensor, appearance_guidance: Sequence[torch.Tensor],
                taps: Optional[Sequence[str]] = None):
        """img_feats (B,C,H,W); text_feats (B,T,P,C); appearance_guidance: 3 tensors (model.py:683-689)."""
        B, T, H, W = self._check(img_feats, text_feats, appearance_guidance)
        lib = _lib.load()
        dev = img_feats.device
        with torch.cuda.device(dev):
            self._ensure_handle(dev)
            self.sync_weights()
            img = self._f32c(img_feats)
            if text_feats is None:
                self._push_vocabulary(dev)
                text_ptr = C.c_void_p(None)
            else:
                text = self._f32c(text_feats)
                text_ptr = C.c_void_p(text.data_ptr())
            g = [self._f32c(x) for x in appearance_guidance]
            need = lib.catseg_workspace_bytes(self._handle, B, T)
            if self._workspace is None or self._workspace.numel() < need or self._workspace.device != dev:
                self._workspace = None
                self._workspace = torch.empty(need, dtype=torch.uint8, device=dev)
            logits = torch.empty(B, T, 4 * H, 4 * W, dtype=torch.float32, device=dev)
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            args = [self._handle, C.c_void_p(img.data_ptr()), text_ptr] + [C.c_void_p(t.data_ptr()) for t in (g[0], g[1], g[2], logits)] + \
                   [C.c_void_p(self._workspace.data_ptr()), self._workspace.numel(), B, T]
            if taps is None:
                rc = lib.catseg_forward(*args, stream)
                if rc != 0:
                    raise self._lib_error(rc)
                return logits
            tap_struct, out = self._make_taps(taps, B, T, H, W, dev)
            rc = lib.catseg_forward_taps(*args, C.byref(tap_struct), stream)
            if rc != 0:
                raise self._lib_error(rc)
            return logits, out

    def forward_class_sharded(self, img_feats: torch.Tensor, text_feats: torch.Tensor,
                              appearance_guidance: Sequence[torch.Tensor], group=None, exchange: str = "allreduce",
                              gather: bool = True, barrier: str = "device") -> torch.Tensor:
        """Class-sharded forward over a torch.distributed process group (NCCL): every rank passes the SAME inputs,
        computes the kept classes [r*Te/world, (r+1)*Te/world) and receives the full [B,T,4H,4W] logits.
        exchange="allreduce": activations stay class-sharded; the only exchange inside the path is one all-reduce (sum) of
        the linear-attention state per class layer (model.py:282-283; SURVEY.md 8e "cheaper equivalent"), issued from the
        library through a callback on the current stream.
        exchange="alltoall": north_star's prescription: the residual stream is transposed class-sharded
        <-> pixel-sharded around each class layer by kernels that store straight into the peers' buffers over NVLink
        (CUDA IPC mappings, ``PeerExchange``); those stores are ordered by a flag barrier through the same peer memory
        (barrier="device", one single-warp kernel) or by a one-element all-reduce of the group (barrier="collective").
        The local logit planes are all-gathered and scattered to their class ids afterwards (gather=False returns the
        compact local planes and the kept-class list instead)."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        B, T, H, W = self._check(img_feats, text_feats, appearance_guidance)
        Te = self.kept_classes(T)
        if Te % world:
            raise RuntimeError(f"kept classes ({Te}) must be a multiple of the shard group size ({world})")
        if exchange not in ("allreduce", "alltoall"):
            raise ValueError(f"exchange must be 'allreduce' or 'alltoall', got {exchange!r}")
        lib = _lib.load()
        dev = img_feats.device
        with torch.cuda.device(dev):
            self._ensure_handle(dev)
            self.sync_weights()
            img, text = self._f32c(img_feats), self._f32c(text_feats)
            g = [self._f32c(x) for x in appearance_guidance]
            need = lib.catseg_workspace_bytes(self._handle, B, T)
            if self._workspace is None or self._workspace.numel() < need or self._workspace.device != dev:
                self._workspace = None
                self._workspace = torch.empty(need, dtype=torch.uint8, device=dev)
            ws = self._workspace
            local = torch.empty(B, Te // world, 4 * H, 4 * W, dtype=torch.float32, device=dev)
            kept = torch.empty(B, Te, dtype=torch.int32, device=dev)
            errors = []
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            ptrs = [C.c_void_p(t.data_ptr()) for t in (img, text, g[0], g[1], g[2], local, kept)]

            if exchange == "allreduce":
                def _allreduce(_ctx, buf, count, _stream):      # called by the library between class_state and class_apply
                    try:
                        off = int(buf) - ws.data_ptr()
                        dist.all_reduce(ws[off:off + 4 * count].view(torch.float32), op=dist.ReduceOp.SUM, group=group)
                        return 0
                    except Exception as e:                      # never let an exception cross the C ABI
                        errors.append(e)
                        return 1

                cb = _lib.ALLREDUCE_FN(_allreduce)
                rc = lib.catseg_forward_class_sharded(self._handle, *ptrs, C.c_void_p(ws.data_ptr()), ws.numel(), B, T, rank, world,
                                                      C.cast(cb, C.c_void_p), None, stream)
            else:
                nbytes = lib.catseg_exchange_buffer_bytes(self._handle, B, T, world)
                lbytes = lib.catseg_exchange_logits_bytes(self._handle, B, T) if gather else 0
                # guidance sharded by image: needs the library's own flag barrier (it also runs on the internal stream)
                gbytes = lib.catseg_exchange_guidance_bytes(self._handle, B, T) if (barrier == "device" and B % world == 0) else 0
                if (self._peer is None or self._peer.nbytes < nbytes or self._peer.lbytes < lbytes or self._peer.gbytes < gbytes
                        or self._peer.group is not group):
                    if self._peer is not None:
                        self._peer.close()
                    self._peer = PeerExchange(nbytes, dev, group, logits_bytes=lbytes, guidance_bytes=gbytes)
                peer = self._peer

                def _barrier(_ctx, _stream):                    # orders the peer stores of the transposition kernels
                    try:
                        dist.all_reduce(peer.flag, group=group)
                        return 0
                    except Exception as e:
                        errors.append(e)
                        return 1

                if barrier not in ("device", "collective"):
                    raise ValueError(f"barrier must be 'device' or 'collective', got {barrier!r}")
                cb = _lib.BARRIER_FN(_barrier)
                rc = lib.catseg_forward_class_sharded_a2a(self._handle, *ptrs, C.c_void_p(ws.data_ptr()), ws.numel(), B, T, rank, world,
                                                          peer.xb, peer.pb, peer.nbytes, peer.lb if gather else None,
                                                          peer.gb if gbytes else None, peer.gbytes,
                                                          C.cast(cb, C.c_void_p) if barrier == "collective" else None, None, stream)
                if rc == 0 and not errors and gather:
                    # peer-direct: every rank's buffer already holds the complete [B,T,4H,4W] logits (a view of the peer
                    # buffer: it is overwritten by the next call)
                    return peer.logits_view(B, T, 4 * H, 4 * W)
            if errors:
                raise errors[0]
            if rc != 0:
                raise self._lib_error(rc)
            if not gather:
                return local, kept
            gathered = torch.empty(world, *local.shape, dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(gathered, local, group=group)
            return assemble_class_sharded(gathered, kept, T)

    def class_shard_healthy(self, B: int, T: int, group=None) -> bool:
        """False if one of this rank's flag barriers of the all-to-all class split ever timed out (a peer died or stalled for
        seconds): every result since then is invalid.  Synchronises the device; call it at the end of a job, not per step."""
        import torch.distributed as dist
        if self._peer is None:
            return True
        out = C.c_int(0)
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        with torch.cuda.device(self._peer.device):
            torch.cuda.synchronize(self._peer.device)
            rc = _lib.load().catseg_exchange_timed_out(self._handle, C.c_void_p(self._peer.pb[rank]),
                                                       C.c_void_p(self._peer.gb[rank]) if self._peer.gbytes else None, B, T, world,
                                                       C.byref(out))
        if rc != 0:
            raise self._lib_error(rc)
        return out.value == 0

    def _make_taps(self, names, B, T, H, W, dev):
        c = self.cfg
        Te, HW, hid = self.kept_classes(T), H * W, c.hidden_dim
        shapes = {
            "corr": ((B, T, c.prompt_channel, HW), torch.float32), "classes": ((B, Te), torch.int32),
            "embed": ((B, Te, HW, hid), torch.float32),
            "app_guidance": ((B, HW, c.appearance_guidance_proj_dim), torch.float32),
            "text_guidance": ((B, Te, c.text_guidance_proj_dim), torch.float32),
            "dec_guidance0": ((B, 4 * HW, c.decoder_guidance_proj_dims[0]), torch.float32),
            "dec_guidance1": ((B, 16 * HW, c.decoder_guidance_proj_dims[1]), torch.float32),
            "up1": ((B, Te, 4 * HW, c.decoder_dims[0]), torch.float32),
            "up2": ((B, Te, 16 * HW, c.decoder_dims[1]), torch.float32),
        }
        for l in range(c.num_layers):
            for k in ("swin_l%d_b1", "swin_l%d_b2", "class_l%d"):
                shapes[k % l] = ((B, Te, HW, hid), torch.float32)
        ts = _lib.CatsegTaps()
        out = {}
        for n in names:
            shape, dt = shapes[n]
            t = torch.zeros(shape, dtype=dt, device=dev)
            out[n] = t
            if n.startswith("swin_l"):
                l = int(n[6])
                getattr(ts, "swin_b1" if n.endswith("b1") else "swin_b2")[l] = t.data_ptr()
            elif n.startswith("class_l"):
                ts.class_out[int(n[7:])] = t.data_ptr()
            else:
                setattr(ts, n, t.data_ptr())
        return ts, out

    # ------------------------------------------------------------------ profiling hooks (bench.py)
    def set_profiling(self, enable: bool) -> None:
        _lib.load().catseg_set_profiling(self._handle, 1 if enable else 0)

    def stage_times(self, reset: bool = True):
        ms = (C.c_float * len(_lib.STAGES))()
        calls = C.c_int(0)
        rc = _lib.load().catseg_stage_times(self._handle, ms, C.byref(calls), 1 if reset else 0)
        if rc != 0:
            raise self._lib_error(rc)
        return {s: ms[i] for i, s in enumerate(_lib.STAGES)}, calls.value

    def last_launch_count(self) -> int:
        return _lib.load().catseg_last_launch_count(self._handle)

    def __del__(self):
        try:
            if self._handle is not None and _lib._LIB is not None:
                _lib._LIB.catseg_destroy(self._handle)
        except Exception:
            pass


class PeerExchange:
    """Two peer-visible buffers per rank (class-sharded X, pixel-sharded P) for the all-to-all class split: allocated by the
    library with cudaMalloc, published as CUDA IPC handles over the process group, mapped on every other rank of the node.
    ``xb`` / ``pb`` are the ctypes pointer arrays catseg_forward_class_sharded_a2a takes (entry r = rank r's buffer)."""

    def __init__(self, nbytes: int, device: torch.device, group=None, logits_bytes: int = 0, guidance_bytes: int = 0):
        import torch.distributed as dist
        lib = _lib.load()
        self.nbytes, self.lbytes, self.gbytes = int(nbytes), int(logits_bytes), int(guidance_bytes)
        self.group, self.device = group, device
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        self._lib, self._own, self._opened = lib, [], []
        handles = []
        # buffers: class-sharded X, pixel-sharded P, full logits (optional), guidance (optional); absent ones are 256 bytes
        sizes = [self.nbytes, self.nbytes, max(self.lbytes, 256), max(self.gbytes, 256)]
        with torch.cuda.device(device):
            for nb in sizes:
                p = C.c_void_p()
                if lib.catseg_peer_alloc(nb, C.byref(p)) != 0:
                    raise RuntimeError("catseg_peer_alloc failed: " + lib.catseg_last_error(None).decode())
                self._own.append(p.value)
                hbuf = C.create_string_buffer(64)
                if lib.catseg_peer_export(C.c_void_p(p.value), hbuf) != 0:
                    raise RuntimeError("catseg_peer_export failed: " + lib.catseg_last_error(None).decode())
                handles.append(hbuf.raw)
            everyone = [None] * world
            dist.all_gather_object(everyone, handles, group=group)
            self.xb, self.pb, self.lb, self.gb = ((C.c_void_p * world)() for _ in range(4))
            for r in range(world):
                for k, arr in enumerate((self.xb, self.pb, self.lb, self.gb)):
                    if r == rank:
                        arr[r] = self._own[k]
                    else:
                        q = C.c_void_p()
                        if lib.catseg_peer_open(everyone[r][k], C.byref(q)) != 0:
                            raise RuntimeError(f"catseg_peer_open (rank {r}) failed: " + lib.catseg_last_error(None).decode())
                        self._opened.append(q.value)
                        arr[r] = q.value
            self.flag = torch.zeros(1, dtype=torch.float32, device=device)

    def logits_view(self, *shape) -> torch.Tensor:
        """This rank's full-logits peer buffer as a torch tensor (no copy; CUDA array interface)."""
        n = 1
        for d in shape:
            n *= d
        assert 4 * n <= self.lbytes

        class _Holder:
            pass

        holder = _Holder()
        holder.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (self._own[2], False), "version": 2}
        holder.owner = self
        return torch.as_tensor(holder, device=self.device)

    def close(self):
        if self._lib is None:
            return
        import torch.distributed as dist
        try:
            torch.cuda.synchronize(self.device)
            with torch.cuda.device(self.device):
                for q in self._opened:
                    self._lib.catseg_peer_close(C.c_void_p(q))
                self._opened = []
                if dist.is_initialized():
                    dist.barrier(self.group)          # nobody frees while a peer still has the mapping open
                for p in self._own:
                    self._lib.catseg_peer_free(C.c_void_p(p))
                self._own = []
        finally:
            self._lib = None


def assemble_class_sharded(gathered: torch.Tensor, kept: torch.Tensor, T: int) -> torch.Tensor:
    """gathered [world, B, Te/world, h, w]: rank r's planes are the kept classes kept[:, r*Te/world:(r+1)*Te/world];
    kept [B, Te] class ids.  Returns [B, T, h, w] with -100 for classes that were not kept (model.py:721-724)."""
    world, B, tl, h, w = gathered.shape
    if gathered.is_cuda:                              # one pass in the CUDA library: copy a gathered plane or fill -100
        lib = _lib.load()
        g = gathered.contiguous()
        k = kept.to(device=g.device, dtype=torch.int32).contiguous()
        out = torch.empty(B, T, h, w, dtype=torch.float32, device=g.device)
        pos = torch.empty(B * T, dtype=torch.int32, device=g.device)
        with torch.cuda.device(g.device):
            rc = lib.catseg_assemble_class_sharded(C.c_void_p(g.data_ptr()), C.c_void_p(k.data_ptr()), C.c_void_p(pos.data_ptr()),
                                                   C.c_void_p(out.data_ptr()), world, B, tl, T, h * w,
                                                   C.c_void_p(torch.cuda.current_stream(g.device).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"catseg_assemble_class_sharded failed ({rc}): {lib.catseg_last_error(None).decode()}")
        return out
    planes = gathered.permute(1, 0, 2, 3, 4).reshape(B, world * tl, h, w)
    out = torch.full((B, T, h, w), -100.0, dtype=gathered.dtype, device=gathered.device)
    out[torch.arange(B, device=gathered.device)[:, None], kept.long()] = planes
    return out
