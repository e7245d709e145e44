"""Host-side feeding of the Aggregator: pinned host inputs are uploaded on a copy stream into one of ``depth``
device slots while the previous batch computes (the data-loader prefetch every serving loop of the reference has
through Detectron2's loader; here it is explicit because the boundary call takes device tensors).

    pipe = HostPipeline(model, device)
    t = pipe.upload(host_tensors)            # async H2D on the copy stream
    y = pipe.run(t)                          # boundary call on the current stream once the upload has landed

Nothing here computes: it orders two CUDA streams with events.
"""
from typing import List, Sequence

import torch


class HostPipeline:
    def __init__(self, model, device, depth: int = 2):
        self.model, self.device, self.depth = model, device, depth
        self.copy_stream = torch.cuda.Stream(device)
        self.slots: List = [None] * depth
        self.uploaded = [torch.cuda.Event() for _ in range(depth)]
        self.consumed = [None] * depth
        self.n = 0

    def upload(self, host: Sequence[torch.Tensor]) -> int:
        """host = (img_feats, text_feats, guidance_1, guidance_2) pinned CPU tensors.  Returns a ticket."""
        s = self.n % self.depth
        self.n += 1
        if self.slots[s] is None:
            self.slots[s] = [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in host]
        with torch.cuda.stream(self.copy_stream):
            if self.consumed[s] is not None:
                self.copy_stream.wait_event(self.consumed[s])      # the batch that used this slot has finished
            for d, h in zip(self.slots[s], host):
                d.copy_(h, non_blocking=True)
            self.uploaded[s].record(self.copy_stream)
        return s

    def upload_sharded(self, host: Sequence[torch.Tensor], rank: int, world: int, group=None) -> int:
        """Class-sharded serving: every rank needs the SAME inputs.  Instead of ``world`` identical PCIe uploads, rank r
        uploads rows [r B/world, (r+1) B/world) of every tensor and the shards are all-gathered over NVLink (in place, on the
        copy stream): 1/world of the host traffic per rank.  All tensors must have the batch as their leading dimension,
        divisible by ``world``."""
        import torch.distributed as dist
        s = self.n % self.depth
        self.n += 1
        if self.slots[s] is None:
            self.slots[s] = [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in host]
        with torch.cuda.stream(self.copy_stream):
            if self.consumed[s] is not None:
                self.copy_stream.wait_event(self.consumed[s])
            for d, h in zip(self.slots[s], host):
                n = d.shape[0] // world
                if n * world != d.shape[0]:
                    raise ValueError(f"batch {d.shape[0]} is not divisible by the group size {world}")
                shard = d[rank * n:(rank + 1) * n]
                shard.copy_(h[rank * n:(rank + 1) * n], non_blocking=True)
                dist.all_gather_into_tensor(d, shard, group=group)          # in place: shard is d's own slice
            self.uploaded[s].record(self.copy_stream)
        return s

    def run(self, ticket: int) -> torch.Tensor:
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self.uploaded[ticket])
        a, b, c, d = self.slots[ticket]
        y = self.model(a, b, [a, c, d])
        ev = torch.cuda.Event()
        ev.record(cur)
        self.consumed[ticket] = ev
        return y


class GraphRunner:
    """Replays the boundary call from a CUDA graph (one capture per input shape): the ~77 kernel launches of a forward
    become one graph launch, which matters when the batch is small and the call is launch-bound (cfg1: one image, 20
    classes).  Inputs are copied into static device buffers; the returned logits tensor is reused by every replay.

        run = GraphRunner(model, img, text, [g0, g1, g2])     # warms up, captures
        logits = run(img2, text2, [g0b, g1b, g2b])             # same shapes
    """

    def __init__(self, model, img_feats, text_feats, appearance_guidance, warmup: int = 2, call=None):
        """call(model, img, text, guidance) -> tensor: the entry point to capture (default: the boundary call itself), e.g.
        ``lambda m, a, b, g: m.forward_class_sharded(a, b, g, exchange="alltoall")`` -- the all-to-all class split has no host
        callback on its data path (peer stores + flag barriers), so every rank can replay its own graph."""
        self.model = model
        if call is None:
            call = lambda m, a, b, g: m(a, b, g)      # noqa: E731
        self.static_in = [img_feats.clone(), text_feats.clone()] + [g.clone() for g in appearance_guidance]
        side = torch.cuda.Stream(img_feats.device)
        side.wait_stream(torch.cuda.current_stream(img_feats.device))
        with torch.cuda.stream(side):                       # warm-up outside capture: lazy handle/stream/attribute setup
            for _ in range(max(1, warmup)):
                call(model, self.static_in[0], self.static_in[1], self.static_in[2:])
        torch.cuda.current_stream(img_feats.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = call(model, self.static_in[0], self.static_in[1], self.static_in[2:])

    def __call__(self, img_feats, text_feats, appearance_guidance):
        for dst, src in zip(self.static_in, [img_feats, text_feats] + list(appearance_guidance)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out
