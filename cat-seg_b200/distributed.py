"""One-process-per-GPU plumbing (torch.distributed; NCCL on GPUs, gloo in the CPU tests).

The reference's only multi-GPU inference mode is data parallel by image: Detectron2 ``launch`` starts one
process per GPU, the dataset is sharded per rank and predictions / confusion matrices are all-gathered
(train_net.py:317-324, SemSegEvaluator(distributed=True) train_net.py:103-107).  The hot path itself has no
collective: images (or sliding windows) are independent, so ranks take contiguous shards of the batch.
"""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def env_ranks() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_from_env(backend: str = "nccl", device: torch.device | None = None) -> Tuple[int, int, int]:
    """Initialises the default process group from RANK / WORLD_SIZE / MASTER_* (torchrun)."""
    rank, local, world = env_ranks()
    if world > 1 and not dist.is_initialized():
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, local, world


def shard_range(n: int, rank: int, world: int) -> range:
    """Contiguous, balanced shard of n items: the first n % world ranks get one extra item."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def max_over_ranks(values: Sequence[float], device: torch.device) -> List[float]:
    """Element-wise max over ranks (device-timed intervals are reported as the slowest rank's)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def gather_in_rank_order(local: torch.Tensor) -> List[torch.Tensor]:
    """All-gathers per-rank results (e.g. label maps) as a list ordered by rank; shards may differ in size."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return [local]
    world = dist.get_world_size()
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    m = int(max(int(s) for s in sizes))
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return [o[: int(s)] for o, s in zip(out, sizes)]


def barrier() -> None:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
