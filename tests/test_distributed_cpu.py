"""CPU: the N > 1 host logic on a world_size-2 gloo group (spawned processes, 127.0.0.1 rendezvous)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cat_seg_b200.distributed import gather_in_rank_order, init_from_env, max_over_ranks, shard_range


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, _, w = init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    # images are sharded contiguously; every image belongs to exactly one rank
    n_img = 5
    mine = shard_range(n_img, rank, world)
    labels = torch.stack([torch.full((4, 4), i, dtype=torch.int32) for i in mine]) if len(mine) else torch.zeros(0, 4, 4, dtype=torch.int32)
    parts = gather_in_rank_order(labels)
    full = torch.cat(parts)
    ok_gather = full.shape[0] == n_img and all(int(full[i, 0, 0]) == i for i in range(n_img))
    # timing: max over ranks
    mx = max_over_ranks([10.0 + rank, 3.0 - rank], torch.device("cpu"))
    dist.barrier()
    q.put((rank, list(mine), ok_gather, mx))
    dist.destroy_process_group()


def test_world2_gloo_sharding_gather_and_max():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 1, 2] and res[1][1] == [3, 4]
    assert all(r[2] for r in res)
    assert res[0][3] == [11.0, 3.0] and res[1][3] == [11.0, 3.0]


def test_shard_range_covers_everything():
    for n in (0, 1, 5, 16, 17):
        for world in (1, 2, 3, 8):
            seen = [i for r in range(world) for i in shard_range(n, r, world)]
            assert seen == list(range(n))
            sizes = [len(shard_range(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_assemble_class_sharded_scatter():
    """Host logic of the class-sharded mode: rank-major planes go back to their kept class ids, the rest is -100."""
    import torch
    from cat_seg_b200.aggregator import assemble_class_sharded
    world, B, tl, T = 2, 2, 3, 10
    kept = torch.tensor([[0, 2, 3, 5, 8, 9], [1, 2, 4, 6, 7, 9]], dtype=torch.int32)
    gathered = torch.arange(world * B * tl * 4, dtype=torch.float32).reshape(world, B, tl, 2, 2)
    out = assemble_class_sharded(gathered, kept, T)
    assert out.shape == (B, T, 2, 2)
    for b in range(B):
        for r in range(world):
            for j in range(tl):
                assert torch.equal(out[b, kept[b, r * tl + j]], gathered[r, b, j])
        dropped = sorted(set(range(T)) - set(kept[b].tolist()))
        assert bool((out[b, dropped] == -100.0).all())
