"""GPU (-m gpu): the class-sharded mode on ONE GPU -- two processes share cuda:0.

NCCL refuses two ranks on one device, so the process group is gloo (its CUDA all_reduce is the state reduction / the
barrier callback); everything else is the product path: both ranks run the CUDA kernels, map each other's exchange buffers
through CUDA IPC (catseg_peer_export / catseg_peer_open) and store into them with the transposition kernels.  The driver's
multi-GPU bench covers the NCCL / NVLink side of the same code.
"""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, q, model: str, B: int, T: int):
    try:
        import torch.distributed as dist
        from cat_seg_b200.aggregator import Aggregator, assemble_class_sharded
        from cat_seg_b200.config import vitb, vitl
        from cat_seg_b200.synth import make_inputs, make_state_dict
        os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        torch.cuda.set_device(0)
        cfg = vitb() if model == "vitb" else vitl()
        sd = make_state_dict(cfg, 0)
        img, text, g = make_inputs(cfg, B, T, 3)
        m = Aggregator(**cfg.ctor_kwargs(), precision="precise")
        m.load_state_dict(sd, strict=False)
        m = m.cuda()
        a = (img.cuda(), text.cuda(), [x.cuda() for x in g])
        out = {}
        for exchange in ("alltoall", "alltoall/collective", "allreduce"):
            local, kept = m.forward_class_sharded(*a, exchange=exchange.split("/")[0], gather=False,
                                                  barrier="collective" if exchange.endswith("collective") else "device")
            torch.cuda.synchronize()
            parts = [torch.empty_like(local.cpu()) for _ in range(world)]
            dist.all_gather(parts, local.cpu())
            out[exchange] = assemble_class_sharded(torch.stack(parts), kept.cpu(), T)
        # a second call re-uses the mapped buffers (hazard check: peers may still be reading them)
        local2, kept2 = m.forward_class_sharded(*a, exchange="alltoall", gather=False)
        torch.cuda.synchronize()
        parts = [torch.empty_like(local2.cpu()) for _ in range(world)]
        dist.all_gather(parts, local2.cpu())
        again = assemble_class_sharded(torch.stack(parts), kept2.cpu(), T)
        # peer-direct result: every rank's buffer holds the full logits (head kernel stores into all ranks, no all-gather)
        full = m.forward_class_sharded(*a, exchange="alltoall").clone()
        torch.cuda.synchronize()
        full2 = m.forward_class_sharded(*a, exchange="alltoall").clone()     # buffers re-used
        torch.cuda.synchronize()
        res = None
        same_full = bool(torch.equal(full, full2))
        if rank != 0:
            ref_r = m(*a)
            same_full = same_full and bool(torch.equal(full, ref_r))           # every rank, not only rank 0, has the result
        flags = [None] * world
        dist.all_gather_object(flags, same_full)
        if rank == 0:
            ref = m(*a).cpu()
            out["peer_direct"] = full.cpu()
            res = {k: (v - ref).abs().max().item() for k, v in out.items()}
            res["mask_equal"] = all(bool(((v == -100.0) == (ref == -100.0)).all()) for v in out.values())
            res["again_equal"] = bool(torch.equal(again, out["alltoall"])) and all(flags)
            # the CUDA assembly kernel (copy a gathered plane / fill -100) against the torch restatement
            stacked = torch.stack(parts)
            res["assemble_equal"] = bool(torch.equal(assemble_class_sharded(stacked.cuda(), kept2, T).cpu(),
                                                     assemble_class_sharded(stacked, kept2.cpu(), T)))
        assert m.class_shard_healthy(B, T), "a flag barrier timed out"
        dist.barrier()
        if m._peer is not None:
            m._peer.close()
        q.put((rank, res, None))
        dist.destroy_process_group()
    except Exception as e:          # noqa: BLE001 -- reported to the parent
        import traceback
        q.put((rank, None, traceback.format_exc()))
        raise e


@pytest.mark.parametrize("model,B,T", [("vitb", 2, 8), ("vitl", 1, 300)])
def test_class_sharded_two_ranks_one_gpu(model, B, T):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, model, B, T)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    for rank, res, err in results:
        assert err is None, f"rank {rank}:\n{err}"
    res = [r for rank, r, _ in results if rank == 0][0]
    # all-to-all: every kernel sees the same operands in the same order as the unsharded run -> bit exact
    assert res["alltoall"] == 0.0 and res["alltoall/collective"] == 0.0 and res["peer_direct"] == 0.0, res
    # all-reduce: only the fp32 summation order of the linear-attention state differs
    assert res["allreduce"] <= 2e-5, res
    assert res["mask_equal"] and res["again_equal"] and res["assemble_equal"], res
