"""Multi-GPU (torchrun): class-sharded forward vs the single-GPU forward on identical inputs.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \\
        tools/check_class_sharded.py [--model vitl --B 2 --T 847 --precision fast --time 5]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from cat_seg_b200.aggregator import Aggregator  # noqa: E402
from cat_seg_b200.config import vitb, vitl  # noqa: E402
from cat_seg_b200.synth import make_inputs, make_state_dict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="fast")
    ap.add_argument("--model", default="vitl")
    ap.add_argument("--B", type=int, default=2)
    ap.add_argument("--T", type=int, default=847)
    ap.add_argument("--time", type=int, default=0)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = (vitb if a.model == "vitb" else vitl)()
    sd = make_state_dict(cfg, 0)
    img, text, g = make_inputs(cfg, a.B, a.T, seed=5, same_text=True)       # identical on every rank
    m = Aggregator(**cfg.ctor_kwargs(), precision=a.precision)
    m.load_state_dict(sd, strict=False)
    m = m.cuda()
    cu = (img.cuda(), text.cuda(), [x.cuda() for x in g])
    ref = m(*cu)
    y = m.forward_class_sharded(*cu)
    torch.cuda.synchronize()
    kept = ref != -100.0
    err = (y - ref).abs().max().item()
    ok_mask = bool(((y != -100.0) == kept).all())
    t_full = t_shard = None
    if a.time:
        def timed(fn):
            for _ in range(2):
                fn()
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.time):
                fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / a.time
        t_full = timed(lambda: m(*cu))
        t_shard = timed(lambda: m.forward_class_sharded(*cu))
    stats = torch.tensor([err, float(ok_mask)], device="cuda")
    dist.all_reduce(stats, op=dist.ReduceOp.MAX if True else None)
    if rank == 0:
        print(f"class-sharded over {world} GPUs ({a.model}, B={a.B}, T={a.T}, precision {a.precision}): "
              f"max-abs vs single-GPU forward {err:.3e}, -100 mask equal {ok_mask}"
              + (f"; single GPU {t_full:.3f} ms/call, class-sharded {t_shard:.3f} ms/call ({t_full / t_shard:.2f}x)" if a.time else ""))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
