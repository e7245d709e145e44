#!/usr/bin/env python
"""Benchmark of the CAT-Seg hot path (Aggregator.forward: cost volume -> logits) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg4]

A step = one boundary call on one batch of synthetic inputs of the workload
(default: BASELINE.json's headline config, ViT-L/14 A-847, 16 images of 336x336 per GPU -> cfg4).
N > 1 (launched by torchrun): images are sharded over ranks (the reference's own data-parallel
inference mode); no data-path collective, weak scaling.
Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

from cat_seg_b200.config import BENCH_CONFIGS, vitb, vitl  # noqa: E402
from cat_seg_b200.synth import make_inputs, make_state_dict  # noqa: E402

METRIC, UNIT = "vitl_a847_aggregator_images_per_sec", "images/s"


# ----------------------------------------------------------------------------- work model
def stage_flops(cfg, B, T):
    """Reference-algorithm FLOPs (2 x MACs of model.py as written) per boundary call, per stage."""
    Te = min(T, cfg.pad_len) if cfg.pad_len > 0 else T
    S = cfg.pad_len if (cfg.pad_len > 0 and Te < cfg.pad_len) else Te
    C, hw = cfg.text_guidance_dim, 576
    npix = hw // (cfg.pooling_size[0] * cfg.pooling_size[1])
    sl = B * Te
    swin_attn = hw * (2 * 256 * 128 + 128 * 128 + 128 * 128) + 4 * 4 * 2 * 144 * 144 * 32   # q,k (K=256), v, proj, QK^T, PV
    swin_mlp = hw * 2 * 128 * 512
    class_tok = 2 * 256 * 128 + 128 * 128 + 2 * 128 * 512 + 4 * (32 * 32 + 32 * 32 + 32)
    dec = 48 * 48 * (96 * 128 + 9 * 64 * 128 + 9 * 64 * 64) + 96 * 96 * (48 * 64 + 9 * 32 * 64 + 9 * 32 * 32 + 9 * 32)
    corr = B * T * cfg.prompt_channel * hw * C * (2 if T > Te else 1)
    prep = corr + B * (hw * 128 * C * 9 + 2304 * 32 * 256 * 9 + 9216 * 16 * 128 * 9) + sl * C * 128
    m = {
        "prep": prep, "embed": sl * hw * 128 * 49 * cfg.prompt_channel,
        "swin": sl * swin_attn * 2 * cfg.num_layers,
        "swin_mlp": sl * swin_mlp * 2 * cfg.num_layers,
        "class": B * npix * S * class_tok * cfg.num_layers,
        "decoder": sl * dec,
    }
    return {k: 2.0 * v for k, v in m.items()}


# (kernel name, launches per boundary call) of the stages whose time is one kernel type launched repeatedly
STAGE_KERNELS = {
    "swin": ("swin_attn_fast_kernel", lambda cfg: 2 * cfg.num_layers),
    "swin_mlp": ("mlp_fast_kernel<GELU>", lambda cfg: 2 * cfg.num_layers),
    "class": ("class_state_fast_kernel + class_apply_fast_kernel", lambda cfg: cfg.num_layers),
    "decoder": ("band_conv_kernel x5 (D1..D5)", lambda cfg: 1),
    "prep": ("prep kernels", lambda cfg: 1), "embed": ("igemm_kernel<EmbedA>", lambda cfg: 1),
}


def ncu_traffic(kernel, workload):
    """dram__bytes_read+write per launch from the committed ncu capture (profiles/ncu_traffic.json), or None."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(path)).get(workload, {}).get(kernel)
    except Exception:
        return None


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx.append(float(s[1]))
                for n, v in zip(names, s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- helpers
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": d.get("bf16_tflops_sustained", 1400.6), "hbm_gbs": d.get("hbm_gbs", 6454.9), "src": "measured"}
    return {"tflops": 1590.0, "hbm_gbs": 6650.0, "src": "fallback"}


def get_cfg(name):
    w = BENCH_CONFIGS[name]
    return (vitb() if w["model"] == "vitb" else vitl()), w["B"], w["T"]


def cpu_port_rate(cfg, T, seed=0, repeats=1):
    """The oracle (CPU fp32 port of the reference) on ONE image of the workload, all host threads."""
    from oracle.aggregator_oracle import aggregator_forward
    torch.set_num_threads(os.cpu_count() or 1)
    sd = make_state_dict(cfg, seed)
    img, text, g = make_inputs(cfg, 1, T, seed)
    ts = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        aggregator_forward(sd, cfg.oracle_cfg(), img, text, g)
        ts.append(time.perf_counter() - t0)
    t = statistics.median(ts)
    return 1.0 / t, t, torch.get_num_threads()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ----------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    cfg, B, T = get_cfg(args.workload)
    steps, warm = args.steps, args.warmup
    from oracle.aggregator_oracle import aggregator_forward
    torch.set_num_threads(os.cpu_count() or 1)
    sd = make_state_dict(cfg, 0)
    img, text, g = make_inputs(cfg, 1, T, 0)
    for _ in range(min(warm, 1)):
        aggregator_forward(sd, cfg.oracle_cfg(), img, text, g)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        aggregator_forward(sd, cfg.oracle_cfg(), img, text, g)
        ts.append(time.perf_counter() - t0)
    tot = sum(ts)
    val = steps / tot
    sample = f"1 image of {args.workload} (B=1,T={T}) per step; CPU fp32 oracle port of model.py:683-725"
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(warm, 1), "ms_per_step": 1e3 * tot / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: CAT-Seg ViT-L/14 A-847, T={T}, pool [1,1]", "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    rank, local, world = dist_env()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from cat_seg_b200 import distributed as cdist
    cdist.init_from_env("nccl", dev)
    from cat_seg_b200.aggregator import Aggregator
    from cat_seg_b200 import sliding_window as sw
    from cat_seg_b200.host_pipeline import HostPipeline

    cfg, B, T = get_cfg(args.workload)
    if args.batch:
        B = args.batch
    sd = make_state_dict(cfg, 0)
    model = Aggregator(**cfg.ctor_kwargs(), precision=args.precision)
    model.load_state_dict(sd, strict=False)
    model = model.to(dev)
    class_par = args.parallel == "class" and world > 1
    img, text, g = make_inputs(cfg, B, T, seed=0 if class_par else rank)
    host = [t.pin_memory() for t in (img, text, g[1], g[2])]
    d_img, d_text, d_g1, d_g2 = [t.to(dev) for t in host]

    sliding = args.workload == "cfg5"           # 5 windows of one 640x640 image + stitch/argmax per step

    graph_run = None

    def step_resident():
        if graph_run is not None:
            y = graph_run(d_img, d_text, [d_img, d_g1, d_g2])
            return sw.stitch(y, 640, 640, want_probs=False, want_labels=True)[1] if sliding else y
        if class_par:
            y = model.forward_class_sharded(d_img, d_text, [d_img, d_g1, d_g2])
        else:
            y = model(d_img, d_text, [d_img, d_g1, d_g2])
        if sliding:
            return sw.stitch(y, 640, 640, want_probs=False, want_labels=True)[1]
        return y

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    if args.cuda_graph and not class_par:
        from cat_seg_b200.host_pipeline import GraphRunner
        graph_run = GraphRunner(model, d_img, d_text, [d_img, d_g1, d_g2])
    for _ in range(args.warmup):
        y = step_resident()
    barrier()
    model.set_profiling(graph_run is None)
    model.stage_times(reset=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        y = step_resident()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    stage_ms, calls = model.stage_times(reset=True)
    model.set_profiling(False)
    launches = model.last_launch_count() * args.steps

    # ---- end to end: pinned host inputs -> device, boundary call, per-image argmax labels -> host
    labels_host = (torch.empty(640 * 640, dtype=torch.int32) if sliding else torch.empty(B, 96 * 96, dtype=torch.int32)).pin_memory()

    # Each step uploads ITS OWN inputs (K uploads inside the timed region); the upload of step i+1 is issued before
    # step i computes (copy stream, two device slots), and the host waits for the labels of step i-1 while step i
    # is queued, so the PCIe copies overlap the kernels the way a prefetching data loader does.
    pipe = HostPipeline(model, dev)
    done = [torch.cuda.Event() for _ in range(2)]

    def e2e_run(nsteps):
        ticket = pipe.upload(host)
        for i in range(nsteps):
            nxt = pipe.upload(host) if i + 1 < nsteps else None
            if class_par:
                a_, b_, c_, d_ = pipe.slots[ticket]
                torch.cuda.current_stream(dev).wait_event(pipe.uploaded[ticket])
                yy = model.forward_class_sharded(a_, b_, [a_, c_, d_])
                ev_ = torch.cuda.Event(); ev_.record(); pipe.consumed[ticket] = ev_
            else:
                yy = pipe.run(ticket)
            if sliding:
                labels_host.copy_(sw.stitch(yy, 640, 640, want_probs=False, want_labels=True)[1].view(-1), non_blocking=True)
            else:
                labels_host.copy_(sw.argmax_batched(yy.view(B, T, -1)), non_blocking=True)
            done[i & 1].record()
            if i > 0:
                done[(i - 1) & 1].synchronize()          # the caller consumes the previous step's labels
            ticket = nxt
        done[(nsteps - 1) & 1].synchronize()

    e2e_run(max(1, min(args.warmup, 2)))
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_run(args.steps)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    ms, ms_e2e = cdist.max_over_ranks([ms, ms_e2e], dev)
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    units = 1 if sliding else B                 # images per step per rank
    nrep = 1 if class_par else world                # class-sharded: all ranks work on the same images
    value = nrep * units * args.steps / (ms / 1e3)
    e2e_value = nrep * units * args.steps / (ms_e2e / 1e3)
    pk = peaks()
    fl = stage_flops(cfg, B, T)
    if class_par:                                   # each rank executes 1/world of the per-(image, class) work
        fl = {k: (v if k == "prep" else v / world) for k, v in fl.items()}
    if stage_ms.get("swin_mlp", 0.0) == 0.0:      # exact path: the FFN half runs inside the Swin block kernel
        fl["swin"] += fl["swin_mlp"]
    # the dominant KERNEL: among the stages that are one kernel launched n times (the class and decoder stages are
    # sequences of different kernels; each of those kernels is shorter than the window-attention kernel, see profiles/)
    single = [k for k in ("swin", "swin_mlp") if stage_ms.get(k, 0.0) > 0.0]
    top = max(single, key=lambda k: stage_ms[k]) if single else max(stage_ms, key=lambda k: stage_ms[k])
    kname, nl_fn = STAGE_KERNELS[top]
    n_per_call = nl_fn(cfg)
    per_launch_ms = stage_ms[top] / max(calls, 1) / n_per_call
    achieved = fl[top] / n_per_call / (per_launch_ms * 1e-3) / 1e12 if per_launch_ms > 0 else 0.0
    roof = {"bound": "tensor", "kernel": kname, "stage": top, "launches_per_step": n_per_call, "achieved": achieved,
            "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"],
            "traffic": ncu_traffic(kname, args.workload), "peak_source": pk["src"] + " bf16 sustained (MEASURED_PEAKS.json)",
            "flops_per_launch": fl[top] / n_per_call, "flops_basis": "reference algorithm as written (SURVEY.md 6.2)",
            "ms_per_launch": per_launch_ms,
            "stage_ms_per_step": {k: v / max(calls, 1) for k, v in stage_ms.items()},
            "stage_tflops": {k: (fl[k] / (stage_ms[k] / max(calls, 1) * 1e-3) / 1e12 if stage_ms[k] > 0 else None)
                             for k in stage_ms}}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, sec, cores = cpu_port_rate(cfg, T)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"1 image of {args.workload} (B=1,T={T}), CPU fp32 oracle, {sec:.1f} s"}
    h2d = sum(t.numel() * t.element_size() for t in host)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if class_par else "weak", "vs_baseline": None,
        "dtype": "f32" if args.precision == "exact" else "bf16", "data": "synthetic",
        "config": {"workload": f"{args.workload}: CAT-Seg {'ViT-L/14 336x336' if cfg.text_guidance_dim == 768 else 'ViT-B/16 384x384'}, "
                               f"T={T} classes (Te={min(T, cfg.pad_len)} kept), B={B} "
                               f"{'sliding windows of one 640x640 image' if sliding else 'images'}/GPU, L=2, pool [1,1], P=1",
                   "precision": args.precision, "parallelism": (f"kept classes sharded over {world} ranks, all-reduce of the linear-attention state per class layer (NCCL), "
                                   f"all-gather of the logit planes" if class_par else f"images sharded over {world} rank(s)"),
                   "l2": "activations (1.2 GB/step) exceed the 126 MB L2; no explicit flush",
                   "e2e_result": "stitched argmax labels [640,640] int32" if sliding else "per-image argmax labels [B,96,96] int32"},
        "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": labels_host.numel() * 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
    }
    print(json.dumps(out), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(BENCH_CONFIGS))
    ap.add_argument("--precision", default="fast", help="exact | fast | fast:<stage>[,<stage>]")
    ap.add_argument("--batch", type=int, default=0, help="override images per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cuda-graph", action="store_true", help="replay the boundary call from a CUDA graph (value only; small batches)")
    ap.add_argument("--parallel", default="image", choices=["image", "class"],
                    help="image: each rank gets its own images (weak scaling, no exchange); class: every rank gets the SAME "
                         "images and a slice of the kept classes, one state all-reduce per class layer (strong scaling)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
