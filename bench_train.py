#!/usr/bin/env python
"""Training-step benchmark of the RSIS hot path: BASELINE.json configs[3] per-rank shard -- 8 images 256x256, T=10,
21 classes: train-mode encoder forward, T decoder steps, `loss.backward()` through both and
the ONE data-parallel gradient all-reduce over a flat buffer (section 8e) when launched on N ranks.

    python bench_train.py --steps K --warmup W                      # 1 GPU
    python -m torch.distributed.run --nproc-per-node N ... bench_train.py --gpus N

Prints ONE JSON line: images/s of whole training steps (forward + backward + all-reduce; the optimiser step and the
criteria / Hungarian matching of train.py:96-176 are outside the hot path), with the forward / backward / all-reduce
split, and -- on rank 0, `--cpu-steps` > 0 -- the same step through autograd over the CPU oracle (the reference's
own torch CPU primitives) as the reported CPU baseline.  This is an extra data point next to bench.py, whose metric
(masks/s of the inference pass) is the one BASELINE.json names.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

B, H, W, T, NUM_CLASSES = 8, 256, 256, 10, 21


def loss_fn(masks, classes, stops):
    """Stand-in for the criteria of train.py:159-176 (outside the hot path): touches every output of every step."""
    loss = 0
    for m, c, s in zip(masks, classes, stops):
        loss = loss + (torch.sigmoid(m) ** 2).mean() + (c ** 2).sum(-1).mean() + (s ** 2).mean()
    return loss


def cpu_step_time(steps):
    from oracle import rsis_oracle as O, synth_weights as sw
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    esd = {k: v.clone() for k, v in sw.encoder_state_dict(1).items()}
    dsd = {k: v.clone() for k, v in sw.decoder_state_dict(1, num_classes=NUM_CLASSES).items()}
    for sd in (esd, dsd):
        for k, v in sd.items():
            if v.is_floating_point() and "running_" not in k:
                v.requires_grad_(True)
    x = sw.synthetic_images(123, B, H, W)
    times = []
    for i in range(steps + 1):
        t0 = time.perf_counter()
        feats = O.feature_extractor(esd, x, bn=O._bn_train)
        hidden = None
        masks, classes, stops = [], [], []
        for _ in range(T):
            m, c, s, hidden = O.rsis_step(dsd, feats, hidden)
            masks.append(m)
            classes.append(c)
            stops.append(s)
        loss_fn(masks, classes, stops).backward()
        for sd in (esd, dsd):
            for v in sd.values():
                v.grad = None
        if i > 0:
            times.append(time.perf_counter() - t0)
    return times, threads


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu-steps", type=int, default=1)
    ap.add_argument("--no-graph", action="store_true", help="time the eager step only")
    ap.add_argument("--precision", default=None, choices=[None, "fp32", "bf16"],
                    help="bf16: single-pass bf16 tensor-core products (BASELINE.json configs[3])")
    a = ap.parse_args()

    import rsis_b200
    from rsis_b200 import dist as rdist, ops
    from rsis_b200.autograd import GradBucket, backward_impl
    from oracle import synth_weights as sw
    from oracle import ref_shims as rs

    if not torch.cuda.is_available():
        raise SystemExit("bench_train.py: no CUDA device; the CUDA path has no CPU fallback")
    rank, local_rank, world = rdist.init_from_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    args = rs.make_args(num_classes=NUM_CLASSES, maxseqlen=T)
    args.hidden_size = int(args.hidden_size)
    args.use_gpu = True
    if a.precision:
        args.precision = a.precision
    enc, dec = rsis_b200.FeatureExtractor(args), rsis_b200.RSIS(args)
    enc.load_state_dict(sw.encoder_state_dict(1))
    dec.load_state_dict(sw.decoder_state_dict(1, num_classes=NUM_CLASSES))
    enc.to(dev).train()
    dec.to(dev).train()
    x = sw.synthetic_images(123 + rank, B, H, W).to(dev)
    from rsis_b200.training import TrainStep
    bucket = GradBucket(list(enc.parameters()) + list(dec.parameters()))

    def timed(step_fn, steps, warmup):
        for _ in range(warmup):
            step_fn()
        torch.cuda.synchronize(dev)
        n0 = ops.launch_count()
        rdist.barrier()
        evs = []
        for _ in range(steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step_fn()
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize(dev)
        rdist.barrier()
        tot = sum(s_.elapsed_time(e_) for s_, e_ in evs) * 1e-3
        return rdist.max_over_ranks(tot), ops.launch_count() - n0

    # ---- CUDA graph: pack + forward + loss + backward captured once, replayed; all-reduce after the replay ----
    graph_s = None
    graph_err = None
    if not a.no_graph:
        try:
            graphed = TrainStep(enc, dec, T, loss_fn, bucket=bucket, cuda_graph=True)
            graph_s, _ = timed(lambda: graphed(x), a.steps, max(a.warmup, 3))
        except Exception as e:  # report, do not hide: the eager number stands
            import traceback
            traceback.print_exc()
            graph_err = f"{type(e).__name__}: {e}"[:300]
            torch.cuda.synchronize(dev)
    # ---- eager: every kernel launched from Python each step (host-bound) ----
    eager = TrainStep(enc, dec, T, loss_fn, bucket=bucket, cuda_graph=False)
    eager_s, eager_launches = timed(lambda: eager(x), a.steps, max(a.warmup, 3))
    # forward / backward split of the eager step (CUDA events around loss.backward())
    split = None
    if rank == 0:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        bucket.zero()
        ev[0].record()
        feats = enc(x)
        hidden = None
        masks, classes, stops = [], [], []
        for _ in range(T):
            m, c, s, hidden = dec(feats, hidden)
            masks.append(m)
            classes.append(c)
            stops.append(s)
        loss = loss_fn(masks, classes, stops)
        ev[1].record()
        loss.backward()
        ev[2].record()
        torch.cuda.synchronize(dev)
        split = {"forward": ev[0].elapsed_time(ev[1]), "backward": ev[1].elapsed_time(ev[2])}
        del feats, hidden, masks, classes, stops, loss
    total_s = graph_s if graph_s is not None else eager_s
    value = world * B * a.steps / total_s
    launches = eager_launches
    # ---- soft-IoU cost matrix of train.py:96-110 (section 8f rank 1): HBM-bound kernel, roofline live ----
    iou = None
    if rank == 0:
        from rsis_b200 import objectives
        G = 20                                            # gt_maxseqlen (args.py)
        gen = torch.Generator().manual_seed(3)
        logits = (torch.randn((B, H * W), generator=gen) * 2).to(dev)
        y = (torch.rand((B, G, H * W), generator=gen) < 0.25).to(dev)
        flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
        peak = 6650.0
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        src = "fallback (B200_PROFILING.md)"
        if os.path.exists(pk):
            peak, src = float(json.load(open(pk))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
        iou = {"peak_GBps": peak, "peak_source": src, "shape": {"B": B, "gtT": G, "HW": H * W}}
        for name, gt, gbytes in (("f32_masks", y.float(), 4), ("u8_masks", y.to(torch.uint8), 1)):
            out = torch.empty((B, G), device=dev)
            evs = []
            for it in range(13):
                flush.fill_(it)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                objectives.soft_iou_cost_matrix(logits, gt, 1.0, out=out)
                e1.record()
                if it >= 3:
                    evs.append((e0, e1))
            torch.cuda.synchronize(dev)
            t = sum(s_.elapsed_time(e_) for s_, e_ in evs) * 1e-3 / len(evs)
            nbytes = 4 * B * H * W + gbytes * B * G * H * W
            iou[name] = {"us": t * 1e6, "alg_bytes": nbytes, "GBps": nbytes / t / 1e9, "frac": nbytes / t / 1e9 / peak}
    cpu = None
    if rank == 0 and a.cpu_steps > 0:
        times, threads = cpu_step_time(a.cpu_steps)
        cpu = {"value": B * len(times) / sum(times), "unit": "images/s", "cores": threads, "kind": "port",
               "sample": f"{len(times)} training step(s) (forward + autograd backward) of the same shard after 1 warm-up, "
                         f"torch CPU fp32, {threads} threads"}
    if rank == 0:
        impl = ops.default_impl()
        print(json.dumps({
            "metric": "training images/sec (forward + backward + gradient all-reduce) at 256x256 T=10",
            "value": value, "unit": "images/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": 1e3 * total_s / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[3] per-rank shard: Pascal VOC training step, batch 8 per GPU, "
                                   "256x256, T=10 (fp32 gradients; split-bf16 tensor-core operands where tcgen05 is used)",
                       "batch_per_gpu": B, "global_batch": B * world, "T": T, "num_classes": NUM_CLASSES,
                       "parallelism": f"dp{world}: one flat-buffer gradient all-reduce per step ({bucket.flat.numel() * 4} bytes)",
                       "forward_impl": "auto" if ops.uses_tcgen05(impl) else "simt",
                       "backward_impl": "auto" if backward_impl(impl) != ops.IMPL_SIMT else "simt",
                       "masks_per_s": value * T},
            "mode": "cuda_graph" if graph_s is not None else "eager",
            "eager_ms_per_step": 1e3 * eager_s / a.steps,
            "graph_ms_per_step": None if graph_s is None else 1e3 * graph_s / a.steps, "graph_error": graph_err,
            "eager_split_ms": split,
            "gpu_launches": launches, "launches_per_step": launches / a.steps, "cpu_baseline": cpu,
            "soft_iou_cost": iou,
        }))
    rdist.barrier()
    rdist.shutdown()
    return 0


if __name__ == "__main__":
    sys.exit(main())
