#!/usr/bin/env python
"""Benchmark of the RSIS hot path (ResNet-101 FeatureExtractor -> T ConvLSTM decoder steps -> sigmoid), the
`test()` loop of /root/reference/src/test.py:16-50, on BASELINE.json configs[1]: batch 8, 256x256, T=10, 21 classes.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path on the host cores (rank 0)

Prints ONE JSON line.  metric = masks/sec (images x T per second).  A "step" is one full test() pass over one batch
of synthetic images per rank (weak scaling: every rank owns its own batch of 8; no data-path collective -- SURVEY.md
section 8e).  `value` is timed with the batch already resident in HBM; `e2e` goes through the public API with HOST
(pinned) buffers: H2D of the images and D2H of masks/classes/stops inside the timed region.
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

METRIC = "masks/sec (images x T) at 256x256 T=10"
UNIT = "masks/s"
B, H, W, T, NUM_CLASSES = 8, 256, 256, 10, 21
WORKLOAD = "BASELINE.json configs[1]: Pascal VOC inference, batch 8 per GPU, 256x256, T=10, ResNet-101 encoder"
# Other BASELINE.json configs can be timed with --workload (extra data points; the default is the headline config).
WORKLOADS = {
    "cfg2": (8, 256, 256, 10, 21, WORKLOAD),
    "cfg3": (4, 512, 1024, 20, 9, "BASELINE.json configs[2]: Cityscapes inference, batch 4, 512x1024, T=20, 9 classes"),
    "cfg5": (32, 512, 512, 16, 21, "BASELINE.json configs[4] per-rank shard at 8 GPUs: batch 32, 512x512, T=16"),
}


def set_workload(name):
    global B, H, W, T, NUM_CLASSES, WORKLOAD, METRIC
    B, H, W, T, NUM_CLASSES, WORKLOAD = WORKLOADS[name]
    if name != "cfg2":
        METRIC = f"masks/sec (images x T) at {H}x{W} T={T}"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                "bf16_tflops_sustained": float(d["bf16_tflops_sustained"]), "source": "measured"}
    # fallback stated in /opt/skills/guides/B200_PROFILING.md
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# --------------------------------------------------------------------------------------------------------------
# algorithmic work of the fused ConvLSTM cell (stated again in DESIGN.md)
# --------------------------------------------------------------------------------------------------------------
def cell_levels(h, w, hidden=128):
    """(Cin, Ch, H_l, W_l) of the five decoder levels for an h x w input (model.py:90-104)."""
    out = []
    for l in range(5):
        ch = hidden >> l
        cin = hidden if l == 0 else 4 * ch
        out.append((cin, ch, (h // 32) << l, (w // 32) << l))
    return out


def cell_alg_bytes(batch, cin, ch, hl, wl, s=4):
    px = batch * hl * wl
    return s * (px * (cin + 2 * ch) + px * 2 * ch + 36 * ch * (cin + ch) + 4 * ch)


def cell_alg_flops(batch, cin, ch, hl, wl):
    return 2 * batch * hl * wl * 4 * ch * 9 * (cin + ch)


class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
            except ValueError:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------------------
# the reference's CPU path (oracle port): the reported CPU baseline and the `--impl reference` arm
# --------------------------------------------------------------------------------------------------------------
def cpu_reference_time(steps, warmup, batch=B):
    """Times the oracle's restatement of test() (oracle/rsis_oracle.py:test_loop, same torch CPU primitives as the
    reference's modules) on all host cores. Returns (seconds per pass, threads)."""
    from oracle import rsis_oracle as O, synth_weights as sw
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    esd, dsd = sw.encoder_state_dict(1), sw.decoder_state_dict(1, num_classes=NUM_CLASSES)
    x = sw.synthetic_images(123, batch, H, W)
    for _ in range(warmup):
        O.test_loop(esd, dsd, x, T)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        O.test_loop(esd, dsd, x, T)
        times.append(time.perf_counter() - t0)
    return times, torch.get_num_threads()


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    times, threads = cpu_reference_time(a.steps, a.warmup)
    total = sum(times)
    v = B * T * a.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * total / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": B, "height": H, "width": W, "T": T,
                   "num_classes": NUM_CLASSES, "device": "host CPU", "note": "reference CPU path: the same torch.nn CPU primitives the reference's "
                   "modules call (src/test.py:16-50), driven by oracle/rsis_oracle.py because /root/reference does "
                   "not exist on the GPU box; one step = one full test() pass over one batch of 8"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{a.steps} full test() passes (B={B}, {H}x{W}, T={T}) after {a.warmup} warm-up"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def torch_gpu_time(dev, passes=5, warmup=2):
    """Secondary baseline: the SAME oracle restatement of test() (stock torch ops: cuDNN / cuBLAS fp32, TF32 off) on the
    same GPU, eager, CUDA-event timed.  Not a product path: it answers "what does stock PyTorch do here"."""
    from oracle import rsis_oracle as O, synth_weights as sw
    tf = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        esd = {k: v.to(dev) for k, v in sw.encoder_state_dict(1).items()}
        dsd = {k: v.to(dev) for k, v in sw.decoder_state_dict(1, num_classes=NUM_CLASSES).items()}
        x = sw.synthetic_images(123, B, H, W).to(dev)
        with torch.no_grad():
            for _ in range(warmup):
                O.test_loop(esd, dsd, x, T)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(passes):
                O.test_loop(esd, dsd, x, T)
            e1.record()
            torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / passes
        return {"value": B * T / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "kind": "stock torch eager on the same "
                "GPU (cuDNN/cuBLAS fp32, TF32 off), the oracle's op sequence", "passes": passes}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf


def train_record(rsis_b200, rdist, dev, rank, world, steps=5, warmup=3, precision=None):
    """BASELINE.json configs[3] per-rank shard (8 images 256x256, T=10): train-mode encoder + T decoder steps +
    loss.backward() as ONE CUDA graph, then the ONE data-parallel collective of the design -- an NCCL all-reduce of the
    flat gradient buffer (replaces nn.DataParallel, /root/reference/src/train.py:269-274).  Timed per step with CUDA
    events (max over ranks); the all-reduce additionally on its own."""
    import bench_train
    from rsis_b200.autograd import GradBucket
    from rsis_b200.training import TrainStep
    from oracle import synth_weights as sw   # deterministic synthetic weights / images only
    from oracle import ref_shims as rs
    args = rs.make_args(num_classes=NUM_CLASSES, maxseqlen=T)
    args.hidden_size = int(args.hidden_size)
    args.use_gpu = True
    enc, dec = rsis_b200.FeatureExtractor(args), rsis_b200.RSIS(args)
    enc.load_state_dict(sw.encoder_state_dict(1))
    dec.load_state_dict(sw.decoder_state_dict(1, num_classes=NUM_CLASSES))
    enc.to(dev).train()
    dec.to(dev).train()
    x = sw.synthetic_images(500 + rank, B, H, W).to(dev)
    bucket = GradBucket(list(enc.parameters()) + list(dec.parameters()))
    step = TrainStep(enc, dec, T, bench_train.loss_fn, bucket=bucket, cuda_graph=True, precision=precision)
    for _ in range(max(warmup, 3)):
        step(x)
    torch.cuda.synchronize(dev)
    rdist.barrier()
    evs = []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(x)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize(dev)
    rdist.barrier()
    step_s = rdist.max_over_ranks(sum(a_.elapsed_time(b_) for a_, b_ in evs) * 1e-3) / steps
    ar = []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        rdist.barrier()
        e0.record()
        bucket.all_reduce()
        e1.record()
        ar.append((e0, e1))
    torch.cuda.synchronize(dev)
    ar_s = rdist.max_over_ranks(sum(a_.elapsed_time(b_) for a_, b_ in ar) * 1e-3) / steps
    nbytes = bucket.flat.numel() * 4
    rec = {"workload": "BASELINE.json configs[3] per-rank shard: training step, 8 images 256x256 per GPU, T=10 "
                       "(global batch = 8 x n_gpus; 64 at 8 GPUs)",
           "ms_per_step": step_s * 1e3, "images_per_s": world * B / step_s, "n_gpus": world,
           "collective": "one NCCL all-reduce (sum, / world) of the flat float32 gradient buffer per step"
                         if world > 1 else "none at 1 GPU (the all-reduce is skipped)",
           "allreduce_ms": ar_s * 1e3 if world > 1 else 0.0, "allreduce_bytes": nbytes,
           "allreduce_busbw_GBps": (2 * (world - 1) / world * nbytes / ar_s / 1e9) if world > 1 and ar_s > 0 else None,
           "precision": ("bf16: single-pass bf16 tensor-core products, fp32 accumulation / master weights / gradients "
                         "(BASELINE.json configs[3])" if precision == "bf16" else
                         "split bf16: fp32-grade products (three / four bf16 passes)"),
           "steps": steps, "cuda_graph": True}
    del step, bucket, enc, dec
    torch.cuda.empty_cache()
    return rec


# --------------------------------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------------------------------
def time_cells(rsis_b200, dec, ws, impl, iters=20):
    """CUDA-event time of the five fused ConvLSTM cell launches of one decoder step, per level, on the real state a
    2-step run left in the decoder workspace `ws` (inputs = its concatenated [up(h) | h_prev] buffers + the hoisted
    skip share of the gates, c_prev = its cell state; outputs go to scratch).  Between iterations a 512 MiB buffer is written to flush L2."""
    ops = rsis_b200.ops
    dev = ws.side.device
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    side = torch.zeros_like(ws.side)
    p = ws.t & 1
    scratch = []
    for l, cell in enumerate(dec.clstm_list):
        x = ws.X[l][p]
        scratch.append((ops.Act.empty(x.n, x.h, x.w, cell.hidden_size, ops.FMT_F32, dev),
                        ops.Act.empty(x.n, x.h, x.w, cell.hidden_size, ops.FMT_F32, dev),
                        ops.Act.empty(x.n, x.h, x.w, cell.hidden_size, ops.FMT_SPLIT_BF16, dev),
                        ws.packs(dec, l)[1]))
    per_level = [[] for _ in dec.clstm_list]
    # One (flush, event, cell, event) group per launch: the ~150 us flush kernel lets the host enqueue the cell
    # and both events before the GPU reaches them, so the interval is kernel time, not host launch latency.
    # time the launch variant the pass itself uses: the default wavefront schedule runs the cells WITHOUT split-K
    import contextlib
    from rsis_b200 import _lib
    unsplit = os.environ.get("RSIS_B200_PIPELINE", "3") in ("2", "3")
    for l, cell in enumerate(dec.clstm_list):
        h, c, h16, pc = scratch[l]
        for it in range(iters + 3):
            flush.fill_(it & 0xFF)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            with (_lib.no_splitk() if unsplit else contextlib.nullcontext()):
                ops.convlstm_cell_x(ws.X[l][p], pc, ws.c[l].t, side, 0, h_out=h, c_out=c, h16_out=h16, impl=impl,
                                    gate_preact=ws.P[l])
            e1.record()
            if it >= 3:
                per_level[l].append((e0, e1))
        torch.cuda.synchronize(dev)
    return [statistics.mean(a.elapsed_time(b) for a, b in lv) * 1e-3 for lv in per_level]


def time_cell_group(rsis_b200, dec, ws, iters=20):
    """CUDA-event time of ONE grouped wavefront launch (`cell_group_kernel`: the five cells of one decoder step, levels
    0-4, side by side -- the launch the default decoder schedule issues in steady state), on the real state a 2-step run
    left in `ws`; outputs go to scratch; L2 flushed (512 MiB fill) before every launch."""
    ops = rsis_b200.ops
    dev = ws.side.device
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    side = torch.zeros_like(ws.side)
    p = ws.t & 1
    offs = [sum(ws.hidden[:l]) for l in range(len(ws.hidden))]
    cells = []
    for l, cell in enumerate(dec.clstm_list):
        x = ws.X[l][p]
        cells.append(dict(x=x, pc=ws.packs(dec, l)[1], c_prev=ws.c[l].t, side_max=side, side_offset=offs[l],
                          h_out=ops.Act.empty(x.n, x.h, x.w, cell.hidden_size, ops.FMT_F32, dev),
                          c_out=ops.Act.empty(x.n, x.h, x.w, cell.hidden_size, ops.FMT_F32, dev),
                          h16_out=ops.Act.empty(x.n, x.h, x.w, cell.hidden_size, ops.FMT_SPLIT_BF16, dev),
                          gate_preact=ws.P[l]))
    evs = []
    for it in range(iters + 3):
        flush.fill_(it & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.convlstm_cell_group(cells)
        e1.record()
        if it >= 3:
            evs.append((e0, e1))
    torch.cuda.synchronize(dev)
    return statistics.mean(a.elapsed_time(b) for a, b in evs) * 1e-3


def run_ours(a):
    import rsis_b200
    from rsis_b200 import dist as rdist, inference, ops
    from oracle import synth_weights as sw  # weights/images generator only (deterministic synthetic data)
    from oracle import ref_shims as rs

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback (use --impl reference)")
    rank, local_rank, world = rdist.init_from_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world != a.gpus and rank == 0:
        print(f"warning: --gpus {a.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    impl = ops.default_impl()

    args = rs.make_args(num_classes=NUM_CLASSES, maxseqlen=T)
    args.hidden_size = int(args.hidden_size)
    args.use_gpu = True
    enc, dec = rsis_b200.FeatureExtractor(args), rsis_b200.RSIS(args)
    enc.load_state_dict(sw.encoder_state_dict(1))
    dec.load_state_dict(sw.decoder_state_dict(1, num_classes=NUM_CLASSES))
    enc.to(dev).eval()
    dec.to(dev).eval()
    # every rank owns its own batch of B images (weak scaling; different images per rank)
    x_host = sw.synthetic_images(123 + rank, B, H, W).pin_memory()
    x_dev = x_host.to(dev)

    sess = inference.InferenceSession(args, enc, dec, x_dev.shape, dev, impl)
    sess.x.copy_(x_dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    # ---- value: inputs resident in HBM, K graph replays, per-step CUDA events, L2 flushed between steps ----
    for i in range(max(a.warmup, 3)):
        flush.fill_(i)
        sess.replay()
    torch.cuda.synchronize(dev)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    rdist.barrier()
    torch.cuda.synchronize(dev)
    evs = []
    for i in range(a.steps):
        flush.fill_(i & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sess.replay()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize(dev)
    rdist.barrier()
    clk = clocks.stop() if rank == 0 else None
    local_s = sum(s.elapsed_time(e) for s, e in evs) * 1e-3
    total_s = rdist.max_over_ranks(local_s)
    value = world * B * T * a.steps / total_s
    launches = sess.launches * a.steps

    # ---- e2e: public API call with HOST buffers (pinned); H2D + graph + D2H inside the timed region ----
    # Results are read back on a copy stream into two alternating pinned buffers, so the D2H of pass i overlaps the
    # compute of pass i+1, and the H2D of pass i+1 is issued under the compute of pass i (every pass still pays its own
    # H2D and D2H inside the timed region).
    nbuf = 2
    masks_h = [torch.empty((B, T, H, W), dtype=torch.float32).pin_memory() for _ in range(nbuf)]
    classes_h = [torch.empty((B, T, NUM_CLASSES), dtype=torch.float32).pin_memory() for _ in range(nbuf)]
    stops_h = [torch.empty((B, T, 1), dtype=torch.float32).pin_memory() for _ in range(nbuf)]
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    copied = [torch.cuda.Event() for _ in range(nbuf)]

    # Inputs: two device buffers filled from pinned host memory on an H2D stream, one step ahead -- the copy for pass
    # i+1 runs under the compute of pass i (every timed pass issues exactly one H2D and one D2H of a full batch).
    h2d_stream = torch.cuda.Stream(device=dev)
    xbuf = [torch.empty_like(x_dev) for _ in range(nbuf)]
    h2d_done = [torch.cuda.Event() for _ in range(nbuf)]
    consumed = [torch.cuda.Event() for _ in range(nbuf)]
    for ev in consumed:
        ev.record(main_stream)

    def issue_h2d(i):
        k = i % nbuf
        with torch.cuda.stream(h2d_stream):
            h2d_stream.wait_event(consumed[k])              # the pass that last read this buffer has finished
            xbuf[k].copy_(x_host, non_blocking=True)
            h2d_done[k].record(h2d_stream)

    def e2e_step(i):
        k = i % nbuf
        issue_h2d(i + 1)                                     # next pass's input, under this pass's compute
        main_stream.wait_event(h2d_done[k])
        m, c, s = rsis_b200.test(args, enc, dec, xbuf[k])    # fresh result tensors (clones of the session outputs)
        consumed[k].record(main_stream)
        done = torch.cuda.Event()
        done.record(main_stream)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done)
            masks_h[k].copy_(m, non_blocking=True)
            classes_h[k].copy_(c, non_blocking=True)
            stops_h[k].copy_(s, non_blocking=True)
            copied[k].record(copy_stream)
        for t in (m, c, s):
            t.record_stream(copy_stream)

    issue_h2d(0)
    for i in range(3):
        e2e_step(i)
    torch.cuda.synchronize(dev)
    rdist.barrier()
    # re-prime: the input of the first timed pass (3 % nbuf) was already issued by the last warm-up pass
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(3, 3 + a.steps):
        e2e_step(i)
    main_stream.wait_stream(copy_stream)   # the timed region ends when the last result has reached host memory
    main_stream.wait_stream(h2d_stream)    # ... and the last issued input copy has landed
    e1.record()
    torch.cuda.synchronize(dev)
    rdist.barrier()
    e2e_s = rdist.max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    e2e_value = world * B * T * a.steps / e2e_s
    h2d = x_host.numel() * 4
    d2h = (masks_h[0].numel() + classes_h[0].numel() + stops_h[0].numel()) * 4

    # ---- roofline of the fused ConvLSTM cell kernel (the kernel BASELINE.json's metric names), rank 0 ----
    roof = None
    cpu = None
    if rank == 0:
        pk = peaks()
        if not ops.uses_tcgen05(impl):
            raise SystemExit("bench.py measures the tcgen05 kernel family (RSIS_B200_IMPL=auto|tcgen05)")
        with torch.no_grad():
            cm = torch.empty((B, T, NUM_CLASSES), device=dev)
            mk = torch.empty((B, T, H, W), device=dev)
            sp = torch.empty((B, T, 1), device=dev)
            ws = inference.run_eager(enc, dec, x_dev, 2, impl, mk, cm, sp)
            lv_s = time_cells(rsis_b200, dec, ws, impl)
            grouped = os.environ.get("RSIS_B200_PIPELINE", "3") == "3"
            group_s = time_cell_group(rsis_b200, dec, ws) if grouped else None
        levels = cell_levels(H, W)
        per_level = []
        tot_b = tot_f = 0
        for (cin, ch, hl, wl), s in zip(levels, lv_s):
            nb, nf = cell_alg_bytes(B, cin, ch, hl, wl), cell_alg_flops(B, cin, ch, hl, wl)
            tot_b += nb
            tot_f += nf
            per_level.append({"level": len(per_level), "us": s * 1e6, "alg_bytes": nb, "alg_flops": nf,
                              "GBps": nb / s / 1e9, "TFLOPs": nf / s / 1e12})
        # the dominant kernel of the pass: under the default (grouped wavefront) schedule ONE launch of cell_group_kernel
        # runs the five cells of a decoder step; under the older schedules the step is five launches, timed one by one
        step_s = group_s if group_s is not None else sum(lv_s)
        ach = tot_b / step_s / 1e9
        # ncu dram__bytes_read.sum + dram__bytes_write.sum of one grouped launch (one `--set full` capture of this kernel,
        # profiles/cell_traffic.json; reads and writes also separately -- a profiler cannot run inside a timed run)
        traffic = traffic_rw = None
        tpath = os.path.join(ROOT, "profiles", "cell_traffic.json")
        if os.path.exists(tpath) and group_s is not None:
            tj = json.load(open(tpath))
            if tj.get("workload") == a.workload:
                traffic = tj.get("dram_bytes_per_step")
                traffic_rw = {"read": tj.get("dram_read_bytes_per_launch"), "write": tj.get("dram_write_bytes_per_launch"),
                              "note": tj.get("note"), "source": tj.get("source")}
        roof = {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                "traffic": traffic, "traffic_read_write": traffic_rw, "peak_source": pk["source"] + " (burst copy figure; kernel timed alone, L2 flushed)",
                "kernel": ("cell_group_kernel: the fused ConvLSTM cells of one decoder step (levels 0-4) in ONE grouped "
                           "launch, as the default wavefront schedule issues them; CUDA-event timed live"
                           if group_s is not None else
                           "fused ConvLSTM cell (5 launches = one decoder step, levels 0-4), CUDA-event timed live, in "
                           "the launch variant the pass uses (no split-K under the wavefront schedule)"),
                "launches_per_step_of_this_kernel": 1 if group_s is not None else 5,
                "single_launch_sum_us": sum(lv_s) * 1e6,
                "alg_bytes_per_step": tot_b, "alg_flops_per_step": tot_f, "step_us": step_s * 1e6,
                "tensor": {"achieved_TFLOPs": tot_f / step_s / 1e12, "peak_bf16_TFLOPs": pk["bf16_tflops"],
                           "frac_of_bf16_peak": tot_f / step_s / 1e12 / pk["bf16_tflops"]},
                "impl": {ops.IMPL_SIMT: "simt-fp32", ops.IMPL_AUTO: "auto", ops.IMPL_TCGEN05: "tcgen05"}[impl],
                "per_level": per_level}
        # ---- CPU baseline: the reference's CPU path (oracle port) on the host cores, bounded sample ----
        # rank 0 at N = 1 only: at N > 1 the other ranks would spin in the next barrier while rank 0 runs the CPU leg
        # (and their spinning would steal the host cores it is timed on)
        if not a.no_cpu_baseline and world == 1:
            times, threads = cpu_reference_time(a.cpu_passes, 1)
            cv = B * T * len(times) / sum(times)
            cpu = {"value": cv, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{len(times)} full test() passes of the same workload (B={B}, {H}x{W}, T={T}) after 1 "
                             f"warm-up; torch CPU fp32, {threads} threads"}
    torch_gpu = None
    if rank == 0 and not a.no_torch_gpu:
        torch_gpu = torch_gpu_time(dev)
    rdist.barrier()
    train = train_bf16 = None
    if not a.no_train and a.workload == "cfg2":
        train = train_record(rsis_b200, rdist, dev, rank, world, steps=a.train_steps)
        train_bf16 = train_record(rsis_b200, rdist, dev, rank, world, steps=a.train_steps, precision="bf16")
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": 1e3 * total_s / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": B * world, "height": H, "width": W,
                       "T": T, "num_classes": NUM_CLASSES, "parallelism": f"dp{world} (batch sharded per image, "
                       "no data-path collective)", "l2": "flushed between timed steps (512 MiB fill); per-step CUDA "
                       "events summed", "cuda_graph": True,
                       "decoder_schedule": {"0": "sequential", "1": "wavefront over (level, step), split-K cells",
                                            "2": "wavefront over (level, step), cells without split-K",
                                            "3": "grouped wavefront: cell (l, t) in wavefront 2l + t, one launch per wavefront, "
                                                 "upsamplings on a side stream beside the next one"
                                            if os.environ.get("RSIS_B200_WAVE_SKEW", "2") == "2" else
                                            "grouped wavefront: one launch per anti-diagonal of (level, step)"}.get(
                           os.environ.get("RSIS_B200_PIPELINE", "3"), "custom"),
                       "impl": {ops.IMPL_SIMT: "simt", ops.IMPL_AUTO: "auto", ops.IMPL_TCGEN05: "tcgen05"}[impl],
                       "tcgen05": bool(ops.has_tcgen05())},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s / a.steps},
            "gpu_launches": launches, "launches_per_step": sess.launches,
            "clocks": clk, "roofline": roof, "cpu_baseline": cpu, "torch_gpu_baseline": torch_gpu, "train": train,
            "train_bf16": train_bf16,
            "weights": "synthetic, conditioned (oracle/synth_weights.py: damped residual branches, calibrated BatchNorm "
                       "statistics; SURVEY.md H3) -- the parity claims are stated on these weights",
        }
        print(json.dumps(line))
    rdist.barrier()
    rdist.shutdown()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-passes", type=int, default=5, help="timed CPU test() passes for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-torch-gpu", action="store_true", help="skip the stock-torch-on-the-same-GPU secondary baseline")
    ap.add_argument("--no-train", action="store_true", help="skip the configs[3] training-step record (the collective)")
    ap.add_argument("--train-steps", type=int, default=5)
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    a = ap.parse_args()
    set_workload(a.workload)
    if a.impl == "reference":
        return run_reference(a)
    return run_ours(a)


if __name__ == "__main__":
    sys.exit(main())
