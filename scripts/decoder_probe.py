#!/usr/bin/env python
"""Development probe: the decoder alone (encoder output resident) under the different schedules, as CUDA-graph
replays, and the grouped wavefront launch timed alone with L2 flushed.  usage: decoder_probe.py [B H W T]"""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import rsis_b200
from rsis_b200 import ops, inference, _lib
from oracle import ref_shims as rs, synth_weights as sw

B, H, W, T = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (8, 256, 256, 10)
args = rs.make_args(maxseqlen=T); args.hidden_size = int(args.hidden_size); args.use_gpu = True
enc, dec = rsis_b200.FeatureExtractor(args), rsis_b200.RSIS(args)
enc.load_state_dict(sw.encoder_state_dict(1)); dec.load_state_dict(sw.decoder_state_dict(1))
enc.cuda().eval(); dec.cuda().eval()
x = sw.synthetic_images(5, B, H, W).cuda()
impl = ops.default_impl()
dev = x.device
masks = torch.empty((B, T, H, W), device=dev); classes = torch.empty((B, T, 21), device=dev); stops = torch.empty((B, T, 1), device=dev)
ws = dec.workspace(B, inference.feature_sizes(H, W), dev)
with torch.no_grad():
    keep = ws.encode_into(enc, dec, x, impl)
torch.cuda.synchronize()
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

def run(mode):
    ws.reset()
    if mode == "3":
        dec.run_wavefront(ws, impl, T, classes, masks, stops)
    elif mode == "2":
        dec.run_pipelined(ws, impl, T, classes, masks, stops, split_k=False)
    elif mode == "1":
        dec.run_pipelined(ws, impl, T, classes, masks, stops, split_k=True)
    else:
        for t in range(T):
            dec.step_ws(ws, impl, None, classes[:, t], T * 21, None, T, mask_prob=masks[:, t], mask_prob_stride=T * H * W, stop_prob=stops[:, t])

res = {}
for mode in sys.argv[5].split(",") if len(sys.argv) > 5 else ["3", "2", "0"]:
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side), torch.no_grad():
        run(mode)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    n0 = ops.launch_count()
    with torch.cuda.graph(g), torch.no_grad():
        run(mode)
    nl = ops.launch_count() - n0
    ts = []
    for it in range(13):
        flush.fill_(it)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        if it >= 3:
            ts.append(e0.elapsed_time(e1) * 1e3)
    res[mode] = masks.clone()
    print(f"decoder schedule {mode}: {statistics.mean(ts):.1f} us per pass (min {min(ts):.1f}), {nl} launches", flush=True)
ks = list(res)
for k in ks[1:]:
    print(f"masks {k} vs {ks[0]}: max abs diff {float((res[k] - res[ks[0]]).abs().max()):.3e}")

# ---- one full wavefront (all five levels) as ONE grouped launch, timed alone, L2 flushed ----
nlev = len(dec.clstm_list)
p = ws.t & 1
cells = []
offs = [sum(ws.hidden[:l]) for l in range(nlev)]
side_keys = torch.zeros_like(ws.side)
for l, cell in enumerate(dec.clstm_list):
    xx = ws.X[l][p]
    cells.append(dict(x=xx, pc=ws.packs(dec, l)[1], c_prev=ws.c[l].t, side_max=side_keys, side_offset=offs[l],
                      h_out=ops.Act.empty(xx.n, xx.h, xx.w, cell.hidden_size, ops.FMT_F32, dev),
                      c_out=ops.Act.empty(xx.n, xx.h, xx.w, cell.hidden_size, ops.FMT_F32, dev),
                      h16_out=ops.Act.empty(xx.n, xx.h, xx.w, cell.hidden_size, ops.FMT_SPLIT_BF16, dev), gate_preact=ws.P[l]))
def time_launch(fn, iters=12):
    ts = []
    for it in range(iters + 3):
        flush.fill_(it & 0xff)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        if it >= 3:
            ts.append((e0, e1))
    torch.cuda.synchronize()
    return statistics.mean(a.elapsed_time(b) for a, b in ts) * 1e3
print(f"grouped wavefront launch (levels 0-4): {time_launch(lambda: ops.convlstm_cell_group(cells)):.1f} us")
for sub in ([0], [1], [2], [3], [4], [0, 1, 2], [3, 4]):
    print(f"  group of levels {sub}: {time_launch(lambda: ops.convlstm_cell_group([cells[i] for i in sub])):.1f} us")
ups = [(ws.h[l], ws.up_view(l + 1, p)) for l in range(nlev - 1)]
print(f"grouped upsample launch (4 levels): {time_launch(lambda: ops.upsample_bilinear_group(ups)):.1f} us")
hl = ws.h[nlev - 1]
print(f"mask head: {time_launch(lambda: ops.upsample_mask_head(hl, 2 * hl.h, 2 * hl.w, dec.conv_out.weight, dec.conv_out.bias, None, masks[:, 0], T * H * W)):.1f} us")
for l in range(nlev):
    c = cells[l]
    with _lib.no_splitk():
        t1 = time_launch(lambda: ops.convlstm_cell_x(c["x"], c["pc"], c["c_prev"], side_keys, offs[l], h_out=c["h_out"], c_out=c["c_out"], h16_out=c["h16_out"], impl=impl, gate_preact=c["gate_preact"]))
    print(f"  single cell launch level {l} (no split-K): {t1:.1f} us")
