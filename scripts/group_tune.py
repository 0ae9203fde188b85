#!/usr/bin/env python
"""Development: the grouped cell launch of one full wavefront (levels 0-4, B=8, 256x256) timed alone with L2 flushed,
under CTA-share / tile-width overrides (RSIS_B200_GROUP_SHARES / RSIS_B200_GROUP_BN, read per call).
usage: group_tune.py "shares;bn" ["shares;bn" ...]   e.g. "8,26,25,29,59;256,64,128,64,32"  ('-' = library default)"""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import rsis_b200
from rsis_b200 import ops, inference, _lib
from oracle import ref_shims as rs, synth_weights as sw
B, H, W = (int(v) for v in os.environ.get("GT_SHAPE", "8,256,256").split(","))  # GT_SHAPE=32,512,512: the configs[4] shard
T = 2
args = rs.make_args(maxseqlen=T); args.hidden_size = int(args.hidden_size); args.use_gpu = True
enc, dec = rsis_b200.FeatureExtractor(args), rsis_b200.RSIS(args)
enc.load_state_dict(sw.encoder_state_dict(1)); dec.load_state_dict(sw.decoder_state_dict(1))
enc.cuda().eval(); dec.cuda().eval()
x = sw.synthetic_images(5, B, H, W).cuda()
impl = ops.default_impl(); dev = x.device
masks = torch.empty((B, T, H, W), device=dev); classes = torch.empty((B, T, 21), device=dev); stops = torch.empty((B, T, 1), device=dev)
ws = dec.workspace(B, inference.feature_sizes(H, W), dev)
with torch.no_grad():
    keep = ws.encode_into(enc, dec, x, impl)
    ws.reset()
    dec.run_wavefront(ws, impl, T, classes, masks, stops)
torch.cuda.synchronize()
nlev = len(dec.clstm_list); p = ws.t & 1
offs = [sum(ws.hidden[:l]) for l in range(nlev)]
side_keys = torch.zeros_like(ws.side)
cells = []
for l, cell in enumerate(dec.clstm_list):
    xx = ws.X[l][p]
    cells.append(dict(x=xx, pc=ws.packs(dec, l)[1], c_prev=ws.c[l].t, side_max=side_keys, side_offset=offs[l],
                      h_out=ops.Act.empty(xx.n, xx.h, xx.w, cell.hidden_size, ops.FMT_F32, dev),
                      c_out=ops.Act.empty(xx.n, xx.h, xx.w, cell.hidden_size, ops.FMT_F32, dev),
                      h16_out=ops.Act.empty(xx.n, xx.h, xx.w, cell.hidden_size, ops.FMT_SPLIT_BF16, dev), gate_preact=ws.P[l]))
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

def time_launch(fn, iters=8):
    ts = []
    for it in range(iters + 2):
        flush.fill_(it & 0xff)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        if it >= 2:
            ts.append((e0, e1))
    torch.cuda.synchronize()
    v = [a.elapsed_time(b) * 1e3 for a, b in ts]
    return statistics.mean(v), min(v)

for spec in sys.argv[1:] or ["-;-"]:
    levels = list(range(nlev))
    if "@" in spec:
        spec, lv = spec.split("@")
        levels = [int(v) for v in lv.split(",")]
    sh, bn = (spec.split(";") + ["-"])[:2]
    for k, v in (("RSIS_B200_GROUP_SHARES", sh), ("RSIS_B200_GROUP_BN", bn)):
        if v == "-":
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    try:
        m, mn = time_launch(lambda: ops.convlstm_cell_group([cells[i] for i in levels]))
        print(f"levels {levels} shares {sh:>16s} bn {bn:>18s}: {m:6.1f} us (min {mn:.1f})", flush=True)
    except Exception as e:
        print(f"levels {levels} shares {sh} bn {bn}: FAILED {e}", flush=True)
