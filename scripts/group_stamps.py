#!/usr/bin/env python
"""Debug (library built with RSIS_B200_BUILD_DEBUG_TIMING=1): per-tile timeline of block 0 of a grouped cell launch on
the real decoder state.  usage: group_stamps.py <levels e.g. 4 or 3,4>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["RSIS_B200_DEBUG_TIMING"] = "1"
import torch
import rsis_b200
from rsis_b200 import ops, inference, _lib
from oracle import ref_shims as rs, synth_weights as sw
B, H, W, T = 8, 256, 256, 2
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
# point the library's debug-stamp buffer at lane 0's workspace: any tcgen05 convolution launched with it does that
_xa = ops.act_from_nchw(torch.rand((1, 64, 8, 16), device=dev), ops.FMT_SPLIT_BF16)
_pc = ops.PackedConv(torch.rand((64, 64, 1, 1), device=dev), None, None, want_umma=True)
ops.conv2d([_xa], _pc, impl=ops.IMPL_TCGEN05)
torch.cuda.synchronize()
wsb = _lib._workspaces[(torch.cuda.current_device(), 0)]
for spec in sys.argv[1:] or ["4", "3", "2", "1", "0", "0,1,2,3,4"]:
    sub = [int(v) for v in spec.split(",")]
    grp = [cells[i] for i in sub]
    for it in range(3):
        ops.convlstm_cell_group(grp)
    torch.cuda.synchronize()
    wsb[2048:2048 + 8 * 144].zero_()
    wsb[2048 + 8 * 199:2048 + 8 * 206].zero_()
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.convlstm_cell_group(grp); e1.record()
    torch.cuda.synchronize()
    st = wsb[2048:2048 + 8 * 12].view(torch.int64).cpu().tolist()
    t0 = st[0]
    ends = wsb[2048 + 8 * 200:2048 + 8 * 205].view(torch.int64).cpu().tolist()
    CLK = 1965.0   # stamps are SM cycles (clock64) of block 0's SM; cells' end times are %globaltimer (not comparable)
    print(f"levels {sub}: event {e0.elapsed_time(e1)*1e3:.1f} us; exit {(st[11]-t0)/CLK:.2f}; cell spans (globaltimer, us from "
          f"the first cell end): " + " ".join(f"{(t - min(ends[:len(sub)]))/1e3:.1f}" for t in ends[:len(sub)]))
    tl = wsb[2048 + 8 * 16:2048 + 8 * 144].view(torch.int64).cpu().view(8, 16)
    for role, rn in enumerate(["A issued", "MMA sees A", "MMAs issued", "epi sees acc", "epi released",
                               "epi sees P", "epi tile done", "epi tile top"]):
        vals = [f"{(int(t) - t0)/CLK:.1f}" for t in tl[role].tolist() if int(t) >= t0]
        print(f"      {rn:13s}: " + " ".join(vals))
