#!/usr/bin/env python
"""Debug (-DRSIS_DEBUG_TIMING build): timeline of every tcgen05 launch of ONE graph replay of the inference pass
(B=8, 256x256, T=10) -- grid start / end (%globaltimer) and block 0's in-kernel stamps, i.e. what a launch costs IN SITU
(L2 state, programmatic dependent launch overlap, neighbours) rather than alone.
  RSIS_B200_BUILD_DEBUG_TIMING=1 RSIS_B200_LIB=build/librsis_dbg.so python -m rsis_b200.build --force
  RSIS_B200_LIB=build/librsis_dbg.so python scripts/pass_trace.py 2> trace_shapes.txt"""
import ctypes, os, re, sys, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import rsis_b200
from rsis_b200 import ops, inference, _lib
from oracle import ref_shims as rs, synth_weights as sw

B, H, W, T = 8, 256, 256, 10
args = rs.make_args(maxseqlen=T); args.hidden_size = int(args.hidden_size); args.use_gpu = True
enc, dec = rsis_b200.FeatureExtractor(args), rsis_b200.RSIS(args)
enc.load_state_dict(sw.encoder_state_dict(1)); dec.load_state_dict(sw.decoder_state_dict(1))
enc.cuda().eval(); dec.cuda().eval()
x = sw.synthetic_images(5, B, H, W).cuda()
lib = _lib.load()
ROWS = 4096
buf = torch.zeros(ROWS * 24, dtype=torch.int64, device="cuda")
lib.rsis_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.rsis_debug_trace.restype = None
# shapes of the launches as they are set up (stderr of the library), captured through a pipe
LOG = os.environ.get("TRACE_LOG", "/tmp/rsis_trace_shapes.txt")
logf = open(LOG, "w")
saved = os.dup(2)
os.dup2(logf.fileno(), 2)
lib.rsis_debug_trace(buf.data_ptr(), ROWS)
s = inference.InferenceSession(args, enc, dec, x.shape, x.device)
torch.cuda.synchronize()
os.dup2(saved, 2)
logf.close()
log = open(LOG).read()
shapes = {}
for m in re.finditer(r"rsis trace (\d+): (.*)", log):
    shapes[int(m.group(1))] = m.group(2)
s.x.copy_(x)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for it in range(3):
    s.replay()
torch.cuda.synchronize()
buf.zero_()
flush.fill_(1)
s.replay()
torch.cuda.synchronize()
t = buf.view(ROWS, 24).cpu()
rows = [(i, t[i].tolist()) for i in range(ROWS) if t[i, 1] != 0]
rows.sort(key=lambda r: (~r[1][0]) & 0xFFFFFFFFFFFFFFFF)  # by grid start
GHZ = 1.965
inv = lambda v: (~v) & 0xFFFFFFFFFFFFFFFF if v >= 0 else (~(v + (1 << 64))) & 0xFFFFFFFFFFFFFFFF
t_first = None
prev_end = None
print(f"{len(rows)} traced launches in one replay; times in us; start/end = grid (all CTAs), rest = block 0, relative to ITS start")
print(f"{'id':>5s} {'start':>8s} {'dur':>6s} {'gap':>6s} | {'b0 start':>8s} {'pdl ok':>6s} {'A0 iss':>6s} {'MMA A':>7s} {'MMAs':>6s} {'exit':>6s} {'1st end':>7s} {'ctas':>4s} {'prod in':>7s} {'dec ok':>6s} {'aempty':>6s} {'exp_tx':>6s} {'epi acc':>7s} | shape")
tot = 0.0
for i, rw in rows:
    u = [v & 0xFFFFFFFFFFFFFFFF for v in rw]
    start = (~u[0]) & 0xFFFFFFFFFFFFFFFF
    end = u[1]
    if t_first is None:
        t_first = start
    st = u[2:18]
    c0 = u[19]
    rel = lambda slot: (st[slot] - c0) / GHZ / 1e3 if st[slot] else float("nan")
    gap = (start - prev_end) / 1e3 if prev_end is not None else 0.0
    print(f"{i:5d} {(start - t_first)/1e3:8.1f} {(end - start)/1e3:6.1f} {gap:6.1f} | {(u[18] - start)/1e3:8.1f} {rel(1):6.1f} {rel(2):6.1f} "
          f"{rel(4):7.1f} {rel(6):6.1f} {(u[21] - c0)/GHZ/1e3:6.1f} {(((~u[22]) & 0xFFFFFFFFFFFFFFFF) - start)/1e3:7.1f} {u[23]:4d} {rel(12):7.1f} {rel(13):6.1f} {rel(14):6.1f} {rel(15):6.1f} {rel(7):7.1f} | {shapes.get(i, '?')}")
    prev_end = end
    tot += (end - start) / 1e3
print(f"sum of grid durations {tot:.1f} us; span {(prev_end - t_first)/1e3:.1f} us")
