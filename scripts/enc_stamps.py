#!/usr/bin/env python
"""Debug: in-kernel cycle stamps (block 0) of the encoder's bottleneck convolutions at batch 8 -- needs the
-DRSIS_DEBUG_TIMING build:  RSIS_B200_LIB=build/dbg/librsis_b200.so python scripts/enc_stamps.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["RSIS_B200_DEBUG_TIMING"] = "1"
import torch
from rsis_b200 import ops, _lib
GHZ = 1.965
NAMES = ["start", "setup", "A0 issued", "A all issued", "A0 full", "B0 full", "MMAs issued", "epi tfull", "parked",
         "partials seen", "epi done", "exit", "prod entry", "prod decoded", "prod aempty ok", "prod expect_tx done"]
CASES = [  # name, Cin, H, W, Cout, k, residual, stride
    ("layer1.conv1", 256, 64, 64, 64, 1, False, 1), ("layer1.conv2", 64, 64, 64, 64, 3, False, 1), ("layer1.conv3", 64, 64, 64, 256, 1, True, 1),
    ("layer2.conv1", 512, 32, 32, 128, 1, False, 1), ("layer2.conv2", 128, 32, 32, 128, 3, False, 1), ("layer2.conv3", 128, 32, 32, 512, 1, True, 1),
    ("layer3.conv1", 1024, 16, 16, 256, 1, False, 1), ("layer3.conv2", 256, 16, 16, 256, 3, False, 1), ("layer3.conv3", 256, 16, 16, 1024, 1, True, 1),
    ("layer4.conv1", 2048, 8, 8, 512, 1, False, 1), ("layer4.conv2", 512, 8, 8, 512, 3, False, 1), ("layer4.conv3", 512, 8, 8, 2048, 1, True, 1),
    ("layer2.ds", 256, 64, 64, 512, 1, False, 2), ("layer3.ds", 512, 32, 32, 1024, 1, False, 2), ("layer2.0.conv2", 128, 64, 64, 128, 3, False, 2),
]
B = 8
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, Cin, H, W, Cout, k, has_res, stride in CASES:
    x = ops.act_from_nchw(torch.rand((B, Cin, H, W), device="cuda") - 0.5, ops.FMT_SPLIT_BF16)
    w = (torch.rand((Cout, Cin, k, k), device="cuda") - 0.5) * 0.05
    pc = ops.PackedConv(w, None, None, want_umma=True)
    y = ops.Act.empty(B, H // stride, W // stride, Cout, ops.FMT_SPLIT_BF16, "cuda")
    res = ops.act_from_nchw(torch.rand((B, Cout, H // stride, W // stride), device="cuda"), ops.FMT_SPLIT_BF16) if has_res else None
    _lib.workspace()
    ws = _lib._workspaces[(torch.cuda.current_device(), 0)]
    for it in range(2):
        ops.conv2d([x], pc, stride=stride, pad=k // 2, relu=True, residual=res, impl=ops.IMPL_TCGEN05, out=y)
    torch.cuda.synchronize()
    ws[2048:2048 + 8 * 208].zero_()
    if not os.environ.get("WARM"):
        flush.fill_(1)
    if os.environ.get("INSITU"):  # as inside the pass: activations (written by the previous launch) in L2, weights cold
        _ = x.t.view(torch.int16).sum()
        if res is not None:
            _ = res.t.view(torch.int16).sum()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.conv2d([x], pc, stride=stride, pad=k // 2, relu=True, residual=res, impl=ops.IMPL_TCGEN05, out=y); e1.record()
    torch.cuda.synchronize()
    st = ws[2048:2048 + 8 * 16].view(torch.int64).cpu().tolist()
    t0 = st[0]
    print(f"{name} ({B}x{Cin}x{H}x{W} -> {Cout} k{k} s{stride}): event {e0.elapsed_time(e1)*1e3:.1f} us; " +
          "; ".join(f"{n} {(t - t0)/GHZ/1e3:.2f}" for n, t in zip(NAMES, st) if t >= t0 and t != 0))
    tl = ws[2048 + 8 * 16:2048 + 8 * 208].view(torch.int64).cpu().view(12, 16)
    for role, rn in enumerate(["A issued", "MMA sees A", "MMAs issued", "epi sees acc", "epi released", "prod decode", "", "",
                               "MMA: A item full", "MMA: B box full", "B box issued", "A item issued"]):
        vals = [f"{(int(t) - t0)/GHZ/1e3:.1f}" for t in tl[role].tolist() if int(t) >= t0 and int(t) != 0]
        if vals:
            print(f"      {rn:13s}: " + " ".join(vals))
