#!/usr/bin/env python
"""Debug: in-kernel %globaltimer stamps of block 0 of one tcgen05 conv launch (library built with -DRSIS_DEBUG_TIMING)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["RSIS_B200_DEBUG_TIMING"] = "1"
import torch
from rsis_b200 import ops, _lib
NAMES = ["start", "setup done", "prod: first A issued", "prod: all issued", "mma: first A full", "mma: first B full",
         "mma: all issued", "epi: tfull", "epi: partial parked", "epi: all partials seen", "epi: done", "exit"]
CASES = {  # N, Cin, H, W, Cout, k
    "l3_conv1": (8, 1024, 16, 16, 256, 1), "l3_conv2": (8, 256, 16, 16, 256, 3), "l3_conv3": (8, 256, 16, 16, 1024, 1),
    "l1_conv2": (8, 64, 64, 64, 64, 3), "tiny": (1, 64, 8, 16, 64, 1),
}
for name, (N, Cin, H, W, Cout, k) in CASES.items():
    x = torch.rand((N, Cin, H, W), device="cuda") - 0.5
    w = (torch.rand((Cout, Cin, k, k), device="cuda") - 0.5) * 0.05
    pc = ops.PackedConv(w, None, None, want_umma=True)
    xa = ops.act_from_nchw(x, ops.FMT_SPLIT_BF16)
    y = ops.Act.empty(N, H, W, Cout, ops.FMT_SPLIT_BF16, "cuda")
    ptr, nbytes = _lib.workspace()
    ws = _lib._workspaces[(torch.cuda.current_device(), 0)]
    for it in range(3):
        ops.conv2d([xa], pc, pad=k // 2, impl=ops.IMPL_TCGEN05, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.conv2d([xa], pc, pad=k // 2, impl=ops.IMPL_TCGEN05, out=y); e1.record()
    torch.cuda.synchronize()
    st = ws[2048:2048 + 8 * 12].view(torch.int64).cpu().tolist()
    t0 = st[0]
    print(f"{name}: event {e0.elapsed_time(e1)*1e3:.1f} us; " + "; ".join(f"{n} {(t - t0)/1e3:.2f}" for n, t in zip(NAMES, st) if t >= t0))
    ws[2048:2048 + 96].zero_()

CELLS = {"L0": (8, 256, 128, 8, 8), "L1": (8, 320, 64, 16, 16), "L2": (8, 160, 32, 32, 32), "L3": (8, 80, 16, 64, 64),
         "L4": (8, 40, 8, 128, 128)}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, (N, Ct, Ch, H, W) in CELLS.items():
    x = ops.act_from_nchw(torch.rand((N, Ct, H, W), device="cuda") - 0.5, ops.FMT_SPLIT_BF16)
    w = (torch.rand((4 * Ch, Ct, 3, 3), device="cuda") - 0.5) * 0.05
    b = torch.rand(4 * Ch, device="cuda") - 0.5
    pc = ops.PackedConv(w, b, None, gate_interleave=True, src_channels=[Ct], want_umma=True)
    cprev = torch.rand((N, H, W, Ch), device="cuda")
    side = torch.zeros((N, 248), dtype=torch.int32, device="cuda")
    h = ops.Act.empty(N, H, W, Ch, ops.FMT_F32, "cuda"); c = ops.Act.empty(N, H, W, Ch, ops.FMT_F32, "cuda")
    h16 = ops.Act.empty(N, H, W, Ch, ops.FMT_SPLIT_BF16, "cuda")
    ws = _lib._workspaces[(torch.cuda.current_device(), 0)]
    for it in range(3):
        ops.convlstm_cell_x(x, pc, cprev, side, 0, h_out=h, c_out=c, h16_out=h16, impl=ops.IMPL_TCGEN05)
    torch.cuda.synchronize()
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.convlstm_cell_x(x, pc, cprev, side, 0, h_out=h, c_out=c, h16_out=h16, impl=ops.IMPL_TCGEN05); e1.record()
    torch.cuda.synchronize()
    st = ws[2048:2048 + 8 * 12].view(torch.int64).cpu().tolist()
    t0 = st[0]
    print(f"{name}: event {e0.elapsed_time(e1)*1e3:.1f} us; " + "; ".join(f"{n} {(t - t0)/1e3:.2f}" for n, t in zip(NAMES, st) if t >= t0))
    tl = ws[2048 + 8 * 16:2048 + 8 * 96].view(torch.int64).cpu().view(5, 16)
    for role, rn in enumerate(["A issued", "MMA sees A", "MMAs issued", "epi sees acc", "epi released"]):
        vals = [f"{(int(t) - t0)/1e3:.1f}" for t in tl[role].tolist() if int(t) >= t0]
        if vals:
            print(f"      {rn:13s}: " + " ".join(vals))
    ws[2048:2048 + 8 * 96].zero_()
