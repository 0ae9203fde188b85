#!/usr/bin/env python
"""Representative ResNet-101 bottleneck convolutions at the configs[1] encoder shapes (B=8, 256x256 input), each run
`--iters` times through rsis_conv2d (tcgen05 family).  Meant to be wrapped in `ncu --set full -k regex:conv_umma`;
without ncu it prints CUDA-event times (L2 flushed before every launch)."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rsis_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--only", default="", help="substring filter on the case names, e.g. layer4")
a = ap.parse_args()
B = a.batch
CASES = [  # name, Cin, H, W, Cout, k, residual
    ("layer1.conv1", 256, 64, 64, 64, 1, False), ("layer1.conv2", 64, 64, 64, 64, 3, False), ("layer1.conv3", 64, 64, 64, 256, 1, True),
    ("layer2.conv1", 512, 32, 32, 128, 1, False), ("layer2.conv2", 128, 32, 32, 128, 3, False), ("layer2.conv3", 128, 32, 32, 512, 1, True),
    ("layer3.conv1", 1024, 16, 16, 256, 1, False), ("layer3.conv2", 256, 16, 16, 256, 3, False), ("layer3.conv3", 256, 16, 16, 1024, 1, True),
    ("layer4.conv1", 2048, 8, 8, 512, 1, False), ("layer4.conv2", 512, 8, 8, 512, 3, False), ("layer4.conv3", 512, 8, 8, 2048, 1, True),
]
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for name, Cin, H, W, Cout, k, has_res in [c for c in CASES if a.only in c[0]]:
    x = ops.act_from_nchw(torch.rand((B, Cin, H, W), device="cuda") - 0.5, ops.FMT_SPLIT_BF16)
    w = (torch.rand((Cout, Cin, k, k), device="cuda") - 0.5) * 0.05
    pc = ops.PackedConv(w, None, None, want_umma=True)
    y = ops.Act.empty(B, H, W, Cout, ops.FMT_SPLIT_BF16, "cuda")
    res = ops.act_from_nchw(torch.rand((B, Cout, H, W), device="cuda"), ops.FMT_SPLIT_BF16) if has_res else None
    ts = []
    for it in range(a.iters):
        mode = os.environ.get("CACHE", "cold")  # cold | insitu (activations in L2, weights from HBM) | warm | wwarm (weights only)
        if mode != "warm":
            flush.fill_(it)
        if mode == "insitu":
            _ = x.t.view(torch.int16).sum()
            if res is not None:
                _ = res.t.view(torch.int16).sum()
        if mode == "wwarm":
            _ = pc.w_umma.view(torch.int16).sum() if hasattr(pc, "w_umma") and pc.w_umma is not None else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.conv2d([x], pc, pad=k // 2, relu=True, residual=res, impl=ops.IMPL_TCGEN05, out=y)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    gflop = 2.0 * B * H * W * Cout * Cin * k * k / 1e9
    print(f"{name}: {B}x{Cin}x{H}x{W} -> {Cout} k{k}: {min(ts):.1f} us ({gflop / min(ts) / 1e3:.1f} TFLOP/s fp32-equivalent)", flush=True)
