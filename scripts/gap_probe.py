#!/usr/bin/env python
"""Measures the per-launch cost of back-to-back dependent kernels inside a CUDA graph (launch gap + kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rsis_b200 import ops

def chain(name, N, C, H, W, Cout, k, n=100):
    x = ops.act_from_nchw(torch.rand((N, C, H, W), device="cuda") - 0.5, ops.FMT_SPLIT_BF16)
    w = (torch.rand((Cout, C, k, k), device="cuda") - 0.5) * 0.05
    pc = ops.PackedConv(w, None, None, want_umma=True)
    y = ops.Act.empty(N, H, W, Cout, ops.FMT_SPLIT_BF16, "cuda")
    ops.conv2d([x], pc, pad=k // 2, impl=ops.IMPL_TCGEN05, out=y)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            ops.conv2d([x], pc, pad=k // 2, impl=ops.IMPL_TCGEN05, out=y)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) * 1e3 / (5 * n):.2f} us per launch in a graph of {n}", flush=True)

chain("tiny 1x1 64->64 (1 CTA)", 1, 64, 8, 16, 64, 1)
chain("l3_conv1 1x1 1024->256", 8, 1024, 16, 16, 256, 1)
chain("l3_conv2 3x3 256->256", 8, 256, 16, 16, 256, 3)
chain("l3_conv3 1x1 256->1024", 8, 256, 16, 16, 1024, 1)
chain("l1_conv2 3x3 64->64 @64x64", 8, 64, 64, 64, 64, 3)
