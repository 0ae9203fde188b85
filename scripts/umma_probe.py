#!/usr/bin/env python
"""Development probe for the tcgen05 convolution: each case runs in its own process (a device trap poisons the CUDA
context) and prints the tensor-relative error against torch CPU fp32.  Usage: python scripts/umma_probe.py [case]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    # name: (N, Cin(list), H, W, Cout, k, stride, relu, residual)
    "1x1_c64": (2, [64], 16, 16, 64, 1, 1, False, False),
    "1x1_c64_n128": (2, [64], 16, 16, 128, 1, 1, False, False),
    "1x1_c256_n512": (2, [256], 16, 16, 512, 1, 1, True, True),
    "3x3_c64": (2, [64], 16, 16, 64, 3, 1, True, False),
    "3x3_c128_big": (8, [128], 32, 32, 128, 3, 1, True, False),
    "3x3_s2": (2, [128], 16, 16, 128, 3, 2, True, False),
    "1x1_s2": (2, [256], 16, 16, 512, 1, 2, False, False),
    "3x3_c40_n24": (1, [40], 5, 5, 24, 3, 1, False, False),
    "3x3_c112_ragged": (2, [112], 12, 12, 48, 3, 1, False, False),
    "3x3_halo_c40_n32": (2, [40], 32, 24, 32, 3, 1, False, False),
    "3x3_halo_c96_n40": (2, [96], 32, 24, 40, 3, 1, True, True),
    "3x3_halo_c256_n256": (8, [256], 16, 16, 256, 3, 1, True, False),
    "3x3_halo_c320_n512": (4, [320], 16, 16, 512, 3, 1, False, False),
    "3x3_small_img": (8, [128], 8, 8, 128, 3, 1, False, False),
    "3x3_wide": (2, [64], 4, 256, 16, 3, 1, False, False),
    "sk5": (2, [2048], 8, 8, 128, 3, 1, False, False),
}


def run_case(name):
    import torch
    import torch.nn.functional as F
    from rsis_b200 import ops
    N, cins, H, W, Cout, k, s, relu, has_res = CASES[name]
    g = torch.Generator().manual_seed(1)
    xs = [torch.rand((N, c, H, W), generator=g) * 2 - 1 for c in cins]
    cin = sum(cins)
    w = (torch.rand((Cout, cin, k, k), generator=g) * 2 - 1) * (3.0 / (cin * k * k)) ** 0.5
    b = torch.rand(Cout, generator=g) - 0.5
    ref = F.conv2d(torch.cat(xs, 1), w, b, stride=s, padding=k // 2)
    res = None
    if has_res:
        res = torch.rand(ref.shape, generator=g) * 2 - 1
        ref = ref + res
    if relu:
        ref = F.relu(ref)
    pc = ops.PackedConv(w.cuda(), b.cuda(), None, src_channels=cins, want_umma=True)
    srcs = [ops.act_from_nchw(x.cuda(), ops.FMT_SPLIT_BF16) for x in xs]
    ra = ops.act_from_nchw(res.cuda(), ops.FMT_SPLIT_BF16) if res is not None else None
    y = ops.conv2d(srcs, pc, stride=s, pad=k // 2, relu=relu, residual=ra, out_fmt=ops.FMT_F32, impl=ops.IMPL_TCGEN05)
    torch.cuda.synchronize()
    got = y.nchw().cpu()
    err = float((got - ref).abs().max() / ref.abs().max())
    ys = ops.conv2d(srcs, pc, stride=s, pad=k // 2, relu=relu, residual=ra, out_fmt=ops.FMT_F32, impl=ops.IMPL_SIMT)
    err_simt = float((ys.nchw().cpu() - ref).abs().max() / ref.abs().max())
    bad = (got - ref).abs() > 1e-3 * ref.abs().max()
    print(f"{name}: rel err tcgen05 {err:.3e} simt {err_simt:.3e} bad {int(bad.sum())}/{bad.numel()}", flush=True)
    if bad.any():
        idx = bad.nonzero()[:5].tolist()
        print("   first bad (n,c,h,w):", idx, "bad channels:", sorted(set(bad.nonzero()[:, 1].tolist()))[:16],
              "bad rows(h):", sorted(set(bad.nonzero()[:, 2].tolist()))[:16], flush=True)


MODES = {  # environment knobs of conv_umma.cu
    "halo": {"RSIS_B200_HALO": "1", "RSIS_B200_HALO_BASEOFF": "0"},
    "halo+baseoff": {"RSIS_B200_HALO": "1", "RSIS_B200_HALO_BASEOFF": "1"},
    "tap": {"RSIS_B200_HALO": "0"},
}

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] in CASES:
        run_case(sys.argv[1])
    else:
        modes = sys.argv[1:] or list(MODES)
        for mode in modes:
            print(f"== mode {mode}", flush=True)
            for name in CASES:
                if mode != "halo" and "3x3" not in name:
                    continue
                env = dict(os.environ)
                env.update(MODES[mode])
                try:
                    r = subprocess.run([sys.executable, os.path.abspath(__file__), name], timeout=120,
                                       capture_output=True, text=True, env=env)
                    out = (r.stdout + r.stderr).strip().splitlines()
                    keep = [l for l in out if l.startswith(name) or "first bad" in l or "rror" in l or "timed out" in l]
                    print("\n".join(keep[-6:]) if keep else f"{name}: no output rc={r.returncode}", flush=True)
                except subprocess.TimeoutExpired:
                    print(f"{name}: TIMEOUT", flush=True)
