#!/usr/bin/env python
"""Debug: the hoisted-gate cell test case (tests/test_gpu_parity.py::test_convlstm_cell_hoisted_gates[shape0]) repeated,
printing the errors of h / c / h16 against the oracle -- to tell a deterministic near-tolerance result from a
timing-dependent one (run plain and under compute-sanitizer)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import rsis_b200
from rsis_b200 import ops
from oracle import rsis_oracle as O, ref_shims as rs

def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).abs().max() / b.abs().max())

shape = tuple(int(v) for v in sys.argv[1].split(",")) if len(sys.argv) > 1 else (3, 32, 32, 16, 16, 24)
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
g = torch.Generator().manual_seed(77)
B, up_c, skip_c, ch, H, W = shape
up = torch.rand((B, up_c, H, W), generator=g) * 2 - 1
skip = torch.rand((B, skip_c, H, W), generator=g) * 4 - 2
hp = torch.rand((B, ch, H, W), generator=g) * 2 - 1
cp = torch.rand((B, ch, H, W), generator=g) * 4 - 2
w = (torch.rand((4 * ch, up_c + skip_c + ch, 3, 3), generator=g) * 2 - 1) * 0.08
b = torch.rand(4 * ch, generator=g) - 0.5
href, cref = O.convlstm_cell(w, b, torch.cat([up, skip], 1), (hp, cp))
a = rs.make_args(); a.hidden_size = int(a.hidden_size)
cell = rsis_b200.ConvLSTMCell(a, up_c + skip_c, ch, 3, 1)
cell.load_state_dict({"Gates.weight": w, "Gates.bias": b})
cell.cuda()
pc_skip, pc_step = cell.packed_hoisted(up_c, skip_c)
F16 = ops.FMT_SPLIT_BF16
first = None
for it in range(reps):
    pre = ops.conv2d([ops.act_from_nchw(skip.cuda(), F16)], pc_skip, pad=1, impl=ops.IMPL_TCGEN05)
    x = ops.Act.zeros(B, H, W, up_c + ch, F16, "cuda")
    ops.convert(ops.act_from_nchw(up.cuda(), F16), F16, out=x.slice(0, up_c))
    ops.convert(ops.act_from_nchw(hp.cuda(), F16), F16, out=x.slice(up_c, ch))
    side = torch.zeros((B, ch), dtype=torch.int32, device="cuda")
    h16 = ops.Act.zeros(B, H, W, ch + 8, F16, "cuda")
    h, c = ops.convlstm_cell_x(x, pc_step, ops.act_from_nchw(cp.cuda(), ops.FMT_F32).t, side, 0,
                               h16_out=h16.slice(8, ch), impl=ops.IMPL_TCGEN05, gate_preact=pre)
    torch.cuda.synchronize()
    hn = h.nchw().cpu().clone()
    same = "first" if first is None else ("identical" if torch.equal(hn, first) else f"DIFFERS by {float((hn-first).abs().max()):.3e}")
    if first is None:
        first = hn
    print(f"rep {it}: h {rel(h.nchw(), href):.3e} c {rel(c.nchw(), cref):.3e} h16 "
          f"{rel(h16.float()[..., 8:].permute(0, 3, 1, 2), href):.3e} pre-vs-first {same}", flush=True)
