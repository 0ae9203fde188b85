"""Bring-up aid: where does the soft-IoU cost call spend its time? (CUDA events, L2 flushed between iterations)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rsis_b200 import objectives, _lib

dev = torch.device("cuda", 0)
B, G, HW = 8, 20, 65536
gen = torch.Generator().manual_seed(3)
logits = (torch.randn((B, HW), generator=gen) * 2).to(dev)
y = (torch.rand((B, G, HW), generator=gen) < 0.25).to(dev)
yf, yu = y.float(), y.to(torch.uint8)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
out = torch.empty((B, G), device=dev)
big = torch.empty_like(yf)


def timed(fn, n=10):
    evs = []
    for it in range(n + 3):
        flush.fill_(it)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        if it >= 3:
            evs.append((e0, e1))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / len(evs) * 1e3


lib = _lib.load()
ws = torch.zeros(4096, device=dev)
st = torch.cuda.current_stream().cuda_stream


def raw(gt, u8, b=B, g=G, hw=HW):
    return lambda: lib.rsis_soft_iou_cost(logits.data_ptr(), gt.data_ptr(), u8, b, g, hw, 1e-6, 1.0, ws.data_ptr(),
                                          out.data_ptr(), g, 1, None, None, st)


print("copy 42 MB (torch)        %.1f us" % timed(lambda: big.copy_(yf)))
print("empty event pair           %.1f us" % timed(lambda: None))
print("wrapper f32                %.1f us" % timed(lambda: objectives.soft_iou_cost_matrix(logits, yf, 1.0, out=out)))
print("raw C call f32             %.1f us" % timed(raw(yf, 0)))
print("raw C call u8              %.1f us" % timed(raw(yu, 1)))
print("raw C call f32 rows mode   %.1f us (160 rows x 65536, G=1)" % timed(raw(yf, 0, b=8, g=1)))
print("torch reference f32        %.1f us" % timed(lambda: (torch.sigmoid(logits)[:, None] * yf).sum(-1)))
for i in range(5):
    import subprocess
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_backward.py", "-m", "gpu", "-q", "-k", "lstm_gates"],
                       capture_output=True, text=True)
    print("lstm_gates run", i, r.stdout.strip().splitlines()[-1])
