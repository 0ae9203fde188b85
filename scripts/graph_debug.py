"""Bring-up aid: which (batch, size, T) lets rsis_b200.training.TrainStep capture, and where capture breaks."""
import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rsis_b200
from rsis_b200.training import TrainStep
from oracle import synth_weights as sw, ref_shims as rs


def loss_fn(masks, classes, stops):
    return sum((torch.sigmoid(m) ** 2).mean() + (c ** 2).sum(-1).mean() + (s ** 2).mean()
               for m, c, s in zip(masks, classes, stops))


for (b, size, T) in [(4, 128, 3), (8, 128, 3), (4, 256, 3), (4, 128, 10), (8, 256, 10)]:
    args = rs.make_args(num_classes=21, maxseqlen=T)
    args.hidden_size = int(args.hidden_size)
    args.use_gpu = True
    enc, dec = rsis_b200.FeatureExtractor(args), rsis_b200.RSIS(args)
    enc.load_state_dict(sw.encoder_state_dict(1))
    dec.load_state_dict(sw.decoder_state_dict(1, num_classes=21))
    enc.cuda().train(); dec.cuda().train()
    x = sw.synthetic_images(123, b, size, size).cuda()
    step = TrainStep(enc, dec, T, loss_fn, cuda_graph=True, all_reduce=False)
    try:
        l = float(step(x)); l2 = float(step(x))
        print(f"OK   B={b} {size}x{size} T={T}: loss {l:.6f} {l2:.6f}", flush=True)
    except Exception as e:
        print(f"FAIL B={b} {size}x{size} T={T}: {type(e).__name__}: {str(e)[:200]}", flush=True)
        tb = traceback.format_exc().splitlines()
        print("\n".join(tb[-25:]), flush=True)
        try:
            torch.cuda.synchronize()
        except Exception as e2:
            print("sync after failure:", str(e2)[:200])
        break
    del step, enc, dec
    torch.cuda.empty_cache()
