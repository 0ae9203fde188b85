"""Design evidence (CPU, ~1 min): how much do the REFERENCE's own gradients move when only the forward convolution
operands are rounded the way the tcgen05 path rounds them (split-bf16, three products)?  Everything else -- torch
autograd, fp32 -- is unchanged.  Result (4 x 128x128, T=2): median relative-L2 change per parameter tensor 2.0e-2,
the same as the whole-step GPU comparison with the tcgen05 forward (profiles/r1x_grad_parity_auto_auto.json: 1.9e-2),
against 2.5e-3 with the exact-fp32 forward: ReLU masks / max-pool arg-maxes flip for values within ~1e-5 of a tie, and
each flip moves single gradient elements by percents.  Usage: python scripts/grad_sensitivity.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402
import train_parity as tp  # noqa: E402
from oracle import rsis_oracle as O, synth_weights as sw  # noqa: E402

kw = dict(batch=4, size=128, T=2, num_classes=5)
ref = tp.grads_through_oracle(**kw)


def run(conv):
    esd = {k: v.clone() for k, v in sw.encoder_state_dict(1).items()}
    dsd = {k: v.clone() for k, v in sw.decoder_state_dict(1, num_classes=5).items()}
    for sd in (esd, dsd):
        for k, v in sd.items():
            if v.is_floating_point() and "running_" not in k:
                v.requires_grad_(True)
    x = sw.synthetic_images(123, 4, 128, 128)
    wm, wc, ws = tp._loss_weights(4, 128, 2, 5)
    feats = O.feature_extractor(esd, x, conv=conv, bn=O._bn_train)
    hidden = None
    M, C, S = [], [], []
    for _ in range(2):
        m, c, s, hidden = O.rsis_step(dsd, feats, hidden, conv=conv)
        M.append(m)
        C.append(c)
        S.append(s)
    loss = tp._loss(M, C, S, wm, wc, ws, True, "cpu")
    loss.backward()
    g = {}
    for pre, sd in (("enc.", esd), ("dec.", dsd)):
        for k, v in sd.items():
            if v.requires_grad and v.grad is not None:
                g[pre + k] = v.grad
    return float(loss.detach()), g


loss, g = run(O.conv_bf16x3)
rows = []
for n, b in ref["grads"].items():
    if b is None or n in tp.ZERO_GRAD or n not in g:
        continue
    rows.append((float((g[n] - b).norm() / b.norm()), n))
rows.sort(reverse=True)
print("loss split-bf16x3 forward", loss, "fp32 forward", ref["loss"])
print("median relative L2 change of the parameter gradients:", rows[len(rows) // 2][0])
print("worst:", rows[:5])
