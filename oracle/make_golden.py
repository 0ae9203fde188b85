"""TEST INFRASTRUCTURE -- generate tests/golden/*.npz by running the UNMODIFIED reference here.

Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden

Each fixture records the seeds / shapes it was produced from, so tests regenerate the
inputs and weights with `oracle/synth_weights.py` and compare outputs only.  Large
tensors are stored as strided samples (`[..., ::s, ::s]`) to keep fixtures small.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from . import ref_shims as rs
from . import synth_weights as sw

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
WEIGHT_SEED = 1
IMAGE_SEED = 123  # reference default seed, args.py:19


def _np(t):
    return t.detach().cpu().numpy().astype(np.float32)


def _build(ref, num_classes, maxseqlen):
    args = rs.make_args(num_classes=num_classes, maxseqlen=maxseqlen)
    enc = ref.FeatureExtractor(args)
    dec = ref.RSIS(args)
    enc.load_state_dict(sw.encoder_state_dict(WEIGHT_SEED))
    dec.load_state_dict(sw.decoder_state_dict(WEIGHT_SEED, num_classes=num_classes))
    enc.eval()
    dec.eval()
    return args, enc, dec


def golden_e2e(ref, name, B, H, W, T, num_classes, stride, keep_feats_full):
    """`test()` (src/test.py:16-50) end to end + encoder feats; B==1 drives the modules directly."""
    args, enc, dec = _build(ref, num_classes, T)
    x = sw.synthetic_images(IMAGE_SEED, B, H, W)
    out = {"meta": np.array([WEIGHT_SEED, IMAGE_SEED, B, H, W, T, num_classes, stride], dtype=np.int64)}
    with torch.no_grad():
        feats = enc(x)
        if B >= 2:
            masks, classes, stops = ref.test(args, enc, dec, x)
        else:
            hidden = None
            ms, cs, ss = [], [], []
            for _ in range(T):
                m, c, s, hidden = dec(feats, hidden)
                ms.append(m)
                cs.append(c.view(1, -1))
                ss.append(s.view(1, -1))
            masks = torch.sigmoid(torch.cat(ms, 1))
            classes = torch.stack(cs, 1)
            stops = torch.sigmoid(torch.stack(ss, 1))
    for i, f in enumerate(feats):
        out[f"feat{i}"] = _np(f) if keep_feats_full else _np(f[:, ::4, ::2, ::2])
    out["masks"] = _np(masks[:, :, ::stride, ::stride])
    out["classes"] = _np(classes)
    out["stops"] = _np(stops)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items()})


def golden_cells(ref, name="cells_teacher_forced"):
    """Reference `ConvLSTMCell.forward` (clstm.py:19-62) on seeded inputs at the five level shapes, two steps each."""
    out = {}
    dsd = sw.decoder_state_dict(WEIGHT_SEED)
    outs = sw.skip_dims_out(128)
    B, S = 2, 8
    for lvl, ch in enumerate(outs):
        cin = 128 if lvl == 0 else 2 * outs[lvl - 1]
        cell = ref.ConvLSTMCell(rs.make_args(), cin, ch, 3, 1)
        cell.load_state_dict({"Gates.weight": dsd[f"clstm_list.{lvl}.Gates.weight"],
                              "Gates.bias": dsd[f"clstm_list.{lvl}.Gates.bias"]})
        x0 = sw._uniform(7, f"cell{lvl}.x0", (B, cin, S, S), -2.0, 2.0)
        x1 = sw._uniform(7, f"cell{lvl}.x1", (B, cin, S, S), -2.0, 2.0)
        with torch.no_grad():
            h0, c0 = cell(x0, None)
            h1, c1 = cell(x1, (h0, c0))
        out[f"l{lvl}_h0"], out[f"l{lvl}_c0"] = _np(h0), _np(c0)
        out[f"l{lvl}_h1"], out[f"l{lvl}_c1"] = _np(h1), _np(c1)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)
    print(name, len(out), "tensors")


def main():
    if not rs.available():
        print("reference tree not present; nothing generated", file=sys.stderr)
        return 1
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.manual_seed(0)
    ref = rs.load_reference()
    golden_cells(ref)
    golden_e2e(ref, "e2e_b2_64x64_t3", 2, 64, 64, 3, 21, 1, True)
    golden_e2e(ref, "cfg1_b1_256x256_t5", 1, 256, 256, 5, 21, 8, False)   # BASELINE.json configs[0]
    golden_e2e(ref, "cfg2_b8_256x256_t10", 8, 256, 256, 10, 21, 16, False)  # BASELINE.json configs[1]
    golden_e2e(ref, "e2e_b2_96x160_t4_c9", 2, 96, 160, 4, 9, 4, False)     # non-square, Cityscapes-like classes
    return 0


if __name__ == "__main__":
    sys.exit(main())
