"""TEST INFRASTRUCTURE -- run the reference's own training iteration, `runIter` (/root/reference/src/train.py:56-197), in
the build container and read back its losses, matching and gradients.

`train.py` as a file cannot be imported under Python 3 (print statements from line 254 on; data-loader / visdom imports
at the top), but the text of `runIter` itself is Python-3 clean.  This module extracts exactly that
function's source text from the file IN MEMORY (nothing is copied into the repository), applies one textual shim --
`.data[0]` (torch <= 0.3 scalar read, train.py:189) -> `.item()` -- and executes it with the names it needs bound to
the UNMODIFIED reference objects (`match`, `softIoU` from utils/hungarian.py with munkres stubbed by scipy, see
ref_shims.py).  Runtime shim: `torch.masked_select` accepts the uint8 masks of utils/objectives.py:13,23,32
(`sw.byte()`), which current torch rejects.  Used by oracle/make_golden.py only.
"""
from __future__ import annotations

import os
from argparse import Namespace

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Variable

from . import ref_shims as rs


def _run_iter_source() -> str:
    path = os.path.join(rs.REFERENCE_ROOT, "src", "train.py")
    lines = open(path).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("def runIter("))
    end = next(i for i, l in enumerate(lines) if i > start and l.startswith("def "))
    return "\n".join(lines[start:end]).replace(".data[0]", ".item()")


def iter_args(maxseqlen=3, gt_maxseqlen=4, **kw) -> Namespace:
    a = Namespace(maxseqlen=maxseqlen, gt_maxseqlen=gt_maxseqlen, curriculum_learning=False, limit_seqlen_to=maxseqlen,
                  iou_weight=1.0, class_weight=0.1, stop_weight=0.5, use_class_loss=True, use_stop_loss=True,
                  use_gpu=False, update_encoder=True)
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def iter_inputs(seed=21, b=2, size=64, gt=4, num_classes=5):
    """Seeded training batch: images, blob-like ground-truth masks [B, gtT, HW] (n_obj real objects per image, zero
    padding rows), classes, and the sample-weight masks of dataset.py:142-146 (first n_obj ones)."""
    from . import synth_weights as sw
    gen = torch.Generator().manual_seed(seed)
    x = sw.synthetic_images(123, b, size, size)
    low = torch.rand((b, gt, size // 8, size // 8), generator=gen)
    y = (torch.nn.functional.interpolate(low, size=(size, size), mode="bilinear", align_corners=True) > 0.55).float()
    n_obj = [gt - 1, max(gt - 2, 1)][:b] + [gt] * max(b - 2, 0)
    sw_mask = torch.zeros((b, gt))
    for i, n in enumerate(n_obj):
        sw_mask[i, :n] = 1
        y[i, n:] = 0
    y_class = torch.randint(1, num_classes, (b, gt), generator=gen) * sw_mask.long()
    return x, y.view(b, gt, -1), y_class, sw_mask, sw_mask.clone()


def run_reference_iter(num_classes=5, args=None):
    """Returns dict(losses, perm_class, perm_mask_sum, grads{name: tensor}) from the reference's runIter in 'train' mode
    (optimisers with lr = 0: `step()` leaves the weights alone, the gradients stay in `.grad`)."""
    from . import synth_weights as sw
    ref = rs.load_reference()
    import hungarian as ref_h           # noqa: E402  (reference files, unmodified)
    import objectives as ref_o          # noqa: E402
    args = args or iter_args()
    margs = rs.make_args(num_classes=num_classes, maxseqlen=args.maxseqlen)
    enc, dec = ref.FeatureExtractor(margs), ref.RSIS(margs)
    enc.load_state_dict(sw.encoder_state_dict(1))
    dec.load_state_dict(sw.decoder_state_dict(1, num_classes=num_classes))
    x, y_mask, y_class, sw_mask, sw_class = iter_inputs(gt=args.gt_maxseqlen, num_classes=num_classes)
    crits = (ref_o.softIoULoss(), ref_o.MaskedNLLLoss(balance_weight=None), ref_o.MaskedBCELoss(balance_weight=None))
    optims = (torch.optim.SGD(enc.parameters(), lr=0.0), torch.optim.SGD(dec.parameters(), lr=0.0))
    ns = {"torch": torch, "nn": nn, "np": np, "Variable": Variable, "match": ref_h.match, "softIoU": ref_h.softIoU}
    exec(compile(_run_iter_source(), "<train.py:runIter>", "exec"), ns)
    orig = torch.masked_select
    torch.masked_select = lambda inp, mask, **kw: orig(inp, mask.bool() if mask.dtype == torch.uint8 else mask, **kw)
    try:
        losses, outs, perms = ns["runIter"](args, enc, dec, Variable(x), Variable(y_mask), Variable(y_class),
                                            Variable(sw_mask), Variable(sw_class), crits, optims, mode="train")
    finally:
        torch.masked_select = orig
    grads = {"enc." + n: p.grad.detach().clone() for n, p in enc.named_parameters() if p.grad is not None}
    grads.update({"dec." + n: p.grad.detach().clone() for n, p in dec.named_parameters() if p.grad is not None})
    return dict(losses=[float(v) for v in losses], perm_mask_sum=perms[0].sum(-1).numpy(), perm_class=perms[1].numpy(),
                grads=grads, bn_batches=int(enc.base.bn1.num_batches_tracked))
