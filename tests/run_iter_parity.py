"""Shared by the CPU pinning test and the GPU test of the whole training iteration (`runIter`, train.py:56-197)."""
from __future__ import annotations

import numpy as np
import torch


def oracle_run_iter(num_classes=5):
    """oracle.run_iter over the CPU oracle with autograd -> (losses, y_class_perm, grads)."""
    from oracle import rsis_oracle as O, synth_weights as sw
    from oracle.run_iter_ref import iter_args, iter_inputs
    args = iter_args()
    esd = {k: v.clone() for k, v in sw.encoder_state_dict(1).items()}
    dsd = {k: v.clone() for k, v in sw.decoder_state_dict(1, num_classes=num_classes).items()}
    for sd in (esd, dsd):
        for k, v in sd.items():
            if v.is_floating_point() and "running_" not in k:
                v.requires_grad_(True)
    x, y_mask, y_class, sw_mask, sw_class = iter_inputs(gt=args.gt_maxseqlen, num_classes=num_classes)
    loss, parts, perm = O.run_iter(lambda xx: O.feature_extractor(esd, xx, bn=O._bn_train),
                                   lambda f, h: O.rsis_step(dsd, f, h), x, y_mask, y_class, sw_mask, sw_class, args)
    loss.backward()
    grads = {"enc." + k: v.grad for k, v in esd.items() if v.requires_grad and v.grad is not None}
    grads.update({"dec." + k: v.grad for k, v in dsd.items() if v.grad is not None})
    return [float(loss)] + [float(p) for p in parts], perm, grads


def modules_run_iter(device, num_classes=5, precision=None):
    """The same recipe over the rsis_b200 modules, the fused soft-IoU cost / loss kernels and the device matching."""
    import rsis_b200
    from rsis_b200 import objectives as OBJ
    from oracle import rsis_oracle as O, synth_weights as sw
    from oracle.run_iter_ref import iter_args, iter_inputs
    from train_parity import _args
    args = iter_args()
    margs = _args(num_classes, args.maxseqlen)
    if precision is not None:
        margs.precision = precision       # "bf16": BASELINE.json configs[3] single-pass training mode
    enc, dec = rsis_b200.FeatureExtractor(margs), rsis_b200.RSIS(margs)
    enc.load_state_dict(sw.encoder_state_dict(1))
    dec.load_state_dict(sw.decoder_state_dict(1, num_classes=num_classes))
    enc.to(device).train()
    dec.to(device).train()
    x, y_mask, y_class, sw_mask, sw_class = (t.to(device) for t in iter_inputs(gt=args.gt_maxseqlen,
                                                                                 num_classes=num_classes))
    on_gpu = torch.device(device).type == "cuda"
    kw = {}
    if on_gpu:  # the CUDA kernels; on CPU (fake ABI) the oracle's own statements stand in for them
        kw = dict(cost_matrix=lambda m, y, w: OBJ.soft_iou_cost_matrix(m, y, w),
                  match_fn=lambda m, c, sc: OBJ.match([m, None], [c, None], sc)[:2],
                  iou_loss=lambda yt, yp, s: OBJ.softIoULoss()(yt, yp, s))
    loss, parts, perm = O.run_iter(enc, dec, x, y_mask, y_class, sw_mask, sw_class, args, **kw)
    loss.backward()
    grads = {"enc." + n: p.grad.detach().float().cpu() for n, p in enc.named_parameters() if p.grad is not None}
    grads.update({"dec." + n: p.grad.detach().float().cpu() for n, p in dec.named_parameters() if p.grad is not None})
    return [float(loss)] + [float(p) for p in parts], perm.cpu(), grads


def digest(g: torch.Tensor):
    flat = g.detach().reshape(-1).double()
    idx = torch.linspace(0, flat.numel() - 1, 32).long()
    return np.concatenate([[float(flat.norm()), float(flat.sum())], flat[idx].numpy()])
