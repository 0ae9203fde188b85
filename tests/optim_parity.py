"""Shared by the CPU host-logic test and the GPU test of rsis_b200.optim.FusedAdam: the reference's optimiser step --
two torch.optim.Adam instances built exactly like train.py:236-240 / utils.py:34-83 build them, duplicates included,
per-parameter loop (`foreach=False`, the only implementation the reference's torch had) -- on CPU copies of the
same modules, against FusedAdam on the modules under test."""
from __future__ import annotations

import copy
import warnings

import torch


def _args():
    from oracle import ref_shims as rs
    a = rs.make_args(num_classes=21, maxseqlen=2)
    a.hidden_size = int(a.hidden_size)
    a.use_gpu = True
    a.lr, a.lr_cnn, a.weight_decay, a.weight_decay_cnn = 1e-3, 1e-6, 1e-6, 1e-6
    return a


def run(device, steps=3, seed=0):
    import rsis_b200
    from rsis_b200 import optim
    from rsis_b200.autograd import GradBucket
    from oracle import synth_weights as sw
    args = _args()
    enc, dec = rsis_b200.FeatureExtractor(args), rsis_b200.RSIS(args)
    enc.load_state_dict(sw.encoder_state_dict(1))
    dec.load_state_dict(sw.decoder_state_dict(1))
    enc_ref, dec_ref = copy.deepcopy(enc), copy.deepcopy(dec)          # CPU twins for torch.optim.Adam
    enc.to(device)
    dec.to(device)
    # --- reference optimisers on the twins (utils.py:72-83; duplicates as get_base_params yields them) ---
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")   # "optimizer contains a parameter group with duplicate parameters"
        dec_opt = torch.optim.Adam([p for p in list(dec_ref.parameters()) + list(optim.get_skip_params(enc_ref))
                                    if p.requires_grad], lr=args.lr, weight_decay=args.weight_decay, foreach=False)
        enc_opt = torch.optim.Adam(list(optim.get_base_params(args, enc_ref)), lr=args.lr_cnn,
                                   weight_decay=args.weight_decay_cnn, foreach=False)
    # --- fused optimiser on the modules under test ---
    bucket = GradBucket(list(enc.parameters()) + list(dec.parameters()), flatten_params=True)
    fused = optim.FusedAdam(bucket, optim.reference_param_groups(args, enc, dec))
    gen = torch.Generator().manual_seed(seed)
    named = list(enc.named_parameters()) + list(dec.named_parameters())
    named_ref = list(enc_ref.named_parameters()) + list(dec_ref.named_parameters())
    for _ in range(steps):
        for (n, p), (_, q) in zip(named, named_ref):
            g = torch.randn(q.shape, generator=gen) * (0.01 + q.detach().abs().mean())
            q.grad = g.clone()
            p.grad.copy_(g.to(p.device))      # the bucket view
        dec_opt.step()
        enc_opt.step()
        fused.step()
    worst = 0.0
    for (n, p), (_, q) in zip(named, named_ref):
        if n.startswith("base.fc."):       # in neither optimiser: untouched
            assert torch.equal(p.detach().cpu(), q.detach()), n
            continue
        d = float((p.detach().cpu() - q.detach()).abs().max())
        s = float(q.detach().abs().max())
        worst = max(worst, d / max(s, 1e-12))
    return worst, fused
