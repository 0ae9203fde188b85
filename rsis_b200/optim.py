"""The optimiser step of the training loop -- counterpart of /root/reference/src/utils/utils.py:34-83
(`get_base_params`, `get_skip_params`, `get_optimizer`) and /root/reference/src/train.py:236-240,185-187
(SURVEY.md section 8f rank 4).

`reference_param_groups(args, encoder, decoder)` reproduces the reference's two Adam optimisers as plain descriptions:
  * decoder group: `list(decoder.parameters()) + list(get_skip_params(encoder))`, lr `args.lr`, weight decay
    `args.weight_decay` (train.py:238-239);
  * encoder group: `get_base_params(args, encoder)`, lr `args.lr_cnn`, weight decay `args.weight_decay_cnn`
    (train.py:236,240) -- a generator that yields every backbone parameter ONCE PER ENCLOSING MODULE (utils.py:45-52 walks
    `b[i].modules()` and, for each, all of its `.parameters()`): 3x for the convolutions / BatchNorms of a Bottleneck, 4x
    inside `downsample`, 1x for the stem.  torch's per-parameter loop (the only implementation in the reference's torch
    0.2; `foreach=False` today) then updates such a parameter that many times per `step()`.

`FusedAdam` applies those updates with ONE elementwise kernel per run of consecutive parameters that share
(lr, weight decay, repeats) over flat parameter / gradient / moment buffers (`GradBucket(..., flatten_params=True)`).
"""
from __future__ import annotations

from collections import Counter
from typing import List

import torch

from . import _lib, ops
from ._lib import check
from .autograd import GradBucket


def get_base_params(args, model):
    """utils/utils.py:34-52, statement for statement (ResNet branch): duplicates included."""
    b = [model.base.conv1, model.base.bn1, model.base.layer1, model.base.layer2, model.base.layer3, model.base.layer4]
    for i in range(len(b)):
        for j in b[i].modules():
            for k in j.parameters():
                if k.requires_grad:
                    yield k


def get_skip_params(model):
    """utils/utils.py:54-71."""
    b = [model.sk1.parameters(), model.sk2.parameters(), model.sk3.parameters(), model.sk4.parameters(),
         model.sk5.parameters(), model.bn1.parameters(), model.bn2.parameters(), model.bn3.parameters(),
         model.bn4.parameters(), model.bn5.parameters()]
    for j in range(len(b)):
        for i in b[j]:
            yield i


def reference_param_groups(args, encoder, decoder, update_encoder: bool = True):
    """[{'params': [...unique, in first-occurrence order...], 'repeats': [...], 'lr': .., 'weight_decay': ..}, ...] for
    `dec_opt` and (when update_encoder, train.py:186-187) `enc_opt`."""
    groups = []
    dec_list = [p for p in list(decoder.parameters()) + list(get_skip_params(encoder)) if p.requires_grad]
    lists = [(dec_list, float(getattr(args, "lr", 1e-3)), float(getattr(args, "weight_decay", 1e-6)))]
    if update_encoder:
        lists.append((list(get_base_params(args, encoder)), float(getattr(args, "lr_cnn", 1e-6)),
                      float(getattr(args, "weight_decay_cnn", 1e-6))))
    for plist, lr, wd in lists:
        cnt = Counter(id(p) for p in plist)
        uniq, seen = [], set()
        for p in plist:
            if id(p) not in seen:
                seen.add(id(p))
                uniq.append(p)
        groups.append({"params": uniq, "repeats": [cnt[id(p)] for p in uniq], "lr": lr, "weight_decay": wd})
    return groups


class FusedAdam:
    """torch.optim.Adam semantics (betas (0.9, 0.999), eps 1e-8, L2 weight decay, per-parameter loop) for parameter groups
    as `reference_param_groups` describes them, on a `GradBucket(..., flatten_params=True)`."""

    def __init__(self, bucket: GradBucket, groups, betas=(0.9, 0.999), eps: float = 1e-8):
        if bucket.flat_params is None:
            raise ValueError("FusedAdam needs GradBucket(..., flatten_params=True)")
        self.bucket = bucket
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        self.m = torch.zeros_like(bucket.flat)
        self.v = torch.zeros_like(bucket.flat)
        hyper = {}
        for g in groups:
            for p, r in zip(g["params"], g["repeats"]):
                if id(p) in hyper:
                    raise ValueError("a parameter may belong to one optimiser group only")
                hyper[id(p)] = (float(g["lr"]), float(g["weight_decay"]), int(r))
        # runs of consecutive bucket parameters with equal hyper-parameters -> one launch each
        self.runs: List[list] = []   # [offset, numel, lr, wd, repeats, steps_done]
        for p in bucket.params:
            h = hyper.get(id(p))
            if h is None:
                continue  # not optimised (e.g. the encoder when update_encoder is off)
            off = bucket.offsets[id(p)]
            last = self.runs[-1] if self.runs else None
            if last is not None and last[0] + last[1] == off and tuple(last[2:5]) == h:
                last[1] += p.numel()
            else:
                self.runs.append([off, p.numel(), h[0], h[1], h[2], 0])

    def step(self):
        lib = _lib.load()
        b = self.bucket
        ops.require_cuda(b.flat, "FusedAdam")
        st = _lib.stream_ptr()
        for run in self.runs:
            off, n, lr, wd, rep, done = run
            check(lib.rsis_adam_step(b.flat_params.data_ptr() + 4 * off, b.flat.data_ptr() + 4 * off,
                                     self.m.data_ptr() + 4 * off, self.v.data_ptr() + 4 * off, n, lr, self.betas[0],
                                     self.betas[1], self.eps, wd, done, rep, st), "adam_step")
            run[5] = done + rep
        _lib.count_launch(len(self.runs))
        # the kernel wrote the parameters through the flat buffer: their `_version`s did not move, so tell the derived
        # weight-pack caches explicitly
        ops.bump_weights_epoch()
