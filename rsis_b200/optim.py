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


def _base_param_multiplicity(encoder):
    """(parameter, how many times utils.py:34-52 yields it), in first-occurrence order.

    The reference walks `m.modules()` of each backbone child (conv1, bn1, layer1..4) and, for every module visited,
    all of its `.parameters()` recursively -- so a parameter is yielded once per module on the path from that child
    down to the parameter's owner.  That is the depth of its qualified name below the child: `layer3.5.conv2.weight`
    -> (layer3, layer3.5, layer3.5.conv2) = 3, `layer2.0.downsample.0.weight` -> 4, `conv1.weight` -> 1.  The first
    visit of a child (the child itself) yields its parameters in `named_parameters()` order, which fixes the
    first-occurrence order.  Pinned by tests/golden/optim_groups.json (made by the unmodified reference)."""
    base = encoder.base
    out = []
    for child in (base.conv1, base.bn1, base.layer1, base.layer2, base.layer3, base.layer4):
        for name, p in child.named_parameters():
            if p.requires_grad:
                out.append((p, name.count(".") + 1))
    return out


def get_base_params(args, model):
    """Same multiset and first-occurrence order as utils/utils.py:34-52 (ResNet branch), duplicates included."""
    pairs = _base_param_multiplicity(model)
    for p, _ in pairs:
        yield p
    for p, r in pairs:
        for _ in range(r - 1):
            yield p


def get_skip_params(model):
    """utils/utils.py:54-71: the five skip convolutions, then the five skip BatchNorms, level 1 to 5."""
    for kind in ("sk", "bn"):
        for level in range(1, 6):
            yield from getattr(model, f"{kind}{level}").parameters()


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

    # ---- checkpointing (utils/utils.py:86-111 saves / restores enc_opt.pt and dec_opt.pt; train.py:314-316 switches
    # the encoder group on mid-run) ---------------------------------------------------------------------------------
    def state_dict(self):
        """First / second moments and step counts per optimised parameter, keyed by the parameter's position in the
        bucket (stable for a given model), plus the hyper-parameters of each run."""
        b = self.bucket
        state = {}
        for idx, p in enumerate(b.params):
            off = b.offsets[id(p)]
            run = self._run_of(off)
            if run is None:
                continue
            state[idx] = {"step": int(run[5]), "exp_avg": self.m[off:off + p.numel()].view_as(p).clone(),
                          "exp_avg_sq": self.v[off:off + p.numel()].view_as(p).clone()}
        return {"state": state, "betas": self.betas, "eps": self.eps,
                "runs": [[int(r[0]), int(r[1]), float(r[2]), float(r[3]), int(r[4]), int(r[5])] for r in self.runs]}

    def _run_of(self, off: int):
        for run in self.runs:
            if run[0] <= off < run[0] + run[1]:
                return run
        return None

    def load_state_dict(self, sd):
        """Restores moments and step counts for every parameter present in BOTH the checkpoint and this optimiser
        (so a checkpoint written before the encoder group was enabled loads into an optimiser that has it)."""
        b = self.bucket
        steps = {}
        with torch.no_grad():
            for idx, st in sd["state"].items():
                p = b.params[int(idx)]
                off = b.offsets[id(p)]
                run = self._run_of(off)
                if run is None:
                    continue
                self.m[off:off + p.numel()].copy_(st["exp_avg"].reshape(-1))
                self.v[off:off + p.numel()].copy_(st["exp_avg_sq"].reshape(-1))
                steps.setdefault(id(run), (run, set()))[1].add(int(st["step"]))
        for run, seen in steps.values():
            if len(seen) != 1:
                raise ValueError("FusedAdam.load_state_dict: parameters of one run carry different step counts")
            run[5] = seen.pop()

    def add_groups(self, groups):
        """Enables further parameter groups (e.g. the encoder when `update_encoder` flips on, train.py:314-316) without
        touching the moments / step counts of the parameters already optimised."""
        b = self.bucket
        hyper = {}
        for g in groups:
            for p, r in zip(g["params"], g["repeats"]):
                hyper[id(p)] = (float(g["lr"]), float(g["weight_decay"]), int(r))
        for p in b.params:
            h = hyper.get(id(p))
            if h is None:
                continue
            off = b.offsets[id(p)]
            if self._run_of(off) is not None:
                raise ValueError("FusedAdam.add_groups: parameter is already optimised")
            self.runs.append([off, p.numel(), h[0], h[1], h[2], 0])
        self.runs.sort(key=lambda r: r[0])
        merged: List[list] = []
        for r in self.runs:
            last = merged[-1] if merged else None
            if last is not None and last[0] + last[1] == r[0] and last[2:] == r[2:]:
                last[1] += r[1]
            else:
                merged.append(r)
        self.runs = merged
