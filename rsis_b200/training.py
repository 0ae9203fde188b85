"""One training step of the hot path -- the forward/backward portion of `runIter()` (/root/reference/src/train.py:71-115,
178-187): train-mode encoder, T decoder steps, a caller-supplied loss over the per-step outputs, `loss.backward()`,
and the ONE data-parallel gradient all-reduce.

    step = TrainStep(encoder, decoder, T, loss_fn, cuda_graph=True)
    loss = step(x)            # gradients are in step.bucket.flat / every parameter's .grad; then optimiser.step()
                              # (or pass optimizer=rsis_b200.optim.FusedAdam(...) and the step includes it)

`loss_fn(masks, classes, stops) -> scalar` receives the lists of per-step outputs (mask logits [B,1,H,W], class
probabilities [B,C], stop logits [B,1]); the reference's criteria (train.py:159-176) are outside the hot path.

cuda_graph=True: the whole step -- weight re-packing, forward, loss, backward -- is captured ONCE per input shape into
a CUDA graph and replayed (the step is ~1900 kernel launches of 5-30 us: eager execution is bound by the host).
Because the packing kernels are inside the graph, a replay always sees the current parameter values (in-place
optimiser updates).  The loss function must then be capturable (no host synchronisation, no data-dependent control
flow); with cuda_graph=False anything goes.  The all-reduce runs after the replay, outside the graph.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import ops
from .autograd import GradBucket, _DgradCache


class TrainStep:
    def __init__(self, encoder, decoder, T: int, loss_fn: Callable, bucket: Optional[GradBucket] = None,
                 cuda_graph: bool = True, all_reduce: bool = True, optimizer=None, precision: Optional[str] = None):
        """optimizer (optional): an `rsis_b200.optim.FusedAdam` over `bucket` (which then must have been built with
        flatten_params=True); its step runs after the all-reduce, outside the captured graph (a handful of launches
        whose bias-correction coefficients change every step)."""
        self.enc, self.dec, self.T, self.loss_fn = encoder, decoder, int(T), loss_fn
        if precision is not None:  # "bf16": BASELINE.json configs[3] (single-pass bf16 products); None: the modules' own
            ops.precision(precision)  # validates the name
            encoder.precision = decoder.precision = precision
        self.optimizer = optimizer
        if optimizer is not None and bucket is None:
            bucket = optimizer.bucket
        self.bucket = bucket if bucket is not None else GradBucket(list(encoder.parameters()) + list(decoder.parameters()))
        self.use_graph = bool(cuda_graph)
        self.do_all_reduce = bool(all_reduce)
        self.graph = None
        self.static_x = None
        self.static_loss = None
        self.static_outs = None

    # -- the step itself (eager kernels on the current stream) -----------------------------------------------------
    def _run(self, x: torch.Tensor):
        self.bucket.zero()
        feats = self.enc(x)
        hidden = None
        masks, classes, stops = [], [], []
        for _ in range(self.T):
            m, c, s, hidden = self.dec(feats, hidden)
            masks.append(m)
            classes.append(c)
            stops.append(s)
        loss = self.loss_fn(masks, classes, stops)
        loss.backward()
        return loss.detach(), (masks, classes, stops)

    def _drop_pack_caches(self):
        """Forget every derived weight pack so that the next run rebuilds (= captures) them."""
        enc, dec = self.enc, self.dec
        enc.base._packed_tr = None
        enc.base._packed = None
        enc._packed_tr = None
        enc._packed = None
        for cell in dec.clstm_list:
            cell._packed = {}
        for m in (enc, dec):
            object.__setattr__(m, "_rsis_dgrad_cache", _DgradCache())

    def _capture(self, x: torch.Tensor):
        ops.require_cuda(x, "TrainStep")
        if not (self.enc.training and self.dec.training):
            raise RuntimeError("TrainStep: encoder and decoder must be in train() mode")
        self.static_x = x.clone()
        # warm-up on a side stream (allocates workspaces, sets kernel attributes); the running statistics the warm-up
        # steps advance are put back afterwards
        saved = [b.clone() for b in self.enc.buffers()]
        cur = torch.cuda.current_stream(x.device)
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(2):
                self._run(self.static_x)
        cur.wait_stream(side)
        torch.cuda.synchronize(x.device)
        with torch.no_grad():
            for b, s in zip(self.enc.buffers(), saved):
                b.copy_(s)
        self._drop_pack_caches()
        g = torch.cuda.CUDAGraph()
        # thread_local: the backward runs on autograd's worker thread (its launches on the capturing stream are
        # recorded all the same); only this thread's own calls are policed for capture-unsafe APIs
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            loss, outs = self._run(self.static_x)
        self.graph, self.static_loss, self.static_outs = g, loss, outs

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if not self.use_graph:
            loss, self.static_outs = self._run(x)
        else:
            if self.graph is None or self.static_x.shape != x.shape:
                self._capture(x)
            self.static_x.copy_(x, non_blocking=True)
            self.graph.replay()
            # the replay advanced every BatchNorm's running statistics on the device without any Python-side signal:
            # eval-side packs / captured inference graphs that folded the old statistics are stale
            ops.bump_bn_stats_epoch()
            loss = self.static_loss
        if self.do_all_reduce:
            self.bucket.all_reduce()
        if self.optimizer is not None:
            self.optimizer.step()
        return loss

    @property
    def outputs(self):
        """(masks, classes, stops) of the last step (static buffers in graph mode: valid until the next call)."""
        return self.static_outs
