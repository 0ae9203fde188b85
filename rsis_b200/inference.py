"""Inference loop -- counterpart of /root/reference/src/test.py:16-50.

`test(args, encoder, decoder, x)` keeps the reference's signature and return value:
    (sigmoid(masks) [B,T,H,W], class_probs [B,T,C], sigmoid(stop) [B,T,1]),   T = args.maxseqlen.

The work is the same sequence the reference runs -- encoder once, T decoder steps, sigmoid -- but driven on NHWC
activations straight into the stacked output tensors (the `torch.cat`/`view`/`sigmoid` of test.py:46-50 are fused
into the mask-head and class/stop-head kernels), and, because every shape is static, captured once per
(shape, T) into a CUDA graph and replayed (`args.cuda_graph`, default True; SURVEY.md H4: the path is
launch-latency-bound at batch 8).
"""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Optional, Tuple

import torch

from . import ops


def _mask_size(h: int, w: int) -> Tuple[int, int]:
    # x1 = stem conv (k7, s2, p3): ceil(h/2); decoder level 4 lives at x1's size and is upsampled x2 (model.py:163)
    return 2 * ((h - 1) // 2 + 1), 2 * ((w - 1) // 2 + 1)


def feature_sizes(h: int, w: int):
    """Spatial sizes of the five decoder levels (x5 .. x1 of vision.py:11-21) for an h x w image: the stem conv
    (k7 s2 p3), the max-pool (k3 s2 p1) and the three stride-2 3x3 convs (p1) each map n -> (n - 1) // 2 + 1."""
    sizes = []
    for _ in range(5):
        h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        sizes.append((h, w))
    return sizes[::-1]


def run_eager(encoder, decoder, x: torch.Tensor, T: int, impl: int, out_masks: torch.Tensor,
              out_classes: torch.Tensor, out_stops: torch.Tensor, feats_op=None, ws=None):
    """Encoder once + T decoder steps, writing into the stacked outputs. Pure kernel launches on the current stream.

    tcgen05 kernel family: runs inside a `DecoderWorkspace` (`ws`, created when not given) -- the time-invariant
    share of every level's gates is computed once, and every step is 13 launches with no allocation.
    Returns the workspace (tcgen05) or the final state list (CUDA-core family)."""
    B, _, H, W = x.shape
    C = out_classes.shape[-1]
    Hm, Wm = _mask_size(H, W)
    if (Hm, Wm) != (H, W):
        # odd input sizes: the decoder's mask lives at 2*ceil(H/2) x 2*ceil(W/2); test.py:39-40 resizes the LOGITS to the
        # input size (nn.UpsamplingBilinear2d = bilinear, align_corners) before the sigmoid of test.py:50
        from . import postprocess
        logits = torch.empty((T, B, Hm, Wm), dtype=torch.float32, device=x.device)   # one contiguous [B,Hm,Wm] per step
        if ops.uses_tcgen05(impl):
            if ws is None:
                ws = decoder.workspace(B, feature_sizes(H, W), x.device)
            if feats_op is None:
                keep = ws.encode_into(encoder, decoder, x, impl)  # noqa: F841
            else:
                ws.load_feats(decoder, feats_op, impl)
            ws.reset()
            for t in range(T):
                decoder.step_ws(ws, impl, logits[t], out_classes[:, t], T * C, None, T, stop_prob=out_stops[:, t])
        else:
            if feats_op is None:
                _, feats_op = encoder.forward_act(x, impl)
            state = None
            for t in range(T):
                state = decoder.step_act(feats_op, state, impl, logits[t], out_classes[:, t], T * C, None, T,
                                         stop_prob=out_stops[:, t])
        resized = postprocess.resize_masks(logits.view(T * B, Hm, Wm), H, W)
        torch.sigmoid(resized.view(T, B, H, W).permute(1, 0, 2, 3), out=out_masks)
        return ws
    if ops.uses_tcgen05(impl):
        if ws is None:
            ws = decoder.workspace(B, feature_sizes(H, W), x.device)
        if feats_op is None:
            ws.reset(on_side_stream=True)                     # zero fills beside the encoder; encode_into joins the stream
            keep = ws.encode_into(encoder, decoder, x, impl)  # noqa: F841 -- alive until the join is enqueued
        else:
            ws.load_feats(decoder, feats_op, impl)
            ws.reset()
        # RSIS_B200_PIPELINE: 0 = one step after the other (12 launches in stream order), 1 = wavefront over
        # (level, step) on per-level streams, 2 = wavefront and no split-K in the cells (measured on B200 at
        # B=8 256x256 T=10: 3.60 / 3.40 / 3.24 ms per pass), 3 (default) = grouped wavefront launches
        mode = os.environ.get("RSIS_B200_PIPELINE", "3")
        last = ws.h[-1]
        if mode == "3" and last.c % 4 == 0 and last.c <= 16 and len(ws.h) <= ops.cell_group_max():
            # grouped wavefront: the independent cells of an anti-diagonal of (level, step) in ONE launch each
            decoder.run_wavefront(ws, impl, T, out_classes, out_masks, out_stops)
            return ws
        if mode != "0" and last.c % 4 == 0 and last.c <= 16:
            # wavefront schedule over (level, step); "2": additionally no split-K in the cells (fewer, longer CTAs)
            caps = os.environ.get("RSIS_B200_PIPE_CAPS", "")   # per-level CTA caps, e.g. "0,0,0,40,96" (tuning aid)
            caps = [int(v) for v in caps.split(",")] if caps else None
            split = os.environ.get("RSIS_B200_PIPE_SPLIT", "")  # per-level split-K switches, e.g. "1,1,0,0,0" (tuning aid)
            split = [int(v) for v in split.split(",")] if split else (mode != "2")
            decoder.run_pipelined(ws, impl, T, out_classes, out_masks, out_stops, split_k=split, cta_caps=caps)
            return ws
        for t in range(T):
            decoder.step_ws(ws, impl, None, out_classes[:, t], T * C, None, T, mask_prob=out_masks[:, t],
                            mask_prob_stride=T * H * W, stop_prob=out_stops[:, t])
        return ws
    if feats_op is None:
        _, feats_op = encoder.forward_act(x, impl)
    state = None
    for t in range(T):
        state = decoder.step_act(feats_op, state, impl, None, out_classes[:, t], T * C, None, T,
                                 mask_prob=out_masks[:, t], mask_prob_stride=T * H * W, stop_prob=out_stops[:, t])
    return state


class InferenceSession:
    """A captured CUDA graph of `test()` for one (batch shape, T): static input buffer -> static outputs."""

    def __init__(self, args, encoder, decoder, shape, device, impl: Optional[int] = None):
        self.T = int(args.maxseqlen)
        self.impl = ops.default_impl() if impl is None else impl
        self.shape = tuple(shape)
        B, _, H, W = self.shape
        self.encoder, self.decoder = encoder, decoder
        self.x = torch.zeros(self.shape, dtype=torch.float32, device=device)
        self.masks = torch.empty((B, self.T, H, W), dtype=torch.float32, device=device)
        self.classes = torch.empty((B, self.T, decoder.num_classes), dtype=torch.float32, device=device)
        self.stops = torch.empty((B, self.T, 1), dtype=torch.float32, device=device)
        self._tensors = [t for m in (encoder, decoder) for t in list(m.parameters()) + list(m.buffers())]
        self.weights_key = self._key()
        self.graph = None
        self.launches = 0
        encoder.eval()
        decoder.eval()
        self.ws = None
        if ops.uses_tcgen05(self.impl):
            with torch.cuda.device(device):
                self.ws = decoder.workspace(B, feature_sizes(H, W), device)
        # warm-up: builds the packed-weight caches and loads every kernel outside the capture
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side), torch.no_grad():
            run_eager(encoder, decoder, self.x, self.T, self.impl, self.masks, self.classes, self.stops, ws=self.ws)
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        g = torch.cuda.CUDAGraph()
        n0 = ops.launch_count()
        # the warm-up above built every weight pack, so inside the capture nothing re-packs: the convolutions may stream
        # their weights while the previous kernel is still running (ops.static_weights)
        with torch.cuda.graph(g), torch.no_grad(), ops.static_weights():
            run_eager(encoder, decoder, self.x, self.T, self.impl, self.masks, self.classes, self.stops, ws=self.ws)
        self.launches = ops.launch_count() - n0
        self.graph = g

    def _key(self):
        # per-tensor (storage, version): a parameter re-pointed without a version bump (GradBucket's `p.data = view`,
        # `.to()` / `.cuda()` re-materialisation) or a reload must invalidate the graph, which reads raw pointers
        return (tuple((t.data_ptr(), t._version) for t in self._tensors), ops.weights_epoch(), ops.bn_stats_epoch())

    def stale(self) -> bool:
        return self._key() != self.weights_key

    def replay(self):
        """Runs the captured path on the current contents of `self.x`; results land in masks/classes/stops."""
        self.graph.replay()

    def __call__(self, x: torch.Tensor):
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.masks, self.classes, self.stops


_sessions: "OrderedDict[tuple, InferenceSession]" = OrderedDict()
_MAX_SESSIONS = int(os.environ.get("RSIS_B200_MAX_SESSIONS", "8"))  # LRU bound: a graph + workspace per input shape


def test(args, encoder, decoder, x):
    """Runs the forward inference loop for the provided batch (test.py:16-50)."""
    ops.require_cuda(x, "test")
    T = int(args.maxseqlen)
    encoder.eval()
    decoder.eval()
    B = x.shape[0]
    if B == 1:
        # the reference itself fails here: `.squeeze()` at model.py:169 drops the batch dim and test.py:47 raises
        raise RuntimeError("rsis_b200.test: batch size 1 is not supported by the reference's test() either "
                           "(model.py:169 squeeze); call encoder/decoder directly or use B >= 2")
    impl = ops.default_impl()
    x = x.float()
    if getattr(args, "cuda_graph", True):
        key = (id(encoder), id(decoder), tuple(x.shape), T, x.device.index, impl)
        s = _sessions.get(key)
        if s is None or s.stale() or s.encoder is not encoder or s.decoder is not decoder:
            _sessions.pop(key, None)
            s = InferenceSession(args, encoder, decoder, x.shape, x.device, impl)
            _sessions[key] = s
            while len(_sessions) > _MAX_SESSIONS:
                _sessions.popitem(last=False)
        else:
            _sessions.move_to_end(key)
        masks, classes, stops = s(x)
        return masks.clone(), classes.clone(), stops.clone()
    H, W = x.shape[-2:]
    masks = torch.empty((B, T, H, W), dtype=torch.float32, device=x.device)
    classes = torch.empty((B, T, decoder.num_classes), dtype=torch.float32, device=x.device)
    stops = torch.empty((B, T, 1), dtype=torch.float32, device=x.device)
    with torch.no_grad():
        run_eager(encoder, decoder, x, T, impl, masks, classes, stops)
    return masks, classes, stops
