"""FeatureExtractor / RSIS -- counterparts of /root/reference/src/modules/model.py:15-70 and :72-184.

Same constructor arguments (`args` namespace), attribute names (`base`, `sk1..sk5`, `bn1..bn5`, `clstm_list`,
`conv_out`, `fc_class`, `fc_stop`), state_dict keys/shapes (661 + 16) and return structure as the reference, so
`train.py` / `eval.py` / `test.py` / `utils/utils.py` can use them unchanged.  The nn.Parameters (OIHW float32)
stay the source of truth; kernel-ready packed copies are derived caches rebuilt when a parameter changes.

Tensors crossing this boundary are logical NCHW float32 (`.shape == [N,C,H,W]`) stored channels-last, which is the
layout the kernels use (NHWC) -- returning them is zero-copy.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from .. import _lib, ops
from ..ops import Act, PackedConv
from .clstm import ConvLSTMCell
from .vision import ResNet101

SKIP_DIMS_IN = {"resnet101": [2048, 1024, 512, 256, 64]}  # utils/utils.py:129-131


def get_skip_dims(model_name):
    if model_name not in SKIP_DIMS_IN:
        raise Exception("The base model you chose is not supported !")  # model.py:37 (only resnet101 is in scope)
    return SKIP_DIMS_IN[model_name]


class FeatureExtractor(nn.Module):
    """Returns base network to extract visual features from image (model.py:15-70)."""

    def __init__(self, args):
        super().__init__()
        skip_dims_in = get_skip_dims(args.base_model)
        # The reference loads ImageNet weights from the network here (model.py:30-31); this build has no network
        # access, so the base starts from torchvision's random initialisation and real weights arrive through
        # load_state_dict (the reference's own checkpoint path, utils/utils.py:97-111).
        self.base = ResNet101()
        # "bf16": the TRAINING-mode forward / backward of this module run single-pass bf16 tensor-core products
        # (BASELINE.json configs[3]; loss-level tolerance).  Inference and the default stay fp32-grade (split bf16).
        self.precision = getattr(args, "precision", None)
        self.hidden_size = int(args.hidden_size)
        self.kernel_size = int(args.kernel_size)
        self.padding = 0 if self.kernel_size == 1 else 1
        hs, k, p = self.hidden_size, self.kernel_size, self.padding
        self.sk5 = nn.Conv2d(skip_dims_in[0], hs, k, padding=p)
        self.sk4 = nn.Conv2d(skip_dims_in[1], hs, k, padding=p)
        self.sk3 = nn.Conv2d(skip_dims_in[2], hs // 2, k, padding=p)
        self.sk2 = nn.Conv2d(skip_dims_in[3], hs // 4, k, padding=p)
        self.sk1 = nn.Conv2d(skip_dims_in[4], hs // 8, k, padding=p)
        self.bn5 = nn.BatchNorm2d(hs)
        self.bn4 = nn.BatchNorm2d(hs)
        self.bn3 = nn.BatchNorm2d(hs // 2)
        self.bn2 = nn.BatchNorm2d(hs // 4)
        self.bn1 = nn.BatchNorm2d(hs // 8)
        self._packed = None
        self._packed_key = None

    def _heads(self):
        return [(self.sk5, self.bn5), (self.sk4, self.bn4), (self.sk3, self.bn3), (self.sk2, self.bn2),
                (self.sk1, self.bn1)]

    def packed_heads(self, want_umma: bool):
        tensors = []
        for sk, bn in self._heads():
            tensors += [sk.weight, sk.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var]
        key = (tuple((t.data_ptr(), t._version) for t in tensors), want_umma, ops.weights_epoch(), ops.bn_stats_epoch())
        if self._packed is None or self._packed_key != key:
            self._packed = [PackedConv(sk.weight, sk.bias, bn, want_umma=want_umma) for sk, bn in self._heads()]
            self._packed_key = key
        return self._packed

    def packed_heads_train(self, want_umma: bool):
        """Skip-head packs without BatchNorm (bias only), for the training-mode forward."""
        key = (tuple((sk.weight.data_ptr(), sk.weight._version, sk.bias._version) for sk, _ in self._heads()), want_umma,
               ops.weights_epoch())
        if getattr(self, "_packed_tr", None) is None or self._packed_tr_key != key:
            self._packed_tr = [PackedConv(sk.weight, sk.bias, None, want_umma=want_umma) for sk, _ in self._heads()]
            self._packed_tr_key = key
        return self._packed_tr

    def forward_act(self, x: torch.Tensor, impl: Optional[int] = None, raw: bool = False, operand_only: bool = False,
                    tape: Optional[dict] = None):
        """Returns (feats f32 Acts, feats in the kernels' operand format or None) -- five entries each.

        operand_only (tcgen05 family): skip the float32 copy of the features (the decoder's fast path consumes the
        split-bf16 operand copy only); returns (None, feats_op).
        tape (training mode): dict that receives what the backward needs (see ResNet101._forward_act_train)."""
        impl = ops.default_impl() if impl is None else impl
        taps = self.base.forward_act(x, impl, tape=tape)
        if raw:
            return taps, None
        fmt = ops.activation_format(impl)
        if self.training:
            # train-mode skip heads (model.py:59-63): conv + bias, then BatchNorm2d with this batch's statistics
            feats, feats_op = [], []
            for hi, (tap, pc, (sk, bn)) in enumerate(zip(taps, self.packed_heads_train(
                    want_umma=(fmt == ops.FMT_SPLIT_BF16)), self._heads())):
                raw = ops.conv2d([tap], pc, pad=self.padding, out_fmt=ops.FMT_F32, impl=impl)
                if tape is None:
                    scale, shift = ops.bn_train_stats(raw, bn)
                else:
                    scale, shift, mean, invstd = ops.bn_train_stats(raw, bn, want_stats=True)
                    tape[f"head{hi}"] = (tap, raw, None, mean, invstd)
                if fmt == ops.FMT_F32:
                    y = ops.affine_act(raw, scale, shift, out_fmt=ops.FMT_F32)
                    y2 = y
                else:
                    y, y2 = ops.affine_act(raw, scale, shift, out_fmt=ops.FMT_F32, out2_fmt=fmt)
                feats.append(y)
                feats_op.append(y2)
            return feats, feats_op
        heads = self.packed_heads(want_umma=(fmt == ops.FMT_SPLIT_BF16))
        if operand_only and fmt == ops.FMT_SPLIT_BF16:
            return None, [ops.conv2d([tap], pc, pad=self.padding, out_fmt=fmt, impl=impl) for tap, pc in zip(taps, heads)]
        feats, feats_op = [], []
        for tap, pc in zip(taps, heads):
            if fmt == ops.FMT_F32:
                y = ops.conv2d([tap], pc, pad=self.padding, out_fmt=ops.FMT_F32, impl=impl)
                feats.append(y)
                feats_op.append(y)
            else:
                y, y2 = ops.conv2d([tap], pc, pad=self.padding, out_fmt=ops.FMT_F32, out2_fmt=fmt, impl=impl)
                feats.append(y)
                feats_op.append(y2)
        return feats, feats_op

    def forward(self, x, semseg=False, raw=False):
        if self.training and torch.is_grad_enabled() and not (semseg or raw):
            # train.py:71-77,184: the outputs carry an autograd node whose backward runs the CUDA backward kernels
            from ..autograd import encoder_forward_train
            ops.require_cuda(x, "FeatureExtractor")
            return encoder_forward_train(self, x)
        if semseg or raw:
            taps, _ = self.forward_act(x, raw=True)
            outs = tuple(ops.act_to_nchw(a) for a in taps)
            return outs[0] if semseg else outs
        feats, feats_op = self.forward_act(x)
        outs = []
        for f, fo in zip(feats, feats_op):
            t = f.nchw()
            if fo is not f:
                ops.attach_operand_copy(t, fo)  # lets the decoder skip re-deriving the split-bf16 copy
            outs.append(t)
        return tuple(outs)


def wavefront_schedule(T: int, nlev: int, skew: int):
    """The decoder's (level, step) loop nest (test.py:37-44 x model.py:129-165) as wavefronts: entry w lists the cells
    (l, t) that run together in grouped launch w.  Cell (l, t) needs (l, t-1) (its own state) and, through the x2
    upsampling, (l-1, t).  skew = 1: wavefront l + t, the upsamplings sit between two launches; skew = 2: wavefront
    2*l + t, the upsampling of (l-1, t) has wavefront 2*(l-1) + t + 1 to itself (a side stream beside that launch) and is
    consumed one launch later.  Pure function (tests/test_schedule_cpu.py checks its hazards)."""
    assert skew in (1, 2) and T >= 1 and nlev >= 1
    return [[(l, w - skew * l) for l in range(nlev) if 0 <= w - skew * l < T] for w in range(T + skew * (nlev - 1))]


class DecoderWorkspace:
    """Device buffers of the tcgen05 decoder for one (batch, feature-map sizes).

    The gate convolution of level l contracts over `[up(h_{l-1}) | skip_l | h_prev_l]` (`torch.cat([hidden, skip], 1)`
    model.py:153 and `torch.cat((input_, prev_hidden), 1)` clstm.py:43).  skip_l (level 0: x5_skip) does not depend on
    the time-step, so its share of the gates (+ bias) is hoisted: `P[l]` = fp32 `[N,H,W,4*Ch]`, computed once per batch
    by `load_feats`.  Per step the fused cell kernel contracts over the ping-pong input buffer
    `X[l][p] = [up(h_{l-1}) | h_prev_l]` (split-bf16) and adds `P[l]` in its epilogue.  Producers write their slice of
    X in place: the bilinear-upsample kernel of step t into `X[l][t % 2]`, the cell epilogue of step t (new hidden
    state) into `X[l][(t + 1) % 2]` (a step reads its neighbours' h_prev with a 3x3 halo, hence the ping-pong).
    Plus float32 h / c per level (c is updated in place), the x2-upsampled last hidden for the mask head and the
    side-feature max keys."""

    def __init__(self, decoder: "RSIS", n: int, sizes, device):
        self.n, self.sizes = n, [tuple(s) for s in sizes]
        cells = decoder.clstm_list
        self.hidden = [c.hidden_size for c in cells]
        self.up_c = [0] + self.hidden[:-1]                              # channels of up(h_{l-1})
        self.skip_c = [c.input_size - u for c, u in zip(cells, self.up_c)]
        self.cin = [u + h for u, h in zip(self.up_c, self.hidden)]      # channels of X[l]
        F16 = ops.FMT_SPLIT_BF16
        self.X = [[ops.Act.zeros(n, h, w, ct, F16, device) for _ in range(2)] for (h, w), ct in zip(self.sizes, self.cin)]
        self.P = [ops.Act.empty(n, h, w, 4 * ch, ops.FMT_F32, device) for (h, w), ch in zip(self.sizes, self.hidden)]
        self.h = [ops.Act.empty(n, h, w, ch, ops.FMT_F32, device) for (h, w), ch in zip(self.sizes, self.hidden)]
        self.c = [ops.Act.empty(n, h, w, ch, ops.FMT_F32, device) for (h, w), ch in zip(self.sizes, self.hidden)]
        self.up_last = None  # only allocated for hidden sizes the fused upsample + mask head does not take
        self.h_last = None   # float32 hidden state of the last level per step (run_wavefront; [2] in the older scheme)
        self.h_last_all = None
        self.h2 = None       # [level][2] float32 hidden states of the other levels, double-buffered (skewed wavefront)
        self.up_stream = None
        self.side = torch.zeros((n, sum(self.hidden)), dtype=torch.int32, device=device)
        self.side_stream = torch.cuda.Stream(device=device)  # forked work: skip heads, class/stop heads
        self.level_streams = None   # pipelined schedule (run_pipelined): one stream per level + the mask head's
        self.mask_stream = None
        self.sides = None           # [T, n, F] side-feature keys, one slab per step (pipelined schedule)
        self.t = 0

    def packs(self, decoder: "RSIS", l: int):
        return decoder.clstm_list[l].packed_hoisted(self.up_c[l], self.skip_c[l])

    def h_view(self, l: int, p: int) -> Act:
        return self.X[l][p].slice(self.up_c[l], self.hidden[l])

    def up_view(self, l: int, p: int) -> Act:
        return self.X[l][p].slice(0, self.up_c[l])

    def encode_into(self, encoder: "FeatureExtractor", decoder: "RSIS", x: torch.Tensor, impl: int):
        """Encoder pass of the fast path.  Each skip head and its hoisted gate convolution only need one backbone tap,
        so they are forked onto the side stream the moment that tap exists and run under the rest of the backbone
        (their own split-K lane; joined before the first decoder step)."""
        heads = encoder.packed_heads(want_umma=True)
        main = torch.cuda.current_stream(x.device)
        side = self.side_stream
        keep = []

        def on_tap(idx, tap):
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(side), _lib.lane(1):
                side.wait_event(ev)
                f = ops.conv2d([tap], heads[idx], pad=encoder.padding, out_fmt=ops.FMT_SPLIT_BF16, impl=impl)
                pc_skip, _ = self.packs(decoder, idx)
                ops.conv2d([f], pc_skip, pad=pc_skip.kh // 2, impl=impl, out=self.P[idx])
                keep.append((tap, f))

        taps = encoder.base.forward_act(x, impl, on_tap=on_tap)
        main.wait_stream(side)
        return taps, keep  # the caller holds these until the join is enqueued (allocator reuse safety)

    def load_feats(self, decoder: "RSIS", feats: Sequence[Act], impl: int):
        """Hoisted, time-invariant part of the gates: P[l] = conv(skip_l, W_gates[:, skip channels]) + bias."""
        for l, f in enumerate(feats):
            if f.fmt != ops.FMT_SPLIT_BF16:
                f = ops.convert(f, ops.FMT_SPLIT_BF16)
            pc_skip, _ = self.packs(decoder, l)
            ops.conv2d([f], pc_skip, pad=pc_skip.kh // 2, impl=impl, out=self.P[l])

    def reset(self, on_side_stream: bool = False):
        """New sequence: hidden state None == zeros (clstm.py:26-37).  on_side_stream: the zero fills (six small launches)
        go to the side stream, ordered after everything enqueued so far -- the caller joins it before the first decoder
        step (`encode_into` does), so they run beside the encoder instead of between encoder and decoder."""
        def fills():
            for l in range(len(self.X)):
                sl = slice(self.up_c[l], self.up_c[l] + self.hidden[l])
                self.X[l][0].t[..., sl].zero_()
            if self.sides is not None:
                self.sides.zero_()
        if on_side_stream:
            self.side_stream.wait_stream(torch.cuda.current_stream(self.side.device))
            with torch.cuda.stream(self.side_stream):
                fills()
        else:
            fills()
        self.t = 0

    def prepare_pipeline(self, T: int):
        dev = self.side.device
        if self.level_streams is None:
            self.level_streams = [torch.cuda.Stream(device=dev) for _ in self.X]
            self.mask_stream = torch.cuda.Stream(device=dev)
        if self.sides is None or self.sides.shape[0] < T:
            self.sides = torch.zeros((T,) + tuple(self.side.shape), dtype=torch.int32, device=dev)


class RSIS(nn.Module):
    """The recurrent decoder (model.py:72-184); `skip_mode='concat'` (the reference default, args.py:109)."""

    def __init__(self, args):
        super().__init__()
        get_skip_dims(args.base_model)
        self.precision = getattr(args, "precision", None)   # see FeatureExtractor
        self.hidden_size = int(args.hidden_size)
        self.num_classes = int(args.num_classes)
        self.kernel_size = int(args.kernel_size)
        padding = 0 if self.kernel_size == 1 else 1
        self.dropout = args.dropout
        self.dropout_stop = args.dropout_stop
        self.dropout_cls = args.dropout_cls
        self.skip_mode = args.skip_mode
        if self.skip_mode != "concat":
            raise NotImplementedError("rsis_b200: only skip_mode='concat' (the reference default) is implemented")
        if self.dropout > 0 or self.dropout_stop > 0 or self.dropout_cls > 0:
            raise NotImplementedError("rsis_b200: dropout > 0 is not implemented (reference defaults are 0.0)")
        hs = self.hidden_size
        skip_dims_out = [hs, hs // 2, hs // 4, hs // 8, hs // 16]
        self.clstm_list = nn.ModuleList()
        for i in range(len(skip_dims_out)):
            clstm_in_dim = hs if i == 0 else skip_dims_out[i - 1] * 2
            self.clstm_list.append(ConvLSTMCell(args, clstm_in_dim, skip_dims_out[i], self.kernel_size,
                                                padding=padding))
        self.conv_out = nn.Conv2d(skip_dims_out[-1], 1, self.kernel_size, padding=padding)
        self.fc_dim = sum(skip_dims_out)
        self.fc_class = nn.Linear(self.fc_dim, self.num_classes)
        self.fc_stop = nn.Linear(self.fc_dim, 1)

    def workspace(self, n: int, sizes, device) -> DecoderWorkspace:
        return DecoderWorkspace(self, n, sizes, device)

    def step_ws(self, ws: DecoderWorkspace, impl: int, mask_logits: Optional[torch.Tensor],
                class_probs: torch.Tensor, class_stride: int, stop_logit: Optional[torch.Tensor],
                stop_stride: int, mask_prob: Optional[torch.Tensor] = None, mask_prob_stride: int = 0,
                stop_prob: Optional[torch.Tensor] = None):
        """One decoder time-step of the tcgen05 fast path, entirely inside `ws` (12 kernel launches, no allocation):
        per level [upsample into the input buffer] + fused cell; then x2 upsample, mask head, class/stop heads."""
        if self.fc_class.in_features != self.fc_dim:
            raise RuntimeError("fc_class.in_features does not match the decoder's side-feature width")
        p = ws.t & 1
        ws.side.zero_()
        off = 0
        nlev = len(self.clstm_list)
        for l, cell in enumerate(self.clstm_list):
            x = ws.X[l][p]
            if l > 0:
                ops.upsample_bilinear(ws.h[l - 1], x.h, x.w, out=ws.up_view(l, p))
            _, pc = ws.packs(self, l)
            ops.convlstm_cell_x(x, pc, ws.c[l].t if ws.t > 0 else None, ws.side, off, h_out=ws.h[l], c_out=ws.c[l],
                                h16_out=ws.h_view(l, 1 - p), impl=impl, gate_preact=ws.P[l])
            off += cell.hidden_size
        hl = ws.h[nlev - 1]
        # the class / stop heads only need the side features: fork them next to the mask head
        main = torch.cuda.current_stream(hl.t.device)
        ws.side_stream.wait_stream(main)
        with torch.cuda.stream(ws.side_stream):
            ops.class_stop_heads(ws.side, self.fc_class.weight, self.fc_class.bias, self.fc_stop.weight,
                                 self.fc_stop.bias, class_probs, class_stride, stop_logit, stop_prob, stop_stride)
        if hl.c % 4 == 0 and hl.c <= 16:
            ops.upsample_mask_head(hl, 2 * hl.h, 2 * hl.w, self.conv_out.weight, self.conv_out.bias, mask_logits,
                                   mask_prob, mask_prob_stride)
        else:  # wide hidden sizes: two kernels through an upsampled buffer
            if ws.up_last is None:
                ws.up_last = ops.Act.empty(hl.n, 2 * hl.h, 2 * hl.w, hl.c, ops.FMT_F32, hl.t.device)
            ops.upsample_bilinear(hl, ws.up_last.h, ws.up_last.w, out=ws.up_last)
            ops.mask_head(ws.up_last, self.conv_out.weight, self.conv_out.bias, mask_logits, mask_prob,
                          mask_prob_stride)
        main.wait_stream(ws.side_stream)
        ws.t += 1

    def run_pipelined(self, ws: DecoderWorkspace, impl: int, T: int, class_probs: torch.Tensor,
                      mask_prob: torch.Tensor, stop_prob: torch.Tensor, split_k: bool = True,
                      cta_caps: Optional[Sequence[int]] = None):
        """All T decoder steps as a WAVEFRONT over (level, step): level l of step t only needs level l-1 of step t and
        level l of step t-1 (model.py:132-153 re-feeds `prev_hidden_list[i]` per level), so level l of step t runs
        concurrently with level l-1 of step t+1.  One stream per level (its upsample + cell in program order), one for
        the mask head, one for the class/stop heads; cross-stream edges are events:
            U(l,t) <- C(l-1,t)                       (reads h[l-1])
            C(l,t) <- U(l+1,t-1) / mask head (t-1)   (they read the h[l] this cell overwrites)
            mask head(t) <- C(4,t);  heads(t) <- C(0..4,t)   (side keys live in one slab per step)
        Outputs: class_probs [B,T,C], mask_prob [B,T,H,W], stop_prob [B,T,1] (the stacked tensors of test.py:46-50).
        Requires ws.reset() before (t = 0 state) and hidden sizes the fused upsample + mask head takes."""
        nlev = len(self.clstm_list)
        hl = ws.h[nlev - 1]
        assert hl.c % 4 == 0 and hl.c <= 16 and ws.t == 0
        ws.prepare_pipeline(T)
        dev = hl.t.device
        main = torch.cuda.current_stream(dev)
        B, _, H, W = mask_prob.shape
        C = class_probs.shape[-1]
        streams = list(ws.level_streams) + [ws.mask_stream, ws.side_stream]
        start = torch.cuda.Event()
        start.record(main)
        for s_ in streams:
            s_.wait_event(start)
        ev_cell, ev_up, ev_mask = {}, {}, {}
        offs = [sum(ws.hidden[:l]) for l in range(nlev)]
        for t in range(T):
            p = t & 1
            for l in range(nlev):
                S = ws.level_streams[l]
                with torch.cuda.stream(S), _lib.lane(2 + l):
                    x = ws.X[l][p]
                    if l > 0:
                        S.wait_event(ev_cell[(l - 1, t)])
                        ops.upsample_bilinear(ws.h[l - 1], x.h, x.w, out=ws.up_view(l, p))
                        ev_up[(l, t)] = torch.cuda.Event()
                        ev_up[(l, t)].record(S)
                    if t > 0:  # the readers of the h[l] this cell is about to overwrite
                        S.wait_event(ev_up[(l + 1, t - 1)] if l + 1 < nlev else ev_mask[t - 1])
                    _, pc = ws.packs(self, l)
                    cap = int(cta_caps[l]) if cta_caps is not None else 0
                    if split_k if isinstance(split_k, bool) else bool(split_k[l]):
                        ops.convlstm_cell_x(x, pc, ws.c[l].t if t > 0 else None, ws.sides[t], offs[l], h_out=ws.h[l],
                                            c_out=ws.c[l], h16_out=ws.h_view(l, 1 - p), impl=impl, gate_preact=ws.P[l])
                    else:
                        with _lib.no_splitk():
                            ops.convlstm_cell_x(x, pc, ws.c[l].t if t > 0 else None, ws.sides[t], offs[l],
                                                h_out=ws.h[l], c_out=ws.c[l], h16_out=ws.h_view(l, 1 - p), impl=impl,
                                                gate_preact=ws.P[l], cta_cap=cap)
                    ev_cell[(l, t)] = torch.cuda.Event()
                    ev_cell[(l, t)].record(S)
            with torch.cuda.stream(ws.mask_stream):
                ws.mask_stream.wait_event(ev_cell[(nlev - 1, t)])
                ops.upsample_mask_head(hl, 2 * hl.h, 2 * hl.w, self.conv_out.weight, self.conv_out.bias, None,
                                       mask_prob[:, t], T * H * W)
                ev_mask[t] = torch.cuda.Event()
                ev_mask[t].record(ws.mask_stream)
            with torch.cuda.stream(ws.side_stream):
                for l in range(nlev):
                    ws.side_stream.wait_event(ev_cell[(l, t)])
                ops.class_stop_heads(ws.sides[t], self.fc_class.weight, self.fc_class.bias, self.fc_stop.weight,
                                     self.fc_stop.bias, class_probs[:, t], T * C, None, stop_prob[:, t], T)
        for s_ in streams:
            main.wait_stream(s_)
        ws.t = T

    def run_wavefront(self, ws: DecoderWorkspace, impl: int, T: int, class_probs: torch.Tensor,
                      mask_prob: torch.Tensor, stop_prob: torch.Tensor):
        """All T decoder steps as GROUPED launches: the independent cells of one wavefront of the (level, step) loop nest
        (model.py:132-153: cell (l, t) needs (l-1, t) through the upsampling and (l, t-1) through its own state; see
        `wavefront_schedule`) run in ONE kernel launch (`ops.convlstm_cell_group`), so the small levels no longer pay a
        launch each -- they share the chip with the large ones.
        Default (skew 2, heads deferred): cell (l, t) in wavefront 2 l + t; the x2 upsamplings of a wavefront's hidden
        states run on a side stream beside the NEXT grouped launch and are consumed by the one after; the mask head
        (fused upsample + conv_out + sigmoid) and the class / stop heads of all T steps run as one launch each after the
        last wavefront.  Per pass at T = 10: 18 + 16 + 2 launches (the sequential schedule: 120).
        RSIS_B200_WAVE_SKEW=1 / RSIS_B200_DEFER_HEADS=0 restore the first form: wavefront l + t, one upsample launch
        between two grouped launches, heads on side streams beside the following wavefronts (the last level's hidden state
        double-buffered for them).  Requires ws.reset() before."""
        nlev = len(self.clstm_list)
        hl = ws.h[nlev - 1]
        assert hl.c % 4 == 0 and hl.c <= 16 and ws.t == 0
        ws.prepare_pipeline(T)
        dev = hl.t.device
        main = torch.cuda.current_stream(dev)
        B, _, H, W = mask_prob.shape
        C = class_probs.shape[-1]
        streams = list(ws.level_streams) + [ws.mask_stream, ws.side_stream]
        start = torch.cuda.Event()
        start.record(main)
        for s_ in streams:
            s_.wait_event(start)
        ev_cell, ev_up, ev_mask = {}, {}, {}
        offs = [sum(ws.hidden[:l]) for l in range(nlev)]
        for t in range(T):
            p = t & 1
            for l in range(nlev):
                S = ws.level_streams[l]
                with torch.cuda.stream(S), _lib.lane(2 + l):
                    x = ws.X[l][p]
                    if l > 0:
                        S.wait_event(ev_cell[(l - 1, t)])
                        ops.upsample_bilinear(ws.h[l - 1], x.h, x.w, out=ws.up_view(l, p))
                        ev_up[(l, t)] = torch.cuda.Event()
                        ev_up[(l, t)].record(S)
                    if t > 0:  # the readers of the h[l] this cell is about to overwrite
                        S.wait_event(ev_up[(l + 1, t - 1)] if l + 1 < nlev else ev_mask[t - 1])
                    _, pc = ws.packs(self, l)
                    cap = int(cta_caps[l]) if cta_caps is not None else 0
                    if split_k if isinstance(split_k, bool) else bool(split_k[l]):
                        ops.convlstm_cell_x(x, pc, ws.c[l].t if t > 0 else None, ws.sides[t], offs[l], h_out=ws.h[l],
                                            c_out=ws.c[l], h16_out=ws.h_view(l, 1 - p), impl=impl, gate_preact=ws.P[l])
                    else:
                        with _lib.no_splitk():
                            ops.convlstm_cell_x(x, pc, ws.c[l].t if t > 0 else None, ws.sides[t], offs[l],
                                                h_out=ws.h[l], c_out=ws.c[l], h16_out=ws.h_view(l, 1 - p), impl=impl,
                                                gate_preact=ws.P[l], cta_cap=cap)
                    ev_cell[(l, t)] = torch.cuda.Event()
                    ev_cell[(l, t)].record(S)
            with torch.cuda.stream(ws.mask_stream):
                ws.mask_stream.wait_event(ev_cell[(nlev - 1, t)])
                ops.upsample_mask_head(hl, 2 * hl.h, 2 * hl.w, self.conv_out.weight, self.conv_out.bias, None,
                                       mask_prob[:, t], T * H * W)
                ev_mask[t] = torch.cuda.Event()
                ev_mask[t].record(ws.mask_stream)
            with torch.cuda.stream(ws.side_stream):
                for l in range(nlev):
                    ws.side_stream.wait_event(ev_cell[(l, t)])
                ops.class_stop_heads(ws.sides[t], self.fc_class.weight, self.fc_class.bias, self.fc_stop.weight,
                                     self.fc_stop.bias, class_probs[:, t], T * C, None, stop_prob[:, t], T)
        for s_ in streams:
            main.wait_stream(s_)
        ws.t = T

    def run_wavefront(self, ws: DecoderWorkspace, impl: int, T: int, class_probs: torch.Tensor,
                      mask_prob: torch.Tensor, stop_prob: torch.Tensor):
        """All T decoder steps as T + 4 GROUPED launches: wavefront w runs the independent cells {(l, t) : l + t = w}
        (model.py:132-153: cell (l, t) needs (l-1, t) through the upsampling and (l, t-1) through its own state) in ONE
        kernel launch (`ops.convlstm_cell_group`), followed by ONE launch with the x2 upsamplings that feed the next
        wavefront.  The mask head of step t (fused upsample + conv_out + sigmoid) and the class/stop heads run on side
        streams beside the following wavefronts.  Per pass at T = 10: 14 + 13 + 10 + 10 launches instead of 120, and the
        small levels no longer pay a launch each -- they share the chip with the large ones.
        The last level's float32 hidden state is double-buffered (`ws.h_last`) so that the mask head of step t may still
        be reading it while cell (4, t+1) runs.  Requires ws.reset() before."""
        nlev = len(self.clstm_list)
        hl = ws.h[nlev - 1]
        assert hl.c % 4 == 0 and hl.c <= 16 and ws.t == 0
        ws.prepare_pipeline(T)
        # The mask / class / stop heads of ALL steps run after the last wavefront, one launch each
        # (`upsample_mask_head_steps`, `class_stop_heads_steps`): the grouped launch leaves no registers for another block
        # on its SM, so heads launched beside the following wavefronts (RSIS_B200_DEFER_HEADS=0, the earlier scheme)
        # held back that wavefront's CTAs on every SM they landed on -- ~140 us per pass (profiles/r2cf_*).  The last
        # level's float32 hidden state of every step is kept for it (T x 4 MB at batch 8).
        defer = os.environ.get("RSIS_B200_DEFER_HEADS", "1") == "1"
        nbuf = T if defer else 2
        if ws.h_last is None or len(ws.h_last) < nbuf:
            if defer:
                ws.h_last_all = ops.Act.empty(T * hl.n, hl.h, hl.w, hl.c, ops.FMT_F32, hl.t.device)
                per = hl.n * hl.h * hl.w * hl.c
                ws.h_last = [ops.Act(ws.h_last_all.t.view(-1)[t * per:(t + 1) * per].view(hl.n, hl.h, hl.w, hl.c), ops.FMT_F32)
                             for t in range(T)]
            else:
                ws.h_last = [hl] + [ops.Act.empty(hl.n, hl.h, hl.w, hl.c, ops.FMT_F32, hl.t.device) for _ in range(nbuf - 1)]
        dev = hl.t.device
        main = torch.cuda.current_stream(dev)
        B, _, H, W = mask_prob.shape
        C = class_probs.shape[-1]
        offs = [sum(ws.hidden[:l]) for l in range(nlev)]
        ev_mask = {}
        # RSIS_B200_WAVE_SKEW=2: cell (l, t) runs in wavefront 2*l + t instead of l + t.  The x2 upsampling of its hidden
        # state then has a whole wavefront to itself: it runs on a side stream BESIDE the next grouped launch and is
        # consumed by the one after, instead of sitting between two grouped launches on the critical path (13 gaps of
        # 15-22 us per pass: profiles/r2be_pass_trace.txt).  T + 2*(nlev-1) grouped launches instead of T + nlev - 1;
        # the float32 hidden states of levels 0..nlev-2 are double-buffered (cell (l, t+1) runs beside the upsampling
        # of (l, t)).
        skew = int(os.environ.get("RSIS_B200_WAVE_SKEW", "2"))
        if skew == 2:
            if ws.h2 is None:
                ws.h2 = [[h, ops.Act.empty(h.n, h.h, h.w, h.c, ops.FMT_F32, dev)] for h in ws.h[:nlev - 1]]
                ws.up_stream = torch.cuda.Stream(device=dev)
            ev_up = {}
        waves = wavefront_schedule(T, nlev, skew)
        n_waves = len(waves)
        for w, wave in enumerate(waves):
            cells, ups = [], []
            for l, t in wave:
                p = t & 1
                _, pc = ws.packs(self, l)
                if l == nlev - 1:
                    h_out = ws.h_last[t if defer else p]
                else:
                    h_out = ws.h2[l][p] if skew == 2 else ws.h[l]
                cells.append(dict(x=ws.X[l][p], pc=pc, c_prev=ws.c[l].t if t > 0 else None, side_max=ws.sides[t],
                                  side_offset=offs[l], h_out=h_out, c_out=ws.c[l], h16_out=ws.h_view(l, 1 - p),
                                  gate_preact=ws.P[l]))
                if l + 1 < nlev:
                    x_next = ws.X[l + 1][p]
                    ups.append((h_out, ws.up_view(l + 1, p)))
                    assert x_next.h == ws.up_view(l + 1, p).h
            t_last = w - skew * (nlev - 1)   # the step whose last level runs in this wavefront
            if t_last >= 2 and t_last < T and not defer:
                main.wait_event(ev_mask[t_last - 2])   # its mask head read the buffer cell (4, t_last) overwrites
            if skew == 2 and (w - 2) in ev_up:
                main.wait_event(ev_up[w - 2])          # the upsamplings this wavefront's cells read
            if cells:   # (T = 1 on the skewed schedule: every other wavefront is empty)
                ops.convlstm_cell_group(cells)
            if 0 <= t_last < T and not defer:
                done = torch.cuda.Event()
                done.record(main)
                with torch.cuda.stream(ws.mask_stream):
                    ws.mask_stream.wait_event(done)
                    src = ws.h_last[t_last & 1]
                    ops.upsample_mask_head(src, 2 * src.h, 2 * src.w, self.conv_out.weight, self.conv_out.bias, None,
                                           mask_prob[:, t_last], T * H * W)
                    ev_mask[t_last] = torch.cuda.Event()
                    ev_mask[t_last].record(ws.mask_stream)
                with torch.cuda.stream(ws.side_stream):
                    ws.side_stream.wait_event(done)
                    ops.class_stop_heads(ws.sides[t_last], self.fc_class.weight, self.fc_class.bias,
                                         self.fc_stop.weight, self.fc_stop.bias, class_probs[:, t_last], T * C, None,
                                         stop_prob[:, t_last], T)
            if ups and skew == 2:
                done_up = torch.cuda.Event()
                done_up.record(main)
                with torch.cuda.stream(ws.up_stream):
                    ws.up_stream.wait_event(done_up)
                    ops.upsample_bilinear_group(ups)
                    ev_up[w] = torch.cuda.Event()
                    ev_up[w].record(ws.up_stream)
            elif ups and w + 1 < n_waves:
                ops.upsample_bilinear_group(ups)
        if skew == 2:
            main.wait_stream(ws.up_stream)
        if defer:
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(ws.side_stream):
                ws.side_stream.wait_event(done)
                ops.class_stop_heads_steps(ws.sides[:T], self.fc_class.weight, self.fc_class.bias, self.fc_stop.weight,
                                           self.fc_stop.bias, class_probs, T * C, C, stop_prob, T, 1)
            src = ws.h_last_all
            if src.n != T * hl.n:   # the workspace was sized for a longer sequence: the first T steps of the buffer
                src = ops.Act(src.t[:T * hl.n], ops.FMT_F32)
            ops.upsample_mask_head_steps(src, T, 2 * src.h, 2 * src.w, self.conv_out.weight, self.conv_out.bias,
                                         mask_prob, T * H * W, H * W)
        if not defer:
            main.wait_stream(ws.mask_stream)
        main.wait_stream(ws.side_stream)
        ws.t = T

    def step_act(self, feats: Sequence[Act], prev, impl: int, mask_logits: torch.Tensor,
                 class_probs: torch.Tensor, class_stride: int, stop_logit: Optional[torch.Tensor],
                 stop_stride: int, mask_prob: Optional[torch.Tensor] = None, mask_prob_stride: int = 0,
                 stop_prob: Optional[torch.Tensor] = None):
        """One decoder time-step on NHWC activations.

        feats: the five skip features in the kernels' operand format; prev: None or list of (h Act, c tensor).
        Outputs are written in place into the given buffers.  Returns the new state list [(h_op Act, h f32 Act,
        c f32 Act)] per level.
        """
        fmt = ops.activation_format(impl)
        n = feats[0].n
        dev = feats[0].t.device
        if self.fc_class.in_features != self.fc_dim:
            raise RuntimeError("fc_class.in_features does not match the decoder's side-feature width")
        side = torch.zeros((n, self.fc_dim), dtype=torch.int32, device=dev)
        inputs: List[Act] = [feats[0]]
        new_state = []
        off = 0
        nlev = len(self.clstm_list)
        for i, cell in enumerate(self.clstm_list):
            ph = pc = None
            if prev is not None:
                ph, pc = prev[i][0], prev[i][2].t
            h, c, hs = cell.step_act(inputs, ph, pc, side, off, impl)
            h_op = hs if hs is not None else h
            new_state.append((h_op, h, c))
            off += cell.hidden_size
            if i + 1 < nlev:
                skip = feats[i + 1]
                up = ops.upsample_bilinear(h, skip.h, skip.w, fmt)
                inputs = [up, skip]
            else:
                up = ops.upsample_bilinear(h, h.h * 2, h.w * 2, ops.FMT_F32)
        ops.mask_head(up, self.conv_out.weight, self.conv_out.bias, mask_logits, mask_prob, mask_prob_stride)
        ops.class_stop_heads(side, self.fc_class.weight, self.fc_class.bias, self.fc_stop.weight, self.fc_stop.bias,
                             class_probs, class_stride, stop_logit, stop_prob, stop_stride)
        return new_state

    def forward(self, skip_feats, prev_hidden_list):
        ops.require_cuda(skip_feats[0], "RSIS")
        if self.training and torch.is_grad_enabled():
            from ..autograd import decoder_forward_train
            return decoder_forward_train(self, skip_feats, prev_hidden_list)
        impl = ops.default_impl()
        fmt = ops.activation_format(impl)
        feats = [ops.act_from_nchw(t, fmt) for t in skip_feats]
        prev = None
        if prev_hidden_list is not None:
            prev = []
            for h_t, c_t in prev_hidden_list:
                prev.append((ops.act_from_nchw(h_t, fmt), None, ops.act_from_nchw(c_t, ops.FMT_F32)))
        n = feats[0].n
        dev = skip_feats[0].device
        last = feats[-1]
        out_mask = torch.empty((n, 1, last.h * 2, last.w * 2), dtype=torch.float32, device=dev)
        class_probs = torch.empty((n, self.num_classes), dtype=torch.float32, device=dev)
        stop = torch.empty((n, 1), dtype=torch.float32, device=dev)
        state = self.step_act(feats, prev, impl, out_mask, class_probs, self.num_classes, stop, 1)
        hidden_list = []
        for h_op, h, c in state:
            ht = h.nchw()
            if h_op is not h:
                ops.attach_operand_copy(ht, h_op)
            hidden_list.append([ht, c.nchw()])
        if n == 1:  # the reference's `.squeeze()` (model.py:169) drops the batch dimension at B=1
            class_probs = class_probs.view(-1)
            stop = stop.view(-1)
        return out_mask, class_probs, stop, hidden_list
