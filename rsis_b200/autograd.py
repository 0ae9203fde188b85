"""Training path: `loss.backward()` (/root/reference/src/train.py:184) through the encoder and the recurrent decoder.

The reference relies on torch autograd over nn.Conv2d / nn.BatchNorm2d / F.interpolate / nn.MaxPool2d nodes.  Here the
two modules stay single autograd nodes each -- `EncoderFunction` (FeatureExtractor.forward in training mode) and
`DecoderStepFunction` (one RSIS.forward call) -- whose backward is a fixed sequence of the CUDA kernels declared in
include/rsis_b200.h ("backward primitives").  torch autograd only links the nodes: back-propagation through time is
the chain of T DecoderStepFunction nodes (hidden_list of step t feeds step t+1, exactly as train.py:94 re-feeds
`hidden`), and the sum of the T per-step gradients of the skip features / parameters is autograd's own accumulation.

Saved for backward, per convolution of the encoder: its input, the raw (pre-BatchNorm) output, the post-activation
output (ReLU mask) and the batch statistics.  Per decoder step and level: the concatenated cell input
`[up(h_{l-1}) | skip_l | h_prev_l]`, the activated gates, c_prev, c and the arg-max of the side max-pool.

Kernel family of the backward convolutions: env RSIS_B200_BWD_IMPL = simt (exact fp32 CUDA cores) | auto (data
gradients on the tcgen05 convolution, weight gradients on the tcgen05 MN-major kernel, wherever their shape rules
allow; gradients then travel as split-bf16 planes).  Default: auto when the forward runs on tcgen05, simt otherwise.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import torch

from . import ops
from .ops import Act, PackedConv

F32 = ops.FMT_F32


def backward_impl(fwd_impl: int) -> int:
    name = os.environ.get("RSIS_B200_BWD_IMPL", "").lower()
    if name == "":
        return ops.IMPL_AUTO if ops.uses_tcgen05(fwd_impl) else ops.IMPL_SIMT
    if name not in ("simt", "auto"):
        raise RuntimeError(f"RSIS_B200_BWD_IMPL must be simt or auto, got {name!r}")
    if name == "auto" and not ops.has_tcgen05():
        return ops.IMPL_SIMT
    return ops.IMPL_SIMT if name == "simt" else ops.IMPL_AUTO


def _grad_fmt(bimpl: int) -> int:
    """Element format of the gradients that feed data-gradient convolutions."""
    return F32 if bimpl == ops.IMPL_SIMT else ops.FMT_SPLIT_BF16


def grad_as_act(t: Optional[torch.Tensor]) -> Optional[Act]:
    """A gradient handed over by autograd (logical [N,C,H,W], any strides) as an NHWC activation; zero-copy when its
    memory already is NHWC (dense or a channel slice of a wider buffer)."""
    if t is None:
        return None
    a = Act.strided_view(t)
    if a is not None:
        return a
    return ops.nchw_to_nhwc(t, F32)


def _grad_target(p: torch.Tensor):
    """Where the gradient of parameter `p` is written.  When `p.grad` already exists as a float32 contiguous tensor
    (GradBucket views, or gradients left by an earlier backward) the kernels ACCUMULATE into it directly and the
    autograd node reports no gradient for `p` -- this removes one elementwise add per parameter and backward node
    (autograd's AccumulateGrad).  Otherwise a fresh tensor is returned to autograd.
    Returns (tensor, accumulate flag, direct flag)."""
    g = p.grad
    if g is not None and g.dtype == torch.float32 and g.is_contiguous() and g.shape == p.shape and g.device == p.device:
        return g, True, True
    return torch.empty_like(p, memory_format=torch.contiguous_format), False, False


class _DgradCache:
    """Packed weights of data-gradient convolutions, keyed by (parameter, channel range); rebuilt when the parameter
    changes (every optimiser step)."""

    def __init__(self):
        self.packs: Dict = {}

    def get(self, weight: torch.Tensor, want_umma: bool, ci0: int = 0, nci: Optional[int] = None) -> PackedConv:
        key = (id(weight), ci0, nci, want_umma)
        ver = (weight.data_ptr(), weight._version, ops.weights_epoch())
        hit = self.packs.get(key)
        if hit is None or hit[0] != ver:
            cout, cin, k, _ = weight.shape
            n_in = cin - ci0 if nci is None else nci    # output channels of the data-gradient convolution
            umma = want_umma and k in (1, 3) and n_in % 4 == 0 and cout % 8 == 0
            hit = (ver, PackedConv(weight, None, None, want_umma=umma, dgrad=(ci0, n_in)))
            self.packs[key] = hit
        return hit[1]


def _dgrad_cache(module) -> _DgradCache:
    c = getattr(module, "_rsis_dgrad_cache", None)
    if c is None:
        c = _DgradCache()
        object.__setattr__(module, "_rsis_dgrad_cache", c)
    return c


def conv_dgrad(cache: _DgradCache, dy: Act, weight: torch.Tensor, stride: int, pad: int, in_h: int, in_w: int,
               bimpl: int, residual: Optional[Act] = None, ci0: int = 0, nci: Optional[int] = None) -> Act:
    """Gradient of conv(x, weight, stride, pad) w.r.t. x (float32 NHWC [N, in_h, in_w, Cin]) given dy, as a forward
    convolution with the rotated / transposed weights (+ `residual`, added in the epilogue)."""
    k = weight.shape[-1]
    pc = cache.get(weight, bimpl != ops.IMPL_SIMT, ci0, nci)
    if stride == 1:
        return ops.conv2d([dy], pc, pad=k - 1 - pad, residual=residual, out_fmt=F32, impl=bimpl)
    assert stride == 2 and residual is None
    if k == 1:  # scatter of the low-resolution result onto the even grid
        low = ops.conv2d([dy], pc, pad=0, out_fmt=F32, impl=bimpl)
        return ops.dilate2x(low, in_h, in_w)
    return ops.conv2d([ops.dilate2x(dy, in_h, in_w, _grad_fmt(bimpl))], pc, pad=k - 1 - pad, out_fmt=F32, impl=bimpl)


# =============================================================================================================
# encoder
# =============================================================================================================
def encoder_backward(enc, tape: dict, dfeats: Sequence[Optional[torch.Tensor]], impl: int) -> Dict[int, torch.Tensor]:
    """Backward of FeatureExtractor.forward (training mode).  Returns {id(parameter): gradient}."""
    bimpl = backward_impl(impl)
    gfmt = _grad_fmt(bimpl)
    cache = _dgrad_cache(enc)
    base = enc.base
    G: Dict[int, torch.Tensor] = {}

    def put(p: torch.Tensor, g: torch.Tensor):
        G[id(p)] = g

    def bn_conv_bwd(tag: str, conv, bn, dy: Act, relu: bool, stride: int, pad: int, want_dres: bool = False,
                    dx_fmt: Optional[int] = None):
        """BatchNorm(+ReLU) backward, then the weight (and bias) gradient of the convolution feeding it.
        Returns (draw, dres)."""
        src, raw, y, mean, invstd = tape[tag]
        gw, _, direct_w = _grad_target(bn.weight)
        gb, _, direct_b = _grad_target(bn.bias)
        draw, dres, dw_bn, db_bn = ops.bn_train_bwd(raw, y if relu else None, dy, bn.weight, mean, invstd,
                                                    dx_fmt=gfmt if dx_fmt is None else dx_fmt, want_dres=want_dres,
                                                    dweight_acc=gw if direct_w else None,
                                                    dbias_acc=gb if direct_b else None)
        if not direct_w:
            put(bn.weight, dw_bn)
        if not direct_b:
            put(bn.bias, db_bn)
        k = conv.weight.shape[-1]
        dw, acc_w, direct_cw = _grad_target(conv.weight)
        ops.conv2d_wgrad(src, draw, k, k, stride, pad, dw, None, accumulate=acc_w)
        if not direct_cw:
            put(conv.weight, dw)
        if conv.bias is not None:
            db, acc_b, direct_cb = _grad_target(conv.bias)
            ops.conv2d_wgrad(src, draw, k, k, stride, pad, None, db, accumulate=acc_b)
            if not direct_cb:
                put(conv.bias, db)
        return draw, dres, src

    # ---- skip heads (model.py:59-63): BatchNorm (no ReLU) + biased 3x3 convolution ----
    head_draw = []
    for hi, (sk, bn) in enumerate(enc._heads()):
        dy = grad_as_act(dfeats[hi])
        src, raw = tape[f"head{hi}"][0], tape[f"head{hi}"][1]
        if dy is None:
            dy = Act.zeros(raw.n, raw.h, raw.w, raw.c, F32, raw.t.device)
        elif not dy.dense:
            dy = ops.convert(dy, F32)
        draw, _, _ = bn_conv_bwd(f"head{hi}", sk, bn, dy, relu=False, stride=1, pad=enc.padding)
        head_draw.append(draw)

    def head_dgrad(hi: int, residual: Optional[Act]) -> Act:
        sk = enc._heads()[hi][0]
        src = tape[f"head{hi}"][0]
        return conv_dgrad(cache, head_draw[hi], sk.weight, 1, enc.padding, src.h, src.w, bimpl, residual=residual)

    # ---- backbone, last layer first ----
    dcur: Optional[Act] = None
    for li in (4, 3, 2, 1):
        dcur = head_dgrad(4 - li, dcur)  # the tap of this layer also feeds its skip head
        layer = getattr(base, f"layer{li}")
        for bi in reversed(range(len(layer))):
            blk = layer[bi]
            p = f"layer{li}.{bi}"
            s = blk.stride
            draw3, dres, out2 = bn_conv_bwd(p + ".conv3", blk.conv3, blk.bn3, dcur, True, 1, 0, want_dres=True)
            dout2 = conv_dgrad(cache, draw3, blk.conv3.weight, 1, 0, out2.h, out2.w, bimpl)
            draw2, _, out1 = bn_conv_bwd(p + ".conv2", blk.conv2, blk.bn2, dout2, True, s, 1)
            dout1 = conv_dgrad(cache, draw2, blk.conv2.weight, s, 1, out1.h, out1.w, bimpl)
            draw1, _, cur = bn_conv_bwd(p + ".conv1", blk.conv1, blk.bn1, dout1, True, 1, 0)
            if blk.downsample is not None:
                drawd, _, _ = bn_conv_bwd(p + ".down", blk.downsample[0], blk.downsample[1], dres, False, s, 0)
                dres = conv_dgrad(cache, drawd, blk.downsample[0].weight, s, 0, cur.h, cur.w, bimpl)
            dcur = conv_dgrad(cache, draw1, blk.conv1.weight, 1, 0, cur.h, cur.w, bimpl, residual=dres)
    # ---- max-pool, x1's skip head, stem ----
    x1, _pooled = tape["pool"]
    dx1 = ops.maxpool3x3s2_bwd(x1, dcur)
    dx1 = head_dgrad(4, dx1)
    bn_conv_bwd("stem", base.conv1, base.bn1, dx1, True, 2, 3, dx_fmt=F32)
    return G


class EncoderFunction(torch.autograd.Function):
    """FeatureExtractor.forward in training mode as ONE autograd node."""

    @staticmethod
    def forward(ctx, enc, x, *params):
        impl = ops.default_impl()
        tape: dict = {}
        ctx.prec = getattr(enc, "precision", None)   # "bf16": BASELINE.json configs[3] single-pass mode (ops.precision)
        with ops.precision(ctx.prec):
            feats, feats_op = enc.forward_act(x, impl, tape=tape)
        ctx.enc, ctx.tape, ctx.impl, ctx.params = enc, tape, impl, params
        ctx.set_materialize_grads(False)
        enc._last_feats_op = feats_op
        return tuple(f.nchw() for f in feats)

    @staticmethod
    def backward(ctx, *dfeats):
        with torch.no_grad(), ops.precision(ctx.prec):
            G = encoder_backward(ctx.enc, ctx.tape, dfeats, ctx.impl)
        ctx.tape = None
        return (None, None) + tuple(G.get(id(p)) if p.requires_grad else None for p in ctx.params)


def encoder_forward_train(enc, x: torch.Tensor):
    params = [p for p in enc.parameters()]
    outs = EncoderFunction.apply(enc, x, *params)
    feats_op = enc._last_feats_op
    enc._last_feats_op = None
    for t, fo in zip(outs, feats_op):
        if fo is not None and fo.fmt != F32:
            ops.attach_operand_copy(t, fo)
    return tuple(outs)


# =============================================================================================================
# decoder step
# =============================================================================================================
def decoder_step_train(dec, feats: Sequence[Act], prev, impl: int):
    """One decoder time-step (model.py:122-184) keeping what the backward needs.  feats: the five skip features in the
    kernels' operand format; prev: None or list of (h Act in operand format, c Act f32).
    Returns (mask logits [N,1,2H,2W], class_probs [N,C], stop [N,1], [(h Act f32, c Act f32)], saved dict)."""
    fmt = ops.activation_format(impl)
    n = feats[0].n
    dev = feats[0].t.device
    if dec.fc_class.in_features != dec.fc_dim:
        raise RuntimeError("fc_class.in_features does not match the decoder's side-feature width")
    packed = torch.zeros((n, dec.fc_dim), dtype=torch.int64, device=dev)
    levels = []
    state = []
    h_below: Optional[Act] = None
    off = 0
    want_umma = fmt == ops.FMT_SPLIT_BF16
    for l, cell in enumerate(dec.clstm_list):
        fl = feats[l]
        ch = cell.hidden_size
        up_c = 0 if l == 0 else h_below.c
        skip_c = fl.c
        if up_c + skip_c != cell.input_size:
            raise RuntimeError(f"decoder level {l}: expected {cell.input_size} input channels, got {up_c + skip_c}")
        ctot = up_c + skip_c + ch
        X = (Act.empty if prev is not None else Act.zeros)(n, fl.h, fl.w, ctot, fmt, dev)
        if l > 0:
            ops.upsample_bilinear(h_below, fl.h, fl.w, out=X.slice(0, up_c))
        ops.convert(fl, fmt, out=X.slice(up_c, skip_c))
        c_prev = None
        if prev is not None:
            ops.convert(prev[l][0], fmt, out=X.slice(up_c + skip_c, ch))
            c_prev = prev[l][1]
        pc = cell.packed_plain(want_umma)
        k = cell.Gates.weight.shape[-1]
        gates = ops.conv2d([X], pc, pad=k // 2, out_fmt=F32, impl=impl)
        h, c = ops.lstm_gates_fwd(gates, c_prev.t if c_prev is not None else None)
        ops.global_maxpool(h, packed, off)
        levels.append(dict(X=X, gates=gates, c_prev=c_prev, c=c, up_c=up_c, skip_c=skip_c, ch=ch, off=off,
                           hw_below=(h_below.h, h_below.w) if h_below is not None else None))
        state.append((h, c))
        off += ch
        h_below = h
    up = ops.upsample_bilinear(h_below, 2 * h_below.h, 2 * h_below.w, F32)
    out_mask = torch.empty((n, 1, up.h, up.w), dtype=torch.float32, device=dev)
    class_probs = torch.empty((n, dec.num_classes), dtype=torch.float32, device=dev)
    stop = torch.empty((n, 1), dtype=torch.float32, device=dev)
    side = torch.empty((n, dec.fc_dim), dtype=torch.float32, device=dev)
    keys, idx = ops.global_maxpool_finish(packed)
    ops.mask_head(up, dec.conv_out.weight, dec.conv_out.bias, out_mask)
    ops.class_stop_heads(keys, dec.fc_class.weight, dec.fc_class.bias, dec.fc_stop.weight, dec.fc_stop.bias,
                         class_probs, dec.num_classes, stop, None, 1, feat_out=side)
    saved = dict(levels=levels, up=up, idx=idx, side=side, class_probs=class_probs, last_hw=(h_below.h, h_below.w))
    return out_mask, class_probs, stop, state, saved


def decoder_step_backward(dec, saved: dict, dmask, dclass, dstop, dh_out: List[Optional[torch.Tensor]],
                          dc_out: List[Optional[torch.Tensor]], has_state: bool, impl: int):
    """Backward of decoder_step_train.  Returns (dfeats[5], dh_prev[5] or None, dc_prev[5] or None, {id(param): grad})."""
    bimpl = backward_impl(impl)
    gfmt = _grad_fmt(bimpl)
    cache = _dgrad_cache(dec)
    levels = saved["levels"]
    nlev = len(levels)
    G: Dict[int, torch.Tensor] = {}
    last = levels[-1]
    n = last["gates"].n
    dev = last["gates"].t.device

    # ---- class / stop heads -> gradient of the side features ----
    dside = None
    if dclass is not None or dstop is not None:
        dc_ = dclass.reshape(n, -1).contiguous().float() if dclass is not None else None
        ds_ = dstop.reshape(n).contiguous().float() if dstop is not None else None
        hp = [dec.fc_class.weight, dec.fc_class.bias, dec.fc_stop.weight, dec.fc_stop.bias]
        tg = [_grad_target(p_) for p_ in hp]
        direct = all(t_[2] for t_ in tg)
        dside, dwc, dbc, dws, dbs = ops.class_stop_heads_bwd(saved["side"], saved["class_probs"], dc_, ds_,
                                                             dec.fc_class.weight, dec.fc_stop.weight,
                                                             into=[t_[0] for t_ in tg] if direct else None)
        if not direct:
            G[id(hp[0])], G[id(hp[1])], G[id(hp[2])], G[id(hp[3])] = dwc, dbc, dws, dbs

    # ---- mask head (conv_out on the x2-upsampled last hidden state) ----
    h_l, w_l = saved["last_hw"]
    dh_base: Optional[Act] = None
    if dmask is not None:
        up = saved["up"]
        dm = dmask.contiguous().float().view(n, up.h, up.w, 1)  # C = 1: NCHW and NHWC coincide
        dm_act = Act(dm, F32)
        k = dec.conv_out.weight.shape[-1]
        dw, acc_w, direct_w = _grad_target(dec.conv_out.weight)
        db, acc_b, direct_b = _grad_target(dec.conv_out.bias)
        if acc_w == acc_b:
            ops.conv2d_wgrad(up, dm_act, k, k, 1, k // 2, dw, db, accumulate=acc_w)
        else:
            ops.conv2d_wgrad(up, dm_act, k, k, 1, k // 2, dw, None, accumulate=acc_w)
            ops.conv2d_wgrad(up, dm_act, k, k, 1, k // 2, None, db, accumulate=acc_b)
        if not direct_w:
            G[id(dec.conv_out.weight)] = dw
        if not direct_b:
            G[id(dec.conv_out.bias)] = db
        d_up = conv_dgrad(cache, dm_act, dec.conv_out.weight, 1, k // 2, up.h, up.w, ops.IMPL_SIMT)
        dh_base = ops.upsample_bilinear_bwd(d_up, h_l, w_l)

    dfeats: List[Optional[torch.Tensor]] = [None] * nlev
    dh_prev: List[Optional[torch.Tensor]] = [None] * nlev
    dc_prev: List[Optional[torch.Tensor]] = [None] * nlev
    for l in range(nlev - 1, -1, -1):
        lv = levels[l]
        cell = dec.clstm_list[l]
        X, gates = lv["X"], lv["gates"]
        if dside is not None:
            if dh_base is None:
                dh_base = Act.zeros(n, X.h, X.w, lv["ch"], F32, dev)
            ops.global_maxpool_bwd(dside, saved["idx"], lv["off"], dh_base)
        dh_b = grad_as_act(dh_out[l])
        dc_n = grad_as_act(dc_out[l])
        if dh_base is None and dh_b is None and dc_n is None:
            continue  # nothing flows into this level (and hence none below it at this step)
        dgates, dcp = ops.lstm_gates_bwd(gates, lv["c_prev"].t if lv["c_prev"] is not None else None, lv["c"].t,
                                         dh_base, dh_b, dc_n, dg_fmt=gfmt)
        w = cell.Gates.weight
        k = w.shape[-1]
        dw, acc_w, direct_w = _grad_target(w)
        db, acc_b, direct_b = _grad_target(cell.Gates.bias)
        ops.conv2d_wgrad(X, dgates, k, k, 1, k // 2, dw, None, accumulate=acc_w)
        ops.conv2d_wgrad(X, dgates, k, k, 1, k // 2, None, db, accumulate=acc_b)
        if not direct_w:
            G[id(w)] = dw
        if not direct_b:
            G[id(cell.Gates.bias)] = db
        dX = conv_dgrad(cache, dgates, w, 1, k // 2, X.h, X.w, bimpl)
        up_c, skip_c, ch = lv["up_c"], lv["skip_c"], lv["ch"]
        dfeats[l] = dX.t[..., up_c:up_c + skip_c].permute(0, 3, 1, 2)
        if has_state:
            dh_prev[l] = dX.t[..., up_c + skip_c:].permute(0, 3, 1, 2)
            dc_prev[l] = dcp.nchw()
        dh_base = None
        if l > 0:
            hb, wb = lv["hw_below"]
            dh_base = ops.upsample_bilinear_bwd(dX.slice(0, up_c), hb, wb)
    return dfeats, dh_prev, dc_prev, G


class DecoderStepFunction(torch.autograd.Function):
    """One RSIS.forward call (training mode) as ONE autograd node.
    apply(dec, has_state, *feats(5), *[h0, c0, ..., h4, c4 if has_state], *params)"""

    @staticmethod
    def forward(ctx, dec, has_state, *tensors):
        impl = ops.default_impl()
        fmt = ops.activation_format(impl)
        nlev = len(dec.clstm_list)
        feats_t = tensors[:nlev]
        state_t = tensors[nlev:nlev + 2 * nlev] if has_state else ()
        params = tensors[nlev + len(state_t):]
        feats = [ops.act_from_nchw(t, fmt) for t in feats_t]
        prev = None
        if has_state:
            prev = [(ops.act_from_nchw(state_t[2 * l], fmt), ops.act_from_nchw(state_t[2 * l + 1], F32))
                    for l in range(nlev)]
        ctx.prec = getattr(dec, "precision", None)
        with ops.precision(ctx.prec):
            out_mask, class_probs, stop, state, saved = decoder_step_train(dec, feats, prev, impl)
        ctx.dec, ctx.saved, ctx.impl, ctx.has_state, ctx.params, ctx.nlev = dec, saved, impl, has_state, params, nlev
        ctx.needs_feats = [t.requires_grad for t in feats_t]
        ctx.set_materialize_grads(False)
        outs = [out_mask, class_probs, stop]
        for h, c in state:
            outs += [h.nchw(), c.nchw()]
        return tuple(outs)

    @staticmethod
    def backward(ctx, dmask, dclass, dstop, *dstate):
        nlev = ctx.nlev
        dh_out = [dstate[2 * l] for l in range(nlev)]
        dc_out = [dstate[2 * l + 1] for l in range(nlev)]
        with torch.no_grad(), ops.precision(ctx.prec):
            dfeats, dh_prev, dc_prev, G = decoder_step_backward(ctx.dec, ctx.saved, dmask, dclass, dstop, dh_out,
                                                                dc_out, ctx.has_state, ctx.impl)
        ctx.saved = None
        grads = [None, None]
        grads += [g if need else None for g, need in zip(dfeats, ctx.needs_feats)]
        if ctx.has_state:
            for l in range(nlev):
                grads += [dh_prev[l], dc_prev[l]]
        grads += [G.get(id(p)) if p.requires_grad else None for p in ctx.params]
        return tuple(grads)


def decoder_forward_train(dec, skip_feats, prev_hidden_list):
    nlev = len(dec.clstm_list)
    has_state = prev_hidden_list is not None
    tensors = list(skip_feats)
    if has_state:
        for h_t, c_t in prev_hidden_list:
            tensors += [h_t, c_t]
    tensors += [p for p in dec.parameters()]
    outs = DecoderStepFunction.apply(dec, has_state, *tensors)
    out_mask, class_probs, stop = outs[0], outs[1], outs[2]
    hidden_list = [[outs[3 + 2 * l], outs[4 + 2 * l]] for l in range(nlev)]
    if out_mask.shape[0] == 1:  # the reference's `.squeeze()` (model.py:169) drops the batch dimension at B=1
        class_probs = class_probs.view(-1)
        stop = stop.view(-1)
    return out_mask, class_probs, stop, hidden_list


# =============================================================================================================
# data-parallel gradient exchange: ONE all-reduce over a flat buffer
# =============================================================================================================
class GradBucket:
    """All trainable parameters' gradients as views of one flat float32 buffer, so that the data-parallel exchange of
    a training step is a single `all_reduce` (NCCL over NVLink on the GPU box; replaces the gradient gather of
    nn.DataParallel, train.py:269-274).  Usage per step: `bucket.zero()` (instead of `zero_grad()`, which would drop
    the views) -> forward -> `loss.backward()` (autograd accumulates into the views in place) ->
    `bucket.all_reduce()` -> optimiser step."""

    def __init__(self, params, flatten_params: bool = False):
        seen = set()
        self.params = []
        for p in params:  # first occurrence only (the reference's parameter lists contain duplicates)
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                self.params.append(p)
        if not self.params:
            raise ValueError("GradBucket: no trainable parameters")
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.offsets = {}
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            self.offsets[id(p)] = off
            off += p.numel()
        self.flat_params = None
        if flatten_params:
            # the parameters themselves become views of ONE buffer, in the same order (what a fused optimiser walks)
            self.flat_params = torch.empty(total, dtype=torch.float32, device=dev)
            with torch.no_grad():
                for p in self.params:
                    o = self.offsets[id(p)]
                    view = self.flat_params[o:o + p.numel()].view_as(p)
                    view.copy_(p.detach())
                    p.data = view

    def zero(self):
        self.flat.zero_()
        off = 0
        for p in self.params:  # re-attach (an optimiser's zero_grad(set_to_none=True) drops the views)
            g = self.flat[off:off + p.numel()].view_as(p)
            if p.grad is None or p.grad.data_ptr() != g.data_ptr():
                p.grad = g
            off += p.numel()

    def all_reduce(self, average: bool = True):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        if average:
            self.flat.mul_(1.0 / dist.get_world_size())
