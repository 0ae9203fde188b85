"""TEST INFRASTRUCTURE -- CPU oracle: a functional restatement of the RSIS hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import this module, and only as the checker / the
reported CPU baseline.  The product package `rsis_b200` never imports it.

What it restates (fp32, NCHW, state_dict in / tensors out, no nn.Module):
  * ResNet-101 taps        -- /root/reference/src/modules/vision.py:11-21 on top of
                              torchvision 0.26.0 `models/resnet.py` (`Bottleneck.forward`,
                              `ResNet._make_layer`; stride on conv2, eps 1e-5). torchvision is a
                              third-party dependency that is NOT vendored in /root/reference and is
                              unpinned there (README.md:16-17); 0.26.0 is what this image ships.
  * FeatureExtractor       -- /root/reference/src/modules/model.py:56-70
  * ConvLSTMCell           -- /root/reference/src/modules/clstm.py:19-62
  * RSIS decoder step      -- /root/reference/src/modules/model.py:122-184 (skip_mode='concat')
  * test() inference loop  -- /root/reference/src/test.py:16-50
  * softIoU / cost matrix  -- /root/reference/src/utils/hungarian.py:64-90, src/train.py:96-110,
                              src/utils/objectives.py:27-34

Pinning: the reference holds no golden vectors or tests for this path (SURVEY.md
section 4), so the oracle is pinned against outputs of the UNMODIFIED reference
modules executed in the build container (`oracle/make_golden.py` ->
`tests/golden/*.npz`; checked by `tests/test_oracle_golden.py`).

Every arithmetic primitive can be swapped through `conv=` so the same code also
serves as a precision emulator (split-bf16 / tf32 operand rounding) for design work.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm2d default, used by torchvision and model.py:50-54
RESNET101_BLOCKS = (3, 4, 23, 3)


def _bn_eval(sd, prefix, x):
    """Eval-mode BatchNorm2d with running statistics."""
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                        sd[prefix + ".weight"], sd[prefix + ".bias"], False, 0.0, BN_EPS)


def _bn_train(sd, prefix, x):
    """Train-mode BatchNorm2d: batch statistics (biased variance), no running-stat update here."""
    return F.batch_norm(x, None, None, sd[prefix + ".weight"], sd[prefix + ".bias"], True, 0.0, BN_EPS)


def bottleneck(sd, p, x, stride, conv=F.conv2d, bn=_bn_eval):
    """torchvision `Bottleneck.forward`: 1x1 -> 3x3(stride) -> 1x1, BN after each, residual add, ReLU."""
    identity = x
    out = F.relu(bn(sd, p + ".bn1", conv(x, sd[p + ".conv1.weight"])))
    out = F.relu(bn(sd, p + ".bn2", conv(out, sd[p + ".conv2.weight"], stride=stride, padding=1)))
    out = bn(sd, p + ".bn3", conv(out, sd[p + ".conv3.weight"]))
    if (p + ".downsample.0.weight") in sd:
        identity = bn(sd, p + ".downsample.1", conv(x, sd[p + ".downsample.0.weight"], stride=stride))
    return F.relu(out + identity)


def resnet101_taps(sd, x, conv=F.conv2d, bn=_bn_eval, prefix="base."):
    """vision.py:11-21 -> (x5, x4, x3, x2, x1)."""
    x = conv(x, sd[prefix + "conv1.weight"], stride=2, padding=3)
    x1 = F.relu(bn(sd, prefix + "bn1", x))
    x = F.max_pool2d(x1, kernel_size=3, stride=2, padding=1)
    taps = []
    for li, nblk in enumerate(RESNET101_BLOCKS, start=1):
        for b in range(nblk):
            stride = 2 if (b == 0 and li > 1) else 1
            x = bottleneck(sd, f"{prefix}layer{li}.{b}", x, stride, conv, bn)
        taps.append(x)
    x2, x3, x4, x5 = taps
    return x5, x4, x3, x2, x1


def feature_extractor(sd, x, conv=F.conv2d, bn=_bn_eval, raw=False):
    """model.py:56-70 -> (x5_skip, x4_skip, x3_skip, x2_skip, x1_skip); heads are conv+bias+BN, no ReLU."""
    taps = resnet101_taps(sd, x, conv, bn)
    if raw:
        return taps
    feats = []
    for n, t in zip("54321", taps):
        k = sd[f"sk{n}.weight"].shape[-1]
        y = conv(t, sd[f"sk{n}.weight"], padding=0 if k == 1 else 1) + sd[f"sk{n}.bias"].view(1, -1, 1, 1)
        feats.append(bn(sd, f"bn{n}", y))
    return tuple(feats)


def convlstm_cell(weight, bias, input_, prev_state, conv=F.conv2d):
    """clstm.py:19-62. Gate order along Cout is [in | remember | out | cell]; Cin order is [input_ | prev_hidden]."""
    ch = weight.shape[0] // 4
    if prev_state is None:  # clstm.py:26-37: zero state materialised
        z = input_.new_zeros((input_.shape[0], ch) + tuple(input_.shape[2:]))
        prev_state = (z, z)
    prev_hidden, prev_cell = prev_state
    stacked = torch.cat((input_, prev_hidden), 1)
    pad = 0 if weight.shape[-1] == 1 else 1
    gates = conv(stacked, weight, padding=pad) + bias.view(1, -1, 1, 1)
    in_gate, remember_gate, out_gate, cell_gate = gates.chunk(4, 1)
    in_gate = torch.sigmoid(in_gate)
    remember_gate = torch.sigmoid(remember_gate)
    out_gate = torch.sigmoid(out_gate)
    cell_gate = torch.tanh(cell_gate)
    cell = remember_gate * prev_cell + in_gate * cell_gate
    hidden = out_gate * torch.tanh(cell)
    return [hidden, cell]


def upsample_bilinear_ac(x, size):
    """nn.UpsamplingBilinear2d == bilinear interpolation with align_corners=True (model.py:149,163)."""
    return F.interpolate(x, size=size, mode="bilinear", align_corners=True)


def rsis_step(sd, feats, prev_hidden_list, conv=F.conv2d):
    """model.py:122-184, dropout 0, skip_mode 'concat'.

    Returns (out_mask logits [B,1,H,W], class_probs [B,C], stop logits [B,1], hidden_list).
    The reference squeezes side features (model.py:169) which drops the batch dim at B=1; this
    restatement keeps [B,...] and the caller squeezes when it needs the reference's B=1 shapes.
    """
    clstm_in = feats[0]
    skips = feats[1:]
    side = []
    hidden_list = []
    for i in range(len(skips) + 1):
        st = convlstm_cell(sd[f"clstm_list.{i}.Gates.weight"], sd[f"clstm_list.{i}.Gates.bias"], clstm_in,
                           None if prev_hidden_list is None else prev_hidden_list[i], conv)
        hidden_list.append(st)
        hidden = st[0]
        side.append(hidden.amax(dim=(2, 3)))  # nn.MaxPool2d(full spatial size), model.py:143
        if i < len(skips):
            skip = skips[i]
            hidden = upsample_bilinear_ac(hidden, tuple(skip.shape[-2:]))
            clstm_in = torch.cat([hidden, skip], 1)
        else:
            clstm_in = upsample_bilinear_ac(hidden, (hidden.shape[-2] * 2, hidden.shape[-1] * 2))
    k = sd["conv_out.weight"].shape[-1]
    out_mask = conv(clstm_in, sd["conv_out.weight"], padding=0 if k == 1 else 1) + sd["conv_out.bias"].view(1, -1, 1, 1)
    side_feats = torch.cat(side, 1)  # [B, 248]
    class_probs = torch.softmax(F.linear(side_feats, sd["fc_class.weight"], sd["fc_class.bias"]), dim=1)
    stop = F.linear(side_feats, sd["fc_stop.weight"], sd["fc_stop.bias"])
    return out_mask, class_probs, stop, hidden_list


def test_loop(enc_sd, dec_sd, x, T, conv=F.conv2d):
    """test.py:16-50: encoder once, T decoder steps, sigmoid on masks and stops.

    Returns (masks [B,T,H,W], classes [B,T,C], stops [B,T,1]).
    """
    with torch.no_grad():
        feats = feature_extractor(enc_sd, x, conv)
        hidden = None
        masks, classes, stops = [], [], []
        for _ in range(T):
            m, c, s, hidden = rsis_step(dec_sd, feats, hidden, conv)
            # test.py:39-40 upsamples to x's size; identity when sizes already match
            if tuple(m.shape[-2:]) != tuple(x.shape[-2:]):
                m = upsample_bilinear_ac(m, tuple(x.shape[-2:]))
            masks.append(m)
            classes.append(c)
            stops.append(s)
        masks = torch.cat(masks, 1)
        classes = torch.stack(classes, 1)
        stops = torch.stack(stops, 1)
        return torch.sigmoid(masks), classes, torch.sigmoid(stops)


def soft_iou(target, out, e=1e-6):
    """utils/hungarian.py:64-90 `softIoU`: row-wise cost 1 - IoU(sigmoid(out), target) for [rows, N] tensors."""
    out = torch.sigmoid(out)
    num = (out * target).sum(1, True)
    den = (out + target - out * target).sum(1, True) + e
    return (1 - num / den).squeeze()


def soft_iou_cost_matrix(out_mask, y_mask, iou_weight=1.0):
    """train.py:96-110: the mask logits of one decoder step [B, HW] against all gtT ground-truth masks [B, gtT, HW]
    -> `c.view(B, gtT)` (the reference repeats the prediction gtT times and calls softIoU on [B*gtT, HW])."""
    b, g, hw = y_mask.shape
    y_pred_i = out_mask.view(b, -1).unsqueeze(0).permute(1, 0, 2).repeat(1, g, 1).view(b * g, hw)
    y_true_p = y_mask.reshape(b * g, hw)
    return (iou_weight * soft_iou(y_true_p, y_pred_i)).view(b, -1)


def soft_iou_loss(y_true, y_pred, sw):
    """utils/objectives.py:27-34 `softIoULoss.forward`."""
    costs = soft_iou(y_true, y_pred).view(-1, 1)
    return torch.mean(torch.masked_select(costs, sw.bool()))


def match(t_mask, t_class, overlaps):
    """utils/hungarian.py:91-125 `match`: per image, the minimum-cost assignment of the [gtT, T] cost matrix
    `overlaps[b]`, then `permute_indices[b, column] = row` and the ground-truth masks / classes gathered by it.
    The reference calls `munkres.Munkres().compute` (munkres==1.0.12 in requirements.txt: a third-party dependency that
    is NOT in /root/reference and not installed in this image); `scipy.optimize.linear_sum_assignment` solves the same
    problem -- Munkres pads a rectangular matrix with zeros to a square one, whose optimum restricted to the original
    entries is the rectangular optimum -- but may pick a different optimal assignment when costs tie.
    Returns (t_mask_perm, t_class_perm, permute_indices [B, gtT] int64, total cost [B])."""
    import numpy as np
    from scipy.optimize import linear_sum_assignment
    b, r, _ = overlaps.shape
    perm = np.zeros((b, r), dtype=np.int64)
    total = np.zeros(b, dtype=np.float64)
    for i in range(b):
        cost = overlaps[i].detach().double().numpy()
        rows, cols = linear_sum_assignment(cost)
        total[i] = cost[rows, cols].sum()
        for row, col in zip(rows, cols):
            if col < r:
                perm[i, col] = row
    idx = torch.from_numpy(perm)
    t_mask_perm = torch.stack([t_mask[i, idx[i]] for i in range(b)])
    t_class_perm = torch.stack([t_class[i, idx[i]] for i in range(b)])
    return t_mask_perm, t_class_perm, idx, torch.from_numpy(total)


def masked_nll(target, probs, balance_weights=None):
    """utils/hungarian.py:10-33 `MaskedNLL`: -log(probs)[target], optionally scaled per class."""
    log_probs = torch.log(probs)
    if balance_weights is not None:
        log_probs = torch.mul(log_probs, balance_weights)
    return -torch.gather(log_probs, dim=1, index=target).squeeze()


def stable_balanced_bce(target, out, balance_weight=None):
    """utils/hungarian.py:35-61 `StableBalancedMaskedBCE` (balance_weight=None: computed from the targets)."""
    if balance_weight is None:
        num_positive = target.sum()
        num_negative = (1 - target).sum()
        balance_weight = num_positive / (num_positive + num_negative)
    max_val = (-out).clamp(min=0)
    loss_values = out - out * target + max_val + ((-max_val).exp() + (-out - max_val).exp()).log()
    losses = (1 - balance_weight) * loss_values * target + balance_weight * loss_values * (1 - target)
    return losses.squeeze()


def masked_nll_loss(y_true, y_pred, sw, balance_weight=None):
    """utils/objectives.py:6-15 `MaskedNLLLoss.forward`: the selected costs (the caller takes the mean, train.py:161)."""
    costs = masked_nll(y_true, y_pred, balance_weight).view(-1, 1)
    return torch.masked_select(costs, sw.to(torch.uint8).bool())


def masked_bce_loss(y_true, y_pred, sw, balance_weight=None):
    """utils/objectives.py:17-25 `MaskedBCELoss.forward`."""
    costs = stable_balanced_bce(y_true, y_pred, balance_weight).view(-1, 1)
    return torch.masked_select(costs, sw.to(torch.uint8).bool())


def run_iter(encode, step, x, y_mask, y_class, sw_mask, sw_class, args, cost_matrix=None, match_fn=None,
             iou_loss=None):
    """The forward + loss portion of `runIter` (train.py:71-176) as a recipe over callables, so that the SAME statement
    sequence runs over the CPU oracle (defaults) and over the modules / kernels under test:
        encode(x) -> feats;  step(feats, hidden) -> (out_mask [B,1,H,W], class_probs [B,C], stop [B,1], hidden);
        cost_matrix(out_mask [B,HW], y_mask [B,gtT,HW], iou_weight) -> [B,gtT]        (train.py:96-110)
        match_fn(y_mask, y_class, scores [B,gtT,T]) -> (y_mask_perm, y_class_perm)    (train.py:137, hungarian.py:91-125)
        iou_loss(y_true [R,HW], y_pred [R,HW], sw [R,1]) -> scalar                    (objectives.py:27-34)
    Returns (loss, [loss_mask_iou, loss_stop, loss_class], y_class_perm)."""
    cost_matrix = cost_matrix or soft_iou_cost_matrix
    match_fn = match_fn or (lambda m, c, sc: match(m, c, sc)[:2])
    iou_loss = iou_loss or soft_iou_loss
    B, gtT = y_mask.shape[0], y_mask.shape[1]
    T = args.maxseqlen
    if args.curriculum_learning:
        T = min(args.maxseqlen, args.limit_seqlen_to)
    feats = encode(x)
    scores = torch.ones(B, args.gt_maxseqlen, args.maxseqlen, device=y_mask.device)
    hidden = None
    out_masks, out_classes, out_stops = [], [], []
    stop_next = False
    for t in range(T):
        if stop_next:
            break
        if float(sw_mask[:, t].sum()) == 0:      # train.py:91 (a host read in the reference as well)
            stop_next = True
        out_mask, out_class, out_stop, hidden = step(feats, hidden)
        out_mask = out_mask.reshape(out_mask.shape[0], -1)   # train.py:96-98 (the resize is an identity at full size)
        scores[:, :, t] = cost_matrix(out_mask.detach(), y_mask, args.iou_weight)
        out_masks.append(out_mask)
        out_classes.append(out_class.reshape(B, -1))
        out_stops.append(out_stop.reshape(B, -1))
    t = len(out_masks)
    out_masks = torch.cat(out_masks, 1).view(B, t, -1)
    out_classes = torch.cat(out_classes, 1).view(B, t, -1)
    out_stops = torch.cat(out_stops, 1).view(B, t, -1)
    sw_mult = sw_mask.unsqueeze(-1).repeat(1, 1, args.maxseqlen)
    sw_mult_t = sw_mask[:, 0:args.maxseqlen].unsqueeze(-1).repeat(1, 1, args.gt_maxseqlen).permute(0, 2, 1)
    sw_mult = sw_mult * sw_mult_t                                        # train.py:121-124 (the byte AND)
    scores = scores * sw_mult + (1 - sw_mult) * 10                       # train.py:125
    y_mask_perm, y_class_perm = match_fn(y_mask, y_class, scores)
    y_mask_perm, y_class_perm = y_mask_perm[:, 0:t], y_class_perm[:, 0:t]
    sw_m = sw_mask[:, 0:t].contiguous().float()
    sw_c = sw_class[:, 0:t].contiguous().float()
    loss_class = masked_nll(y_class_perm.reshape(-1, 1).long(), out_classes.view(-1, out_classes.size()[-1])).view(-1, 1)
    loss_class = torch.mean(torch.masked_select(loss_class, sw_m.view(-1, 1).bool()))
    loss_mask_iou = iou_loss(y_mask_perm.reshape(-1, y_mask_perm.size()[-1]), out_masks.view(-1, out_masks.size()[-1]),
                             sw_m.view(-1, 1))
    loss_stop = stable_balanced_bce(sw_m, out_stops.squeeze(-1)).view(-1, 1)
    loss_stop = torch.mean(torch.masked_select(loss_stop, sw_c.view(-1, 1).bool()))
    loss = args.iou_weight * loss_mask_iou
    if args.use_class_loss:
        loss = loss + args.class_weight * loss_class
    if args.use_stop_loss:
        loss = loss + args.stop_weight * loss_stop
    return loss, [loss_mask_iou, loss_stop, loss_class], y_class_perm


# ----------------------------------------------------------------------------------------------
# precision emulators (design aids; not part of any parity claim)
# ----------------------------------------------------------------------------------------------
def split_bf16(t):
    hi = t.to(torch.bfloat16).to(torch.float32)
    lo = (t - hi).to(torch.bfloat16).to(torch.float32)
    return hi, lo


def conv_bf16x3(x, w, **kw):
    """a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi with bf16 operands and fp32 accumulation."""
    xh, xl = split_bf16(x)
    wh, wl = split_bf16(w)
    return F.conv2d(xh, wh, **kw) + F.conv2d(xh, wl, **kw) + F.conv2d(xl, wh, **kw)


def conv_bf16(x, w, **kw):
    return F.conv2d(split_bf16(x)[0], split_bf16(w)[0], **kw)


def _tf32_trunc(t):
    return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)


def conv_tf32_trunc(x, w, **kw):
    return F.conv2d(_tf32_trunc(x.contiguous()), _tf32_trunc(w.contiguous()), **kw)
