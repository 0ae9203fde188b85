"""The criteria of the training loop -- counterparts of /root/reference/src/utils/hungarian.py:10-90 (`MaskedNLL`,
`StableBalancedMaskedBCE`, `softIoU`), /root/reference/src/utils/objectives.py:6-34 (`MaskedNLLLoss`, `MaskedBCELoss`,
`softIoULoss`) and the per-step cost matrix of /root/reference/src/train.py:96-110 (the
component right after the decoder step in the training loop).  One fused, HBM-bound CUDA kernel per call (`csrc/objectives.cu`) instead of the reference's
`repeat` + ~8 elementwise / reduction kernels; ground-truth masks may stay uint8 on the device (4x fewer bytes).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import check

_workspaces = {}


def _workspace(dev, b: int, g: int) -> torch.Tensor:
    n = _lib.load().rsis_soft_iou_workspace_bytes(b, g) // 4
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < n:
        ws = torch.zeros(max(n, 4096), dtype=torch.float32, device=dev)
        _workspaces[key] = ws
    return ws


def _gt(t: torch.Tensor):
    if t.dtype == torch.uint8 or t.dtype == torch.bool:
        return t.contiguous().view(torch.uint8), 1
    return t.contiguous().float(), 0


def soft_iou_cost_matrix(out_mask: torch.Tensor, y_mask: torch.Tensor, iou_weight: float = 1.0, e: float = 1e-6,
                         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """train.py:96-110 in one kernel: `c = iou_weight * softIoU(y_true_p, y_pred_i).view(B, gtT)` for the mask logits
    of ONE decoder step `out_mask` [B, HW] (or [B,1,H,W]) against all ground-truth masks `y_mask` [B, gtT, HW]
    (float 0/1 as in the reference, or uint8 / bool).  `out` may be a strided [B, gtT] view (e.g. `scores[:, :, t]` on
    the device).  No autograd (the reference detaches it too: `c.cpu().data`)."""
    ops.require_cuda(out_mask, "soft_iou_cost_matrix")
    lib = _lib.load()
    b = out_mask.shape[0]
    logits = out_mask.detach().reshape(b, -1).contiguous().float()
    hw = logits.shape[1]
    g = y_mask.shape[1]
    gt, is_u8 = _gt(y_mask.detach().reshape(b, g, hw))
    if out is None:
        out = torch.empty((b, g), dtype=torch.float32, device=logits.device)
    assert out.shape == (b, g) and out.dtype == torch.float32 and out.device == logits.device
    ws = _workspace(logits.device, b, g)
    check(lib.rsis_soft_iou_cost(logits.data_ptr(), gt.data_ptr(), is_u8, b, g, hw, float(e), float(iou_weight),
                                 ws.data_ptr(), out.data_ptr(), out.stride(0), out.stride(1), None, None,
                                 _lib.stream_ptr()), "soft_iou_cost")
    _lib.count_launch(1)
    return out


class _SoftIoURows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, target, out, e):
        lib = _lib.load()
        rows = out.shape[0]
        logits = out.detach().reshape(rows, -1).contiguous().float()
        hw = logits.shape[1]
        gt, is_u8 = _gt(target.detach().reshape(rows, hw))
        dev = logits.device
        cost = torch.empty(rows, dtype=torch.float32, device=dev)
        num = torch.empty(rows, dtype=torch.float32, device=dev)
        den = torch.empty(rows, dtype=torch.float32, device=dev)
        ws = _workspace(dev, rows, 1)
        check(lib.rsis_soft_iou_cost(logits.data_ptr(), gt.data_ptr(), is_u8, rows, 1, hw, float(e), 1.0, ws.data_ptr(),
                                     cost.data_ptr(), 1, 1, num.data_ptr(), den.data_ptr(), _lib.stream_ptr()),
              "soft_iou_cost")
        _lib.count_launch(1)
        ctx.saved = (logits, gt, is_u8, num, den, tuple(out.shape))
        return cost

    @staticmethod
    def backward(ctx, dcost):
        lib = _lib.load()
        logits, gt, is_u8, num, den, shape = ctx.saved
        rows, hw = logits.shape
        d = torch.empty_like(logits)
        dc = dcost.contiguous().float()
        check(lib.rsis_soft_iou_bwd(logits.data_ptr(), gt.data_ptr(), is_u8, rows, hw, num.data_ptr(), den.data_ptr(),
                                    dc.data_ptr(), 1.0, d.data_ptr(), _lib.stream_ptr()), "soft_iou_bwd")
        _lib.count_launch(1)
        return None, d.view(shape), None


def softIoU(target: torch.Tensor, out: torch.Tensor, e: float = 1e-6) -> torch.Tensor:
    """utils/hungarian.py:64-90: row-wise `1 - IoU(sigmoid(out), target)` for [rows, N] logits / binary targets;
    differentiable w.r.t. `out` (one forward kernel pass, one backward kernel)."""
    ops.require_cuda(out, "softIoU")
    return _SoftIoURows.apply(target, out, e)


class softIoULoss(nn.Module):
    """utils/objectives.py:27-34: mean of the soft-IoU cost over the rows selected by `sw`."""

    def forward(self, y_true, y_pred, sw):
        costs = softIoU(y_true, y_pred).view(-1, 1)
        return torch.mean(torch.masked_select(costs, sw.bool()))


# ---- masked class / stop losses (objectives.py:6-25) ------------------------------------------------------------------
def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().contiguous().float()


class _MaskedNLLRows(torch.autograd.Function):
    """Per-row `-balance[target] * log(probs[row, target])` (0 on rows whose `sw.byte()` is 0) + (sum, count) of the
    selected rows, one kernel; backward one kernel."""

    @staticmethod
    def forward(ctx, target, probs, sw, balance):
        lib = _lib.load()
        rows, c = probs.shape
        p = _f32(probs)
        t = target.detach().reshape(-1).contiguous().long()
        m = _f32(sw.reshape(-1))
        assert t.numel() == rows and m.numel() == rows
        bal = None if balance is None else _f32(balance).to(p.device)
        cost = torch.empty(rows, dtype=torch.float32, device=p.device)
        sc = torch.empty(2, dtype=torch.float32, device=p.device)
        check(lib.rsis_masked_nll_fwd(p.data_ptr(), t.data_ptr(), m.data_ptr(), None if bal is None else bal.data_ptr(),
                                      rows, c, cost.data_ptr(), sc.data_ptr(), _lib.stream_ptr()), "masked_nll_fwd")
        _lib.count_launch(1)
        ctx.saved = (p, t, m, bal)
        ctx.mark_non_differentiable(sc)
        return cost, sc

    @staticmethod
    def backward(ctx, dcost, _dsc):
        lib = _lib.load()
        p, t, m, bal = ctx.saved
        rows, c = p.shape
        d = torch.empty_like(p)
        g = dcost.contiguous().float()
        check(lib.rsis_masked_nll_bwd(p.data_ptr(), t.data_ptr(), m.data_ptr(), None if bal is None else bal.data_ptr(),
                                      g.data_ptr(), 1, rows, c, d.data_ptr(), _lib.stream_ptr()), "masked_nll_bwd")
        _lib.count_launch(1)
        return None, d, None, None


class _MaskedBCERows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, target, out, sw, balance_weight):
        lib = _lib.load()
        o = _f32(out.reshape(-1))
        t = _f32(target.reshape(-1))
        m = _f32(sw.reshape(-1))
        n = o.numel()
        assert t.numel() == n and m.numel() == n
        cost = torch.empty(n, dtype=torch.float32, device=o.device)
        sc = torch.empty(3, dtype=torch.float32, device=o.device)
        bw = -1.0 if balance_weight is None else float(balance_weight)
        check(lib.rsis_masked_bce_fwd(t.data_ptr(), o.data_ptr(), m.data_ptr(), bw, n, cost.data_ptr(), sc.data_ptr(),
                                      _lib.stream_ptr()), "masked_bce_fwd")
        _lib.count_launch(1)
        ctx.saved = (t, o, m, sc, tuple(out.shape))
        ctx.mark_non_differentiable(sc)
        return cost, sc

    @staticmethod
    def backward(ctx, dcost, _dsc):
        lib = _lib.load()
        t, o, m, sc, shape = ctx.saved
        d = torch.empty_like(o)
        g = dcost.contiguous().float()
        check(lib.rsis_masked_bce_bwd(t.data_ptr(), o.data_ptr(), m.data_ptr(), sc.data_ptr() + 8, g.data_ptr(), 1,
                                      o.numel(), d.data_ptr(), _lib.stream_ptr()), "masked_bce_bwd")
        _lib.count_launch(1)
        return None, d.view(shape), None, None


def masked_mean(cost_rows: torch.Tensor, sum_count: torch.Tensor, group=None) -> torch.Tensor:
    """`torch.mean(masked_select(costs, sw.byte()))` (train.py:161,168) from the per-row costs (0 on unselected rows) and
    the device-side selected count: no masked_select, no host synchronisation (capturable).

    Data-parallel rule: the reference's DataParallel criteria return the UN-reduced selected
    vectors, which are concatenated and averaged globally.  With one process per GPU and gradients AVERAGED over ranks,
    each rank must therefore contribute `local_sum * world / n_valid_global`: one scalar all-reduce of the count."""
    import torch.distributed as dist
    count = sum_count[1]
    scale = 1.0
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        count = count.clone()
        dist.all_reduce(count, op=dist.ReduceOp.SUM, group=group)
        scale = float(dist.get_world_size(group))
    return cost_rows.sum() * scale / count


class MaskedNLLLoss(nn.Module):
    """utils/objectives.py:6-15: the vector of selected `MaskedNLL` costs (its mean is taken by the caller,
    train.py:161).  `mean(...)` is the fused, synchronisation-free form of `torch.mean(self(...))`."""

    def __init__(self, balance_weight=None):
        super().__init__()
        self.balance_weight = balance_weight

    def rows(self, y_true, y_pred, sw):
        ops.require_cuda(y_pred, "MaskedNLLLoss")
        return _MaskedNLLRows.apply(y_true, y_pred, sw, self.balance_weight)

    def forward(self, y_true, y_pred, sw):
        costs, _ = self.rows(y_true, y_pred, sw)
        return torch.masked_select(costs.view(-1, 1), sw.reshape(-1, 1).to(torch.uint8).bool())

    def mean(self, y_true, y_pred, sw, group=None):
        costs, sc = self.rows(y_true, y_pred, sw)
        return masked_mean(costs, sc, group)


class MaskedBCELoss(nn.Module):
    """utils/objectives.py:17-25 over `StableBalancedMaskedBCE` (hungarian.py:34-59); `balance_weight=None` derives it
    from the targets like the reference does."""

    def __init__(self, balance_weight=None):
        super().__init__()
        self.balance_weight = balance_weight

    def rows(self, y_true, y_pred, sw):
        ops.require_cuda(y_pred, "MaskedBCELoss")
        return _MaskedBCERows.apply(y_true, y_pred, sw, self.balance_weight)

    def forward(self, y_true, y_pred, sw):
        costs, _ = self.rows(y_true, y_pred, sw)
        return torch.masked_select(costs.view(-1, 1), sw.reshape(-1, 1).to(torch.uint8).bool())

    def mean(self, y_true, y_pred, sw, group=None):
        costs, sc = self.rows(y_true, y_pred, sw)
        return masked_mean(costs, sc, group)


def hungarian_match(overlaps: torch.Tensor):
    """Minimum-cost matching per image on the device.  overlaps: float32 [B, gtT, T] costs (any strides).
    Returns (permute_indices int32 [B, gtT] -- `permute_indices[b, t] = ground-truth row matched to prediction t` for
    t < min(gtT, T), 0 elsewhere, the convention of utils/hungarian.py:113-121 --, total cost [B])."""
    ops.require_cuda(overlaps, "hungarian_match")
    lib = _lib.load()
    c = overlaps.detach().float()
    b, r, t = c.shape
    perm = torch.empty((b, r), dtype=torch.int32, device=c.device)
    total = torch.empty(b, dtype=torch.float32, device=c.device)
    check(lib.rsis_hungarian_match(c.data_ptr(), c.stride(0), c.stride(1), c.stride(2), b, r, t, perm.data_ptr(), r,
                                   total.data_ptr(), _lib.stream_ptr()), "hungarian_match")
    _lib.count_launch(1)
    return perm, total


def match(masks, classes, overlaps):
    """utils/hungarian.py:91-125 with everything left on the device: the ground-truth masks [B, gtT, N] and classes
    [B, gtT] permuted by the minimum-cost matching of `overlaps` [B, gtT, T].  Returns (t_mask, t_class,
    permute_indices) as device tensors (the reference returns numpy arrays after a D2H copy of all masks)."""
    t_mask, _p_mask = masks
    t_class, _p_class = classes
    perm, _ = hungarian_match(overlaps)
    idx = perm.long()
    t_mask_perm = torch.gather(t_mask, 1, idx.unsqueeze(-1).expand(-1, -1, t_mask.shape[2]))
    t_class_perm = torch.gather(t_class, 1, idx if t_class.dim() == 2 else idx.unsqueeze(-1).expand_as(t_class))
    return t_mask_perm, t_class_perm, perm
