"""Evaluation post-processing on the device -- counterpart of `resize_mask` in /root/reference/src/eval.py:97-127 for
masks that are already at the output size: threshold, ignore mask, area, and the COCO
run-length encoding (`mask.encode(np.asfortranarray(segmentation))`, coco/common/maskApi.c:32-41 + :203-215).
Only the run lengths (a few hundred integers per instance) are copied to the host instead of the float masks.
The resize to the original image size (eval.py:111-115: `scipy.ndimage.zoom(pred_mask, [h/H, w/W, 1], order=1)`) is the
align-corners bilinear interpolation the decoder already uses (`rsis_upsample_bilinear`): zoom with its default
`grid_mode=False` maps output index i to input coordinate i*(in-1)/(out-1).
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import _lib, ops
from ._lib import check


def resize_masks(masks: torch.Tensor, height: int, width: int) -> torch.Tensor:
    """eval.py:111-115 on the device: [n, H, W] float masks -> [n, height, width], order-1 (bilinear) interpolation with
    scipy.ndimage.zoom's default coordinate mapping (= align_corners=True)."""
    ops.require_cuda(masks, "resize_masks")
    n, h, w = masks.shape
    if (h, w) == (int(height), int(width)):
        return masks
    # the interpolation kernel works on NHWC activations with a multiple-of-4 channel count: 4 masks per "pixel"
    pad = (-n) % 4
    m = masks.detach().float()
    if pad:
        m = torch.cat([m, m.new_zeros((pad, h, w))], 0)
    groups = (n + pad) // 4
    x = ops.Act(m.view(groups, 4, h, w).permute(0, 2, 3, 1).contiguous(), ops.FMT_F32)
    y = ops.upsample_bilinear(x, int(height), int(width), ops.FMT_F32)
    return y.t.permute(0, 3, 1, 2).reshape(groups * 4, int(height), int(width))[:n].contiguous()


def rle_encode(masks: torch.Tensor, threshold: float = 0.5, ignore: Optional[torch.Tensor] = None,
               max_runs: Optional[int] = None):
    """masks: float32 CUDA tensor [n, H, W] (e.g. `test()`'s [B,T,H,W] flattened).  Returns (counts int32 [n, max_runs],
    n_runs int32 [n], areas int32 [n]) on the device; counts[i, :n_runs[i]] are the COCO run lengths of
    `masks[i] > threshold` in column-major order, areas[i] = number of ones."""
    ops.require_cuda(masks, "rle_encode")
    lib = _lib.load()
    m = masks.detach().contiguous().float()
    n, h, w = m.shape
    if max_runs is None:
        max_runs = min(h * w + 1, 4096)
    dev = m.device
    ws = torch.empty(lib.rsis_rle_workspace_bytes(n, h, w), dtype=torch.uint8, device=dev)
    counts = torch.zeros((n, max_runs), dtype=torch.int32, device=dev)
    n_runs = torch.empty(n, dtype=torch.int32, device=dev)
    areas = torch.empty(n, dtype=torch.int32, device=dev)
    ig = None
    if ignore is not None:
        ig = ignore.detach().to(torch.uint8).expand(n, h, w).contiguous()
    check(lib.rsis_rle_encode(m.data_ptr(), float(threshold), None if ig is None else ig.data_ptr(), n, h, w,
                              ws.data_ptr(), counts.data_ptr(), max_runs, n_runs.data_ptr(), areas.data_ptr(),
                              _lib.stream_ptr()), "rle_encode")
    _lib.count_launch(2)
    return counts, n_runs, areas


def rle_to_string(cnts) -> bytes:
    """maskApi.c:203-215 `rleToString` (host side; the COCO JSON `counts` field): LEB128-like, 6 bits per char, ascii
    48-111, every count after the third stored as the difference to the count two places before."""
    out = bytearray()
    c = [int(v) for v in cnts]
    for i, v in enumerate(c):
        x = v - c[i - 2] if i > 2 else v
        more = True
        while more:
            ch = x & 0x1f
            x >>= 5
            more = (x != -1) if (ch & 0x10) else (x != 0)
            if more:
                ch |= 0x20
            out.append(ch + 48)
    return bytes(out)


def encode_instances(masks: torch.Tensor, threshold: float = 0.5, ignore: Optional[torch.Tensor] = None,
                     size=None) -> List[dict]:
    """COCO-style segmentations of `masks` [n, H, W]: [{'size': [H, W], 'counts': bytes, 'area': int}, ...] (what
    eval.py:122 obtains from pycocotools), with one small D2H copy for all instances.  size=(height, width): resize
    first (`resize_mask` of eval.py:97-127 end to end)."""
    if size is not None:
        masks = resize_masks(masks, size[0], size[1])
    n, h, w = masks.shape
    counts, n_runs, areas = rle_encode(masks, threshold, ignore, max_runs=h * w + 1 if h * w < 4096 else None)
    nr = n_runs.cpu().tolist()
    if max(nr) > counts.shape[1]:  # rare: extremely fragmented mask -- redo with room for every run
        counts, n_runs, areas = rle_encode(masks, threshold, ignore, max_runs=max(nr))
    width = max(nr)
    host = counts[:, :width].cpu().numpy()
    ar = areas.cpu().tolist()
    return [{"size": [h, w], "counts": rle_to_string(host[i, :nr[i]]), "area": ar[i]} for i in range(n)]
