"""Thin functional wrappers over the C ABI (include/rsis_b200.h).  torch supplies device memory and streams only."""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import FMT_F32, FMT_SPLIT_BF16, IMPL_AUTO, IMPL_SIMT, IMPL_TCGEN05, Act, check  # noqa: F401


_weights_epoch = 0


def weights_epoch() -> int:
    """Part of every derived weight-pack cache key next to the parameters' (data_ptr, _version): bumped by writers that
    change parameter VALUES without going through torch (the fused optimiser writes through a flat buffer)."""
    return _weights_epoch


def bump_weights_epoch():
    global _weights_epoch
    _weights_epoch += 1


_bn_stats_epoch = 0


def bn_stats_epoch() -> int:
    """Part of every EVAL-side cache key (BatchNorm-folded packs, captured inference graphs): bumped whenever a
    train-mode BatchNorm kernel advances running_mean / running_var through raw pointers (torch's `_version` does not
    move then).  Train-mode packs do not fold the statistics and do not key on it."""
    return _bn_stats_epoch


def bump_bn_stats_epoch():
    global _bn_stats_epoch
    _bn_stats_epoch += 1


PRECISIONS = {"fp32": 0, "split": 0, "split-bf16": 0, "bf16": 1}


def get_precision() -> str:
    return "bf16" if _lib.load().rsis_get_precision() == 1 else "fp32"


class precision:
    """`with ops.precision("bf16"):` -- the tcgen05 convolutions set up inside run single-pass bf16 (BASELINE.json
    configs[3] "training step bf16": hi planes only, fp32 accumulation, fp32 master weights and gradients) instead of the
    default split-bf16 fp32-grade products.  Process-wide (the autograd worker thread must see it too); a CUDA graph keeps
    the mode it was captured in.  None / "fp32" = leave the default."""

    def __init__(self, mode: Optional[str]):
        mode = "fp32" if mode is None else str(mode).lower()
        if mode not in PRECISIONS:
            raise ValueError(f"rsis_b200: precision must be one of {sorted(PRECISIONS)}, got {mode!r}")
        self.mode = PRECISIONS[mode]
        self.prev = None

    def __enter__(self):
        self.prev = _lib.load().rsis_set_precision(self.mode)
        check(min(self.prev, 0), "set_precision")
        return self

    def __exit__(self, *exc):
        _lib.load().rsis_set_precision(self.prev)
        return False


class static_weights:
    """`with ops.static_weights():` -- launches set up inside may stream their packed weights before the previous
    kernel of the stream has finished (rsis_set_static_weights): the caller promises that nothing enqueued in between
    re-packs them.  Used by the inference pass; training steps re-pack in-stream and stay outside."""

    def __init__(self, on: bool = True):
        self.on = 1 if on else 0
        self.prev = 0

    def __enter__(self):
        self.prev = _lib.load().rsis_set_static_weights(self.on)
        return self

    def __exit__(self, *exc):
        _lib.load().rsis_set_static_weights(self.prev)
        return False


def launch_count() -> int:
    """Number of CUDA kernels this process has enqueued through the C ABI so far."""
    return _lib._launches


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"rsis_b200.{what}: tensors must live on a CUDA device (B200); there is no CPU path")


_IMPL_NAMES = {"auto": IMPL_AUTO, "simt": IMPL_SIMT, "tcgen05": IMPL_TCGEN05}


def default_impl() -> int:
    """Kernel family used by the nn.Module surface: env RSIS_B200_IMPL = auto (default) | simt | tcgen05."""
    name = os.environ.get("RSIS_B200_IMPL", "auto").lower()
    if name not in _IMPL_NAMES:
        raise RuntimeError(f"RSIS_B200_IMPL must be one of {sorted(_IMPL_NAMES)}, got {name!r}")
    return _IMPL_NAMES[name]


def has_tcgen05() -> bool:
    return bool(_lib.load().rsis_has_tcgen05())


def activation_format(impl: int) -> int:
    """Element format activations travel in between kernels for a kernel family: float32 for the CUDA-core
    path, split-bf16 (hi|lo planes) when the tcgen05 convolutions consume them."""
    if impl == IMPL_SIMT:
        return FMT_F32
    if impl == IMPL_TCGEN05 and not has_tcgen05():
        raise RuntimeError("rsis_b200: this build of librsis_b200.so has no tcgen05 kernels")
    return FMT_SPLIT_BF16 if has_tcgen05() else FMT_F32


def act_from_nchw(t: torch.Tensor, fmt: int) -> "Act":
    """Logical [N,C,H,W] float32 tensor -> NHWC activation in `fmt` (zero-copy for channels-last float32)."""
    require_cuda(t, "act_from_nchw")
    cached = getattr(t, "_rsis_operand", None)
    if cached is not None and cached.fmt == fmt and cached.valid_for(t):
        return cached
    if t.dim() != 4 or t.dtype != torch.float32:
        raise RuntimeError("rsis_b200: expected a float32 [N,C,H,W] tensor")
    xl = t.permute(0, 2, 3, 1)
    a = Act(xl, FMT_F32) if xl.is_contiguous() else nchw_to_nhwc(t, FMT_F32)
    return a if fmt == FMT_F32 else convert(a, fmt)


def act_to_nchw(a: "Act") -> torch.Tensor:
    """NHWC activation -> logical [N,C,H,W] float32 tensor (channels-last memory; zero-copy for float32)."""
    return (a if a.fmt == FMT_F32 else convert(a, FMT_F32)).nchw()


def attach_operand_copy(t: torch.Tensor, a: "Act"):
    """Remembers that `a` holds the same values as `t` in the kernels' operand format (derived cache)."""
    a.source_key = (t.data_ptr(), t._version, tuple(t.shape))
    t._rsis_operand = a


class PackedConv:
    """Kernel-ready copy of one convolution's parameters: K-major packed weights + folded bias/BatchNorm affine.

    Derived cache only -- the nn.Parameters (OIHW float32, reference layout) stay the source of truth.
    """

    def __init__(self, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, bn=None, gate_interleave=False,
                 src_channels: Optional[Sequence[int]] = None, want_umma: bool = False, dgrad=None):
        """dgrad=(ci0, nci): pack the DATA-GRADIENT convolution of `weight` w.r.t. its input channels [ci0, ci0+nci)
        (rotated, in/out swapped; see rsis_conv_pack_all) straight from the stored parameter."""
        lib = _lib.load()
        require_cuda(weight, "PackedConv")
        w = weight.detach().contiguous().float()
        w_cout, w_cin, kh, kw = w.shape
        if dgrad is not None:
            ci0, nci = dgrad
            cout, cin = nci, w_cout
            assert bias is None and bn is None and not gate_interleave
        else:
            ci0, nci = 0, 0
            cout, cin = w_cout, w_cin
        dev = w.device
        self.cout, self.cin, self.kh, self.kw = cout, cin, kh, kw
        self.gate_interleave = bool(gate_interleave)
        self.src_channels = list(src_channels) if src_channels is not None else [cin]
        assert sum(self.src_channels) == cin
        self.w_kc = torch.empty(lib.rsis_conv_pack_bytes_simt(cout, cin, kh, kw) // 4, dtype=torch.float32, device=dev)
        na = lib.rsis_conv_pack_bytes_affine(cout) // 4
        self.scale = torch.empty(na, dtype=torch.float32, device=dev)
        self.shift = torch.empty(na, dtype=torch.float32, device=dev)
        b = None if bias is None else bias.detach().contiguous().float()
        bnp = [None] * 4
        eps = 0.0
        if bn is not None:
            bnp = [bn.weight.detach().contiguous().float(), bn.bias.detach().contiguous().float(),
                   bn.running_mean.detach().contiguous().float(), bn.running_var.detach().contiguous().float()]
            eps = float(bn.eps)
        st = _lib.stream_ptr()
        sc = (C.c_int32 * len(self.src_channels))(*self.src_channels)
        self.w_umma = None
        self.k_pad = 0
        if want_umma:
            nbytes = lib.rsis_conv_pack_bytes_umma(cout, kh, kw, len(self.src_channels), sc)
            self.k_pad = lib.rsis_conv_umma_kpad(kh, kw, len(self.src_channels), sc)
            self.w_umma = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=dev)
        if os.environ.get("RSIS_B200_FUSED_PACK", "1") != "0":
            # ONE launch: CUDA-core pack + folded affine + tcgen05 pack (+ the data-gradient index mapping)
            check(lib.rsis_conv_pack_all(_ptr(w), w_cout, w_cin, kh, kw, int(dgrad is not None), ci0, nci, _ptr(b),
                                         _ptr(bnp[0]), _ptr(bnp[1]), _ptr(bnp[2]), _ptr(bnp[3]), eps,
                                         int(self.gate_interleave), len(self.src_channels), sc, _ptr(self.w_kc),
                                         _ptr(self.scale), _ptr(self.shift), _ptr(self.w_umma), st), "conv_pack_all")
            _lib.count_launch(1)
        else:
            if dgrad is not None:
                w = dgrad_weights(weight, ci0, nci)
            check(lib.rsis_conv_pack(_ptr(w), _ptr(b), _ptr(bnp[0]), _ptr(bnp[1]), _ptr(bnp[2]), _ptr(bnp[3]), eps, cout,
                                     cin, kh, kw, int(self.gate_interleave), _ptr(self.w_kc), _ptr(self.scale),
                                     _ptr(self.shift), st), "conv_pack")
            _lib.count_launch(2)
            if want_umma:
                check(lib.rsis_conv_pack_umma(_ptr(w), cout, cin, kh, kw, len(self.src_channels), sc,
                                              int(self.gate_interleave), _ptr(self.w_umma), st), "conv_pack_umma")
                _lib.count_launch(1)
        # plane-interleaved copy [cout_pad][2][k_pad] for the swapped-operand cell kernel of the narrow decoder levels
        self.w_umma_il = None
        if (self.w_umma is not None and self.gate_interleave and cout in (32, 64) and len(self.src_channels) == 1
                and cin <= 64 and kh == 3 and kw == 3):
            cp = lib.rsis_conv_umma_coutpad(cout)
            self.w_umma_il = self.w_umma.view(2, cp, self.k_pad).permute(1, 0, 2).contiguous()
        self.desc = _lib.ConvWeights(_ptr(self.w_kc), _ptr(self.w_umma), _ptr(self.scale), _ptr(self.shift), cout, cin,
                                     kh, kw, int(self.gate_interleave), _ptr(self.w_umma_il))

    def ref(self):
        return C.byref(self.desc)


def _src_array(srcs: Sequence[Act]):
    arr = (_lib.Tensor * len(srcs))()
    for i, s in enumerate(srcs):
        arr[i] = s.desc
    return arr


def conv2d(srcs: Sequence[Act], pc: PackedConv, stride: int = 1, pad: int = 0, relu: bool = False,
           residual: Optional[Act] = None, out_fmt: int = FMT_F32, out2_fmt: Optional[int] = None,
           impl: int = IMPL_AUTO, out: Optional[Act] = None, out2: Optional[Act] = None):
    """conv (+folded BN/bias) (+residual) (+ReLU); inputs concatenated along C. Returns y or (y, y2).
    `out` / `out2` may be pitched channel-slice views (the result lands inside a wider NHWC buffer)."""
    lib = _lib.load()
    x = srcs[0]
    ho = (x.h + 2 * pad - pc.kh) // stride + 1
    wo = (x.w + 2 * pad - pc.kw) // stride + 1
    dev = x.t.device
    y = out if out is not None else Act.empty(x.n, ho, wo, pc.cout, out_fmt, dev)
    y2 = out2 if out2 is not None else (Act.empty(x.n, ho, wo, pc.cout, out2_fmt, dev) if out2_fmt is not None else None)
    check(lib.rsis_conv2d(_src_array(srcs), len(srcs), pc.ref(), residual.ref() if residual is not None else None,
                          y.ref(), y2.ref() if y2 is not None else None, stride, pad, int(relu), impl,
                          *_lib.workspace(), _lib.stream_ptr()), "conv2d")
    _lib.count_launch(1)
    return (y, y2) if y2 is not None else y


def maxpool3x3s2(x: Act) -> Act:
    lib = _lib.load()
    y = Act.empty(x.n, (x.h - 1) // 2 + 1, (x.w - 1) // 2 + 1, x.c, x.fmt, x.t.device)
    check(lib.rsis_maxpool3x3s2(x.ref(), y.ref(), _lib.stream_ptr()), "maxpool3x3s2")
    _lib.count_launch(1)
    return y


def nchw_to_nhwc(x: torch.Tensor, fmt: int = FMT_F32) -> Act:
    lib = _lib.load()
    require_cuda(x, "nchw_to_nhwc")
    x = x.contiguous().float()
    n, c, h, w = x.shape
    y = Act.empty(n, h, w, c, fmt, x.device)
    check(lib.rsis_nchw_to_nhwc(x.data_ptr(), y.ref(), _lib.stream_ptr()), "nchw_to_nhwc")
    _lib.count_launch(1)
    return y


def im2col(x: Act, kh: int, kw: int, stride: int, pad: int, cy: int) -> Act:
    """float32 NHWC -> split-bf16 [N, Ho, Wo, cy] patch matrix (cy >= kh*kw*C, multiple of 8; the tail is zero)."""
    lib = _lib.load()
    ho = (x.h + 2 * pad - kh) // stride + 1
    wo = (x.w + 2 * pad - kw) // stride + 1
    y = Act.empty(x.n, ho, wo, cy, FMT_SPLIT_BF16, x.t.device)
    check(lib.rsis_im2col(x.ref(), kh, kw, stride, pad, y.ref(), _lib.stream_ptr()), "im2col")
    _lib.count_launch(1)
    return y


def convert(x: Act, fmt: int, out: Optional[Act] = None) -> Act:
    """Element-format conversion and/or channel-slice copy (`out` may be a pitched view of a wider buffer)."""
    lib = _lib.load()
    y = out if out is not None else Act.empty(x.n, x.h, x.w, x.c, fmt, x.t.device)
    check(lib.rsis_convert(x.ref(), y.ref(), _lib.stream_ptr()), "convert")
    _lib.count_launch(1)
    return y


_bn_workspaces = {}


def _bn_workspace(dev, c: int) -> torch.Tensor:
    # per (device, stream): two train forwards on different streams must not share the accumulator
    key = (dev.index, _lib.stream_ptr())
    ws = _bn_workspaces.get(key)
    if ws is None or ws.numel() < 2 * c:
        ws = torch.zeros(max(2 * c, 8192), dtype=torch.float64, device=dev)
        _bn_workspaces[key] = ws
    return ws


def bn_train_stats(x: Act, bn, want_stats: bool = False):
    """Train-mode nn.BatchNorm2d statistics of `x` (float32 dense NHWC): returns this batch's (scale, shift) and
    updates bn.running_mean / running_var / num_batches_tracked in place, like the module's own forward would.
    want_stats: also return (batch mean, batch 1/sqrt(var + eps)) -- what the backward needs."""
    lib = _lib.load()
    assert x.fmt == FMT_F32 and x.dense
    dev = x.t.device
    c = x.c
    ws = _bn_workspace(dev, c)
    scale = torch.empty(c, dtype=torch.float32, device=dev)
    shift = torch.empty(c, dtype=torch.float32, device=dev)
    mean = torch.empty(c, dtype=torch.float32, device=dev) if want_stats else None
    invstd = torch.empty(c, dtype=torch.float32, device=dev) if want_stats else None
    track = bn.track_running_stats and bn.running_mean is not None
    momentum = -1.0 if bn.momentum is None else float(bn.momentum)
    check(lib.rsis_bn_train_stats(x.ref(), _ptr(bn.weight), _ptr(bn.bias), float(bn.eps), momentum,
                                  _ptr(bn.running_mean) if track else None, _ptr(bn.running_var) if track else None,
                                  _ptr(bn.num_batches_tracked) if track else None, ws.data_ptr(), scale.data_ptr(),
                                  shift.data_ptr(), _ptr(mean), _ptr(invstd), _lib.stream_ptr()), "bn_train_stats")
    _lib.count_launch(2)
    if track:
        bump_bn_stats_epoch()  # eval-side packs fold these statistics: they are stale now
    return (scale, shift, mean, invstd) if want_stats else (scale, shift)


def affine_act(x: Act, scale: torch.Tensor, shift: torch.Tensor, residual: Optional[Act] = None, relu: bool = False,
               out_fmt: int = FMT_F32, out2_fmt: Optional[int] = None):
    """y = [relu](x * scale[c] + shift[c] [+ residual]) -- the normalise/ReLU/residual tail of a train-mode block."""
    lib = _lib.load()
    dev = x.t.device
    y = Act.empty(x.n, x.h, x.w, x.c, out_fmt, dev)
    y2 = Act.empty(x.n, x.h, x.w, x.c, out2_fmt, dev) if out2_fmt is not None else None
    check(lib.rsis_affine_act(x.ref(), scale.data_ptr(), shift.data_ptr(), residual.ref() if residual is not None else None,
                              int(relu), y.ref(), y2.ref() if y2 is not None else None, _lib.stream_ptr()), "affine_act")
    _lib.count_launch(1)
    return (y, y2) if y2 is not None else y


def uses_tcgen05(impl: int) -> bool:
    return activation_format(impl) == FMT_SPLIT_BF16


def convlstm_cell_x(x: Act, pc: PackedConv, c_prev: Optional[torch.Tensor], side_max: Optional[torch.Tensor],
                    side_offset: int = 0, h_out: Optional[Act] = None, c_out: Optional[Act] = None,
                    h16_out: Optional[Act] = None, impl: int = IMPL_AUTO, gate_preact: Optional[Act] = None,
                    cta_cap: int = 0):
    """One fused ConvLSTM step on the already-concatenated input buffer `x` = [input_ | prev_hidden] (split-bf16,
    all w.cin channels; the prev_hidden slice is zeros when the state is None).  `h16_out` (optional, may be a
    pitched view) receives the new hidden state in the operand format, e.g. the prev_hidden slice of the next
    step's input buffer.  c_out may alias c_prev.  Returns (h Act f32, c Act f32)."""
    lib = _lib.load()
    ch = pc.cout // 4
    dev = x.t.device
    h = h_out if h_out is not None else Act.empty(x.n, x.h, x.w, ch, FMT_F32, dev)
    c = c_out if c_out is not None else Act.empty(x.n, x.h, x.w, ch, FMT_F32, dev)
    stride = side_max.shape[1] if side_max is not None else 0
    srcs = _src_array([x])
    pre = None
    if gate_preact is not None:
        assert gate_preact.fmt == FMT_F32 and gate_preact.dense and gate_preact.c == pc.cout
        pre = gate_preact.t.data_ptr()
    check(lib.rsis_convlstm_cell(srcs, 1, pc.ref(), _ptr(c_prev), pre, h.ref(), h16_out.ref() if h16_out else None,
                                 c.ref(), _ptr(side_max), stride, side_offset, impl | ((cta_cap & 0xffff) << 8),
                                 *_lib.workspace(),
                                 _lib.stream_ptr()),
          "convlstm_cell")
    _lib.count_launch(1)
    return h, c


def cell_group_max() -> int:
    return int(_lib.load().rsis_convlstm_cell_group_max())


def convlstm_cell_group(cells: Sequence[dict]):
    """One wavefront of the decoder in ONE launch (rsis_convlstm_cell_group): `cells` = dicts with the keyword arguments
    of convlstm_cell_x (x, pc, c_prev, side_max, side_offset, h_out, c_out, h16_out, gate_preact), all buffers given."""
    lib = _lib.load()
    arr = (_lib.CellArgs * len(cells))()
    keep = []
    for i, c in enumerate(cells):
        x, pc = c["x"], c["pc"]
        pre = c.get("gate_preact")
        if pre is not None:
            assert pre.fmt == FMT_F32 and pre.dense and pre.c == pc.cout
        side = c.get("side_max")
        h16 = c.get("h16_out")
        a = arr[i]
        a.x = C.pointer(x.desc)
        a.w = C.pointer(pc.desc)
        a.c_prev = _ptr(c.get("c_prev"))
        a.gate_preact = None if pre is None else pre.t.data_ptr()
        a.h_out = C.pointer(c["h_out"].desc)
        a.h_split = C.pointer(h16.desc) if h16 is not None else None
        a.c_out = C.pointer(c["c_out"].desc)
        a.side_max = _ptr(side)
        a.side_stride = side.shape[-1] if side is not None else 0
        a.side_offset = int(c.get("side_offset", 0))
        keep.append((x, pc, pre, h16))
    check(lib.rsis_convlstm_cell_group(arr, len(cells), _lib.stream_ptr()), "convlstm_cell_group")
    _lib.count_launch(1)


def upsample_bilinear_group(pairs: Sequence[tuple]):
    """[(x Act f32 dense, out Act -- may be a pitched slice)] -> one launch (rsis_upsample_bilinear_group)."""
    lib = _lib.load()
    n = len(pairs)
    xs = (_lib.Tensor * n)()
    ys = (_lib.Tensor * n)()
    for i, (x, y) in enumerate(pairs):
        xs[i] = x.desc
        ys[i] = y.desc
    check(lib.rsis_upsample_bilinear_group(xs, ys, n, _lib.stream_ptr()), "upsample_bilinear_group")
    _lib.count_launch(1)


def convlstm_cell(srcs: Sequence[Act], pc: PackedConv, c_prev: Optional[torch.Tensor], side_max: Optional[torch.Tensor],
                  side_offset: int = 0, want_split: bool = False, impl: int = IMPL_AUTO):
    """One fused ConvLSTM step. `srcs` = [input_ parts..., prev_hidden] (prev_hidden omitted when the state is None).

    Returns (h Act f32, c Act f32, h_split Act or None). side_max: int32/uint32-viewed [N, F] key buffer (zeroed).
    The CUDA-core kernel reads the parts directly; the tcgen05 kernel takes one concatenated buffer, so the parts
    are first copied into their channel slices of it (general-purpose entry; the decoder's fast path writes those
    slices in place instead)."""
    lib = _lib.load()
    x = srcs[0]
    ch = pc.cout // 4
    dev = x.t.device
    if uses_tcgen05(impl) and all(s.fmt == FMT_SPLIT_BF16 for s in srcs):
        have = sum(s.c for s in srcs)
        buf = (Act.zeros if have < pc.cin else Act.empty)(x.n, x.h, x.w, pc.cin, FMT_SPLIT_BF16, dev)
        off = 0
        for s_ in srcs:
            convert(s_, FMT_SPLIT_BF16, out=buf.slice(off, s_.c))
            off += s_.c
        hs = Act.empty(x.n, x.h, x.w, ch, FMT_SPLIT_BF16, dev) if want_split else None
        h, c = convlstm_cell_x(buf, pc, c_prev, side_max, side_offset, h16_out=hs, impl=impl)
        return h, c, hs
    h = Act.empty(x.n, x.h, x.w, ch, FMT_F32, dev)
    c = Act.empty(x.n, x.h, x.w, ch, FMT_F32, dev)
    hs = Act.empty(x.n, x.h, x.w, ch, FMT_SPLIT_BF16, dev) if want_split else None
    stride = side_max.shape[1] if side_max is not None else 0
    check(lib.rsis_convlstm_cell(_src_array(srcs), len(srcs), pc.ref(), _ptr(c_prev), None, h.ref(),
                                 hs.ref() if hs else None,
                                 c.ref(), _ptr(side_max), stride, side_offset, impl, *_lib.workspace(),
                                 _lib.stream_ptr()),
          "convlstm_cell")
    _lib.count_launch(1)
    return h, c, hs


def upsample_bilinear(x: Act, ho: int, wo: int, fmt: int = FMT_F32, out: Optional[Act] = None) -> Act:
    lib = _lib.load()
    y = out if out is not None else Act.empty(x.n, ho, wo, x.c, fmt, x.t.device)
    check(lib.rsis_upsample_bilinear(x.ref(), y.ref(), _lib.stream_ptr()), "upsample_bilinear")
    _lib.count_launch(1)
    return y


def mask_head(x: Act, weight: torch.Tensor, bias: Optional[torch.Tensor], logits: Optional[torch.Tensor],
              prob_out: Optional[torch.Tensor] = None, prob_stride_n: int = 0):
    """conv_out: writes logits [N,H,W] (any tensor with N*H*W contiguous floats) and optional sigmoid copy."""
    lib = _lib.load()
    ks = weight.shape[-1]
    check(lib.rsis_mask_head(x.ref(), weight.data_ptr(), _ptr(bias), ks, _ptr(logits), _ptr(prob_out),
                             prob_stride_n, _lib.stream_ptr()), "mask_head")
    _lib.count_launch(1)


def upsample_mask_head(h: Act, out_h: int, out_w: int, weight: torch.Tensor, bias: Optional[torch.Tensor],
                       logits: Optional[torch.Tensor], prob_out: Optional[torch.Tensor] = None, prob_stride_n: int = 0):
    """upsample_bilinear(h -> out_h x out_w) + conv_out in one launch; the upsampled tensor never exists."""
    lib = _lib.load()
    ks = weight.shape[-1]
    check(lib.rsis_upsample_mask_head(h.ref(), weight.data_ptr(), _ptr(bias), ks, out_h, out_w, _ptr(logits),
                                      _ptr(prob_out), prob_stride_n, _lib.stream_ptr()), "upsample_mask_head")
    _lib.count_launch(1)


def upsample_mask_head_steps(h_all: Act, steps: int, out_h: int, out_w: int, weight: torch.Tensor,
                             bias: Optional[torch.Tensor], prob_out: torch.Tensor, prob_stride_n: int, prob_stride_t: int):
    """`upsample_mask_head` of all `steps` time-steps in one launch: `h_all` holds steps * B images, step-major."""
    lib = _lib.load()
    ks = weight.shape[-1]
    check(lib.rsis_upsample_mask_head_steps(h_all.ref(), steps, weight.data_ptr(), _ptr(bias), ks, out_h, out_w,
                                            prob_out.data_ptr(), prob_stride_n, prob_stride_t, _lib.stream_ptr()),
          "upsample_mask_head_steps")
    _lib.count_launch(1)


def class_stop_heads_steps(side_max_all: torch.Tensor, w_class, b_class, w_stop, b_stop, class_probs: torch.Tensor,
                           class_stride: int, class_stride_t: int, stop_prob: torch.Tensor, stop_stride: int,
                           stop_stride_t: int):
    """`class_stop_heads` of all time-steps in one launch: `side_max_all` is [steps, n, f] keys."""
    lib = _lib.load()
    steps, n, f = side_max_all.shape
    check(lib.rsis_class_stop_heads_steps(side_max_all.data_ptr(), n, steps, f, w_class.data_ptr(), b_class.data_ptr(),
                                          w_class.shape[0], w_stop.data_ptr(), b_stop.data_ptr(), class_probs.data_ptr(),
                                          class_stride, class_stride_t, stop_prob.data_ptr(), stop_stride, stop_stride_t,
                                          _lib.stream_ptr()), "class_stop_heads_steps")
    _lib.count_launch(1)


def class_stop_heads(side_max: torch.Tensor, w_class, b_class, w_stop, b_stop, class_probs: torch.Tensor,
                     class_stride: int, stop_logit: Optional[torch.Tensor], stop_prob: Optional[torch.Tensor],
                     stop_stride: int, feat_out: Optional[torch.Tensor] = None):
    lib = _lib.load()
    n, f = side_max.shape
    check(lib.rsis_class_stop_heads(side_max.data_ptr(), n, f, w_class.data_ptr(), b_class.data_ptr(),
                                    w_class.shape[0], w_stop.data_ptr(), b_stop.data_ptr(), _ptr(feat_out),
                                    class_probs.data_ptr(), class_stride, _ptr(stop_logit), _ptr(stop_prob),
                                    stop_stride, _lib.stream_ptr()), "class_stop_heads")
    _lib.count_launch(1)


# ---------------------------------------------------------------------------------------------------------
# backward primitives (loss.backward() of train.py:184); see include/rsis_b200.h
# ---------------------------------------------------------------------------------------------------------
def dgrad_weights(weight: torch.Tensor, ci0: int = 0, nci: Optional[int] = None) -> torch.Tensor:
    """OIHW weights of the convolution that computes the data gradient w.r.t. input channels [ci0, ci0 + nci)."""
    lib = _lib.load()
    w = weight.detach().contiguous().float()
    cout, cin, kh, kw = w.shape
    nci = cin - ci0 if nci is None else nci
    out = torch.empty((nci, cout, kh, kw), dtype=torch.float32, device=w.device)
    check(lib.rsis_conv_dgrad_weights(w.data_ptr(), cout, cin, kh, kw, ci0, nci, out.data_ptr(), _lib.stream_ptr()),
          "conv_dgrad_weights")
    _lib.count_launch(1)
    return out


def conv2d_wgrad(x: Act, dy: Act, kh: int, kw: int, stride: int, pad: int, dw: Optional[torch.Tensor],
                 dbias: Optional[torch.Tensor] = None, accumulate: bool = False, impl: int = IMPL_AUTO):
    """dw (OIHW, float32 contiguous) / dbias (+)= the weight / bias gradient of conv(x) given dy.
    impl AUTO: the tcgen05 kernel when x and dy are split-bf16 and the convolution is a stride-1 1x1 / 3x3."""
    lib = _lib.load()
    if dw is not None:
        assert dw.is_contiguous() and dw.dtype == torch.float32 and tuple(dw.shape) == (dy.c, x.c, kh, kw)
    ws = _lib.wgrad_workspace() if (impl != IMPL_SIMT and x.fmt == FMT_SPLIT_BF16 and dy.fmt == FMT_SPLIT_BF16) \
        else (None, 0)
    check(lib.rsis_conv2d_wgrad(x.ref(), dy.ref(), kh, kw, stride, pad, _ptr(dw), _ptr(dbias), int(accumulate), impl,
                                ws[0], ws[1], _lib.stream_ptr()), "conv2d_wgrad")
    _lib.count_launch((2 if dw is not None else 0) + (1 if dbias is not None else 0))


def dilate2x(x: Act, ho: int, wo: int, fmt: int = FMT_F32) -> Act:
    lib = _lib.load()
    y = Act.empty(x.n, ho, wo, x.c, fmt, x.t.device)
    check(lib.rsis_dilate2x(x.ref(), y.ref(), _lib.stream_ptr()), "dilate2x")
    _lib.count_launch(1)
    return y


def bn_train_bwd(raw: Act, y_act: Optional[Act], dy: Act, weight: Optional[torch.Tensor], mean: torch.Tensor,
                 invstd: torch.Tensor, dx_fmt: int = FMT_F32, want_dres: bool = False,
                 dweight_acc: Optional[torch.Tensor] = None, dbias_acc: Optional[torch.Tensor] = None):
    """Backward of train-mode BatchNorm (+ReLU when y_act is given).  Returns (dx, dres or None, dweight, dbias);
    dweight_acc / dbias_acc (optional) are incremented in place by this call's sums."""
    lib = _lib.load()
    dev = raw.t.device
    c = raw.c
    dx = Act.empty(raw.n, raw.h, raw.w, c, dx_fmt, dev)
    dres = Act.empty(raw.n, raw.h, raw.w, c, FMT_F32, dev) if want_dres else None
    dweight = torch.empty(c, dtype=torch.float32, device=dev)
    dbias = torch.empty(c, dtype=torch.float32, device=dev)
    check(lib.rsis_bn_train_bwd(raw.ref(), y_act.ref() if y_act is not None else None, dy.ref(), _ptr(weight),
                                mean.data_ptr(), invstd.data_ptr(), _bn_workspace(dev, c).data_ptr(),
                                dweight.data_ptr(), dbias.data_ptr(), _ptr(dweight_acc), _ptr(dbias_acc), dx.ref(),
                                dres.ref() if dres is not None else None, _lib.stream_ptr()), "bn_train_bwd")
    _lib.count_launch(3)
    return dx, dres, dweight, dbias


def maxpool3x3s2_bwd(x: Act, dy: Act) -> Act:
    lib = _lib.load()
    dx = Act.empty(x.n, x.h, x.w, x.c, FMT_F32, x.t.device)
    check(lib.rsis_maxpool3x3s2_bwd(x.ref(), dy.ref(), dx.ref(), _lib.stream_ptr()), "maxpool3x3s2_bwd")
    _lib.count_launch(1)
    return dx


def lstm_gates_fwd(gates: Act, c_prev: Optional[torch.Tensor], h_out2: Optional[Act] = None):
    """gates (pre-activations, overwritten with the activated gates) -> (h Act f32, c Act f32)."""
    lib = _lib.load()
    ch = gates.c // 4
    dev = gates.t.device
    h = Act.empty(gates.n, gates.h, gates.w, ch, FMT_F32, dev)
    c = Act.empty(gates.n, gates.h, gates.w, ch, FMT_F32, dev)
    check(lib.rsis_lstm_gates_fwd(gates.ref(), _ptr(c_prev), h.ref(), h_out2.ref() if h_out2 is not None else None,
                                  c.ref(), _lib.stream_ptr()), "lstm_gates_fwd")
    _lib.count_launch(1)
    return h, c


def lstm_gates_bwd(gates: Act, c_prev: Optional[torch.Tensor], c_new: torch.Tensor, dh_a: Optional[Act],
                   dh_b: Optional[Act], dc_next: Optional[Act], dg_fmt: int = FMT_F32):
    """Returns (dgates Act [N,H,W,4Ch] in dg_fmt, dc_prev Act f32)."""
    lib = _lib.load()
    dev = gates.t.device
    dg = Act.empty(gates.n, gates.h, gates.w, gates.c, dg_fmt, dev)
    dcp = Act.empty(gates.n, gates.h, gates.w, gates.c // 4, FMT_F32, dev)
    r = lambda a: a.ref() if a is not None else None  # noqa: E731
    check(lib.rsis_lstm_gates_bwd(gates.ref(), _ptr(c_prev), c_new.data_ptr(), r(dh_a), r(dh_b), r(dc_next), dg.ref(),
                                  dcp.t.data_ptr(), _lib.stream_ptr()), "lstm_gates_bwd")
    _lib.count_launch(1)
    return dg, dcp


def global_maxpool(h: Act, side_packed: torch.Tensor, side_offset: int):
    """Folds the per-(image, channel) maximum of h and its pixel index into side_packed (int64 [N, F], zeroed)."""
    lib = _lib.load()
    assert side_packed.dtype == torch.int64 and side_packed.is_contiguous()
    check(lib.rsis_global_maxpool(h.ref(), side_packed.data_ptr(), side_packed.shape[1], side_offset,
                                  _lib.stream_ptr()), "global_maxpool")
    _lib.count_launch(1)


def global_maxpool_finish(side_packed: torch.Tensor):
    """Unpacks side_packed into (keys int32 [N,F] -- the input of class_stop_heads --, pixel indices int32 [N,F])."""
    lib = _lib.load()
    n, f = side_packed.shape
    keys = torch.empty((n, f), dtype=torch.int32, device=side_packed.device)
    idx = torch.empty((n, f), dtype=torch.int32, device=side_packed.device)
    check(lib.rsis_global_maxpool_finish(side_packed.data_ptr(), n, f, keys.data_ptr(), idx.data_ptr(),
                                         _lib.stream_ptr()), "global_maxpool_finish")
    _lib.count_launch(1)
    return keys, idx


def global_maxpool_bwd(dside: torch.Tensor, side_idx: torch.Tensor, side_offset: int, dh: Act):
    lib = _lib.load()
    assert dside.is_contiguous() and dside.shape == side_idx.shape
    check(lib.rsis_global_maxpool_bwd(dside.data_ptr(), side_idx.data_ptr(), side_idx.shape[1], side_offset, dh.ref(),
                                      _lib.stream_ptr()), "global_maxpool_bwd")
    _lib.count_launch(1)


def upsample_bilinear_bwd(dy: Act, h: int, w: int) -> Act:
    lib = _lib.load()
    dx = Act.empty(dy.n, h, w, dy.c, FMT_F32, dy.t.device)
    check(lib.rsis_upsample_bilinear_bwd(dy.ref(), dx.ref(), _lib.stream_ptr()), "upsample_bilinear_bwd")
    _lib.count_launch(1)
    return dx


def class_stop_heads_bwd(feat: torch.Tensor, class_probs: torch.Tensor, dclass: Optional[torch.Tensor],
                         dstop: Optional[torch.Tensor], w_class: torch.Tensor, w_stop: torch.Tensor, into=None):
    """Returns (dfeat [N,F], dw_class, db_class, dw_stop, db_stop).  `into` (optional): four float32 contiguous
    buffers (dw_class, db_class, dw_stop, db_stop) that are incremented in place instead of fresh zero tensors."""
    lib = _lib.load()
    n, f = feat.shape
    nc = w_class.shape[0]
    dev = feat.device
    dfeat = torch.empty((n, f), dtype=torch.float32, device=dev)
    if into is not None:
        dwc, dbc, dws, dbs = into
    else:
        dwc = torch.zeros((nc, f), dtype=torch.float32, device=dev)
        dbc = torch.zeros(nc, dtype=torch.float32, device=dev)
        dws = torch.zeros((1, f), dtype=torch.float32, device=dev)
        dbs = torch.zeros(1, dtype=torch.float32, device=dev)
    scratch = torch.empty(n * (nc + 1), dtype=torch.float32, device=dev)
    check(lib.rsis_class_stop_heads_bwd(feat.data_ptr(), class_probs.data_ptr(), _ptr(dclass), _ptr(dstop), n, f,
                                        w_class.data_ptr(), nc, w_stop.data_ptr(), scratch.data_ptr(),
                                        dfeat.data_ptr(), dwc.data_ptr(), dbc.data_ptr(), dws.data_ptr(),
                                        dbs.data_ptr(), _lib.stream_ptr()), "class_stop_heads_bwd")
    _lib.count_launch(2)
    return dfeat, dwc, dbc, dws, dbs
