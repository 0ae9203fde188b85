"""ctypes binding of the C ABI declared in include/rsis_b200.h.

The CUDA library is the product: if it is missing or fails to load this module raises -- there is no
PyTorch/CPU fallback for any primitive.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import build as _build

FMT_F32 = 0
FMT_SPLIT_BF16 = 1
IMPL_AUTO = 0
IMPL_SIMT = 1
IMPL_TCGEN05 = 2
ABI_VERSION = 21


class Tensor(C.Structure):
    """`rsis_tensor`: NHWC activation view (cstride = pixel pitch in elements, 0 = dense)."""
    _fields_ = [("data", C.c_void_p), ("fmt", C.c_int32), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("c", C.c_int32), ("cstride", C.c_int32)]


class ConvWeights(C.Structure):
    """`rsis_conv_weights`."""
    _fields_ = [("w_kc", C.c_void_p), ("w_umma", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p),
                ("cout", C.c_int32), ("cin", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32),
                ("gate_interleaved", C.c_int32), ("w_umma_il", C.c_void_p)]


class CellArgs(C.Structure):
    """`rsis_cell_args`: one cell of a grouped (wavefront) launch."""
    _fields_ = [("x", C.POINTER(Tensor)), ("w", C.POINTER(ConvWeights)), ("c_prev", C.c_void_p),
                ("gate_preact", C.c_void_p), ("h_out", C.POINTER(Tensor)), ("h_split", C.POINTER(Tensor)),
                ("c_out", C.POINTER(Tensor)), ("side_max", C.c_void_p), ("side_stride", C.c_int32),
                ("side_offset", C.c_int32)]


_P = C.c_void_p
_I = C.c_int
_TP = C.POINTER(Tensor)
_WP = C.POINTER(ConvWeights)

# name -> (restype, argtypes); must list every symbol include/rsis_b200.h declares (tests/test_abi.py checks)
SIGNATURES = {
    "rsis_abi_version": (_I, []),
    "rsis_strerror": (C.c_char_p, [_I]),
    "rsis_last_cuda_error": (C.c_char_p, []),
    "rsis_device_check": (_I, []),
    "rsis_has_tcgen05": (_I, []),
    "rsis_conv_pack_bytes_simt": (C.c_size_t, [_I, _I, _I, _I]),
    "rsis_conv_pack_bytes_affine": (C.c_size_t, [_I]),
    "rsis_conv_pack": (_I, [_P, _P, _P, _P, _P, _P, C.c_float, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "rsis_conv_umma_kpad": (_I, [_I, _I, _I, C.POINTER(C.c_int32)]),
    "rsis_conv_umma_coutpad": (_I, [_I]),
    "rsis_conv_pack_bytes_umma": (C.c_size_t, [_I, _I, _I, _I, C.POINTER(C.c_int32)]),
    "rsis_conv_pack_umma": (_I, [_P, _I, _I, _I, _I, _I, C.POINTER(C.c_int32), _I, _P, _P]),
    "rsis_conv_pack_all": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, C.c_float, _I, _I,
                                C.POINTER(C.c_int32), _P, _P, _P, _P, _P]),
    "rsis_nchw_to_nhwc": (_I, [_P, _TP, _P]),
    "rsis_convert": (_I, [_TP, _TP, _P]),
    "rsis_im2col": (_I, [_TP, _I, _I, _I, _I, _TP, _P]),
    "rsis_conv_workspace_bytes": (C.c_size_t, []),
    "rsis_conv2d": (_I, [_TP, _I, _WP, _TP, _TP, _TP, _I, _I, _I, _I, _P, C.c_size_t, _P]),
    "rsis_maxpool3x3s2": (_I, [_TP, _TP, _P]),
    "rsis_bn_workspace_bytes": (C.c_size_t, [_I]),
    "rsis_bn_train_stats": (_I, [_TP, _P, _P, C.c_float, C.c_float, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "rsis_affine_act": (_I, [_TP, _P, _P, _TP, _I, _TP, _TP, _P]),
    "rsis_convlstm_cell": (_I, [_TP, _I, _WP, _P, _P, _TP, _TP, _TP, _P, _I, _I, _I, _P, C.c_size_t, _P]),
    "rsis_upsample_bilinear": (_I, [_TP, _TP, _P]),
    "rsis_mask_head": (_I, [_TP, _P, _P, _I, _P, _P, C.c_int64, _P]),
    "rsis_upsample_mask_head": (_I, [_TP, _P, _P, _I, _I, _I, _P, _P, C.c_int64, _P]),
    "rsis_class_stop_heads": (_I, [_P, _I, _I, _P, _P, _I, _P, _P, _P, _P, C.c_int64, _P, _P, C.c_int64, _P]),
    "rsis_upsample_mask_head_steps": (_I, [_TP, _I, _P, _P, _I, _I, _I, _P, C.c_int64, C.c_int64, _P]),
    "rsis_class_stop_heads_steps": (_I, [_P, _I, _I, _I, _P, _P, _I, _P, _P, _P, C.c_int64, C.c_int64, _P, C.c_int64,
                                         C.c_int64, _P]),
    # backward primitives
    "rsis_conv_dgrad_weights": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "rsis_wgrad_workspace_bytes": (C.c_size_t, []),
    "rsis_conv2d_wgrad": (_I, [_TP, _TP, _I, _I, _I, _I, _P, _P, _I, _I, _P, C.c_size_t, _P]),
    "rsis_dilate2x": (_I, [_TP, _TP, _P]),
    "rsis_bn_train_bwd": (_I, [_TP, _TP, _TP, _P, _P, _P, _P, _P, _P, _P, _P, _TP, _TP, _P]),
    "rsis_maxpool3x3s2_bwd": (_I, [_TP, _TP, _TP, _P]),
    "rsis_lstm_gates_fwd": (_I, [_TP, _P, _TP, _TP, _TP, _P]),
    "rsis_lstm_gates_bwd": (_I, [_TP, _P, _P, _TP, _TP, _TP, _TP, _P, _P]),
    "rsis_global_maxpool": (_I, [_TP, _P, _I, _I, _P]),
    "rsis_global_maxpool_finish": (_I, [_P, _I, _I, _P, _P, _P]),
    "rsis_global_maxpool_bwd": (_I, [_P, _P, _I, _I, _TP, _P]),
    "rsis_upsample_bilinear_bwd": (_I, [_TP, _TP, _P]),
    "rsis_soft_iou_workspace_bytes": (C.c_size_t, [_I, _I]),
    "rsis_soft_iou_cost": (_I, [_P, _P, _I, _I, _I, C.c_int64, C.c_float, C.c_float, _P, _P, C.c_int64, C.c_int64, _P,
                                _P, _P]),
    "rsis_soft_iou_bwd": (_I, [_P, _P, _I, _I, C.c_int64, _P, _P, _P, C.c_float, _P, _P]),
    "rsis_hungarian_match": (_I, [_P, C.c_int64, C.c_int64, C.c_int64, _I, _I, _I, _P, _I, _P, _P]),
    "rsis_set_precision": (_I, [_I]),
    "rsis_get_precision": (_I, []),
    "rsis_set_static_weights": (_I, [_I]),
    "rsis_convlstm_cell_group_max": (_I, []),
    "rsis_convlstm_cell_group": (_I, [C.POINTER(CellArgs), _I, _P]),
    "rsis_upsample_bilinear_group": (_I, [_TP, _TP, _I, _P]),
    "rsis_masked_nll_fwd": (_I, [_P, _P, _P, _P, _I, _I, _P, _P, _P]),
    "rsis_masked_nll_bwd": (_I, [_P, _P, _P, _P, _P, C.c_int64, _I, _I, _P, _P]),
    "rsis_masked_bce_fwd": (_I, [_P, _P, _P, C.c_float, C.c_int64, _P, _P, _P]),
    "rsis_masked_bce_bwd": (_I, [_P, _P, _P, _P, _P, C.c_int64, C.c_int64, _P, _P]),
    "rsis_rle_workspace_bytes": (C.c_size_t, [_I, _I, _I]),
    "rsis_rle_encode": (_I, [_P, C.c_float, _P, _I, _I, _I, _P, _P, _I, _P, _P, _P]),
    "rsis_adam_step": (_I, [_P, _P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int64,
                            _I, _P]),
    "rsis_class_stop_heads_bwd": (_I, [_P, _P, _P, _P, _I, _I, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
}

_lib = None


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """Loads librsis_b200.so (building it with nvcc when absent). Raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if not os.path.exists(path):
        _build.build()
    try:
        lib = C.CDLL(path)
    except OSError as e:  # pragma: no cover
        raise RuntimeError(f"rsis_b200: cannot load the CUDA library {path}: {e}. There is no CPU fallback; "
                           f"build it with `python -m rsis_b200.build`.") from e
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.rsis_abi_version() != ABI_VERSION:
        raise RuntimeError("rsis_b200: ABI version mismatch between the Python package and librsis_b200.so")
    _lib = lib
    return lib


_launches = 0


def count_launch(n: int = 1):
    """Book-keeping for bench.py's `gpu_launches`: every ABI call that enqueues kernels reports how many."""
    global _launches
    _launches += n


def check(status: int, what: str):
    if status != 0:
        lib = load()
        msg = lib.rsis_strerror(status).decode()
        cuda = lib.rsis_last_cuda_error().decode()
        raise RuntimeError(f"rsis_b200.{what} failed: {msg}" + (f" [{cuda}]" if status == -3 and cuda else ""))


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


_workspaces = {}
_lane = 0


class lane:
    """Context manager: launches issued inside use split-K exchange buffer number `n` of the device.  Launches that
    may run concurrently (a forked side stream) must not share a buffer; everything else uses lane 0."""

    def __init__(self, n: int):
        self.n = n

    def __enter__(self):
        global _lane
        self.prev, _lane = _lane, self.n
        return self

    def __exit__(self, *exc):
        global _lane
        _lane = self.prev
        return False


_no_splitk = False


class no_splitk:
    """Context manager: launches issued inside get no split-K exchange buffer (the kernels then keep every tile's K
    loop in one CTA: fewer, longer CTAs -- what a pipelined schedule of concurrent small launches wants)."""

    def __enter__(self):
        global _no_splitk
        self.prev, _no_splitk = _no_splitk, True
        return self

    def __exit__(self, *exc):
        global _no_splitk
        _no_splitk = self.prev
        return False


def workspace():
    """(pointer, bytes) of the split-K exchange buffer of the current (device, lane): allocated and zero-filled once;
    the kernels leave its counters at zero.  All launches of a lane are issued in one stream order."""
    if _no_splitk:
        return None, 0
    dev = torch.cuda.current_device()
    key = (dev, _lane)
    ws = _workspaces.get(key)
    if ws is None:
        n = load().rsis_conv_workspace_bytes()
        ws = torch.zeros(n, dtype=torch.uint8, device=torch.device("cuda", dev))
        _workspaces[key] = ws
    return ws.data_ptr(), ws.numel()


_wgrad_workspaces = {}


def wgrad_workspace():
    """(pointer, bytes) of the device's weight-gradient scratch (zero-filled once; the kernels leave it zeroed)."""
    dev = torch.cuda.current_device()
    ws = _wgrad_workspaces.get(dev)
    if ws is None:
        n = load().rsis_wgrad_workspace_bytes()
        ws = torch.zeros(n, dtype=torch.uint8, device=torch.device("cuda", dev))
        _wgrad_workspaces[dev] = ws
    return ws.data_ptr(), ws.numel()


class Act:
    """An NHWC activation held in a torch tensor: float32 [N,H,W,C] or split-bf16 [2,N,H,W,C] (hi|lo planes).

    `Act.slice(c0, c)` is a pitched channel-slice view of the same storage (rsis_tensor.cstride): producers write
    their part of a concatenated buffer in place, so `torch.cat(..., 1)` never runs."""
    __slots__ = ("t", "fmt", "n", "h", "w", "c", "desc", "source_key", "c0", "pitch")

    def __init__(self, t: torch.Tensor, fmt: int, c0: int = 0, c: int = None):
        if fmt == FMT_F32:
            assert t.dtype == torch.float32 and t.dim() == 4 and t.is_contiguous()
            n, h, w, pitch = t.shape
            esize = 4
        else:
            assert t.dtype == torch.bfloat16 and t.dim() == 5 and t.shape[0] == 2 and t.is_contiguous()
            _, n, h, w, pitch = t.shape
            esize = 2
        c = pitch - c0 if c is None else c
        assert 0 <= c0 and c > 0 and c0 + c <= pitch
        self.t, self.fmt, self.n, self.h, self.w, self.c = t, fmt, n, h, w, c
        self.c0, self.pitch = c0, pitch
        self.desc = Tensor(t.data_ptr() + c0 * esize, fmt, n, h, w, c, 0 if c == pitch else pitch)
        self.source_key = None

    @property
    def dense(self) -> bool:
        return self.c == self.pitch

    def slice(self, c0: int, c: int) -> "Act":
        return Act(self.t, self.fmt, self.c0 + c0, c)

    def valid_for(self, src: torch.Tensor) -> bool:
        """True while `src` (the float32 tensor this operand copy was derived from) is unchanged."""
        return self.source_key == (src.data_ptr(), src._version, tuple(src.shape))

    @staticmethod
    def empty(n, h, w, c, fmt, device) -> "Act":
        if fmt == FMT_F32:
            return Act(torch.empty((n, h, w, c), dtype=torch.float32, device=device), fmt)
        return Act(torch.empty((2, n, h, w, c), dtype=torch.bfloat16, device=device), fmt)

    @staticmethod
    def zeros(n, h, w, c, fmt, device) -> "Act":
        if fmt == FMT_F32:
            return Act(torch.zeros((n, h, w, c), dtype=torch.float32, device=device), fmt)
        return Act(torch.zeros((2, n, h, w, c), dtype=torch.bfloat16, device=device), fmt)

    @staticmethod
    def strided_view(t: torch.Tensor) -> "Optional[Act]":
        """Zero-copy Act over a logical [N,C,H,W] float32 tensor whose memory is NHWC with a pixel pitch >= C (a dense
        channels-last tensor, or a channel slice of a wider NHWC buffer -- e.g. a gradient returned as a slice of a
        concatenated gradient buffer).  Returns None when the strides do not have that form."""
        if t.dim() != 4 or t.dtype != torch.float32:
            return None
        n, c, h, w = t.shape
        sn, sc, sh, sw = t.stride()
        pitch = sw if w > 1 else (sh if h > 1 else (sn // max(h * w, 1) if n > 1 else c))
        if pitch < c or pitch % 4 != 0 or t.data_ptr() % 16 != 0:
            return None
        want = {"c": 1, "w": pitch, "h": w * pitch, "n": h * w * pitch}
        if (c > 1 and sc != want["c"]) or (w > 1 and sw != want["w"]) or (h > 1 and sh != want["h"]) or \
                (n > 1 and sn != want["n"]):
            return None
        a = object.__new__(Act)
        a.t, a.fmt, a.n, a.h, a.w, a.c = t, FMT_F32, n, h, w, c
        a.c0, a.pitch = 0, pitch
        a.desc = Tensor(t.data_ptr(), FMT_F32, n, h, w, c, 0 if c == pitch else pitch)
        a.source_key = None
        return a

    @staticmethod
    def from_nchw_view(x: torch.Tensor) -> "Act":
        """Zero-copy when `x` ([N,C,H,W] float32) is channels_last-contiguous; otherwise one torch copy."""
        assert x.dim() == 4 and x.dtype == torch.float32
        xl = x.permute(0, 2, 3, 1)
        if not xl.is_contiguous():
            xl = xl.contiguous()
        return Act(xl, FMT_F32)

    def nchw(self) -> torch.Tensor:
        """float32 activations as the reference's logical [N,C,H,W] (channels_last memory, zero-copy)."""
        assert self.fmt == FMT_F32 and self.dense
        return self.t.permute(0, 3, 1, 2)

    def ref(self):
        return C.byref(self.desc)

    def float(self) -> torch.Tensor:
        """[N,H,W,C] float32 values (a torch op; tests/debug only)."""
        sl = slice(self.c0, self.c0 + self.c)
        if self.fmt == FMT_F32:
            return self.t[..., sl]
        return self.t[0][..., sl].float() + self.t[1][..., sl].float()
