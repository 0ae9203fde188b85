"""TEST INFRASTRUCTURE -- RLE encoding of binary masks, the COCO wire format eval.py:97-127 produces through
pycocotools (`mask.encode(np.asfortranarray(segmentation))`).

Two checkers, used by tests/ only:
  * `rle_encode_ref`  -- the reference's OWN C code, /root/reference/src/coco/common/maskApi.c:32-41 (`rleEncode`) and
                         :203-207 (`rleArea`), compiled from the sources where they lie by oracle/Makefile into
                         oracle/_ref/libmaskapi.so (git-ignored; travels to the GPU box with the snapshot) and bound
                         here with ctypes;
  * `rle_encode`      -- a numpy restatement of the same loop (maskApi.c:36-39): column-major scan, the first count is
                         the number of leading ZEROS (0 when the mask starts with a one).
tests/test_oracle_golden.py pins the restatement against the compiled reference.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_REF_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libmaskapi.so")


class _RLE(C.Structure):  # maskApi.h:13  typedef struct { siz h, w, m; uint *cnts; } RLE;
    _fields_ = [("h", C.c_ulong), ("w", C.c_ulong), ("m", C.c_ulong), ("cnts", C.POINTER(C.c_uint))]


def ref_available() -> bool:
    return os.path.exists(_REF_SO)


def rle_encode_ref(masks: np.ndarray):
    """masks: uint8 [n, h, w] (row-major, 0/1).  Runs the reference's rleEncode on the column-major (Fortran) bytes, as
    eval.py:122 does.  Returns (list of count arrays, areas)."""
    lib = C.CDLL(_REF_SO)
    n, h, w = masks.shape
    fortran = np.ascontiguousarray(masks.transpose(0, 2, 1)).astype(np.uint8)  # [n][x][y] == column-major bytes
    R = (_RLE * n)()
    lib.rleEncode(R, fortran.ctypes.data_as(C.POINTER(C.c_ubyte)), C.c_ulong(h), C.c_ulong(w), C.c_ulong(n))
    areas = (C.c_uint * n)()
    lib.rleArea(R, C.c_ulong(n), areas)
    out = [np.array([R[i].cnts[k] for k in range(R[i].m)], dtype=np.uint32) for i in range(n)]
    for i in range(n):
        lib.rleFree(C.byref(R[i]))
    return out, np.array(list(areas), dtype=np.uint32)


def rle_to_string_ref(cnts: np.ndarray, h: int, w: int) -> bytes:
    """The reference's rleToString (maskApi.c:203-215) on one count array."""
    lib = C.CDLL(_REF_SO)
    lib.rleToString.restype = C.c_void_p
    arr = (C.c_uint * len(cnts))(*[int(v) for v in cnts])
    R = _RLE()
    lib.rleInit(C.byref(R), C.c_ulong(h), C.c_ulong(w), C.c_ulong(len(cnts)), arr)
    ptr = lib.rleToString(C.byref(R))
    out = C.string_at(ptr)
    C.CDLL(None).free(C.c_void_p(ptr))
    lib.rleFree(C.byref(R))
    return out


def rle_encode(masks: np.ndarray):
    """numpy restatement of maskApi.c:32-41 / :203-207."""
    n, h, w = masks.shape
    out, areas = [], []
    for i in range(n):
        t = masks[i].T.reshape(-1).astype(np.uint8)          # column-major scan order
        prev = np.concatenate([[0], t[:-1]])                  # p starts at 0 (maskApi.c:36)
        pos = np.flatnonzero(t != prev)
        edges = np.concatenate([[0], pos, [t.size]])
        out.append(np.diff(edges).astype(np.uint32))          # cnts[0] = leading zeros (0 if the mask starts with 1)
        areas.append(int(t.sum()))
    return out, np.array(areas, dtype=np.uint32)
