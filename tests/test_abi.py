"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares
(no compute calls -- there is no GPU here), and the ctypes table covers the header."""
import ctypes
import os
import re

import pytest

from rsis_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rsis_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(rsis_[a-z0-9_]+)\s*\(", src)
    return sorted(set(n for n in names if n not in ("rsis_tensor", "rsis_conv_weights")))


def test_header_functions_exported():
    lib = ctypes.CDLL(_lib.lib_path()) if os.path.exists(_lib.lib_path()) else None
    if lib is None:
        _lib.load()
        lib = ctypes.CDLL(_lib.lib_path())
    names = declared_functions()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rsis_b200.h but not exported by librsis_b200.so"


def test_ctypes_table_covers_header():
    assert sorted(_lib.SIGNATURES.keys()) == declared_functions()


def test_host_only_entry_points():
    lib = _lib.load()
    assert lib.rsis_abi_version() == _lib.ABI_VERSION
    assert lib.rsis_strerror(0) == b"ok"
    for code in (-1, -2, -3, -4, -5, -99):
        assert len(lib.rsis_strerror(code)) > 0
    assert lib.rsis_last_cuda_error() == b"" or isinstance(lib.rsis_last_cuda_error(), bytes)
    # packed sizes: [KH*KW*Cin][cout_pad64] float32
    assert lib.rsis_conv_pack_bytes_simt(512, 256, 3, 3) == 9 * 256 * 512 * 4
    assert lib.rsis_conv_pack_bytes_simt(16, 64, 3, 3) == 9 * 64 * 64 * 4
    assert lib.rsis_conv_pack_bytes_affine(21) == 64 * 4
    assert lib.rsis_conv_pack_bytes_simt(0, 1, 1, 1) == 0
    src_c = (ctypes.c_int32 * 3)(128, 128, 64)
    assert lib.rsis_conv_umma_kpad(3, 3, 3, src_c) == 9 * (2 + 2 + 1) * 64
    assert lib.rsis_conv_umma_coutpad(40) == 48
    assert lib.rsis_has_tcgen05() in (0, 1)


def test_device_check_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    assert lib.rsis_device_check() != 0  # no silent CPU fallback: the check reports an error


def test_status_raises():
    with pytest.raises(RuntimeError, match="bad argument"):
        _lib.check(-1, "unit")


def test_static_weights_mode_is_a_host_side_switch():
    """rsis_set_static_weights (ABI v20) returns the previous setting; ops.static_weights restores it on exit."""
    from rsis_b200 import ops
    lib = _lib.load()
    prev = lib.rsis_set_static_weights(0)
    try:
        assert lib.rsis_set_static_weights(1) == 0
        assert lib.rsis_set_static_weights(0) == 1
        with ops.static_weights():
            assert lib.rsis_set_static_weights(1) == 1   # on inside the context
        assert lib.rsis_set_static_weights(0) == 0       # restored to off
    finally:
        lib.rsis_set_static_weights(prev)
