"""RLE oracle pinning (CPU): the numpy restatement and the host-side string packer against the reference's own C code
(coco/common/maskApi.c compiled by oracle/Makefile into oracle/_ref/libmaskapi.so; skipped when that file is absent
and /root/reference is not there to build it)."""
import os
import subprocess

import numpy as np
import pytest

from oracle import rle_oracle as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rle_cases():
    rng = np.random.default_rng(0)
    m = (rng.random((8, 23, 17)) < 0.4).astype(np.uint8)
    m[1] = 0                      # empty mask: one run
    m[2] = 1                      # full mask: [0, a]
    m[3] = 0
    m[3, 0, 0] = 1                # starts with a one
    m[4] = 0
    m[4, -1, -1] = 1              # ends with a one
    m[5, :, ::2] = 1              # column stripes
    m[5, :, 1::2] = 0
    blob = np.zeros((64, 48), dtype=np.uint8)
    blob[10:40, 5:30] = 1
    return m, blob[None]


@pytest.fixture(scope="module")
def ref_lib():
    if not R.ref_available():
        if not os.path.isdir("/root/reference/src/coco/common"):
            pytest.skip("oracle/_ref/libmaskapi.so absent and no reference tree to build it from")
        subprocess.run(["make", "-f", os.path.join(ROOT, "oracle", "Makefile")], check=True, cwd=ROOT)
    return R


def test_numpy_restatement_matches_reference_c(ref_lib):
    for masks in rle_cases():
        a, areas_a = R.rle_encode_ref(masks)
        b, areas_b = R.rle_encode(masks)
        assert (areas_a == areas_b).all()
        for x, y in zip(a, b):
            assert len(x) == len(y) and (x == y).all()
            assert int(x.sum()) == masks.shape[1] * masks.shape[2]


def test_string_packer_matches_reference_c(ref_lib):
    from rsis_b200.postprocess import rle_to_string
    for masks in rle_cases():
        cnts, _ = R.rle_encode(masks)
        for c in cnts:
            assert rle_to_string(c) == R.rle_to_string_ref(c, masks.shape[1], masks.shape[2])
    big = np.array([0, 70000, 3, 70001, 1, 5, 1234567], dtype=np.uint32)   # multi-char and negative differences
    assert rle_to_string(big) == R.rle_to_string_ref(big, 1200, 1200)
