"""GPU parity of the device RLE / area kernel against the numpy restatement of
maskApi.c:32-41 and, when oracle/_ref/libmaskapi.so travelled with the snapshot, against the reference's own C code.
Bit-exact: integer work."""
import numpy as np
import pytest
import torch

from test_rle_cpu import rle_cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def PP():
    from rsis_b200 import _lib, postprocess
    assert _lib.load().rsis_device_check() == 0
    return postprocess


def _check(PP, masks_u8, oracle_fn, probs=None, th=0.5, ignore=None):
    want, want_area = oracle_fn(masks_u8)
    p = torch.from_numpy(masks_u8.astype(np.float32)) if probs is None else probs
    counts, n_runs, areas = PP.rle_encode(p.cuda(), th, None if ignore is None else torch.from_numpy(ignore).cuda(),
                                          max_runs=masks_u8.shape[1] * masks_u8.shape[2] + 1)
    counts, n_runs, areas = counts.cpu().numpy(), n_runs.cpu().numpy(), areas.cpu().numpy()
    assert (areas.astype(np.uint32) == want_area).all()
    for i, w in enumerate(want):
        assert n_runs[i] == len(w)
        assert (counts[i, :len(w)].astype(np.uint32) == w).all()


def test_rle_matches_oracle_edge_cases(PP):
    from oracle import rle_oracle as R
    for masks in rle_cases():
        _check(PP, masks, R.rle_encode)
        if R.ref_available():
            _check(PP, masks, R.rle_encode_ref)


@pytest.mark.parametrize("shape", [(80, 256, 256), (3, 100, 37), (2, 512, 1024), (1, 1, 1)])
def test_rle_threshold_ignore_and_sizes(PP, shape):
    from oracle import rle_oracle as R
    n, h, w = shape
    gen = torch.Generator().manual_seed(h + w)
    # smooth blobs: a low-resolution random field upsampled, like sigmoid(mask logits)
    low = torch.rand((n, 1, max(h // 16, 1), max(w // 16, 1)), generator=gen)
    probs = torch.nn.functional.interpolate(low, size=(h, w), mode="bilinear", align_corners=True)[:, 0].contiguous()
    ignore = (torch.rand((n, h, w), generator=gen) < 0.05).to(torch.uint8).numpy()
    th = 0.55
    seg = (probs.numpy() > th).astype(np.uint8)
    seg[ignore == 1] = 0
    _check(PP, seg, R.rle_encode, probs=probs, th=th, ignore=ignore)
    if R.ref_available():
        _check(PP, seg, R.rle_encode_ref, probs=probs, th=th, ignore=ignore)


def test_encode_instances_strings(PP):
    from oracle import rle_oracle as R
    masks, _ = rle_cases()
    out = PP.encode_instances(torch.from_numpy(masks.astype(np.float32)).cuda(), 0.5)
    want, areas = R.rle_encode(masks)
    for o, w, a in zip(out, want, areas):
        assert o["size"] == [masks.shape[1], masks.shape[2]] and o["area"] == int(a)
        assert o["counts"] == PP.rle_to_string(w)
        if R.ref_available():
            assert o["counts"] == R.rle_to_string_ref(w, masks.shape[1], masks.shape[2])


@pytest.mark.parametrize("shape", [(5, 64, 64, 96, 80), (3, 48, 64, 37, 101), (8, 256, 256, 375, 500)])
def test_resize_matches_scipy_zoom_and_rle_of_it(PP, shape):
    """eval.py:97-127 end to end: zoom(order=1) -> threshold -> RLE.  The interpolation runs in fp32 on the device and in
    fp64 in scipy, so pixels whose interpolated value is within 1e-5 of the threshold may differ: the values are
    compared with a tolerance, the RLE against the oracle applied to the DEVICE's own resized mask."""
    from scipy.ndimage import zoom
    from oracle import rle_oracle as R
    n, h, w, H, W = shape
    gen = torch.Generator().manual_seed(H)
    low = torch.rand((n, 1, max(h // 8, 2), max(w // 8, 2)), generator=gen)
    probs = torch.nn.functional.interpolate(low, size=(h, w), mode="bicubic", align_corners=True)[:, 0].contiguous()
    want = np.stack([zoom(probs[i].numpy().reshape(h, w, 1), [float(H) / h, float(W) / w, 1], order=1)[:, :, 0]
                     for i in range(n)])
    got = PP.resize_masks(probs.cuda(), H, W)
    assert tuple(got.shape) == (n, H, W) == want.shape
    assert float(np.abs(got.cpu().numpy() - want).max()) <= 2e-5
    th = 0.5
    out = PP.encode_instances(probs.cuda(), th, size=(H, W))
    seg = (got.cpu().numpy() > th).astype(np.uint8)
    cnts, areas = R.rle_encode(seg)
    for o, c, a in zip(out, cnts, areas):
        assert o["size"] == [H, W] and o["area"] == int(a) and o["counts"] == PP.rle_to_string(c)
    agree = ((want > th) == (seg == 1)).mean()
    assert agree >= 0.9999
