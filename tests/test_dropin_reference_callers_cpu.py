"""Drop-in proof: the reference's OWN callers -- `test()` (/root/reference/src/test.py:16-50) and
`runIter()` (/root/reference/src/train.py:54-197) -- executed UNMODIFIED against `rsis_b200.FeatureExtractor` /
`rsis_b200.RSIS`, and compared with the goldens the unmodified reference MODULES produced
(tests/golden/e2e_b2_64x64_t3.npz, run_iter.npz).

Runs in the build container only (needs /root/reference for the callers' text; skipped elsewhere).  The ABI is the CPU
stand-in of tests/fake_abi.py, so what is proven here is the module SURFACE: constructor arguments, state_dict keys,
call signatures, argument / return shapes and dtypes, train / eval switching, autograd wiring, `.data`, `.size()`,
`torch.cat` on the returned tensors -- everything test.py / train.py touch.  The CUDA kernels behind the same surface
are checked against the same goldens on the GPU box (tests/test_gpu_parity.py, tests/test_gpu_backward.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_shims as rs

pytestmark = pytest.mark.skipif(not rs.available(), reason="reference tree not present (GPU box)")


@pytest.fixture
def fake(monkeypatch):
    import fake_abi
    return fake_abi.install(monkeypatch)


def rel(a, b):
    a = torch.as_tensor(a, dtype=torch.float32)
    b = torch.as_tensor(b, dtype=torch.float32)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_reference_test_loop_runs_on_repo_modules(fake, golden_dir):
    """`test(args, encoder, decoder, x)` of src/test.py, imported as is, driving the repo's modules."""
    import rsis_b200
    from oracle import synth_weights as sw
    ref = rs.load_reference()
    g = np.load(os.path.join(golden_dir, "e2e_b2_64x64_t3.npz"))
    wseed, iseed, B, H, W, T, ncls, stride = [int(v) for v in g["meta"]]
    args = rs.make_args(num_classes=ncls, maxseqlen=T)          # the FloorInt hidden_size the reference needs (Py2 `/`)
    enc, dec = rsis_b200.FeatureExtractor(args), rsis_b200.RSIS(args)
    enc.load_state_dict(sw.encoder_state_dict(wseed))           # the reference's own checkpoint format
    dec.load_state_dict(sw.decoder_state_dict(wseed, num_classes=ncls))
    x = sw.synthetic_images(iseed, B, H, W)
    masks, classes, stops = ref.test(args, enc, dec, x)         # <- the reference's function, unmodified
    assert not enc.training and not dec.training                # test.py:33-34 switched the modules to eval
    assert tuple(masks.shape) == (B, T, H, W) and tuple(classes.shape) == (B, T, ncls) and tuple(stops.shape) == (B, T, 1)
    assert rel(masks[:, :, ::stride, ::stride], g["masks"]) < 2e-5
    assert rel(classes, g["classes"]) < 2e-5
    assert rel(stops, g["stops"]) < 2e-5
    # and the repo's own test() returns the same thing through the same surface
    args.cuda_graph = False
    m2, c2, s2 = rsis_b200.test(args, enc, dec, x)
    assert rel(m2, masks) < 1e-6 and rel(c2, classes) < 1e-6 and rel(s2, stops) < 1e-6


def test_reference_run_iter_runs_on_repo_modules(fake, golden_dir):
    """`runIter` of src/train.py (text executed as oracle/run_iter_ref.py does for the golden), in 'train' mode, with the
    reference's own criteria / matching and torch optimisers, on the repo's modules: losses, matched classes and the
    gradient digests equal those the unmodified reference modules produced."""
    import rsis_b200
    import torch.nn as nn
    from torch.autograd import Variable
    from oracle import run_iter_ref as rr, synth_weights as sw
    from oracle.make_golden import _grad_digest
    rs.load_reference()
    import hungarian as ref_h           # noqa: E402  (reference files, unmodified)
    import objectives as ref_o          # noqa: E402
    g = np.load(os.path.join(golden_dir, "run_iter.npz"))
    num_classes = 5
    args = rr.iter_args()
    margs = rs.make_args(num_classes=num_classes, maxseqlen=args.maxseqlen)
    enc, dec = rsis_b200.FeatureExtractor(margs), rsis_b200.RSIS(margs)
    enc.load_state_dict(sw.encoder_state_dict(1))
    dec.load_state_dict(sw.decoder_state_dict(1, num_classes=num_classes))
    x, y_mask, y_class, sw_mask, sw_class = rr.iter_inputs(gt=args.gt_maxseqlen, num_classes=num_classes)
    crits = (ref_o.softIoULoss(), ref_o.MaskedNLLLoss(balance_weight=None), ref_o.MaskedBCELoss(balance_weight=None))
    optims = (torch.optim.SGD(enc.parameters(), lr=0.0), torch.optim.SGD(dec.parameters(), lr=0.0))
    ns = {"torch": torch, "nn": nn, "np": np, "Variable": Variable, "match": ref_h.match, "softIoU": ref_h.softIoU}
    exec(compile(rr._run_iter_source(), "<train.py:runIter>", "exec"), ns)
    orig = torch.masked_select
    torch.masked_select = lambda inp, mask, **kw: orig(inp, mask.bool() if mask.dtype == torch.uint8 else mask, **kw)
    try:
        losses, outs, perms = ns["runIter"](args, enc, dec, Variable(x), Variable(y_mask), Variable(y_class),
                                            Variable(sw_mask), Variable(sw_class), crits, optims, mode="train")
    finally:
        torch.masked_select = orig
    assert enc.training and dec.training
    assert np.abs(np.array([float(v) for v in losses]) - g["losses"]).max() <= 2e-5
    assert (perms[1].numpy() == g["perm_class"]).all()
    assert int(enc.base.bn1.num_batches_tracked) == 1           # train-mode BatchNorm updated its running statistics
    grads = {"enc." + n: p.grad for n, p in enc.named_parameters() if p.grad is not None}
    grads.update({"dec." + n: p.grad for n, p in dec.named_parameters() if p.grad is not None})
    names = [str(n) for n in g["names"]]
    assert sorted(grads) == sorted(names)
    worst = 0.0
    for n, dig in zip(names, g["digests"]):
        if abs(dig[0]) < 1e-7:   # a bias in front of a train-mode BatchNorm: its gradient is rounding noise (~1e-9) in the
            continue             # reference itself (enc.sk*.bias)
        mine = _grad_digest(grads[n])
        scale = max(abs(dig[0]), 1e-12)                          # the tensor's norm
        worst = max(worst, abs(mine[0] - dig[0]) / scale, float(np.abs(mine[2:] - dig[2:]).max()) / scale)
    assert worst <= 5e-3, worst


def test_odd_input_size_resize_matches_the_reference_caller(fake):
    """test.py:39-40 on an odd-sized input: the reference's own loop (its nn.UpsamplingBilinear2d on the module's
    mask logits) and `rsis_b200.test()` (resize of the logits, then the sigmoid) agree."""
    import rsis_b200
    from oracle import synth_weights as sw
    ref = rs.load_reference()
    B, H, W, T = 2, 63, 49, 2
    args = rs.make_args(num_classes=21, maxseqlen=T)
    enc, dec = rsis_b200.FeatureExtractor(args), rsis_b200.RSIS(args)
    enc.load_state_dict(sw.encoder_state_dict(1))
    dec.load_state_dict(sw.decoder_state_dict(1))
    x = sw.synthetic_images(3, B, H, W)
    want = ref.test(args, enc, dec, x)
    args.cuda_graph = False
    got = rsis_b200.test(args, enc, dec, x)
    assert tuple(got[0].shape) == (B, T, H, W)
    for a, b in zip(got, want):
        assert rel(a, b) < 2e-5
