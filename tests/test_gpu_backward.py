"""GPU parity of the training path: every backward primitive teacher-forced against torch CPU autograd of the op it
differentiates, then one whole training step (train-mode encoder + T decoder steps + loss.backward()) against autograd
over the CPU oracle.  Run on the B200 box with `-m gpu`.

Tolerances: the exact-fp32 kernels (CUDA-core family, RSIS_B200_BWD_IMPL=simt) are held to 2e-5 tensor-relative
per primitive; data gradients on the tcgen05 family to 1e-3 (north_star).  The whole-step comparison uses a relative
L2 metric for the encoder (ReLU / max-pool discontinuities, see tests/train_parity.py) and the max-norm metric for
the decoder-only BPTT.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from train_parity import (compare_grads, decoder_grads_through_modules, decoder_grads_through_oracle,
                          grads_through_modules, grads_through_oracle)

pytestmark = pytest.mark.gpu

TOL = 1e-3
TOL_FP32 = 2e-5


def rel(a, b):
    a = torch.as_tensor(a, dtype=torch.float32).cpu()
    b = torch.as_tensor(b, dtype=torch.float32).cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def R():
    import rsis_b200
    from rsis_b200 import _lib
    assert _lib.load().rsis_device_check() == 0, "not a B200 (sm_100) device"
    return rsis_b200


def _act(R, t_nchw, fmt=0):
    """CPU [N,C,H,W] -> device NHWC Act."""
    return R.ops.act_from_nchw(t_nchw.cuda().contiguous(), fmt)


def _nchw(a):
    return a.float().permute(0, 3, 1, 2).cpu()


WGRAD_CASES = [
    # (N, Cin, H, W, Cout, k, stride, pad)
    (2, 3, 32, 32, 64, 7, 2, 3),      # stem
    (2, 64, 16, 16, 64, 1, 1, 0),     # bottleneck conv1
    (2, 64, 12, 20, 64, 3, 1, 1),     # conv2 s1
    (2, 128, 16, 16, 128, 3, 2, 1),   # conv2 s2
    (2, 256, 16, 16, 512, 1, 2, 0),   # downsample 1x1 s2
    (3, 40, 9, 7, 24, 3, 1, 1),       # ragged everything
    (2, 8, 32, 32, 1, 3, 1, 1),       # conv_out (Cout = 1)
    (1, 320, 8, 8, 256, 3, 1, 1),     # ConvLSTM gates level 1
    (2, 2048, 4, 4, 128, 3, 1, 1),    # sk5
]


@pytest.mark.parametrize("case", WGRAD_CASES)
@pytest.mark.parametrize("fmt", [0, 1])
def test_conv_wgrad(R, case, fmt):
    ops = R.ops
    if fmt == 1 and not ops.has_tcgen05():
        pytest.skip("no split format without tcgen05")
    N, Cin, H, W, Cout, k, s, p = case
    g = torch.Generator().manual_seed(hash(case) & 0xFFFF)
    x = torch.randn((N, Cin, H, W), generator=g)
    w = torch.randn((Cout, Cin, k, k), generator=g, requires_grad=True)
    b = torch.zeros(Cout, requires_grad=True)
    y = F.conv2d(x, w, b, stride=s, padding=p)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    xa = _act(R, x, fmt)
    dya = _act(R, dy, fmt if dy.shape[1] % 4 == 0 else 0)
    dw = torch.empty((Cout, Cin, k, k), device="cuda")
    db = torch.empty(Cout, device="cuda")
    ops.conv2d_wgrad(xa, dya, k, k, s, p, dw, db)
    tol = TOL_FP32 if fmt == 0 else 1e-4   # split-bf16 operands carry ~2^-16 relative rounding
    assert rel(dw, w.grad) <= tol
    assert rel(db, b.grad) <= tol
    ops.conv2d_wgrad(xa, dya, k, k, s, p, dw, db, accumulate=True)
    assert rel(dw, 2 * w.grad) <= tol


DGRAD_CASES = [
    # (N, Cin, H, W, Cout, k, stride, pad)
    (2, 64, 16, 16, 256, 1, 1, 0),
    (2, 64, 16, 16, 64, 3, 1, 1),
    (2, 128, 16, 16, 128, 3, 2, 1),
    (2, 256, 16, 16, 512, 1, 2, 0),
    (2, 128, 15, 13, 128, 3, 2, 1),   # odd sizes
    (2, 64, 15, 13, 128, 1, 2, 0),
    (2, 8, 32, 32, 1, 3, 1, 1),       # conv_out
    (1, 320, 8, 8, 256, 3, 1, 1),     # gates
]


@pytest.mark.parametrize("case", DGRAD_CASES)
@pytest.mark.parametrize("bimpl", ["simt", "auto"])
def test_conv_dgrad(R, case, bimpl):
    from rsis_b200 import autograd as ag
    ops = R.ops
    if bimpl == "auto" and not ops.has_tcgen05():
        pytest.skip("library built without tcgen05 kernels")
    N, Cin, H, W, Cout, k, s, p = case
    g = torch.Generator().manual_seed(hash(case) & 0xFFFF)
    x = torch.randn((N, Cin, H, W), generator=g, requires_grad=True)
    w = torch.randn((Cout, Cin, k, k), generator=g) * (1.0 / (Cin * k * k)) ** 0.5
    y = F.conv2d(x, w, stride=s, padding=p)
    dy = torch.randn(y.shape, generator=g)
    res = torch.randn(x.shape, generator=g)
    y.backward(dy)
    impl = ops.IMPL_SIMT if bimpl == "simt" else ops.IMPL_AUTO
    fmt = 0 if (bimpl == "simt" or Cout % 8 != 0) else 1
    wd = w.cuda()
    cache = ag._DgradCache()
    dya = _act(R, dy, fmt)
    use_res = s == 1
    out = ag.conv_dgrad(cache, dya, wd, s, p, H, W, impl, residual=_act(R, res) if use_res else None)
    want = x.grad + (res if use_res else 0)
    assert rel(_nchw(out), want) <= (TOL_FP32 if bimpl == "simt" else TOL)


@pytest.mark.parametrize("shape", [(2, 64, 16, 16), (3, 32, 9, 7), (8, 2048, 2, 2)])
@pytest.mark.parametrize("relu", [True, False])
def test_bn_train_bwd(R, shape, relu):
    ops = R.ops
    N, C, H, W = shape
    g = torch.Generator().manual_seed(C + H)
    x = (torch.randn(shape, generator=g) * 2 + 0.5).requires_grad_(True)
    wt = (torch.rand(C, generator=g) + 0.5).requires_grad_(True)
    bs = (torch.rand(C, generator=g) - 0.5).requires_grad_(True)
    y = F.batch_norm(x, None, None, wt, bs, True, 0.0, 1e-5)
    out = F.relu(y) if relu else y
    dy = torch.randn(shape, generator=g)
    out.backward(dy)
    bn = torch.nn.BatchNorm2d(C).cuda().train()
    with torch.no_grad():
        bn.weight.copy_(wt)
        bn.bias.copy_(bs)
    raw = _act(R, x.detach())
    scale, shift, mean, invstd = ops.bn_train_stats(raw, bn, want_stats=True)
    ya = ops.affine_act(raw, scale, shift, relu=relu)
    assert rel(_nchw(ya), out.detach()) <= TOL_FP32
    for dx_fmt in ([0, 1] if ops.has_tcgen05() else [0]):
        dx, dres, dw, db = ops.bn_train_bwd(raw, ya if relu else None, _act(R, dy), bn.weight, mean, invstd,
                                            dx_fmt=dx_fmt, want_dres=True)
        assert rel(_nchw(dx), x.grad) <= 5e-5
        assert rel(dw, wt.grad) <= 5e-5
        assert rel(db, bs.grad) <= 5e-5
        want_res = dy * (out.detach() > 0) if relu else dy
        assert rel(_nchw(dres), want_res) <= 1e-6
    # accumulation targets (gradient buffers that already hold a value)
    acc_w, acc_b = torch.ones(C, device="cuda"), torch.full((C,), 2.0, device="cuda")
    ops.bn_train_bwd(raw, ya if relu else None, _act(R, dy), bn.weight, mean, invstd, dweight_acc=acc_w, dbias_acc=acc_b)
    assert rel(acc_w, wt.grad + 1) <= 5e-5
    assert rel(acc_b, bs.grad + 2) <= 5e-5


@pytest.mark.parametrize("shape", [(2, 64, 16, 16), (2, 8, 15, 11), (1, 4, 2, 2)])
def test_maxpool_bwd_first_max_ties(R, shape):
    ops = R.ops
    g = torch.Generator().manual_seed(3)
    # quantised values (and a ReLU) create many exact ties: the first maximum in scan order must win
    x = torch.relu(torch.randint(-3, 4, shape, generator=g).float()).requires_grad_(True)
    y = F.max_pool2d(x, 3, 2, 1)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    dx = ops.maxpool3x3s2_bwd(_act(R, x.detach()), _act(R, dy))
    assert rel(_nchw(dx), x.grad) <= 1e-6


@pytest.mark.parametrize("shape", [(2, 8, 4, 4, 8, 8), (2, 16, 5, 7, 10, 14), (1, 4, 3, 3, 7, 5), (2, 8, 1, 1, 2, 2),
                                   (2, 128, 2, 2, 4, 4)])
def test_upsample_bilinear_bwd_is_the_adjoint(R, shape):
    ops = R.ops
    N, C, H, W, Ho, Wo = shape
    g = torch.Generator().manual_seed(5)
    x = torch.randn((N, C, H, W), generator=g, requires_grad=True)
    y = F.interpolate(x, size=(Ho, Wo), mode="bilinear", align_corners=True)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    dx = ops.upsample_bilinear_bwd(_act(R, dy), H, W)
    assert rel(_nchw(dx), x.grad) <= TOL_FP32
    # pitched input: dy as a channel slice of a wider buffer
    wide = torch.randn((N, C + 8, Ho, Wo), generator=g)
    wide[:, 4:4 + C] = dy
    wa = _act(R, wide)
    dx2 = ops.upsample_bilinear_bwd(wa.slice(4, C), H, W)
    assert rel(_nchw(dx2), x.grad) <= TOL_FP32


@pytest.mark.parametrize("shape", [(2, 8, 16, 16), (3, 128, 3, 5), (1, 32, 8, 8)])
@pytest.mark.parametrize("has_state", [True, False])
def test_lstm_gates_fwd_bwd(R, shape, has_state):
    ops = R.ops
    N, Ch, H, W = shape
    g = torch.Generator().manual_seed(Ch)
    pre = torch.randn((N, 4 * Ch, H, W), generator=g, requires_grad=True)
    cp = torch.randn((N, Ch, H, W), generator=g, requires_grad=True)
    i, f, o, gg = pre.chunk(4, 1)
    i, f, o, gg = torch.sigmoid(i), torch.sigmoid(f), torch.sigmoid(o), torch.tanh(gg)
    c = f * (cp if has_state else 0) + i * gg
    h = o * torch.tanh(c)
    dh1, dh2, dc = (torch.randn((N, Ch, H, W), generator=g) for _ in range(3))
    (h * (dh1 + dh2) + c * dc).sum().backward()
    gates = _act(R, pre.detach().clone())
    cpa = _act(R, cp.detach())
    ha, ca = ops.lstm_gates_fwd(gates, cpa.t if has_state else None)
    assert rel(_nchw(ha), h.detach()) <= TOL_FP32
    assert rel(_nchw(ca), c.detach()) <= TOL_FP32
    assert rel(_nchw(gates), torch.cat([i, f, o, gg], 1).detach()) <= TOL_FP32
    for fmt in ([0, 1] if ops.has_tcgen05() else [0]):
        dg, dcp = ops.lstm_gates_bwd(gates, cpa.t if has_state else None, ca.t, _act(R, dh1), _act(R, dh2), _act(R, dc),
                                     dg_fmt=fmt)
        assert rel(_nchw(dg), pre.grad) <= 5e-5
        if has_state:
            assert rel(_nchw(dcp), cp.grad) <= 5e-5
    # missing pieces are zeros
    dg, _ = ops.lstm_gates_bwd(gates, cpa.t if has_state else None, ca.t, _act(R, dh1 + dh2), None, None)
    pre2 = pre.detach().clone().requires_grad_(True)
    i, f, o, gg = pre2.chunk(4, 1)
    c2 = torch.sigmoid(f) * (cp.detach() if has_state else 0) + torch.sigmoid(i) * torch.tanh(gg)
    (torch.sigmoid(o) * torch.tanh(c2) * (dh1 + dh2)).sum().backward()
    assert rel(_nchw(dg), pre2.grad) <= 5e-5


def test_global_maxpool_argmax_and_scatter(R):
    ops = R.ops
    g = torch.Generator().manual_seed(9)
    chans = [128, 64, 8]
    sizes = [(2, 2), (4, 8), (32, 32)]
    N = 3
    F_ = sum(chans)
    packed = torch.zeros((N, F_), dtype=torch.int64, device="cuda")
    dside = torch.randn((N, F_), generator=g)
    off = 0
    hs = []
    for C, (H, W) in zip(chans, sizes):
        h = torch.randint(-4, 5, (N, C, H, W), generator=g).float().requires_grad_(True)  # ties on purpose
        m = F.max_pool2d(h, (H, W))
        m.backward(dside[:, off:off + C].reshape(N, C, 1, 1))
        ops.global_maxpool(_act(R, h.detach()), packed, off)
        hs.append((h, off))
        off += C
    keys, idx = ops.global_maxpool_finish(packed)
    for (h, off), C, (H, W) in zip(hs, chans, sizes):
        dh = R.ops.Act.zeros(N, H, W, C, 0, "cuda")
        ops.global_maxpool_bwd(dside.cuda(), idx, off, dh)
        assert rel(_nchw(dh), h.grad) <= 1e-6
    # the keys decode to the maxima (rsis_class_stop_heads reads them)
    feat = torch.empty((N, F_), device="cuda")
    probs = torch.empty((N, 3), device="cuda")
    stop = torch.empty((N, 1), device="cuda")
    wc, bc, ws, bs = torch.zeros((3, F_), device="cuda"), torch.zeros(3, device="cuda"), \
        torch.zeros((1, F_), device="cuda"), torch.zeros(1, device="cuda")
    ops.class_stop_heads(keys, wc, bc, ws, bs, probs, 3, stop, None, 1, feat_out=feat)
    assert float(feat.abs().max()) <= 4.0


def test_class_stop_heads_bwd(R):
    ops = R.ops
    g = torch.Generator().manual_seed(13)
    N, F_, NC = 5, 248, 21
    feat = torch.randn((N, F_), generator=g, requires_grad=True)
    wc = (torch.randn((NC, F_), generator=g) * 0.1).requires_grad_(True)
    bc = torch.randn(NC, generator=g).requires_grad_(True)
    ws = (torch.randn((1, F_), generator=g) * 0.1).requires_grad_(True)
    bs = torch.randn(1, generator=g).requires_grad_(True)
    p = torch.softmax(F.linear(feat, wc, bc), 1)
    s = F.linear(feat, ws, bs)
    dp, dsv = torch.randn((N, NC), generator=g), torch.randn((N, 1), generator=g)
    ((p * dp).sum() + (s * dsv).sum()).backward()
    out = ops.class_stop_heads_bwd(feat.detach().cuda(), p.detach().cuda().contiguous(), dp.cuda(),
                                   dsv.view(-1).cuda().contiguous(), wc.detach().cuda(), ws.detach().cuda())
    for got, want in zip(out, (feat.grad, wc.grad, bc.grad, ws.grad, bs.grad)):
        assert rel(got, want) <= 5e-5


# ---------------------------------------------------------------------------------------------------------
# whole training step
# ---------------------------------------------------------------------------------------------------------
@pytest.fixture(params=[("simt", "simt"), ("auto", "simt"), ("auto", "auto")])
def families(request, monkeypatch, R):
    fwd, bwd = request.param
    if "auto" in request.param and not R.ops.has_tcgen05():
        pytest.skip("library built without tcgen05 kernels")
    monkeypatch.setenv("RSIS_B200_IMPL", fwd)
    monkeypatch.setenv("RSIS_B200_BWD_IMPL", bwd)
    return request.param


@pytest.mark.parametrize("shape", [(2, 2, 2, 3), (1, 1, 3, 2), (2, 4, 4, 2)])
def test_decoder_bptt_gradients(R, families, shape):
    b, h0, w0, T = shape
    res = decoder_grads_through_modules("cuda", b, h0, w0, T, num_classes=5)
    ref = decoder_grads_through_oracle(b, h0, w0, T, num_classes=5)
    compare_grads(res, ref, tol=2e-4 if families == ("simt", "simt") else TOL, metric="max", verbose=True)


_ref_cache = {}


def test_train_step_gradients_match_oracle_autograd(R, families):
    """Whole step, 4 images 128x128, T=2 (layer4 still has 64 samples per BatchNorm channel: train-mode statistics of
    a handful of samples amplify forward rounding differences into percents of gradient change -- not a kernel
    property).  Relative-L2 per parameter tensor; see tests/train_parity.py for why not max-norm here."""
    import json
    import os
    kw = dict(batch=4, size=128, T=2, num_classes=5)
    res = grads_through_modules(device="cuda", **kw)
    if "ref" not in _ref_cache:
        _ref_cache["ref"] = grads_through_oracle(**kw)
    ref = _ref_cache["ref"]
    tol = 2e-2 if families == ("simt", "simt") else 5e-2
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        worst = compare_grads(res, ref, tol=tol, metric="l2", verbose=True)
    finally:
        rows = []
        for name, gref in ref["grads"].items():
            g = res["grads"].get(name)
            if g is None or gref is None or float(gref.norm()) == 0:
                continue
            rows.append((float((g - gref).norm() / gref.norm()), name))
        rows.sort(reverse=True)
        if os.path.isdir(out):
            with open(os.path.join(out, f"grad_parity_{families[0]}_{families[1]}.json"), "w") as f:
                json.dump({"config": kw, "families": families, "loss": [res["loss"], ref["loss"]],
                           "median_rel_l2": rows[len(rows) // 2][0], "worst": rows[:10]}, f, indent=1)
    print("worst relative L2 over all parameter tensors:", worst)


def test_train_step_cfg4_shard_shape(R, monkeypatch):
    """configs[3] per-rank shard: 8 images 256x256, T=10 -- runs, finite gradients everywhere, running statistics
    advance, and the loss of a second identical step is reproduced (no state leaks between steps)."""
    import rsis_b200
    from oracle import synth_weights as sw
    from train_parity import _args
    args = _args(21, 10)
    enc, dec = rsis_b200.FeatureExtractor(args), rsis_b200.RSIS(args)
    enc.load_state_dict(sw.encoder_state_dict(1))
    dec.load_state_dict(sw.decoder_state_dict(1, num_classes=21))
    enc.cuda().train()
    dec.cuda().train()
    x = sw.synthetic_images(123, 8, 256, 256).cuda()
    losses = []
    for _ in range(2):
        for p in list(enc.parameters()) + list(dec.parameters()):
            p.grad = None
        feats = enc(x)
        hidden = None
        loss = 0
        for _t in range(10):
            m, c, s, hidden = dec(feats, hidden)
            loss = loss + (torch.sigmoid(m) ** 2).mean() + (c ** 2).sum(1).mean() + (s ** 2).mean()
        loss.backward()
        losses.append(float(loss))
    n_with = 0
    for name, p in list(enc.named_parameters()) + list(dec.named_parameters()):
        if name.startswith("base.fc."):
            assert p.grad is None
            continue
        assert p.grad is not None and bool(torch.isfinite(p.grad).all()), name
        n_with += 1
    assert n_with > 300
    assert int(enc.base.bn1.num_batches_tracked) == 2
    assert abs(losses[0] - losses[1]) <= 1e-5 * abs(losses[0])


def test_train_step_cuda_graph_matches_eager(R):
    """rsis_b200.training.TrainStep: the captured step (pack + forward + loss + backward in ONE CUDA graph) reproduces
    the eager step, keeps doing so after an in-place parameter update (the packing kernels are inside the graph),
    and advances the BatchNorm running statistics once per call."""
    import rsis_b200
    from rsis_b200.training import TrainStep
    from oracle import synth_weights as sw
    from train_parity import _args
    args = _args(5, 3)

    def make():
        enc, dec = rsis_b200.FeatureExtractor(args), rsis_b200.RSIS(args)
        enc.load_state_dict(sw.encoder_state_dict(1))
        dec.load_state_dict(sw.decoder_state_dict(1, num_classes=5))
        return enc.cuda().train(), dec.cuda().train()

    def loss_fn(masks, classes, stops):
        return sum((torch.sigmoid(m) ** 2).mean() + (c ** 2).sum(-1).mean() + (s ** 2).mean()
                   for m, c, s in zip(masks, classes, stops))

    x = sw.synthetic_images(123, 4, 128, 128).cuda()
    enc_e, dec_e = make()
    enc_g, dec_g = make()
    eager = TrainStep(enc_e, dec_e, 3, loss_fn, cuda_graph=False, all_reduce=False)
    graph = TrainStep(enc_g, dec_g, 3, loss_fn, cuda_graph=True, all_reduce=False)
    for it in range(3):
        le, lg = float(eager(x)), float(graph(x))
        assert abs(le - lg) <= 1e-5 * abs(le), (it, le, lg)
        ge, gg = eager.bucket.flat, graph.bucket.flat
        assert float((ge - gg).norm() / ge.norm()) <= 1e-3, it
        with torch.no_grad():  # the same in-place "optimiser step" on both
            for pe, pg in zip(eager.bucket.params, graph.bucket.params):
                pe.add_(pe.grad, alpha=-1e-4)
                pg.add_(pe.grad, alpha=-1e-4)
    assert int(enc_g.base.bn1.num_batches_tracked) == int(enc_e.base.bn1.num_batches_tracked) == 3
    assert float((enc_g.base.bn1.running_mean - enc_e.base.bn1.running_mean).abs().max()) <= 1e-5


def test_run_iter_whole_training_iteration(R, families):
    """`runIter` (train.py:56-197) end to end on the GPU -- train-mode encoder, T decoder steps, the fused soft-IoU cost
    matrix per step, Hungarian matching on the device, softIoULoss (fused forward + backward), class / stop losses,
    `loss.backward()` -- against the oracle recipe, which tests/test_oracle_golden.py pins to the reference's own
    runIter.  Losses to 1e-4; matched classes exactly; gradients in the relative-L2 metric of the whole-step test."""
    from run_iter_parity import modules_run_iter, oracle_run_iter
    if "ri" not in _ref_cache:
        _ref_cache["ri"] = oracle_run_iter()
    want_l, want_perm, want_g = _ref_cache["ri"]
    got_l, got_perm, got_g = modules_run_iter("cuda")
    assert max(abs(a - b) for a, b in zip(got_l, want_l)) <= 1e-4, (got_l, want_l)
    assert (got_perm == want_perm).all()
    tol = 2e-2 if families == ("simt", "simt") else 1e-1   # 2 x 64x64: layer4 normalises over 8 samples (see above)
    compare_grads({"loss": got_l[0], "grads": got_g}, {"loss": want_l[0], "grads": want_g}, tol=tol, metric="l2",
                  verbose=True)


def _cfg4_step(R, sw, impl_env, monkeypatch, precision=None, batch=8, size=256, T=10):
    """One training step at the configs[3] per-rank shard shape under a kernel family; returns (loss, grads)."""
    import rsis_b200
    from train_parity import _args
    for k, v in impl_env.items():
        monkeypatch.setenv(k, v)
    args = _args(21, T)
    if precision is not None:
        args.precision = precision
    enc, dec = rsis_b200.FeatureExtractor(args), rsis_b200.RSIS(args)
    enc.load_state_dict(sw.encoder_state_dict(1))
    dec.load_state_dict(sw.decoder_state_dict(1, num_classes=21))
    enc.cuda().train()
    dec.cuda().train()
    x = sw.synthetic_images(123, batch, size, size).cuda()
    feats = enc(x)
    hidden = None
    loss = 0
    for _t in range(T):
        m, c, s, hidden = dec(feats, hidden)
        loss = loss + (torch.sigmoid(m) ** 2).mean() + (c ** 2).sum(1).mean() + (s ** 2).mean()
    loss.backward()
    grads = {n: p.grad.detach().float().cpu() for n, p in list(enc.named_parameters()) + list(dec.named_parameters())
             if p.grad is not None}
    return float(loss), grads


def _rel_l2_table(got, ref):
    rows = []
    for n, g in ref.items():
        # a convolution bias in front of a train-mode BatchNorm has an exactly zero gradient; what the kernels return
        # there is rounding noise of either family (sk1..sk5.bias)
        if float(g.norm()) < 1e-7 or (n.startswith("sk") and n.endswith(".bias")):
            continue
        rows.append((float((got[n] - g).norm() / g.norm()), n))
    rows.sort(reverse=True)
    return rows


def test_cfg4_shard_gradients_tcgen05_vs_exact_fp32_family(R, monkeypatch):
    """GRADIENT parity (not finiteness) at the configs[3] per-rank shard shape, 8 x 256x256, T=10: the tcgen05 family
    (split-bf16 products, forward and backward) against the exact-fp32 CUDA-core family of the same library on the same
    inputs.  Both run the same algorithm; they differ by operand rounding (~2^-16 per product) and summation order, which
    train-mode BatchNorm / ReLU / arg-max discontinuities amplify (tests/train_parity.py; the reference's own fp32 vs
    fp64 autograd differ as much).  Measured on B200: median relative L2 over the parameter tensors 2.1e-2, worst
    3.1e-2; bounds 5e-2 / 1.5e-1; loss to 1e-4."""
    from oracle import synth_weights as sw
    if not R.ops.has_tcgen05():
        pytest.skip("library built without tcgen05 kernels")
    l_ref, g_ref = _cfg4_step(R, sw, {"RSIS_B200_IMPL": "simt", "RSIS_B200_BWD_IMPL": "simt"}, monkeypatch)
    l_tc, g_tc = _cfg4_step(R, sw, {"RSIS_B200_IMPL": "auto", "RSIS_B200_BWD_IMPL": "auto"}, monkeypatch)
    assert abs(l_tc - l_ref) <= 1e-4 * abs(l_ref), (l_tc, l_ref)
    rows = _rel_l2_table(g_tc, g_ref)
    median = rows[len(rows) // 2][0]
    print("cfg4 shard, tcgen05 vs exact fp32: median rel-L2", median, "worst", rows[:5])
    assert len(rows) > 300 and median <= 5e-2 and rows[0][0] <= 1.5e-1, (median, rows[:5])


def test_cfg5_geometry_gradients_tcgen05_vs_exact_fp32_family(R, monkeypatch):
    """The same gradient-parity statement at the configs[4] geometry (512x512, T=16; two images of the 32 a rank holds):
    the full-length training iteration (forward, BPTT through 16 decoder steps, train-mode ResNet-101 backward) on the
    tcgen05 family against the exact-fp32 CUDA-core family.  Same bounds as at the configs[3] shard shape."""
    from oracle import synth_weights as sw
    if not R.ops.has_tcgen05():
        pytest.skip("library built without tcgen05 kernels")
    l_ref, g_ref = _cfg4_step(R, sw, {"RSIS_B200_IMPL": "simt", "RSIS_B200_BWD_IMPL": "simt"}, monkeypatch, batch=2,
                              size=512, T=16)
    l_tc, g_tc = _cfg4_step(R, sw, {"RSIS_B200_IMPL": "auto", "RSIS_B200_BWD_IMPL": "auto"}, monkeypatch, batch=2,
                            size=512, T=16)
    assert abs(l_tc - l_ref) <= 1e-4 * abs(l_ref), (l_tc, l_ref)
    rows = _rel_l2_table(g_tc, g_ref)
    median = rows[len(rows) // 2][0]
    print("cfg5 geometry, tcgen05 vs exact fp32: median rel-L2", median, "worst", rows[:5])
    assert len(rows) > 300 and all(bool(torch.isfinite(g).all()) for g in g_tc.values())
    assert median <= 5e-2 and rows[0][0] <= 1.5e-1, (median, rows[:5])


def test_bf16_training_mode_loss_level_parity(R, monkeypatch):
    """BASELINE.json configs[3] "training step bf16": `args.precision = "bf16"` runs the training-mode forward and
    backward with single-pass bf16 tensor-core products (fp32 accumulation, fp32 master weights and gradients).
    LOSS-LEVEL tolerance, stated here: loss within 5e-2 relative of the fp32-grade step (measured: 2.2e-2
    through 104 train-mode convolutions + BatchNorms and a 10-step recurrence), every gradient tensor finite, and the
    gradient direction preserved -- cosine similarity of the concatenated gradient >= 0.85 (measured 0.90) against the
    split-bf16 (fp32-grade) step at the configs[3] shard shape.  The ARITHMETIC of the mode is pinned exactly by
    test_bf16_mode_primitives_equal_bf16_rounded_operands."""
    from oracle import synth_weights as sw
    if not R.ops.has_tcgen05():
        pytest.skip("library built without tcgen05 kernels")
    env = {"RSIS_B200_IMPL": "auto", "RSIS_B200_BWD_IMPL": "auto"}
    l_ref, g_ref = _cfg4_step(R, sw, env, monkeypatch)
    l_bf, g_bf = _cfg4_step(R, sw, env, monkeypatch, precision="bf16")
    assert R.ops.get_precision() == "fp32"          # the mode does not leak out of the modules' nodes
    assert abs(l_bf - l_ref) <= 5e-2 * abs(l_ref), (l_bf, l_ref)
    assert l_bf != l_ref                             # ... and it really was a different arithmetic
    names = sorted(g_ref)
    a = torch.cat([g_bf[n].reshape(-1) for n in names]).double()
    b = torch.cat([g_ref[n].reshape(-1) for n in names]).double()
    assert bool(torch.isfinite(a).all())
    cos = float((a * b).sum() / (a.norm() * b.norm()))
    rows = _rel_l2_table(g_bf, g_ref)
    median = rows[len(rows) // 2][0]
    print("bf16 training mode vs fp32-grade: loss", l_bf, l_ref, "cosine", cos, "median rel-L2", median, rows[:3])
    assert cos >= 0.85, (cos, median)


def test_bf16_run_iter_losses_against_reference_golden(R, golden_dir):
    """The whole training iteration (runIter recipe: soft-IoU costs, Hungarian matching, the three criteria) in bf16
    mode against the golden of the reference's own runIter: identical matching, losses within 1e-2 absolute."""
    import os
    from run_iter_parity import modules_run_iter
    if not R.ops.has_tcgen05():
        pytest.skip("library built without tcgen05 kernels")
    g = np.load(os.path.join(golden_dir, "run_iter.npz"))
    got_l, got_perm, got_g = modules_run_iter("cuda", precision="bf16")
    assert (got_perm.numpy() == g["perm_class"]).all()
    assert np.abs(np.array(got_l) - g["losses"]).max() <= 1e-2, (got_l, g["losses"])
    assert all(bool(torch.isfinite(v).all()) for v in got_g.values())


def test_eval_after_train_sees_updated_batchnorm_statistics(R, monkeypatch):
    """a train-mode step advances running_mean / running_var through raw pointers; the eval-side packs
    (BatchNorm folded) and captured inference graphs must be rebuilt.  Eval features after a train step must equal the
    oracle's eval forward with the UPDATED statistics, and differ from the pre-step ones."""
    import rsis_b200
    from oracle import rsis_oracle as O, synth_weights as sw
    from train_parity import _args
    args = _args(21, 2)
    enc, dec = rsis_b200.FeatureExtractor(args), rsis_b200.RSIS(args)
    enc.load_state_dict(sw.encoder_state_dict(1))
    dec.load_state_dict(sw.decoder_state_dict(1, num_classes=21))
    enc.cuda()
    dec.cuda()
    x = sw.synthetic_images(7, 2, 64, 64).cuda()
    xt = sw.synthetic_images(8, 2, 64, 64).cuda() * 1.7 + 0.3
    enc.eval()
    dec.eval()
    with torch.no_grad():
        before = [f.clone() for f in enc(x)]
        m0, c0, s0 = rsis_b200.test(args, enc, dec, x)          # captures an inference graph with the old statistics
    enc.train()
    feats = enc(xt)                                             # train-mode forward: running statistics move
    sum(f.sum() for f in feats).backward()
    enc.eval()
    with torch.no_grad():
        after = enc(x)
        m1, c1, s1 = rsis_b200.test(args, enc, dec, x)
    esd = {k: v.detach().cpu() for k, v in enc.state_dict().items()}
    want = O.feature_extractor(esd, x.cpu())
    assert int(enc.base.bn1.num_batches_tracked) == 1
    for i, (a, b, w) in enumerate(zip(after, before, want)):
        assert rel(a, w) < 1e-3, f"feat{i} does not use the updated statistics"
        assert rel(a, b) > 1e-3, f"feat{i} unchanged: stale eval pack"
    assert float((m1 - m0).abs().max()) > 1e-6                  # the captured graph was rebuilt, too


@pytest.mark.parametrize("case", [(2, 64, 16, 16, 64, 3, 1, 1), (2, 256, 16, 16, 512, 1, 1, 0), (2, 128, 32, 16, 128, 3, 1, 1),
                                  (1, 320, 16, 16, 256, 3, 1, 1)])
def test_bf16_mode_primitives_equal_bf16_rounded_operands(R, case):
    """RSIS_PRECISION_BF16 is single-pass bf16: the product of the bf16-ROUNDED operands with fp32 accumulation.  That
    arithmetic can be emulated exactly on the CPU (round both operands to bf16, convolve in fp32), so the mode is held
    to the exact-kernel tolerance against the emulation: forward convolution, data gradient and weight gradient."""
    from rsis_b200 import autograd as ag
    ops = R.ops
    if not ops.has_tcgen05():
        pytest.skip("library built without tcgen05 kernels")
    N, Cin, H, W, Cout, k, s, p = case
    g = torch.Generator().manual_seed(31)
    bf = lambda t: t.bfloat16().float()
    x = torch.randn((N, Cin, H, W), generator=g)
    w = torch.randn((Cout, Cin, k, k), generator=g) * (1.0 / (Cin * k * k)) ** 0.5
    xr = bf(x).requires_grad_(True)
    wr = bf(w).requires_grad_(True)
    y = F.conv2d(xr, wr, stride=s, padding=p)
    dy = torch.randn(y.shape, generator=g)
    F32B = ops.FMT_SPLIT_BF16
    pc = ops.PackedConv(w.cuda(), None, None, want_umma=True)
    with ops.precision("bf16"):
        got = ops.conv2d([_act(R, x, F32B)], pc, stride=s, pad=p, out_fmt=ops.FMT_F32, impl=ops.IMPL_TCGEN05)
    assert ops.get_precision() == "fp32"
    assert rel(_nchw(got), y.detach()) <= 2e-5
    # backward: dx = dgrad(bf16(dy), bf16(w)), dw = wgrad(bf16(x), bf16(dy))
    (bf(dy) * F.conv2d(xr, wr.detach(), stride=s, padding=p)).sum().backward()       # dL/dx with dy rounded
    dx_want = xr.grad.clone()
    xr.grad = None
    (bf(dy) * F.conv2d(xr.detach(), wr, stride=s, padding=p)).sum().backward()
    dw_want = wr.grad.clone()
    cache = ag._DgradCache()
    with ops.precision("bf16"):
        dx = ag.conv_dgrad(cache, _act(R, dy, F32B), w.cuda(), s, p, H, W, ops.IMPL_AUTO)
        dw = torch.empty((Cout, Cin, k, k), device="cuda")
        ops.conv2d_wgrad(_act(R, x, F32B), _act(R, dy, F32B), k, k, s, p, dw, None)
    assert rel(_nchw(dx), dx_want) <= 2e-5
    assert rel(dw, dw_want) <= 2e-5
    # and the default mode on the same operands is fp32-grade (NOT the bf16-rounded result)
    full = ops.conv2d([_act(R, x, F32B)], pc, stride=s, pad=p, out_fmt=ops.FMT_F32, impl=ops.IMPL_TCGEN05)
    assert rel(_nchw(full), F.conv2d(x, w, stride=s, padding=p)) <= 1e-4 and rel(_nchw(full), y.detach()) > 1e-4
