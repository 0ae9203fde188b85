"""GPU parity tests proper: the CUDA path (through the C ABI of include/rsis_b200.h) against the CPU oracle and the
committed reference-generated golden vectors.  Run on the B200 box with `-m gpu`.

Tolerance (BASELINE.json north_star): every output tensor within 1e-3 *relative* fp32 of the reference path, measured
tensor-relative: max|new - ref| / max|ref| <= 1e-3 (SURVEY.md section 8c "parity metric").  The exact-fp32 CUDA-core
kernels are held to a tighter 2e-5 so that a regression in them cannot hide inside the tensor-core budget.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = 1e-3        # north_star tolerance, tensor-relative
TOL_FP32 = 2e-5   # exact-fp32 kernels (reassociation only)


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


@pytest.fixture(scope="module")
def sw():
    from oracle import synth_weights
    return synth_weights


@pytest.fixture(scope="module")
def O():
    from oracle import rsis_oracle
    return rsis_oracle


def _args(**kw):
    from oracle import ref_shims as rs
    a = rs.make_args(**kw)
    a.hidden_size = int(a.hidden_size)
    a.use_gpu = True
    return a


def _models(R, sw, num_classes=21, T=10, seed=1):
    args = _args(num_classes=num_classes, maxseqlen=T)
    enc, dec = R.FeatureExtractor(args), R.RSIS(args)
    enc.load_state_dict(sw.encoder_state_dict(seed))
    dec.load_state_dict(sw.decoder_state_dict(seed, num_classes=num_classes))
    return args, enc.cuda().eval(), dec.cuda().eval()


IMPLS = ["simt", "auto"]


@pytest.fixture(params=IMPLS)
def impl(request, monkeypatch, R):
    if request.param != "simt" and not R.ops.has_tcgen05():
        pytest.skip("library built without tcgen05 kernels")
    monkeypatch.setenv("RSIS_B200_IMPL", request.param)
    return request.param


def _tol(impl):
    return TOL_FP32 if impl == "simt" else TOL


# ---------------------------------------------------------------------------------------------------------
# primitives, teacher-forced against torch CPU fp32 (the oracle's arithmetic)
# ---------------------------------------------------------------------------------------------------------
CONV_CASES = [
    # (N, Cin, H, W, Cout, k, stride, pad, bias, bn, relu, residual)
    (2, 3, 64, 64, 64, 7, 2, 3, False, True, True, False),     # stem, vision.py:12-14
    (2, 64, 16, 16, 64, 1, 1, 0, False, True, True, False),    # Bottleneck conv1
    (2, 64, 16, 16, 64, 3, 1, 1, False, True, True, False),    # Bottleneck conv2 s1
    (2, 128, 16, 16, 128, 3, 2, 1, False, True, True, False),  # Bottleneck conv2 s2
    (2, 256, 16, 16, 512, 1, 2, 0, False, True, False, False),  # downsample 1x1 s2
    (2, 64, 16, 16, 256, 1, 1, 0, False, True, True, True),    # conv3 + residual + relu
    (3, 256, 9, 7, 32, 3, 1, 1, True, True, False, False),     # skip head sk2, ragged M
    (2, 64, 20, 12, 16, 3, 1, 1, True, True, False, False),    # skip head sk1 (Cout 16)
    (1, 2048, 4, 4, 128, 3, 1, 1, True, True, False, False),   # sk5, K = 18432
    (1, 40, 5, 5, 24, 3, 1, 1, True, False, False, False),     # odd channel counts (multiples of 4)
    (2, 512, 8, 8, 128, 1, 1, 0, True, True, False, False),    # kernel_size=1 heads (args.kernel_size=1)
    (2, 96, 32, 24, 40, 3, 1, 1, True, True, True, True),      # halo-staged 3x3 (W % 8 == 0, H % 16 == 0), 1.5 chunks
    (8, 256, 16, 16, 256, 3, 1, 1, False, True, True, False),  # layer3 conv2 at batch 8: 16 pixel tiles, narrow BN
    (2, 40, 32, 32, 32, 3, 1, 1, True, False, False, False),   # level-4-like: 40 channels = 3 K steps of one chunk
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_matches_oracle(R, impl, case):
    ops = R.ops
    N, Cin, H, W, Cout, k, s, p, has_bias, has_bn, relu, has_res = case
    g = torch.Generator().manual_seed(hash(case) & 0xFFFF)
    x = torch.rand((N, Cin, H, W), generator=g) * 2 - 1
    w = (torch.rand((Cout, Cin, k, k), generator=g) * 2 - 1) * (3.0 / (Cin * k * k)) ** 0.5
    bias = torch.rand(Cout, generator=g) - 0.5 if has_bias else None
    bn = None
    if has_bn:
        bn = torch.nn.BatchNorm2d(Cout)
        with torch.no_grad():
            bn.weight.copy_(torch.rand(Cout, generator=g) + 0.5)
            bn.bias.copy_(torch.rand(Cout, generator=g) - 0.5)
            bn.running_mean.copy_(torch.rand(Cout, generator=g) - 0.5)
            bn.running_var.copy_(torch.rand(Cout, generator=g) + 0.5)
        bn.eval()
    ref = F.conv2d(x, w, bias, stride=s, padding=p)
    if bn is not None:
        with torch.no_grad():
            ref = bn(ref)
    res = None
    if has_res:
        res = torch.rand(ref.shape, generator=g) * 2 - 1
        ref = ref + res
    if relu:
        ref = F.relu(ref)
    which = ops.default_impl()
    fmt = ops.activation_format(which)
    if Cin == 3:
        which, in_fmt = ops.IMPL_SIMT, ops.FMT_F32  # the stem always runs on the CUDA cores
    else:
        in_fmt = fmt
    pc = ops.PackedConv(w.cuda(), None if bias is None else bias.cuda(), None if bn is None else bn.cuda(),
                        want_umma=(fmt == ops.FMT_SPLIT_BF16 and Cin != 3))
    xa = ops.act_from_nchw(x.cuda(), in_fmt)
    ra = ops.act_from_nchw(res.cuda(), fmt) if res is not None else None
    y = ops.conv2d([xa], pc, stride=s, pad=p, relu=relu, residual=ra, out_fmt=ops.FMT_F32, impl=which)
    got = y.nchw()
    assert tuple(got.shape) == tuple(ref.shape)
    assert rel(got, ref) < _tol(impl)
    # second output in the other element format carries the same values
    y1, y2 = ops.conv2d([xa], pc, stride=s, pad=p, relu=relu, residual=ra, out_fmt=ops.FMT_F32,
                        out2_fmt=ops.FMT_SPLIT_BF16, impl=which)
    assert torch.equal(y1.t, y.t)
    assert rel(y2.float().permute(0, 3, 1, 2), ref) < max(_tol(impl), 2e-5)


def test_conv2d_concat_sources(R, impl):
    """Inputs concatenated along C without materialising the concat (model.py:153 `cat([hidden, skip])`)."""
    ops = R.ops
    g = torch.Generator().manual_seed(5)
    a = torch.rand((2, 64, 12, 12), generator=g) - 0.5
    b = torch.rand((2, 32, 12, 12), generator=g) - 0.5
    c = torch.rand((2, 16, 12, 12), generator=g) - 0.5
    w = (torch.rand((48, 112, 3, 3), generator=g) - 0.5) * 0.1
    ref = F.conv2d(torch.cat([a, b, c], 1), w, padding=1)
    which = ops.default_impl()
    fmt = ops.activation_format(which)
    pc = ops.PackedConv(w.cuda(), src_channels=[64, 32, 16], want_umma=(fmt == ops.FMT_SPLIT_BF16))
    srcs = [ops.act_from_nchw(t.cuda(), fmt) for t in (a, b, c)]
    y = ops.conv2d(srcs, pc, pad=1, impl=which)
    assert rel(y.nchw(), ref) < _tol(impl)


def test_conv2d_on_channel_slices(R):
    """tcgen05 convolution reading a pitched channel slice of a wider NHWC buffer and writing its result into slices
    of two other buffers -- how the decoder's concatenated cell inputs are filled in place (rsis_tensor.cstride)."""
    ops = R.ops
    if not ops.has_tcgen05():
        pytest.skip("library built without tcgen05 kernels")
    g = torch.Generator().manual_seed(21)
    wide = torch.rand((2, 104, 16, 24), generator=g) * 2 - 1
    w = (torch.rand((32, 64, 3, 3), generator=g) - 0.5) * 0.1
    b = torch.rand(32, generator=g) - 0.5
    ref = F.conv2d(wide[:, 24:88], w, b, padding=1)
    src = ops.act_from_nchw(wide.cuda(), ops.FMT_SPLIT_BF16).slice(24, 64)
    pc = ops.PackedConv(w.cuda(), b.cuda(), None, want_umma=True)
    dst0 = ops.Act.zeros(2, 16, 24, 80, ops.FMT_SPLIT_BF16, "cuda")
    dst1 = ops.Act.zeros(2, 16, 24, 40, ops.FMT_F32, "cuda")
    ops.conv2d([src], pc, pad=1, impl=ops.IMPL_TCGEN05, out=dst0.slice(16, 32), out2=dst1.slice(8, 32))
    got0 = dst0.float().permute(0, 3, 1, 2)
    got1 = dst1.float().permute(0, 3, 1, 2)
    assert rel(got0[:, 16:48], ref) < TOL and rel(got1[:, 8:40], ref) < TOL
    assert float(got0[:, :16].abs().max()) == 0.0 and float(got0[:, 48:].abs().max()) == 0.0  # neighbours untouched
    assert float(got1[:, :8].abs().max()) == 0.0
    # upsample into a slice
    x = torch.rand((2, 16, 8, 12), generator=g)
    up = ops.Act.zeros(2, 16, 24, 40, ops.FMT_SPLIT_BF16, "cuda")
    ops.upsample_bilinear(ops.act_from_nchw(x.cuda(), ops.FMT_F32), 16, 24, out=up.slice(8, 16))
    refu = F.interpolate(x, size=(16, 24), mode="bilinear", align_corners=True)
    gu = up.float().permute(0, 3, 1, 2)
    assert rel(gu[:, 8:24], refu) < 2e-5 and float(gu[:, :8].abs().max()) == 0.0 and float(gu[:, 24:].abs().max()) == 0.0


def test_maxpool_and_upsample(R):
    ops = R.ops
    g = torch.Generator().manual_seed(3)
    x = torch.rand((2, 64, 17, 22), generator=g) * 4 - 2
    xa = ops.act_from_nchw(x.cuda(), ops.FMT_F32)
    y = ops.maxpool3x3s2(xa)
    assert torch.equal(y.nchw().cpu(), F.max_pool2d(x, 3, 2, 1))  # nn.MaxPool2d(3, 2, 1), vision.py:15 -- exact
    for size in [(34, 44), (17, 22), (40, 31), (18, 23)]:
        ref = F.interpolate(x, size=size, mode="bilinear", align_corners=True)  # nn.UpsamplingBilinear2d
        got = ops.upsample_bilinear(xa, size[0], size[1], ops.FMT_F32).nchw()
        assert rel(got, ref) < 1e-6, size
        got2 = ops.upsample_bilinear(xa, size[0], size[1], ops.FMT_SPLIT_BF16).float().permute(0, 3, 1, 2)
        assert rel(got2, ref) < 2e-5, size
    same = ops.upsample_bilinear(xa, 17, 22, ops.FMT_F32).nchw()
    assert torch.equal(same.cpu(), x)  # same-size upsample is an exact identity


@pytest.mark.parametrize("ks", [3, 1])
def test_mask_head(R, ks):
    ops = R.ops
    g = torch.Generator().manual_seed(11)
    x = torch.rand((3, 8, 20, 28), generator=g) * 2 - 1
    w = torch.rand((1, 8, ks, ks), generator=g) - 0.5
    b = torch.rand(1, generator=g)
    ref = F.conv2d(x, w, b, padding=ks // 2)
    xa = ops.act_from_nchw(x.cuda(), ops.FMT_F32)
    logits = torch.empty((3, 1, 20, 28), device="cuda")
    T = 4
    probs = torch.zeros((3, T, 20, 28), device="cuda")
    ops.mask_head(xa, w.cuda(), b.cuda(), logits, probs[:, 2], T * 20 * 28)
    assert rel(logits, ref) < TOL_FP32
    assert rel(probs[:, 2], torch.sigmoid(ref[:, 0])) < TOL_FP32
    assert float(probs[:, 0].abs().max()) == 0.0 and float(probs[:, 3].abs().max()) == 0.0


@pytest.mark.parametrize("ks", [3, 1])
def test_fused_upsample_mask_head(R, ks):
    """model.py:163-167 in one launch: equals UpsamplingBilinear2d + conv_out, and is bit-identical to the two-kernel
    form (same interpolation and accumulation order)."""
    ops = R.ops
    g = torch.Generator().manual_seed(17)
    x = torch.rand((3, 8, 20, 28), generator=g) * 2 - 1
    w = torch.rand((1, 8, ks, ks), generator=g) - 0.5
    b = torch.rand(1, generator=g)
    up = F.interpolate(x, size=(40, 56), mode="bilinear", align_corners=True)
    ref = F.conv2d(up, w, b, padding=ks // 2)
    xa = ops.act_from_nchw(x.cuda(), ops.FMT_F32)
    T = 3
    logits = torch.empty((3, 1, 40, 56), device="cuda")
    probs = torch.zeros((3, T, 40, 56), device="cuda")
    ops.upsample_mask_head(xa, 40, 56, w.cuda(), b.cuda(), logits, probs[:, 1], T * 40 * 56)
    assert rel(logits, ref) < TOL_FP32 and rel(probs[:, 1], torch.sigmoid(ref[:, 0])) < TOL_FP32
    assert float(probs[:, 0].abs().max()) == 0.0 and float(probs[:, 2].abs().max()) == 0.0
    two = torch.empty_like(logits)
    ops.mask_head(ops.upsample_bilinear(xa, 40, 56, ops.FMT_F32), w.cuda(), b.cuda(), two)
    assert torch.equal(two, logits)


def test_class_stop_heads_and_side_keys(R, sw):
    """fc_class + Softmax + fc_stop (model.py:169-182) on max-pooled features delivered as order-preserving keys."""
    ops = R.ops
    g = torch.Generator().manual_seed(13)
    feat = torch.rand((5, 248), generator=g) * 4 - 2
    feat[0, 0], feat[1, 1], feat[2, 2] = 0.0, -0.0, -3.5
    bits = feat.view(torch.int32)
    keys = torch.where(bits < 0, ~bits, bits | torch.tensor(-2 ** 31, dtype=torch.int32))  # common.cuh float_to_key
    dsd = sw.decoder_state_dict(1)
    wc, bc, ws, bs = (dsd[k].cuda() for k in ("fc_class.weight", "fc_class.bias", "fc_stop.weight", "fc_stop.bias"))
    cls = torch.empty((5, 21), device="cuda")
    stop = torch.empty((5, 1), device="cuda")
    stop_p = torch.empty((5, 1), device="cuda")
    fout = torch.empty((5, 248), device="cuda")
    ops.class_stop_heads(keys.cuda(), wc, bc, ws, bs, cls, 21, stop, stop_p, 1, feat_out=fout)
    assert torch.equal(fout.cpu().view(torch.int32), feat.view(torch.int32))  # keys decode bit-exactly (incl. -0.0)
    ref_c = torch.softmax(F.linear(feat, dsd["fc_class.weight"], dsd["fc_class.bias"]), 1)
    ref_s = F.linear(feat, dsd["fc_stop.weight"], dsd["fc_stop.bias"])
    assert rel(cls, ref_c) < TOL_FP32 and rel(stop, ref_s) < TOL_FP32 and rel(stop_p, torch.sigmoid(ref_s)) < TOL_FP32


@pytest.mark.parametrize("ks", [3, 1])
def test_heads_of_all_steps_in_one_launch_equal_the_per_step_calls(R, sw, ks):
    """`upsample_mask_head_steps` / `class_stop_heads_steps` (the T steps of test.py:37-50 in one launch each, step-major
    inputs, [b][t] outputs) are bit-identical to T per-step calls."""
    ops = R.ops
    g = torch.Generator().manual_seed(23)
    T, B, C, h, w_ = 4, 3, 8, 12, 20
    x = (torch.rand((T * B, C, h, w_), generator=g) * 2 - 1).cuda()
    wt = (torch.rand((1, C, ks, ks), generator=g) - 0.5).cuda()
    b = torch.rand(1, generator=g).cuda()
    xa = ops.act_from_nchw(x, ops.FMT_F32)
    H, W = 2 * h, 2 * w_
    per_step = torch.zeros((B, T, H, W), device="cuda")
    for t in range(T):
        xt = ops.act_from_nchw(x[t * B:(t + 1) * B].contiguous(), ops.FMT_F32)
        ops.upsample_mask_head(xt, H, W, wt, b, None, per_step[:, t], T * H * W)
    batched = torch.zeros((B, T, H, W), device="cuda")
    ops.upsample_mask_head_steps(xa, T, H, W, wt, b, batched, T * H * W, H * W)
    assert torch.equal(batched, per_step)

    feat = torch.rand((T, B, 248), generator=g) * 4 - 2
    bits = feat.view(torch.int32)
    keys = torch.where(bits < 0, ~bits, bits | torch.tensor(-2 ** 31, dtype=torch.int32)).cuda()
    dsd = sw.decoder_state_dict(1)
    wc, bc, ws, bs = (dsd[k].cuda() for k in ("fc_class.weight", "fc_class.bias", "fc_stop.weight", "fc_stop.bias"))
    cls_a, stop_a = torch.zeros((B, T, 21), device="cuda"), torch.zeros((B, T, 1), device="cuda")
    for t in range(T):
        ops.class_stop_heads(keys[t], wc, bc, ws, bs, cls_a[:, t], T * 21, None, stop_a[:, t], T)
    cls_b, stop_b = torch.zeros_like(cls_a), torch.zeros_like(stop_a)
    ops.class_stop_heads_steps(keys, wc, bc, ws, bs, cls_b, T * 21, 21, stop_b, T, 1)
    assert torch.equal(cls_a, cls_b) and torch.equal(stop_a, stop_b)


# ---------------------------------------------------------------------------------------------------------
# ConvLSTM cell: the reference's own outputs (tests/golden/cells_teacher_forced.npz, made by oracle/make_golden.py)
# ---------------------------------------------------------------------------------------------------------
def test_convlstm_cell_matches_reference_golden(R, sw, impl, golden_dir):
    g = np.load(os.path.join(golden_dir, "cells_teacher_forced.npz"))
    dsd = sw.decoder_state_dict(1)
    outs = sw.skip_dims_out(128)
    for lvl, ch in enumerate(outs):
        cin = 128 if lvl == 0 else 2 * outs[lvl - 1]
        cell = R.ConvLSTMCell(_args(), cin, ch, 3, 1)
        cell.load_state_dict({"Gates.weight": dsd[f"clstm_list.{lvl}.Gates.weight"],
                              "Gates.bias": dsd[f"clstm_list.{lvl}.Gates.bias"]})
        cell.cuda()
        x0 = sw._uniform(7, f"cell{lvl}.x0", (2, cin, 8, 8), -2.0, 2.0).cuda()
        x1 = sw._uniform(7, f"cell{lvl}.x1", (2, cin, 8, 8), -2.0, 2.0).cuda()
        h0, c0 = cell(x0, None)                      # clstm.py:26-37 zero state
        assert tuple(h0.shape) == (2, ch, 8, 8)
        assert rel(h0, g[f"l{lvl}_h0"]) < _tol(impl) and rel(c0, g[f"l{lvl}_c0"]) < _tol(impl), lvl
        # teacher-forced second step: feed the REFERENCE's state so errors do not compound
        hr = torch.from_numpy(g[f"l{lvl}_h0"]).cuda()
        cr = torch.from_numpy(g[f"l{lvl}_c0"]).cuda()
        h1, c1 = cell(x1, (hr, cr))
        assert rel(h1, g[f"l{lvl}_h1"]) < _tol(impl) and rel(c1, g[f"l{lvl}_c1"]) < _tol(impl), lvl
        # explicit zero state == None state (the reference materialises zeros)
        z = torch.zeros_like(h0)
        h0z, c0z = cell(x0, (z, z.clone()))
        assert rel(h0z, h0) < 1e-6 and rel(c0z, c0) < 1e-6


@pytest.mark.parametrize("shape", [(1, 128, 128, 3, 5), (3, 64, 16, 17, 9), (2, 32, 8, 33, 31), (2, 512, 32, 6, 6)])
def test_convlstm_cell_matches_oracle_ragged(R, O, impl, shape):
    """Ragged spatial sizes / odd batch; global max side feature (model.py:143) fused in the epilogue."""
    ops = R.ops
    B, cin, ch, H, W = shape
    g = torch.Generator().manual_seed(B * 1000 + cin)
    x = torch.rand((B, cin, H, W), generator=g) * 4 - 2
    hp = torch.rand((B, ch, H, W), generator=g) * 2 - 1
    cp = torch.rand((B, ch, H, W), generator=g) * 4 - 2
    a = 2.0 / (9 * (cin + ch)) ** 0.5
    w = (torch.rand((4 * ch, cin + ch, 3, 3), generator=g) * 2 - 1) * a
    b = torch.rand(4 * ch, generator=g) - 0.5
    href, cref = O.convlstm_cell(w, b, x, (hp, cp))
    cell = R.ConvLSTMCell(_args(), cin, ch, 3, 1)
    cell.load_state_dict({"Gates.weight": w, "Gates.bias": b})
    cell.cuda()
    which = ops.default_impl()
    fmt = ops.activation_format(which)
    side = torch.zeros((B, ch + 7), dtype=torch.int32, device="cuda")
    h, c, _ = cell.step_act([ops.act_from_nchw(x.cuda(), fmt)], ops.act_from_nchw(hp.cuda(), fmt),
                            ops.act_from_nchw(cp.cuda(), ops.FMT_F32).t, side, 3, which)
    assert rel(h.nchw(), href) < _tol(impl) and rel(c.nchw(), cref) < _tol(impl)
    # decode the keys on the host: must equal the max of the h the kernel itself wrote, bit for bit
    k = side[:, 3:3 + ch].cpu()
    dec = torch.where(k < 0, k & 0x7FFFFFFF, ~k).view(torch.float32)
    assert torch.equal(dec, h.nchw().amax(dim=(2, 3)).cpu())
    assert int(side[:, :3].abs().sum()) == 0 and int(side[:, 3 + ch:].abs().sum()) == 0


@pytest.mark.parametrize("shape", [(3, 32, 32, 16, 16, 24),    # pixels-as-M kernel (map height not a multiple of 32)
                                   (2, 16, 16, 8, 32, 16),     # level-4-like: 32 gate columns -> swapped-operand kernel
                                   (2, 32, 32, 16, 64, 24),    # level-3-like: 64 gate columns -> swapped-operand kernel
                                   (5, 16, 16, 8, 96, 40)])    # several tiles per CTA, resident weights
def test_convlstm_cell_hoisted_gates(R, O, shape):
    """The decoder's fast path computes the time-invariant skip share of the gates once (rsis_conv2d with the skip
    columns of Gates.weight, gate-interleaved, + bias) and feeds it to the cell kernel as `gate_preact`; the step then
    contracts over [up(h_below) | prev_hidden] only.  Must equal the reference cell on cat([up, skip]) (model.py:153)."""
    ops = R.ops
    if not ops.has_tcgen05():
        pytest.skip("library built without tcgen05 kernels")
    g = torch.Generator().manual_seed(77)
    B, up_c, skip_c, ch, H, W = shape
    up = torch.rand((B, up_c, H, W), generator=g) * 2 - 1
    skip = torch.rand((B, skip_c, H, W), generator=g) * 4 - 2
    hp = torch.rand((B, ch, H, W), generator=g) * 2 - 1
    cp = torch.rand((B, ch, H, W), generator=g) * 4 - 2
    w = (torch.rand((4 * ch, up_c + skip_c + ch, 3, 3), generator=g) * 2 - 1) * 0.08
    b = torch.rand(4 * ch, generator=g) - 0.5
    href, cref = O.convlstm_cell(w, b, torch.cat([up, skip], 1), (hp, cp))
    cell = R.ConvLSTMCell(_args(), up_c + skip_c, ch, 3, 1)
    cell.load_state_dict({"Gates.weight": w, "Gates.bias": b})
    cell.cuda()
    pc_skip, pc_step = cell.packed_hoisted(up_c, skip_c)
    F16 = ops.FMT_SPLIT_BF16
    pre = ops.conv2d([ops.act_from_nchw(skip.cuda(), F16)], pc_skip, pad=1, impl=ops.IMPL_TCGEN05)
    x = ops.Act.zeros(B, H, W, up_c + ch, F16, "cuda")
    ops.convert(ops.act_from_nchw(up.cuda(), F16), F16, out=x.slice(0, up_c))
    ops.convert(ops.act_from_nchw(hp.cuda(), F16), F16, out=x.slice(up_c, ch))
    side = torch.zeros((B, ch), dtype=torch.int32, device="cuda")
    h16 = ops.Act.zeros(B, H, W, ch + 8, F16, "cuda")
    h, c = ops.convlstm_cell_x(x, pc_step, ops.act_from_nchw(cp.cuda(), ops.FMT_F32).t, side, 0,
                               h16_out=h16.slice(8, ch), impl=ops.IMPL_TCGEN05, gate_preact=pre)
    assert rel(h.nchw(), href) < TOL_FP32 * 5 and rel(c.nchw(), cref) < TOL_FP32 * 5
    # the operand-format copy of h (next step's input slice) and the global max-pool side feature
    assert rel(h16.float()[..., 8:].permute(0, 3, 1, 2), href) < 2e-5 and float(h16.float()[..., :8].abs().max()) == 0.0
    k = side.cpu()
    dec = torch.where(k < 0, k & 0x7FFFFFFF, ~k).view(torch.float32)
    assert torch.equal(dec, h.nchw().amax(dim=(2, 3)).cpu())


# ---------------------------------------------------------------------------------------------------------
# module surface: encoder features, decoder step, test() end to end -- against reference-generated goldens
# ---------------------------------------------------------------------------------------------------------
def _golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    meta = [int(v) for v in g["meta"]]
    return g, meta


@pytest.mark.parametrize("name", ["e2e_b2_64x64_t3", "e2e_b2_96x160_t4_c9", "cfg2_b8_256x256_t10"])
def test_e2e_matches_reference_golden(R, sw, impl, golden_dir, name):
    g, (wseed, iseed, B, H, W, T, ncls, stride) = _golden(golden_dir, name)
    args, enc, dec = _models(R, sw, ncls, T, wseed)
    x = sw.synthetic_images(iseed, B, H, W).cuda()
    tol = _tol(impl)
    with torch.no_grad():
        feats = enc(x)
    full = name == "e2e_b2_64x64_t3"
    for i, f in enumerate(feats):
        assert f.shape[0] == B and f.dim() == 4
        got = f if full else f[:, ::4, ::2, ::2]
        assert rel(got, g[f"feat{i}"]) < tol, f"feat{i}"
    for graph in (True, False):
        args.cuda_graph = graph
        masks, classes, stops = R.test(args, enc, dec, x)
        assert tuple(masks.shape) == (B, T, H, W) and tuple(classes.shape) == (B, T, ncls)
        assert tuple(stops.shape) == (B, T, 1)
        assert rel(masks[:, :, ::stride, ::stride], g["masks"]) < tol
        assert rel(classes, g["classes"]) < tol
        assert rel(stops, g["stops"]) < tol
        if graph:
            again = R.test(args, enc, dec, x)  # replay of the captured graph
            assert torch.equal(again[0], masks) and torch.equal(again[1], classes) and torch.equal(again[2], stops)


def test_cfg1_single_image_through_modules(R, sw, impl, golden_dir):
    """BASELINE.json configs[0]: 1 image 256x256, T=5.  The reference's test() cannot run B=1, so the
    golden was made by calling encoder/decoder directly; same here, including the squeezed B=1 return shapes."""
    g, (wseed, iseed, B, H, W, T, ncls, stride) = _golden(golden_dir, "cfg1_b1_256x256_t5")
    args, enc, dec = _models(R, sw, ncls, T, wseed)
    x = sw.synthetic_images(iseed, B, H, W).cuda()
    with torch.no_grad():
        feats = enc(x)
        hidden = None
        ms, cs, ss = [], [], []
        for _ in range(T):
            m, c, s, hidden = dec(feats, hidden)
            assert tuple(m.shape) == (1, 1, H, W) and tuple(c.shape) == (ncls,) and tuple(s.shape) == (1,)
            assert len(hidden) == 5 and all(len(hc) == 2 for hc in hidden)
            ms.append(m)
            cs.append(c.view(1, -1))
            ss.append(s.view(1, -1))
    masks = torch.sigmoid(torch.cat(ms, 1))
    tol = _tol(impl)
    assert rel(masks[:, :, ::stride, ::stride], g["masks"]) < tol
    assert rel(torch.stack(cs, 1), g["classes"]) < tol
    assert rel(torch.sigmoid(torch.stack(ss, 1)), g["stops"]) < tol
    with pytest.raises(RuntimeError, match="batch size 1"):
        R.test(args, enc, dec, x)


def test_decoder_step_teacher_forced(R, O, sw, impl):
    """RSIS.forward on the ORACLE's features and hidden state (NCHW-contiguous CPU-made tensors moved to the GPU)."""
    args, enc, dec = _models(R, sw)
    esd, dsd = sw.encoder_state_dict(1), sw.decoder_state_dict(1)
    x = sw.synthetic_images(9, 2, 64, 96)
    with torch.no_grad():
        feats = O.feature_extractor(esd, x)
        m0, c0, s0, hid0 = O.rsis_step(dsd, feats, None)
        m1, c1, s1, hid1 = O.rsis_step(dsd, feats, hid0)
        gf = [f.cuda() for f in feats]
        gm0, gc0, gs0, ghid0 = dec(gf, None)
        gm1, gc1, gs1, ghid1 = dec(gf, [[h.cuda(), c.cuda()] for h, c in hid0])
    tol = _tol(impl)
    for got, ref in ((gm0, m0), (gc0, c0), (gs0, s0), (gm1, m1), (gc1, c1), (gs1, s1)):
        assert rel(got, ref) < tol
    for lvl in range(5):
        for j in range(2):
            assert tuple(ghid1[lvl][j].shape) == tuple(hid1[lvl][j].shape)
            assert rel(ghid0[lvl][j], hid0[lvl][j]) < tol and rel(ghid1[lvl][j], hid1[lvl][j]) < tol


def test_encoder_raw_taps(R, O, sw, impl):
    """`forward(x, raw=True)` returns the backbone taps (model.py:65-68); `base(x)` is vision.py:11-21."""
    args, enc, dec = _models(R, sw)
    x = sw.synthetic_images(4, 2, 64, 64)
    with torch.no_grad():
        ref = O.feature_extractor(sw.encoder_state_dict(1), x, raw=True)
        got = enc(x.cuda(), raw=True)
        got2 = enc.base(x.cuda())
    assert [tuple(t.shape) for t in got] == [tuple(t.shape) for t in ref]
    for a, b, r in zip(got, got2, ref):
        assert rel(a, r) < _tol(impl) and torch.equal(a, b)


# ---------------------------------------------------------------------------------------------------------
# training-mode forward (train.py:71-77): BatchNorm2d with batch statistics
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("momentum", [0.1, None])
def test_bn_train_stats_and_affine(R, momentum):
    ops = R.ops
    g = torch.Generator().manual_seed(31)
    x = torch.rand((3, 64, 9, 13), generator=g) * 4 - 1.5
    res = torch.rand((3, 64, 9, 13), generator=g) - 0.5
    bn = torch.nn.BatchNorm2d(64, momentum=momentum)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(64, generator=g) + 0.5)
        bn.bias.copy_(torch.rand(64, generator=g) - 0.5)
        bn.running_mean.copy_(torch.rand(64, generator=g) - 0.5)
        bn.running_var.copy_(torch.rand(64, generator=g) + 0.5)
    import copy
    ref_bn = copy.deepcopy(bn).train()
    with torch.no_grad():
        ref = F.relu(ref_bn(x) + res)
        ref_bn(x * 0.5 + 1.0)  # a second batch: running statistics and num_batches_tracked advance again
    bn = bn.cuda().train()
    xa = ops.act_from_nchw(x.cuda(), ops.FMT_F32)
    scale, shift = ops.bn_train_stats(xa, bn)
    y, y2 = ops.affine_act(xa, scale, shift, residual=ops.act_from_nchw(res.cuda(), ops.FMT_SPLIT_BF16), relu=True,
                           out_fmt=ops.FMT_F32, out2_fmt=ops.FMT_SPLIT_BF16)
    assert rel(y.nchw(), ref) < TOL_FP32 and rel(y2.float().permute(0, 3, 1, 2), ref) < TOL_FP32
    ops.bn_train_stats(ops.act_from_nchw((x * 0.5 + 1.0).cuda(), ops.FMT_F32), bn)
    assert rel(bn.running_mean, ref_bn.running_mean) < 1e-5 and rel(bn.running_var, ref_bn.running_var) < 1e-5
    assert int(bn.num_batches_tracked) == int(ref_bn.num_batches_tracked) == 2


def test_encoder_train_mode_forward(R, O, sw, impl):
    """encoder.train(): every BatchNorm2d normalises with the statistics of the batch (oracle: F.batch_norm(training))
    and advances its running statistics; the decoder has no BatchNorm (dropout 0), so runIter's forward
    (train.py:71-115) is this + the same decoder steps.  Backward is not built."""
    args, enc, dec = _models(R, sw)
    x = sw.synthetic_images(7, 4, 64, 96)
    esd = sw.encoder_state_dict(1)
    with torch.no_grad():
        ref = O.feature_extractor(esd, x, bn=O._bn_train)
        conv1 = F.conv2d(x, esd["base.conv1.weight"], stride=2, padding=3)
    enc.train()
    dec.train()
    rm0 = enc.base.bn1.running_mean.clone()
    with torch.no_grad():
        feats = enc(x.cuda())
        m, c, s, hidden = dec(feats, None)
    tol = _tol(impl) * (5 if impl == "simt" else 1)  # batch statistics amplify fp32 reassociation noise a little
    for f, r in zip(feats, ref):
        assert rel(f, r) < tol
    rm = 0.9 * rm0.cpu() + 0.1 * conv1.mean(dim=(0, 2, 3))
    assert rel(enc.base.bn1.running_mean, rm) < 1e-4 and int(enc.base.bn1.num_batches_tracked) == 1
    with torch.no_grad():
        rm_, rc_, rs_, _ = O.rsis_step(sw.decoder_state_dict(1), ref, None)
    assert rel(m, rm_) < _tol(impl) * 5 and rel(c, rc_) < _tol(impl) * 5
    enc.eval()
    dec.eval()


# ---------------------------------------------------------------------------------------------------------
# size-independent properties at BASELINE.json's full sizes
# ---------------------------------------------------------------------------------------------------------
def test_batch_sharding_is_exact_cfg2(R, sw, impl):
    """The path shards per image: a run is deterministic bit for bit (the only atomics are
    order-independent maxima; split-K partials are summed in a fixed order), and running a batch of 8 in one piece
    or as two shards of 4 gives the same images.  The CUDA-core family is bit-identical across shardings; the
    tcgen05 family picks its tile width / K split from the batch size, which reorders fp32 sums, so there the
    shards agree to fp32 rounding (5e-5 tensor-relative, 20x below the parity tolerance)."""
    args, enc, dec = _models(R, sw, 21, 10)
    args.cuda_graph = False
    x = sw.synthetic_images(123, 8, 256, 256).cuda()
    full = R.test(args, enc, dec, x)
    again = R.test(args, enc, dec, x)
    for a, b in zip(full, again):
        assert torch.equal(a, b)
    parts = [R.test(args, enc, dec, x[i:i + 4].contiguous()) for i in (0, 4)]
    for k in range(3):
        joined = torch.cat([parts[0][k], parts[1][k]], 0)
        if impl == "simt":
            assert torch.equal(full[k], joined)
        else:
            assert rel(joined, full[k]) < 5e-5
    masks, classes, stops = full
    assert float(masks.min()) >= 0.0 and float(masks.max()) <= 1.0
    assert float((classes.sum(-1) - 1).abs().max()) < 1e-5  # Softmax rows
    assert bool(torch.isfinite(masks).all()) and bool(torch.isfinite(classes).all())


def test_cityscapes_shape_cfg3_smoke(R, O, sw, impl):
    """BASELINE.json configs[2] geometry (512x1024, T=20, 9 classes) at batch 1 of the 4: shapes, finiteness, and the
    first decoder step against the oracle (the full T=20 CPU run is too slow for a unit test)."""
    args, enc, dec = _models(R, sw, 9, 20)
    x = sw.synthetic_images(123, 2, 512, 1024)
    masks, classes, stops = R.test(args, enc, dec, x.cuda())
    assert tuple(masks.shape) == (2, 20, 512, 1024) and tuple(classes.shape) == (2, 20, 9)
    assert bool(torch.isfinite(masks).all())
    with torch.no_grad():
        feats = O.feature_extractor(sw.encoder_state_dict(1), x)
        m0, c0, s0, _ = O.rsis_step(sw.decoder_state_dict(1, num_classes=9), feats, None)
    assert rel(masks[:, 0], torch.sigmoid(m0[:, 0])) < _tol(impl)
    assert rel(classes[:, 0], c0) < _tol(impl)
    assert rel(stops[:, 0], torch.sigmoid(s0)) < _tol(impl)


@pytest.mark.parametrize("cfg", [("configs[2] Cityscapes", 2, 512, 1024, 20, 9), ("configs[4] shard", 2, 512, 512, 16, 21)])
def test_full_length_passes_at_the_large_baseline_shapes_match_the_oracle(R, O, sw, impl, cfg):
    """BASELINE.json configs[2] (512x1024, T=20, 9 classes) and the configs[4] geometry (512x512, T=16), two images each:
    EVERY step of the pass against the oracle's test() loop (test.py:16-50) -- masks, class probabilities and stop
    probabilities of all T steps, so the recurrence's error growth over the full length is inside the tolerance too."""
    _, B, H, W, T, C = cfg
    args, enc, dec = _models(R, sw, C, T)
    x = sw.synthetic_images(321, B, H, W)
    masks, classes, stops = R.test(args, enc, dec, x.cuda())
    rm, rc, rs_ = O.test_loop(sw.encoder_state_dict(1), sw.decoder_state_dict(1, num_classes=C), x, T)
    assert tuple(masks.shape) == (B, T, H, W) and tuple(classes.shape) == (B, T, C) and tuple(stops.shape) == (B, T, 1)
    tol = _tol(impl)
    assert rel(masks, rm) < tol and rel(classes, rc) < tol and rel(stops, rs_) < tol
    # per step, so that a late step cannot hide behind the tensor-wide maximum
    for t in (0, T // 2, T - 1):
        assert rel(masks[:, t], rm[:, t]) < tol and rel(classes[:, t], rc[:, t]) < tol


@pytest.mark.parametrize("T", [1, 2])
def test_short_sequences_match_the_oracle(R, O, sw, impl, T):
    """maxseqlen = 1 and 2: on the skewed wavefront schedule T = 1 leaves every other wavefront without a cell (found by
    tests/test_schedule_cpu.py); the pass must still equal the oracle's loop (test.py:16-50)."""
    args, enc, dec = _models(R, sw, 21, T)
    x = sw.synthetic_images(77, 2, 128, 128)
    masks, classes, stops = R.test(args, enc, dec, x.cuda())
    rm, rc, rs_ = O.test_loop(sw.encoder_state_dict(1), sw.decoder_state_dict(1, num_classes=21), x, T)
    tol = _tol(impl)
    assert tuple(masks.shape) == (2, T, 128, 128)
    assert rel(masks, rm) < tol and rel(classes, rc) < tol and rel(stops, rs_) < tol


def test_decoder_workspace_reused_for_a_shorter_sequence(R, sw):
    """One DecoderWorkspace, first T = 4, then T = 2 (the all-steps heads read a prefix of the per-step hidden-state buffer
    that was sized for the longer sequence): the second pass equals a pass on a fresh workspace bit for bit."""
    from rsis_b200 import inference, ops
    if not ops.has_tcgen05():
        pytest.skip("library built without tcgen05 kernels")
    impl = ops.default_impl()
    args, enc, dec = _models(R, sw, 21, 4)
    x = sw.synthetic_images(55, 2, 128, 128).cuda()

    def run(T, ws):
        m = torch.zeros((2, T, 128, 128), device="cuda")
        c = torch.zeros((2, T, 21), device="cuda")
        s_ = torch.zeros((2, T, 1), device="cuda")
        with torch.no_grad():
            ws = inference.run_eager(enc, dec, x, T, impl, m, c, s_, ws=ws)
        torch.cuda.synchronize()
        return (m, c, s_), ws

    _, ws = run(4, None)
    second, _ = run(2, ws)
    fresh, _ = run(2, None)
    for a, b in zip(second, fresh):
        assert torch.equal(a, b)


# ---------------------------------------------------------------------------------------------------------
# error behaviour across the ABI
# ---------------------------------------------------------------------------------------------------------
def test_errors_are_reported_not_swallowed(R):
    ops = R.ops
    w = torch.zeros((8, 16, 3, 3), device="cuda")
    pc = ops.PackedConv(w)
    x = ops.act_from_nchw(torch.zeros((1, 12, 8, 8), device="cuda"), ops.FMT_F32)  # 12 != 16 input channels
    with pytest.raises(RuntimeError, match="bad argument"):
        ops.conv2d([x], pc, pad=1, impl=ops.IMPL_SIMT)
    cell = R.ConvLSTMCell(_args(), 16, 8, 3, 1).cuda()
    with pytest.raises(RuntimeError, match="input channels"):
        cell(torch.zeros((1, 12, 8, 8), device="cuda"), None)
    from rsis_b200 import _lib
    assert _lib.load().rsis_device_check() == 0


@pytest.mark.parametrize("size", [(63, 97), (64, 95)])
def test_odd_input_sizes_follow_the_reference_resize(R, O, sw, impl, size):
    """test.py:39-40: the mask logits (produced at 2*ceil(H/2) x 2*ceil(W/2)) are resized to the INPUT size with
    nn.UpsamplingBilinear2d before the sigmoid.  Odd H / W against the oracle's test loop."""
    H, W = size
    T, B = 2, 2
    args, enc, dec = _models(R, sw, 21, T, 1)
    x = sw.synthetic_images(11, B, H, W)
    want_m, want_c, want_s = O.test_loop(sw.encoder_state_dict(1), sw.decoder_state_dict(1), x, T)
    for graph in (False, True):
        args.cuda_graph = graph
        masks, classes, stops = R.test(args, enc, dec, x.cuda())
        assert tuple(masks.shape) == (B, T, H, W)
        tol = _tol(impl)
        assert rel(masks, want_m) < tol and rel(classes, want_c) < tol and rel(stops, want_s) < tol
