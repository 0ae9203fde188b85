"""Host logic of the training path (rsis_b200/autograd.py) on CPU: which tensors are saved, the order of the backward
primitives, gradient slicing / accumulation through autograd -- checked against autograd over the CPU oracle.

The ABI entry points are replaced by tests/fake_abi.py (torch-CPU restatements of the header contracts); the CUDA
kernels themselves are checked on the GPU box by tests/test_gpu_backward.py with the very same comparison.
"""
import types

import pytest
import torch

from train_parity import (compare_grads, decoder_grads_through_modules, decoder_grads_through_oracle,
                          grads_through_modules, grads_through_oracle)


@pytest.fixture
def fake(monkeypatch):
    import fake_abi
    return fake_abi.install(monkeypatch)


def test_train_step_gradients_match_oracle_autograd(fake):
    res = grads_through_modules(device="cpu", batch=2, size=64, T=3, num_classes=5)
    ref = grads_through_oracle(batch=2, size=64, T=3, num_classes=5)
    compare_grads(res, ref, tol=1e-2, metric="l2")


@pytest.mark.parametrize("shape", [(2, 2, 2, 3), (1, 1, 3, 2)])
def test_decoder_bptt_gradients_strict(fake, shape):
    """Decoder only, T steps of BPTT: max-norm parity (no ReLU / max-pool discontinuity on this part of the path)."""
    b, h0, w0, T = shape
    res = decoder_grads_through_modules("cpu", b, h0, w0, T, num_classes=5)
    ref = decoder_grads_through_oracle(b, h0, w0, T, num_classes=5)
    compare_grads(res, ref, tol=2e-4, metric="max")


def test_single_image_squeeze_and_missing_heads(fake):
    """B=1 (the reference's `.squeeze()` shapes) with a loss that ignores the class and stop outputs."""
    res = grads_through_modules(device="cpu", batch=1, size=64, T=2, num_classes=4, use_heads=False)
    ref = grads_through_oracle(batch=1, size=64, T=2, num_classes=4, use_heads=False)
    compare_grads(res, ref, tol=1e-2, metric="l2")
    assert res["grads"]["dec.fc_class.weight"] is None or float(res["grads"]["dec.fc_class.weight"].abs().max()) == 0


def test_grad_bucket_views_receive_the_gradients(fake):
    from rsis_b200.autograd import GradBucket
    res = grads_through_modules(device="cpu", batch=2, size=64, T=2, num_classes=5, bucket=True)
    flat = res["bucket"].flat
    assert float(flat.abs().sum()) > 0
    off = 0
    for p in res["bucket"].params:
        assert p.grad.data_ptr() == flat[off:off + p.numel()].data_ptr()
        off += p.numel()
    ref = grads_through_oracle(batch=2, size=64, T=2, num_classes=5)
    compare_grads(res, ref, tol=1e-2, metric="l2")
    assert isinstance(res["bucket"], GradBucket)


def test_train_step_object_eager_matches_manual_loop(fake):
    """rsis_b200.training.TrainStep (eager mode) = the manual loop of train.py:71-115 + backward, gradients in the
    bucket."""
    import rsis_b200
    from rsis_b200.training import TrainStep
    from oracle import synth_weights as sw
    from train_parity import _args
    args = _args(5, 2)
    enc, dec = rsis_b200.FeatureExtractor(args), rsis_b200.RSIS(args)
    enc.load_state_dict(sw.encoder_state_dict(1))
    dec.load_state_dict(sw.decoder_state_dict(1, num_classes=5))
    enc.train()
    dec.train()
    x = sw.synthetic_images(123, 2, 64, 64)

    def loss_fn(masks, classes, stops):
        return sum((m ** 2).mean() + (c ** 2).sum() + (s ** 2).sum() for m, c, s in zip(masks, classes, stops))

    step = TrainStep(enc, dec, 2, loss_fn, cuda_graph=False, all_reduce=False)
    l1 = float(step(x))
    g1 = step.bucket.flat.clone()
    l2 = float(step(x))   # bucket.zero() at the start of every step: gradients do not pile up
    assert abs(l1 - l2) <= 1e-4 * abs(l1)  # running statistics move, the train-mode outputs do not
    assert float((step.bucket.flat - g1).abs().max()) <= 1e-3 * float(g1.abs().max())
    assert float(g1.abs().sum()) > 0


def test_train_step_with_fused_adam_moves_the_parameters(fake):
    """TrainStep(optimizer=FusedAdam): forward + backward + optimiser step in one call; the next forward uses the
    updated parameters (pack caches follow ops.weights_epoch) and the loss goes down on a fixed batch."""
    import rsis_b200
    from rsis_b200 import optim
    from rsis_b200.autograd import GradBucket
    from rsis_b200.training import TrainStep
    from oracle import synth_weights as sw
    from train_parity import _args
    args = _args(5, 2)
    args.lr, args.lr_cnn, args.weight_decay, args.weight_decay_cnn = 1e-3, 1e-6, 1e-6, 1e-6
    enc, dec = rsis_b200.FeatureExtractor(args), rsis_b200.RSIS(args)
    enc.load_state_dict(sw.encoder_state_dict(1))
    dec.load_state_dict(sw.decoder_state_dict(1, num_classes=5))
    enc.train()
    dec.train()
    x = sw.synthetic_images(123, 2, 64, 64)

    def loss_fn(masks, classes, stops):
        return sum((m ** 2).mean() + (s ** 2).sum() for m, c, s in zip(masks, classes, stops))

    bucket = GradBucket(list(enc.parameters()) + list(dec.parameters()), flatten_params=True)
    fused = optim.FusedAdam(bucket, optim.reference_param_groups(args, enc, dec))
    step = TrainStep(enc, dec, 2, loss_fn, cuda_graph=False, all_reduce=False, optimizer=fused)
    w0 = dec.conv_out.weight.detach().clone()
    losses = [float(step(x)) for _ in range(4)]
    assert float((dec.conv_out.weight.detach() - w0).abs().max()) > 0
    assert losses[-1] < losses[0], losses


def test_run_iter_through_modules_matches_oracle(fake):
    """The whole training iteration of train.py:56-197 through the modules' autograd nodes (host logic, fake ABI)."""
    from run_iter_parity import modules_run_iter, oracle_run_iter
    want_l, want_perm, want_g = oracle_run_iter()
    got_l, got_perm, got_g = modules_run_iter("cpu")
    assert max(abs(a - b) for a, b in zip(got_l, want_l)) <= 1e-5
    assert (got_perm == want_perm).all()
    res = {"loss": got_l[0], "grads": got_g}
    ref = {"loss": want_l[0], "grads": want_g}
    compare_grads(res, ref, tol=1e-2, metric="l2")
