"""Optimiser-step host logic on CPU (fake ABI): parameter groups against the reference-generated golden, FusedAdam
against torch.optim.Adam with the reference's duplicated parameter lists."""
import json
import os

import pytest


@pytest.fixture
def fake(monkeypatch):
    import fake_abi
    return fake_abi.install(monkeypatch)


def test_param_groups_match_reference_golden(golden_dir):
    """Same parameters, same order, same multiplicities as the UNMODIFIED utils.py:34-71 + train.py:236-240 produce on the
    UNMODIFIED reference modules (tests/golden/optim_groups.json, oracle/make_golden.py::golden_optim_groups)."""
    import rsis_b200
    from rsis_b200 import optim
    from optim_parity import _args
    g = json.load(open(os.path.join(golden_dir, "optim_groups.json")))
    args = _args()
    enc, dec = rsis_b200.FeatureExtractor(args), rsis_b200.RSIS(args)
    names = {id(p): "enc." + n for n, p in enc.named_parameters()}
    names.update({id(p): "dec." + n for n, p in dec.named_parameters()})
    dec_group, enc_group = optim.reference_param_groups(args, enc, dec)
    assert [names[id(p)] for p in dec_group["params"]] == g["dec_opt_params"]
    assert all(r == 1 for r in dec_group["repeats"])
    got = {names[id(p)]: r for p, r in zip(enc_group["params"], enc_group["repeats"])}
    assert got == g["enc_opt_counts"]
    assert (dec_group["lr"], enc_group["lr"]) == (1e-3, 1e-6)
    assert sorted(set(got.values())) == [1, 3, 4]


def test_fused_adam_matches_torch_adam_with_duplicates(fake):
    from optim_parity import run
    worst, fused = run("cpu", steps=2)
    assert worst <= 2e-6, worst
    assert len(fused.runs) < 40          # a handful of launches, not one per parameter
    assert {r[4] for r in fused.runs} == {1, 3, 4}


def test_fused_adam_checkpoint_and_late_encoder_group(fake):
    """utils.py:86-111 saves / restores enc_opt.pt / dec_opt.pt and train.py:314-316 enables the encoder mid-run:
    state_dict -> load_state_dict round trip continues bit-identically, and add_groups() keeps the decoder moments."""
    import torch
    import rsis_b200
    from rsis_b200 import optim
    from rsis_b200.autograd import GradBucket
    from optim_parity import _args
    args = _args()

    def make(update_encoder):
        torch.manual_seed(0)
        enc, dec = rsis_b200.FeatureExtractor(args), rsis_b200.RSIS(args)
        bucket = GradBucket(list(enc.parameters()) + list(dec.parameters()), flatten_params=True)
        return enc, dec, bucket, optim.FusedAdam(bucket, optim.reference_param_groups(args, enc, dec, update_encoder))

    def step(bucket, opt, seed):
        g = torch.Generator().manual_seed(seed)
        bucket.flat.copy_(torch.randn(bucket.flat.shape, generator=g) * 0.01)
        opt.step()

    enc, dec, bucket, opt = make(False)      # decoder + skip heads only (the reference's first epochs)
    step(bucket, opt, 1)
    step(bucket, opt, 2)
    sd = opt.state_dict()
    assert all(v["step"] == 2 for v in sd["state"].values())
    params_after_2 = bucket.flat_params.clone()
    # (1) resume: a fresh optimiser loaded from the checkpoint takes the same third step
    enc2, dec2, bucket2, opt2 = make(False)
    bucket2.flat_params.copy_(params_after_2)
    opt2.load_state_dict(sd)
    step(bucket, opt, 3)
    step(bucket2, opt2, 3)
    assert torch.equal(bucket.flat_params, bucket2.flat_params)
    # (2) enabling the encoder group later keeps the decoder's moments and step counts
    enc3, dec3, bucket3, opt3 = make(False)
    bucket3.flat_params.copy_(params_after_2)
    opt3.load_state_dict(sd)
    n_before = len(opt3.runs)
    opt3.add_groups(optim.reference_param_groups(args, enc3, dec3, True)[1:])
    assert len(opt3.runs) > n_before and {r[4] for r in opt3.runs} == {1, 3, 4}
    step(bucket3, opt3, 3)
    dec_offs = [bucket3.offsets[id(p)] for p in dec3.parameters()]
    for p, p_ref in zip(dec3.parameters(), dec.parameters()):
        assert torch.equal(p.detach(), p_ref.detach())
    # the checkpoint of a decoder-only optimiser also loads into one that optimises the encoder from the start
    enc4, dec4, bucket4, opt4 = make(True)
    opt4.load_state_dict(sd)
    assert max(r[5] for r in opt4.runs) == 2 and min(r[5] for r in opt4.runs) == 0
    assert dec_offs
