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
