"""Host-side mirror of the reference's module surface: constructors, attribute names and the
state_dict contract -- and that nothing silently falls back to the CPU."""
import os

import pytest
import torch

import rsis_b200
from oracle import ref_shims as rs
from oracle import synth_weights as sw
from rsis_b200.dist import shard_range


@pytest.fixture(scope="module")
def models():
    args = rs.make_args()
    return args, rsis_b200.FeatureExtractor(args), rsis_b200.RSIS(args)


def test_state_dict_contract(models):
    _, enc, dec = models
    esd, dsd = sw.encoder_state_dict(1), sw.decoder_state_dict(1)
    assert list(enc.state_dict().keys()) == list(esd.keys())  # 661 keys, reference order
    assert list(dec.state_dict().keys()) == list(dsd.keys())  # 16 keys
    for k, v in enc.state_dict().items():
        assert v.shape == esd[k].shape and v.dtype == esd[k].dtype, k
    for k, v in dec.state_dict().items():
        assert v.shape == dsd[k].shape, k
    enc.load_state_dict(esd)
    dec.load_state_dict(dsd)


def test_attribute_surface(models):
    args, enc, dec = models
    # utils/utils.py:39-66 walks these
    for name in ("conv1", "bn1", "layer1", "layer2", "layer3", "layer4"):
        assert hasattr(enc.base, name)
    for name in ("sk1", "sk2", "sk3", "sk4", "sk5", "bn1", "bn2", "bn3", "bn4", "bn5"):
        assert len(list(getattr(enc, name).parameters())) == 2
    assert dec.fc_class.weight.size()[1] == 248  # train.py:250
    assert len(dec.clstm_list) == 5
    assert tuple(dec.clstm_list[0].Gates.weight.shape) == (512, 256, 3, 3)
    assert tuple(dec.clstm_list[4].Gates.weight.shape) == (32, 40, 3, 3)
    n_dec = sum(p.numel() for p in dec.parameters())
    n_enc = sum(p.numel() for p in enc.parameters())
    assert n_dec == 2165391 and n_enc == 48467064
    enc.train(False)
    dec.eval()
    dec.zero_grad()


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/modules"), reason="reference tree only in the build container")
def test_same_keys_as_reference_modules(models):
    _, enc, dec = models
    ref = rs.load_reference()
    args = rs.make_args()
    renc, rdec = ref.FeatureExtractor(args), ref.RSIS(args)
    assert list(renc.state_dict().keys()) == list(enc.state_dict().keys())
    assert list(rdec.state_dict().keys()) == list(dec.state_dict().keys())
    assert [n for n, _ in renc.named_parameters()] == [n for n, _ in enc.named_parameters()]


def test_no_cpu_fallback(models):
    args, enc, dec = models
    enc.eval()
    x = torch.zeros(2, 3, 64, 64)
    with pytest.raises(RuntimeError, match="CUDA"):
        enc(x)
    with pytest.raises(RuntimeError, match="CUDA"):
        dec([torch.zeros(2, 128, 2, 2)] * 5, None)
    with pytest.raises(RuntimeError, match="CUDA"):
        rsis_b200.test(args, enc, dec, x)
    with pytest.raises(RuntimeError, match="CUDA"):
        dec.clstm_list[0](torch.zeros(2, 128, 2, 2), None)


def test_unsupported_modes_fail_loudly():
    with pytest.raises(NotImplementedError):
        rsis_b200.RSIS(rs.make_args(skip_mode="sum"))
    with pytest.raises(NotImplementedError):
        rsis_b200.RSIS(rs.make_args(dropout=0.5))
    with pytest.raises(Exception, match="not supported"):
        rsis_b200.FeatureExtractor(rs.make_args(base_model="vgg16"))


def test_shard_range_partitions_batch():
    for n in (0, 1, 7, 8, 64, 256):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                b, e = shard_range(n, r, world)
                assert 0 <= b <= e <= n
                cover += list(range(b, e))
            assert cover == list(range(n))
            sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
