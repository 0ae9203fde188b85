"""GPU parity of the fused Adam step against torch.optim.Adam on CPU with the
reference's duplicated parameter lists, and the interplay with the derived weight-pack caches."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fused_adam_matches_torch_adam_with_duplicates():
    from optim_parity import run
    worst, fused = run("cuda", steps=3)
    assert worst <= 5e-6, worst
    assert {r[4] for r in fused.runs} == {1, 3, 4}


def test_forward_sees_the_updated_parameters():
    """The kernel writes the parameters through the flat buffer (no torch version bump): the packed-weight caches and
    the captured inference graph must still notice (ops.weights_epoch)."""
    import rsis_b200
    from rsis_b200 import optim
    from rsis_b200.autograd import GradBucket
    from oracle import rsis_oracle as O, synth_weights as sw
    from optim_parity import _args
    args = _args()
    enc, dec = rsis_b200.FeatureExtractor(args), rsis_b200.RSIS(args)
    enc.load_state_dict(sw.encoder_state_dict(1))
    dec.load_state_dict(sw.decoder_state_dict(1))
    enc.cuda().eval()
    dec.cuda().eval()
    x = sw.synthetic_images(123, 2, 64, 64)
    before = [t.clone() for t in rsis_b200.test(args, enc, dec, x.cuda())]
    bucket = GradBucket(list(enc.parameters()) + list(dec.parameters()), flatten_params=True)
    same = rsis_b200.test(args, enc, dec, x.cuda())          # flattening moved the storage, not the values
    assert float((same[0] - before[0]).abs().max()) <= 1e-6
    groups = optim.reference_param_groups(args, enc, dec)
    groups[0]["lr"] = 1e-2                                    # a visible decoder update
    fused = optim.FusedAdam(bucket, groups)
    bucket.flat.fill_(1.0)
    fused.step()
    after = rsis_b200.test(args, enc, dec, x.cuda())
    assert float((after[0] - before[0]).abs().max()) > 1e-3
    esd = {k: v.detach().cpu() for k, v in enc.state_dict().items()}
    dsd = {k: v.detach().cpu() for k, v in dec.state_dict().items()}
    want = O.test_loop(esd, dsd, x, 2)
    for a, w in zip(after, want):
        assert float((a.cpu() - w).abs().max() / w.abs().max()) <= 1e-3
