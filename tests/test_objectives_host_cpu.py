"""Host logic of rsis_b200.objectives / rsis_b200.postprocess on CPU through the fake ABI: argument marshalling
(strides, uint8 masks, padding to groups of four masks), autograd wiring, the reference conventions of match()."""
import numpy as np
import pytest
import torch


@pytest.fixture
def fake(monkeypatch):
    import fake_abi
    f = fake_abi.install(monkeypatch)
    from rsis_b200 import objectives
    monkeypatch.setattr(objectives, "_workspace", lambda dev, b, g: torch.zeros(b * g * 2 + b + 1))
    return f


def test_cost_matrix_loss_and_matching_follow_the_oracle(fake):
    from rsis_b200 import objectives as OBJ
    from oracle import rsis_oracle as O
    from oracle.make_golden import match_inputs, soft_iou_inputs
    logits, y_mask, sw = soft_iou_inputs()
    b, g, hw = y_mask.shape
    scores = torch.full((b, g, 4), 7.0)
    OBJ.soft_iou_cost_matrix(logits.view(b, 1, -1, 4), y_mask.to(torch.uint8), 0.5, out=scores[:, :, 1])
    assert float((scores[:, :, 1] - O.soft_iou_cost_matrix(logits, y_mask, 0.5)).abs().max()) <= 1e-6
    assert float(scores[:, :, 0].min()) == 7.0
    pred = logits.unsqueeze(1).repeat(1, g, 1).view(b * g, hw)
    p1 = pred.clone().requires_grad_(True)
    p2 = pred.clone().requires_grad_(True)
    OBJ.softIoULoss()(y_mask.view(b * g, hw), p1, sw).backward()
    O.soft_iou_loss(y_mask.view(b * g, hw), p2, sw).backward()
    assert float((p1.grad - p2.grad).abs().max()) <= 1e-6 * float(p2.grad.abs().max()) + 1e-12
    t_mask, t_class, overlaps = match_inputs()
    pm, pc, perm = OBJ.match([t_mask, None], [t_class, None], overlaps)
    wm, wc, wperm, _ = O.match(t_mask, t_class, overlaps)
    assert torch.equal(perm.long(), wperm) and torch.equal(pc, wc) and torch.equal(pm, wm)


def test_masked_losses_follow_the_reference_golden(fake, golden_dir):
    """MaskedNLLLoss / MaskedBCELoss host logic (argument marshalling, autograd wiring, the fused mean) through the fake
    ABI against the goldens made by the unmodified utils/objectives.py:6-25."""
    import os
    from rsis_b200 import objectives as OBJ
    from oracle.make_golden import masked_loss_inputs
    g = np.load(os.path.join(golden_dir, "masked_losses.npz"))
    probs, target, sw, sw_class, stop_logits, balance = masked_loss_inputs()
    for tag, bal in (("none", None), ("bal", balance)):
        for fused in (False, True):
            p = probs.clone().requires_grad_(True)
            crit = OBJ.MaskedNLLLoss(balance_weight=bal)
            if fused:
                loss = crit.mean(target, p, sw.view(-1, 1))
            else:
                sel = crit(target, p, sw.view(-1, 1))
                assert np.abs(sel.detach().numpy() - g[f"nll_{tag}_sel"]).max() <= 1e-6
                loss = torch.mean(sel)
            assert abs(float(loss) - float(g[f"nll_{tag}_sel"].mean())) <= 1e-6
            loss.backward()
            assert np.abs(p.grad.numpy() - g[f"nll_{tag}_grad"]).max() <= 2e-6 * np.abs(g[f"nll_{tag}_grad"]).max()
    for tag, bw in (("half", 0.5), ("none", None)):
        for fused in (False, True):
            o = stop_logits.clone().requires_grad_(True)
            crit = OBJ.MaskedBCELoss(balance_weight=bw)
            if fused:
                loss = crit.mean(sw, o.squeeze(), sw_class.view(-1, 1))
            else:
                sel = crit(sw, o.squeeze(), sw_class.view(-1, 1))          # train.py:167
                assert np.abs(sel.detach().numpy() - g[f"bce_{tag}_sel"]).max() <= 1e-6
                loss = torch.mean(sel)
            loss.backward()
            assert np.abs(o.grad.numpy() - g[f"bce_{tag}_grad"]).max() <= 2e-6 * np.abs(g[f"bce_{tag}_grad"]).max()


def test_resize_and_encode_instances(fake):
    from scipy.ndimage import zoom
    from rsis_b200 import postprocess as PP
    from oracle import rle_oracle as R
    gen = torch.Generator().manual_seed(2)
    n, h, w, H, W = 5, 24, 32, 37, 50        # 5 masks: padded to two groups of four inside resize_masks
    probs = torch.nn.functional.interpolate(torch.rand((n, 1, 4, 4), generator=gen), size=(h, w), mode="bilinear",
                                            align_corners=True)[:, 0].contiguous()
    got = PP.resize_masks(probs, H, W)
    want = np.stack([zoom(probs[i].numpy().reshape(h, w, 1), [float(H) / h, float(W) / w, 1], order=1)[:, :, 0]
                     for i in range(n)])
    assert tuple(got.shape) == (n, H, W) and float(np.abs(got.numpy() - want).max()) <= 2e-5
    assert PP.resize_masks(probs, h, w) is probs
    out = PP.encode_instances(probs, 0.5, size=(H, W))
    cnts, areas = R.rle_encode((got.numpy() > 0.5).astype(np.uint8))
    for o, c, a in zip(out, cnts, areas):
        assert o == {"size": [H, W], "counts": PP.rle_to_string(c), "area": int(a)}
