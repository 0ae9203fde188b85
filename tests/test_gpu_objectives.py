"""GPU parity of the fused soft-IoU cost / loss kernels against the CPU oracle's
restatement of utils/hungarian.py:64-90, train.py:96-110 and utils/objectives.py:27-34, and against the
reference-generated golden fixture.  Tolerance: 1e-5 tensor-relative (fp32 reductions in a different order)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = torch.as_tensor(a, dtype=torch.float32).cpu()
    b = torch.as_tensor(b, dtype=torch.float32).cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def OBJ():
    from rsis_b200 import _lib, objectives
    assert _lib.load().rsis_device_check() == 0
    return objectives


def test_cost_matrix_matches_reference_golden(OBJ, golden_dir):
    from oracle.make_golden import soft_iou_inputs
    g = np.load(os.path.join(golden_dir, "soft_iou.npz"))
    logits, y_mask, _ = soft_iou_inputs()
    for gt in (y_mask.cuda(), y_mask.cuda().to(torch.uint8), y_mask.cuda().bool()):
        cost = OBJ.soft_iou_cost_matrix(logits.cuda(), gt, 1.0)
        assert rel(cost, g["cost"]) <= 1e-5


@pytest.mark.parametrize("shape", [(8, 20, 256 * 256), (2, 1, 64), (3, 33, 1028), (1, 40, 4096)])
@pytest.mark.parametrize("u8", [False, True])
def test_cost_matrix_matches_oracle(OBJ, shape, u8):
    from oracle import rsis_oracle as O
    b, g, hw = shape
    gen = torch.Generator().manual_seed(b * 131 + g)
    logits = torch.randn((b, hw), generator=gen) * 2
    y = (torch.rand((b, g, hw), generator=gen) < 0.3).float()
    y[:, -1] = 0
    want = O.soft_iou_cost_matrix(logits, y, 0.7)
    gt = y.cuda().to(torch.uint8) if u8 else y.cuda()
    # written straight into a strided [B, gtT] view of scores[B, gtT, T] (train.py:110)
    scores = torch.full((b, g, 5), 7.0, device="cuda")
    OBJ.soft_iou_cost_matrix(logits.cuda().view(b, 1, -1, 4), gt, 0.7, out=scores[:, :, 2])
    assert rel(scores[:, :, 2], want) <= 1e-5
    assert float(scores[:, :, 1].min()) == 7.0 and float(scores[:, :, 3].max()) == 7.0
    # the workspace is left zeroed: a second call gives the same answer
    again = OBJ.soft_iou_cost_matrix(logits.cuda(), gt, 0.7)
    assert rel(again, want) <= 1e-5


@pytest.mark.parametrize("u8", [False, True])
def test_soft_iou_loss_forward_backward(OBJ, u8, golden_dir):
    from oracle import rsis_oracle as O
    from oracle.make_golden import soft_iou_inputs
    g = np.load(os.path.join(golden_dir, "soft_iou.npz"))
    logits, y_mask, sw = soft_iou_inputs()
    b, gt, hw = y_mask.shape
    pred = logits.unsqueeze(1).repeat(1, gt, 1).view(b * gt, hw)
    y_rows = y_mask.view(b * gt, hw)
    p_cpu = pred.clone().requires_grad_(True)
    want = O.soft_iou_loss(y_rows, p_cpu, sw)
    want.backward()
    p_gpu = pred.cuda().requires_grad_(True)
    y_dev = y_rows.cuda().to(torch.uint8) if u8 else y_rows.cuda()
    loss = OBJ.softIoULoss()(y_dev, p_gpu, sw.cuda())
    loss.backward()
    assert abs(float(loss) - float(want)) <= 1e-5 * abs(float(want))
    assert abs(float(loss) - float(g["loss"])) <= 1e-5
    assert rel(p_gpu.grad, p_cpu.grad) <= 1e-5
    assert rel(p_gpu.grad.cpu()[:, ::16], g["grad"]) <= 1e-5
    # row-wise function on its own, weighted sum of rows
    w = torch.rand(b * gt)
    p2 = pred.cuda().requires_grad_(True)
    (OBJ.softIoU(y_dev, p2) * w.cuda()).sum().backward()
    p3 = pred.clone().requires_grad_(True)
    (O.soft_iou(y_rows, p3) * w).sum().backward()
    assert rel(p2.grad, p3.grad) <= 1e-5


def test_hungarian_match_matches_reference_golden(OBJ, golden_dir):
    from oracle.make_golden import match_inputs
    g = np.load(os.path.join(golden_dir, "match.npz"))
    t_mask, t_class, overlaps = match_inputs()
    pm, pc, perm = OBJ.match([t_mask.cuda(), None], [t_class.cuda(), None], overlaps.cuda())
    assert (perm.cpu().numpy() == g["perm"]).all()
    assert (pc.cpu().numpy() == g["t_class"]).all()
    assert np.abs(pm.sum(-1).cpu().numpy() - g["t_mask_sum"]).max() == 0


@pytest.mark.parametrize("shape", [(8, 20, 10), (3, 10, 20), (5, 7, 7), (2, 1, 1), (4, 32, 32), (6, 20, 1)])
def test_hungarian_match_matches_oracle(OBJ, shape):
    from oracle import rsis_oracle as O
    b, r, t = shape
    gen = torch.Generator().manual_seed(r * 37 + t)
    big = torch.rand((b, r, t + 3), generator=gen)
    overlaps = big[:, :, 1:t + 1]                      # a strided view, like scores[B, gtT, T] slices
    t_mask = torch.rand((b, r, 16), generator=gen)
    t_class = torch.randint(0, 21, (b, r), generator=gen)
    _, _, want_perm, want_total = O.match(t_mask, t_class, overlaps)
    perm, total = OBJ.hungarian_match(big.cuda()[:, :, 1:t + 1])
    assert (perm.cpu().long() == want_perm).all()
    assert float((total.cpu().double() - want_total).abs().max()) <= 1e-5


def test_hungarian_match_with_ties_is_optimal(OBJ):
    """train.py:125: masked-out pairs all cost 10 -> many optimal assignments; the total must be the optimum."""
    from oracle import rsis_oracle as O
    gen = torch.Generator().manual_seed(4)
    b, r, t = 8, 20, 10
    scores = torch.rand((b, r, t), generator=gen)
    n_obj = torch.randint(1, 12, (b,), generator=gen)
    for i in range(b):
        scores[i, int(n_obj[i]):, :] = 10.0
        scores[i, :, int(n_obj[i]):] = 10.0
    _, _, _, want_total = O.match(torch.zeros(b, r, 1), torch.zeros(b, r, dtype=torch.long), scores)
    perm, total = OBJ.hungarian_match(scores.cuda())
    assert float((total.cpu().double() - want_total).abs().max()) <= 1e-4
    p = perm.cpu().long()
    for i in range(b):   # a valid assignment: the first T entries are distinct rows, the tail is zero
        assert len(set(p[i, :t].tolist())) == t and int(p[i, t:].abs().sum()) == 0
        assert abs(float(scores[i, p[i, :t], torch.arange(t)].sum()) - float(want_total[i])) <= 1e-4


@pytest.mark.parametrize("fused", [False, True])
def test_masked_losses_match_reference_golden(OBJ, golden_dir, fused):
    """MaskedNLLLoss / MaskedBCELoss kernels (forward + backward) against goldens produced by the UNMODIFIED
    utils/objectives.py:6-25 called as train.py:159-168 calls them.  Tolerance 1e-5 (logf / expf vs torch CPU)."""
    from oracle.make_golden import masked_loss_inputs
    g = np.load(os.path.join(golden_dir, "masked_losses.npz"))
    probs, target, sw, sw_class, stop_logits, balance = masked_loss_inputs()
    for tag, bal in (("none", None), ("bal", balance)):
        p = probs.cuda().requires_grad_(True)
        crit = OBJ.MaskedNLLLoss(balance_weight=None if bal is None else bal.cuda())
        if fused:
            loss = crit.mean(target.cuda(), p, sw.view(-1, 1).cuda())
        else:
            sel = crit(target.cuda(), p, sw.view(-1, 1).cuda())
            assert rel(sel.detach(), g[f"nll_{tag}_sel"]) <= 1e-5
            loss = torch.mean(sel)
        assert abs(float(loss.detach()) - float(g[f"nll_{tag}_sel"].mean())) <= 1e-5
        loss.backward()
        assert rel(p.grad, g[f"nll_{tag}_grad"]) <= 1e-5
    for tag, bw in (("half", 0.5), ("none", None)):
        o = stop_logits.cuda().requires_grad_(True)
        crit = OBJ.MaskedBCELoss(balance_weight=bw)
        if fused:
            loss = crit.mean(sw.cuda(), o.squeeze(), sw_class.view(-1, 1).cuda())
        else:
            sel = crit(sw.cuda(), o.squeeze(), sw_class.view(-1, 1).cuda())
            assert rel(sel.detach(), g[f"bce_{tag}_sel"]) <= 1e-5
            loss = torch.mean(sel)
        assert abs(float(loss.detach()) - float(g[f"bce_{tag}_sel"].mean())) <= 1e-5
        loss.backward()
        assert rel(o.grad, g[f"bce_{tag}_grad"]) <= 1e-5


def test_masked_losses_match_oracle_large(OBJ):
    """configs[3] scale: B*T = 640 rows, 21 classes; every row masked in a seeded pattern incl. all-unselected tail."""
    from oracle import rsis_oracle as O
    gen = torch.Generator().manual_seed(21)
    rows, c = 640, 21
    probs = torch.softmax(torch.randn((rows, c), generator=gen) * 3, -1)
    target = torch.randint(0, c, (rows, 1), generator=gen)
    sw = (torch.rand((rows, 1), generator=gen) < 0.4).float()
    sw[500:] = 0
    p1, p2 = probs.cuda().requires_grad_(True), probs.clone().requires_grad_(True)
    OBJ.MaskedNLLLoss().mean(target.cuda(), p1, sw.cuda()).backward()
    torch.mean(O.masked_nll_loss(target, p2, sw)).backward()
    assert rel(p1.grad, p2.grad) <= 1e-5
    logits = torch.randn((64, 10), generator=gen) * 4
    tgt = (torch.rand((64, 10), generator=gen) < 0.3).float()
    swc = (torch.rand((640, 1), generator=gen) < 0.7).float()
    o1, o2 = logits.cuda().requires_grad_(True), logits.clone().requires_grad_(True)
    l1 = OBJ.MaskedBCELoss(None).mean(tgt.cuda(), o1, swc.cuda())
    l2 = torch.mean(O.masked_bce_loss(tgt, o2, swc, None))
    l1.backward()
    l2.backward()
    assert abs(float(l1.detach()) - float(l2.detach())) <= 1e-5 * abs(float(l2.detach()))
    assert rel(o1.grad, o2.grad) <= 1e-5
