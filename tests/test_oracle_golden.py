"""Pin the CPU oracle (oracle/rsis_oracle.py) to outputs of the unmodified reference (tests/golden/*.npz).

The goldens were produced by oracle/make_golden.py from /root/reference's own modules; these tests
need only the committed fixtures.  Tolerance: 1e-5 tensor-relative (observed 2e-7; both sides are the
same fp32 torch CPU primitives, differences come from thread-count dependent reduction order only).
"""
import os

import numpy as np
import pytest
import torch

from oracle import rsis_oracle as O
from oracle import synth_weights as sw

TOL = 1e-5


def rel(a, b):
    a = torch.as_tensor(a, dtype=torch.float32)
    b = torch.as_tensor(b, dtype=torch.float32)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("name", ["e2e_b2_64x64_t3", "cfg1_b1_256x256_t5", "e2e_b2_96x160_t4_c9", "cfg2_b8_256x256_t10"])
def test_oracle_matches_reference_e2e(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    wseed, iseed, B, H, W, T, ncls, stride = [int(v) for v in g["meta"]]
    esd = sw.encoder_state_dict(wseed)
    dsd = sw.decoder_state_dict(wseed, num_classes=ncls)
    x = sw.synthetic_images(iseed, B, H, W)
    with torch.no_grad():
        feats = O.feature_extractor(esd, x)
    full = name == "e2e_b2_64x64_t3"
    for i, f in enumerate(feats):
        got = f if full else f[:, ::4, ::2, ::2]
        assert got.shape == g[f"feat{i}"].shape
        assert rel(got, g[f"feat{i}"]) < TOL, f"feat{i}"
    masks, classes, stops = O.test_loop(esd, dsd, x, T)
    assert rel(masks[:, :, ::stride, ::stride], g["masks"]) < TOL
    assert rel(classes, g["classes"]) < TOL
    assert rel(stops, g["stops"]) < TOL
    assert tuple(masks.shape) == (B, T, H, W) and tuple(classes.shape) == (B, T, ncls) and tuple(stops.shape) == (B, T, 1)


def test_oracle_cell_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "cells_teacher_forced.npz"))
    dsd = sw.decoder_state_dict(1)
    outs = sw.skip_dims_out(128)
    for lvl, ch in enumerate(outs):
        cin = 128 if lvl == 0 else 2 * outs[lvl - 1]
        w, b = dsd[f"clstm_list.{lvl}.Gates.weight"], dsd[f"clstm_list.{lvl}.Gates.bias"]
        x0 = sw._uniform(7, f"cell{lvl}.x0", (2, cin, 8, 8), -2.0, 2.0)
        x1 = sw._uniform(7, f"cell{lvl}.x1", (2, cin, 8, 8), -2.0, 2.0)
        h0, c0 = O.convlstm_cell(w, b, x0, None)
        h1, c1 = O.convlstm_cell(w, b, x1, (h0, c0))
        for nm, t in (("h0", h0), ("c0", c0), ("h1", h1), ("c1", c1)):
            assert rel(t, g[f"l{lvl}_{nm}"]) < TOL, (lvl, nm)


def test_synthetic_weights_contract():
    esd = sw.encoder_state_dict(1)
    dsd = sw.decoder_state_dict(1)
    assert len(esd) == 661 and len(dsd) == 16  # SURVEY.md section 8b state_dict contract
    assert tuple(dsd["clstm_list.0.Gates.weight"].shape) == (512, 256, 3, 3)
    assert tuple(dsd["clstm_list.4.Gates.weight"].shape) == (32, 40, 3, 3)
    assert tuple(dsd["fc_class.weight"].shape) == (21, 248)
    assert tuple(esd["base.fc.weight"].shape) == (1000, 2048)
    # bit-stable generator: a fixed checksum guards against silent RNG drift between boxes
    s = float(esd["base.layer3.7.conv2.weight"].double().sum())
    assert abs(s - float(sw.encoder_state_dict(1)["base.layer3.7.conv2.weight"].double().sum())) == 0.0


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/modules"), reason="reference tree only exists in the build container")
def test_reference_state_dict_keys_match():
    from oracle import ref_shims as rs
    ref = rs.load_reference()
    args = rs.make_args()
    enc, dec = ref.FeatureExtractor(args), ref.RSIS(args)
    esd, dsd = sw.encoder_state_dict(1), sw.decoder_state_dict(1)
    assert list(enc.state_dict().keys()) == list(esd.keys())
    assert list(dec.state_dict().keys()) == list(dsd.keys())
    for k, v in enc.state_dict().items():
        assert v.shape == esd[k].shape and v.dtype == esd[k].dtype, k
    for k, v in dec.state_dict().items():
        assert v.shape == dsd[k].shape, k


def test_soft_iou_oracle_matches_reference_golden(golden_dir):
    """oracle.soft_iou / soft_iou_cost_matrix / soft_iou_loss == the unmodified utils/hungarian.py:64-90,
    train.py:96-110 and utils/objectives.py:27-34 run in the build container (oracle/make_golden.py)."""
    import numpy as np
    import torch
    from oracle import rsis_oracle as O
    from oracle.make_golden import soft_iou_inputs
    g = np.load(os.path.join(golden_dir, "soft_iou.npz"))
    logits, y_mask, sw = soft_iou_inputs()
    b, gt, hw = y_mask.shape
    cost = O.soft_iou_cost_matrix(logits, y_mask, 1.0)
    assert np.abs(cost.numpy() - g["cost"]).max() <= 1e-6
    pred = logits.unsqueeze(1).repeat(1, gt, 1).view(b * gt, hw).clone().requires_grad_(True)
    loss = O.soft_iou_loss(y_mask.view(b * gt, hw), pred, sw)
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) <= 1e-6
    assert np.abs(pred.grad.numpy()[:, ::16] - g["grad"]).max() <= 1e-9 + 1e-5 * np.abs(g["grad"]).max()


def test_masked_losses_oracle_matches_reference_golden(golden_dir):
    """oracle.masked_nll_loss / masked_bce_loss == the unmodified utils/objectives.py:6-25 (over hungarian.py:10-59) called
    as train.py:159-168 calls them; selected costs and the gradients of their means."""
    import numpy as np
    import torch
    from oracle import rsis_oracle as O
    from oracle.make_golden import masked_loss_inputs
    g = np.load(os.path.join(golden_dir, "masked_losses.npz"))
    probs, target, sw, sw_class, stop_logits, balance = masked_loss_inputs()
    for tag, bal in (("none", None), ("bal", balance)):
        p = probs.clone().requires_grad_(True)
        sel = O.masked_nll_loss(target, p, sw.view(-1, 1), bal)
        torch.mean(sel).backward()
        assert np.abs(sel.detach().numpy() - g[f"nll_{tag}_sel"]).max() <= 1e-6
        assert np.abs(p.grad.numpy() - g[f"nll_{tag}_grad"]).max() <= 1e-6 * np.abs(g[f"nll_{tag}_grad"]).max()
    assert np.abs(g["nll_bal_sel"] - g["nll_none_sel"]).max() > 1e-3      # the class weights really were applied
    for tag, bw in (("half", 0.5), ("none", None)):
        o = stop_logits.clone().requires_grad_(True)
        sel = O.masked_bce_loss(sw, o, sw_class.view(-1, 1), bw)
        torch.mean(sel).backward()
        assert np.abs(sel.detach().numpy() - g[f"bce_{tag}_sel"]).max() <= 1e-6
        assert np.abs(o.grad.numpy() - g[f"bce_{tag}_grad"]).max() <= 1e-6 * np.abs(g[f"bce_{tag}_grad"]).max()


def test_match_oracle_matches_reference_golden(golden_dir):
    """oracle.match == the unmodified utils/hungarian.py:91-125 `match` (its Munkres stubbed by scipy: see
    oracle/make_golden.py::golden_match) -- pins the permutation / gather conventions."""
    import numpy as np
    from oracle import rsis_oracle as O
    from oracle.make_golden import match_inputs
    g = np.load(os.path.join(golden_dir, "match.npz"))
    t_mask, t_class, overlaps = match_inputs()
    pm, pc, perm, total = O.match(t_mask, t_class, overlaps)
    assert (perm.numpy() == g["perm"]).all()
    assert (pc.numpy() == g["t_class"]).all()
    assert np.abs(pm.sum(-1).numpy() - g["t_mask_sum"]).max() == 0
    assert (perm[:, overlaps.shape[2]:] == 0).all()   # the zero-initialised tail of permute_indices


def test_run_iter_oracle_matches_the_reference_runIter(golden_dir):
    """oracle.run_iter (forward, soft-IoU costs, matching, the three losses) + torch autograd == the reference's OWN
    `runIter` (train.py:56-197) executed in the build container by oracle/run_iter_ref.py: losses, matched classes and
    every parameter gradient (norm + 32 samples per tensor)."""
    import numpy as np
    from run_iter_parity import digest, oracle_run_iter
    from train_parity import ZERO_GRAD
    g = np.load(os.path.join(golden_dir, "run_iter.npz"))
    losses, perm, grads = oracle_run_iter()
    assert np.abs(np.array(losses) - g["losses"]).max() <= 1e-6
    assert (perm.numpy() == g["perm_class"]).all()
    names = [str(n) for n in g["names"]]
    assert sorted(grads) == names
    for n, d in zip(names, g["digests"]):
        if n in ZERO_GRAD:
            continue
        got = digest(grads[n])
        assert abs(got[0] - d[0]) <= 1e-4 * d[0] + 1e-12, n                        # gradient norm
        assert np.abs(got[2:] - d[2:]).max() <= 1e-4 * np.abs(d[2:]).max() + 1e-9, n   # samples
