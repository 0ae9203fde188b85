"""TEST INFRASTRUCTURE (oracle side) -- deterministic synthetic weights for the RSIS hot path.

No pretrained weights are reachable offline (the reference downloads them at
`src/modules/model.py:30-31`; `models/README.md` is empty), so every parity
check in this repo runs on a *synthetic, well-conditioned* weight set that can
be regenerated bit-identically on any box from a seed:

* values come from numpy's Philox bit generator keyed on (seed, crc32(key)), are
  uniform doubles transformed with plain arithmetic and cast to float32 -- no
  libm calls, no dependence on tensor iteration order or SIMD width;
* backbone convolutions use He/fan-in scaling, BatchNorm running statistics are
  set analytically (mean ~ 0, var ~ 1) and every Bottleneck `bn3.weight` is
  damped to ~0.2 (SURVEY.md section 8c "required conditioning") so eval-mode
  activations stay O(1) through all 33 blocks without a calibration pass.

The key set / shapes restate the reference's state_dict contract:
  encoder (661 keys): `base.*` = torchvision ResNet(Bottleneck,[3,4,23,3]) as
      subclassed at /root/reference/src/modules/vision.py:6-9, plus
      `sk{5..1}`/`bn{5..1}` from /root/reference/src/modules/model.py:43-54
      with dims from /root/reference/src/utils/utils.py:129-131;
  decoder (16 keys): `clstm_list.{0-4}.Gates`, `conv_out`, `fc_class`,
      `fc_stop` from /root/reference/src/modules/model.py:90-120 and
      /root/reference/src/modules/clstm.py:17.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict

import numpy as np
import torch

RESNET101_BLOCKS = (3, 4, 23, 3)
RESNET101_PLANES = (64, 128, 256, 512)
SKIP_DIMS_IN = (2048, 1024, 512, 256, 64)  # utils.py:129-131 (resnet101)


def _rng(seed: int, key: str) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=[seed & 0xFFFFFFFF, zlib.crc32(key.encode())]))


def _uniform(seed: int, key: str, shape, lo: float, hi: float) -> torch.Tensor:
    n = int(np.prod(shape)) if len(shape) else 1
    u = _rng(seed, key).random(n)  # float64 in [0,1), exact integer->double construction
    v = (lo + (hi - lo) * u).astype(np.float32).reshape(shape)
    return torch.from_numpy(v)


def _conv_he(seed, key, cout, cin, k, gain=1.0):
    fan_in = cin * k * k
    a = gain * math.sqrt(3.0) * math.sqrt(2.0 / fan_in)
    return _uniform(seed, key, (cout, cin, k, k), -a, a)


def _bn(sd, seed, prefix, c, gamma=(0.8, 1.2)):
    sd[prefix + ".weight"] = _uniform(seed, prefix + ".weight", (c,), *gamma)
    sd[prefix + ".bias"] = _uniform(seed, prefix + ".bias", (c,), -0.1, 0.1)
    sd[prefix + ".running_mean"] = _uniform(seed, prefix + ".running_mean", (c,), -0.1, 0.1)
    sd[prefix + ".running_var"] = _uniform(seed, prefix + ".running_var", (c,), 0.8, 1.2)
    sd[prefix + ".num_batches_tracked"] = torch.zeros((), dtype=torch.int64)


def skip_dims_out(hidden_size: int):
    h = int(hidden_size)
    return [h, h // 2, h // 4, h // 8, h // 16]  # model.py:91-93 (Py2 integer division)


def encoder_state_dict(seed: int = 1, hidden_size: int = 128, kernel_size: int = 3) -> "OrderedDict[str, torch.Tensor]":
    """661-key FeatureExtractor state_dict in the reference's own key order."""
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    sd["base.conv1.weight"] = _conv_he(seed, "base.conv1.weight", 64, 3, 7)
    _bn(sd, seed, "base.bn1", 64)
    inplanes = 64
    for li, (nblk, planes) in enumerate(zip(RESNET101_BLOCKS, RESNET101_PLANES), start=1):
        for b in range(nblk):
            p = f"base.layer{li}.{b}"
            sd[p + ".conv1.weight"] = _conv_he(seed, p + ".conv1.weight", planes, inplanes, 1)
            _bn(sd, seed, p + ".bn1", planes)
            sd[p + ".conv2.weight"] = _conv_he(seed, p + ".conv2.weight", planes, planes, 3)
            _bn(sd, seed, p + ".bn2", planes)
            sd[p + ".conv3.weight"] = _conv_he(seed, p + ".conv3.weight", planes * 4, planes, 1)
            _bn(sd, seed, p + ".bn3", planes * 4, gamma=(0.15, 0.25))  # damped residual branch
            if b == 0:  # stride != 1 or inplanes != planes*4: true for the first block of every layer
                sd[p + ".downsample.0.weight"] = _conv_he(seed, p + ".downsample.0.weight", planes * 4, inplanes, 1)
                _bn(sd, seed, p + ".downsample.1", planes * 4)
            inplanes = planes * 4
    # constructed by torchvision, never run by the reference forward (vision.py:11-21)
    sd["base.fc.weight"] = _uniform(seed, "base.fc.weight", (1000, 2048), -0.02, 0.02)
    sd["base.fc.bias"] = _uniform(seed, "base.fc.bias", (1000,), -0.02, 0.02)
    outs = skip_dims_out(hidden_size)
    sk_out = [outs[0], outs[0], outs[1], outs[2], outs[3]]  # model.py:43-47
    names = ["5", "4", "3", "2", "1"]
    k = int(kernel_size)
    # per-head gains: backbone taps are post-ReLU with a non-zero mean (E[x^2] ~ 12 at x5), the heads
    # have no ReLU; these keep the five decoder inputs at roughly unit scale so the gates do not saturate
    gains = [0.2, 0.22, 0.45, 0.45, 1.0]
    for n, cin, cout, g in zip(names, SKIP_DIMS_IN, sk_out, gains):
        sd[f"sk{n}.weight"] = _conv_he(seed, f"sk{n}.weight", cout, cin, k, gain=g)
        sd[f"sk{n}.bias"] = _uniform(seed, f"sk{n}.bias", (cout,), -0.1, 0.1)
    for n, cout in zip(names, sk_out):
        _bn(sd, seed, f"bn{n}", cout)
    return sd


def decoder_state_dict(seed: int = 1, hidden_size: int = 128, kernel_size: int = 3,
                       num_classes: int = 21) -> "OrderedDict[str, torch.Tensor]":
    """16-key RSIS state_dict (skip_mode='concat')."""
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    outs = skip_dims_out(hidden_size)
    k = int(kernel_size)
    for i, ch in enumerate(outs):
        cin = int(hidden_size) if i == 0 else 2 * outs[i - 1]  # model.py:99-104
        fan_in = (cin + ch) * k * k
        a = 2.0 / math.sqrt(fan_in)
        p = f"clstm_list.{i}.Gates"
        sd[p + ".weight"] = _uniform(seed, p + ".weight", (4 * ch, cin + ch, k, k), -a, a)
        sd[p + ".bias"] = _uniform(seed, p + ".bias", (4 * ch,), -0.5, 0.5)
    fan_in = outs[-1] * k * k
    a = 4.0 / math.sqrt(fan_in)
    sd["conv_out.weight"] = _uniform(seed, "conv_out.weight", (1, outs[-1], k, k), -a, a)
    sd["conv_out.bias"] = _uniform(seed, "conv_out.bias", (1,), -0.5, 0.5)
    fc_dim = sum(outs)  # model.py:115-117 -> 248
    sd["fc_class.weight"] = _uniform(seed, "fc_class.weight", (num_classes, fc_dim), -0.2, 0.2)
    sd["fc_class.bias"] = _uniform(seed, "fc_class.bias", (num_classes,), -0.5, 0.5)
    sd["fc_stop.weight"] = _uniform(seed, "fc_stop.weight", (1, fc_dim), -0.2, 0.2)
    sd["fc_stop.bias"] = _uniform(seed, "fc_stop.bias", (1,), -0.5, 0.5)
    return sd


def synthetic_images(seed: int, batch: int, height: int, width: int) -> torch.Tensor:
    """x ~ U(-1,1), fp32 [B,3,H,W] (SURVEY.md section 8d 'synthetic inputs')."""
    return _uniform(seed, f"images.{batch}x{height}x{width}", (batch, 3, height, width), -1.0, 1.0)
