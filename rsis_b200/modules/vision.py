"""ResNet-101 feature taps -- counterpart of /root/reference/src/modules/vision.py:6-21.

The reference subclasses torchvision's `ResNet(Bottleneck, [3, 4, 23, 3], 1000)` and returns the five taps
`(x5, x4, x3, x2, x1)`.  Here the module tree only *holds parameters* (same names, shapes and state_dict keys as
torchvision's, including the never-executed `avgpool`/`fc`); the arithmetic runs through `rsis_b200.ops`
(conv + folded BatchNorm (+residual) (+ReLU) in one CUDA kernel per convolution, NHWC activations).
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from .. import ops
from ..ops import Act, PackedConv

RESNET101_BLOCKS = (3, 4, 23, 3)


class Bottleneck(nn.Module):
    """Parameter container laid out like torchvision's `Bottleneck` (stride on conv2, expansion 4)."""
    expansion = 4

    def __init__(self, inplanes: int, planes: int, stride: int = 1, downsample: bool = False):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = None
        if downsample:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes * 4, 1, stride=stride, bias=False),
                                            nn.BatchNorm2d(planes * 4))
        self.stride = stride

    def forward(self, x):  # pragma: no cover - never called; ResNet101.forward_act drives the packed kernels
        raise RuntimeError("rsis_b200 Bottleneck holds parameters only; call the enclosing ResNet101")


class ResNet101(nn.Module):
    """Returns intermediate features from ResNet-101 (vision.py:6-21): `forward(x) -> (x5, x4, x3, x2, x1)`."""

    def __init__(self):
        super().__init__()
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(64, RESNET101_BLOCKS[0], 1)
        self.layer2 = self._make_layer(128, RESNET101_BLOCKS[1], 2)
        self.layer3 = self._make_layer(256, RESNET101_BLOCKS[2], 2)
        self.layer4 = self._make_layer(512, RESNET101_BLOCKS[3], 2)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))  # constructed by torchvision, unused by the reference forward
        self.fc = nn.Linear(512 * 4, 1000)            # idem; present in the state_dict
        for m in self.modules():  # torchvision's initialisation
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        self._packed = None
        self._packed_key = None

    def _make_layer(self, planes: int, blocks: int, stride: int) -> nn.Sequential:
        layers = [Bottleneck(self.inplanes, planes, stride, downsample=(stride != 1 or self.inplanes != planes * 4))]
        self.inplanes = planes * 4
        for _ in range(1, blocks):
            layers.append(Bottleneck(self.inplanes, planes))
        return nn.Sequential(*layers)

    # ---- packed-weight cache (derived from the nn.Parameters; rebuilt when they change) ----------------------
    def _weights_key(self):
        return (ops.weights_epoch(), ops.bn_stats_epoch()) + tuple((p.data_ptr(), p._version)
                                             for p in list(self.parameters()) + list(self.buffers()))

    def packed(self, want_umma: bool):
        key = (self._weights_key(), want_umma)
        if self._packed is None or self._packed_key != key:
            pk = {"stem": PackedConv(self.conv1.weight, None, self.bn1)}
            if want_umma and os.environ.get("RSIS_B200_TC_STEM", "1") != "0":
                # the stem as a 1x1 tensor-core convolution over the im2col'd image: K = (kh, kw, c) order of
                # rsis_im2col, zero padded from 147 to 152 channels
                w = self.conv1.weight.detach()
                k = w.shape[1] * w.shape[2] * w.shape[3]
                cy = (k + 7) // 8 * 8
                w1 = torch.zeros((w.shape[0], cy, 1, 1), dtype=torch.float32, device=w.device)
                w1[:, :k, 0, 0] = w.permute(0, 2, 3, 1).reshape(w.shape[0], k)
                pk["stem_tc"] = PackedConv(w1, None, self.bn1, want_umma=True)
            for li in range(1, 5):
                for bi, blk in enumerate(getattr(self, f"layer{li}")):
                    p = f"layer{li}.{bi}"
                    pk[p + ".conv1"] = PackedConv(blk.conv1.weight, None, blk.bn1, want_umma=want_umma)
                    pk[p + ".conv2"] = PackedConv(blk.conv2.weight, None, blk.bn2, want_umma=want_umma)
                    pk[p + ".conv3"] = PackedConv(blk.conv3.weight, None, blk.bn3, want_umma=want_umma)
                    if blk.downsample is not None:
                        pk[p + ".down"] = PackedConv(blk.downsample[0].weight, None, blk.downsample[1],
                                                     want_umma=want_umma)
            self._packed, self._packed_key = pk, key
        return self._packed

    def packed_train(self, want_umma: bool):
        """Packs WITHOUT BatchNorm (raw convolution output): in training mode the statistics are those of the batch
        and are applied after the convolution (ops.bn_train_stats + ops.affine_act)."""
        convs = [self.conv1] + [m for li in range(1, 5) for blk in getattr(self, f"layer{li}")
                                for m in ([blk.conv1, blk.conv2, blk.conv3] +
                                          ([blk.downsample[0]] if blk.downsample is not None else []))]
        key = (tuple((c.weight.data_ptr(), c.weight._version) for c in convs), want_umma, ops.weights_epoch())
        if getattr(self, "_packed_tr", None) is None or self._packed_tr_key != key:
            pk = {"stem": PackedConv(self.conv1.weight, None, None)}
            for li in range(1, 5):
                for bi, blk in enumerate(getattr(self, f"layer{li}")):
                    p = f"layer{li}.{bi}"
                    pk[p + ".conv1"] = PackedConv(blk.conv1.weight, None, None, want_umma=want_umma)
                    pk[p + ".conv2"] = PackedConv(blk.conv2.weight, None, None, want_umma=want_umma)
                    pk[p + ".conv3"] = PackedConv(blk.conv3.weight, None, None, want_umma=want_umma)
                    if blk.downsample is not None:
                        pk[p + ".down"] = PackedConv(blk.downsample[0].weight, None, None, want_umma=want_umma)
            self._packed_tr, self._packed_tr_key = pk, key
        return self._packed_tr

    def _forward_act_train(self, x: torch.Tensor, impl: int, tape: Optional[dict] = None) -> List[Act]:
        """Training-mode forward (train.py:71-77): every BatchNorm2d uses the statistics of this batch and updates its
        running statistics; conv -> bn_train_stats -> affine_act (+residual, +ReLU).
        tape (optional dict): receives, per convolution, what `loss.backward()` needs -- (input, raw conv output,
        post-activation output, batch mean, batch 1/std) -- consumed by rsis_b200.autograd.encoder_backward."""
        fmt = ops.activation_format(impl)
        pk = self.packed_train(want_umma=(fmt == ops.FMT_SPLIT_BF16))
        F32 = ops.FMT_F32

        def conv_bn(tag, src, pc, bn, stride=1, pad=0, relu=True, residual=None, which=impl):
            raw = ops.conv2d([src], pc, stride=stride, pad=pad, out_fmt=F32, impl=which)
            if tape is None:
                scale, shift = ops.bn_train_stats(raw, bn)
                return ops.affine_act(raw, scale, shift, residual=residual, relu=relu, out_fmt=fmt)
            scale, shift, mean, invstd = ops.bn_train_stats(raw, bn, want_stats=True)
            y = ops.affine_act(raw, scale, shift, residual=residual, relu=relu, out_fmt=fmt)
            tape[tag] = (src, raw, y, mean, invstd)
            return y

        xa = ops.act_from_nchw(x, F32)
        x1 = conv_bn("stem", xa, pk["stem"], self.bn1, stride=2, pad=3, which=ops.IMPL_SIMT)
        cur = ops.maxpool3x3s2(x1)
        if tape is not None:
            tape["pool"] = (x1, cur)
        taps = []
        for li in range(1, 5):
            for bi, blk in enumerate(getattr(self, f"layer{li}")):
                p = f"layer{li}.{bi}"
                out = conv_bn(p + ".conv1", cur, pk[p + ".conv1"], blk.bn1)
                out = conv_bn(p + ".conv2", out, pk[p + ".conv2"], blk.bn2, stride=blk.stride, pad=1)
                identity = cur
                if blk.downsample is not None:
                    identity = conv_bn(p + ".down", cur, pk[p + ".down"], blk.downsample[1], stride=blk.stride,
                                       relu=False)
                cur = conv_bn(p + ".conv3", out, pk[p + ".conv3"], blk.bn3, residual=identity)
            taps.append(cur)
        x2, x3, x4, x5 = taps
        return [x5, x4, x3, x2, x1]

    def forward_act(self, x: torch.Tensor, impl: Optional[int] = None, on_tap=None, tape: Optional[dict] = None) -> List[Act]:
        """x: float32 [N,3,H,W] (any memory format) -> the five taps as NHWC activations [x5, x4, x3, x2, x1].
        on_tap(index in that list, tap): called as soon as a tap exists (eval mode), so that work which only needs
        that tap (its skip head) can be forked onto a side stream while the rest of the backbone runs."""
        ops.require_cuda(x, "ResNet101")
        impl = ops.default_impl() if impl is None else impl
        if self.training:
            return self._forward_act_train(x, impl, tape)
        fmt = ops.activation_format(impl)
        pk = self.packed(want_umma=(fmt == ops.FMT_SPLIT_BF16))
        xa = ops.act_from_nchw(x, ops.FMT_F32)
        if "stem_tc" in pk:
            st = pk["stem_tc"]
            cols = ops.im2col(xa, self.conv1.kernel_size[0], self.conv1.kernel_size[1], 2, 3, st.cin)
            x1 = ops.conv2d([cols], st, relu=True, out_fmt=fmt, impl=impl)
        else:
            x1 = ops.conv2d([xa], pk["stem"], stride=2, pad=3, relu=True, out_fmt=fmt, impl=ops.IMPL_SIMT)
        if on_tap is not None:
            on_tap(4, x1)
        cur = ops.maxpool3x3s2(x1)
        taps = []
        for li in range(1, 5):
            for bi, blk in enumerate(getattr(self, f"layer{li}")):
                p = f"layer{li}.{bi}"
                out = ops.conv2d([cur], pk[p + ".conv1"], relu=True, out_fmt=fmt, impl=impl)
                out = ops.conv2d([out], pk[p + ".conv2"], stride=blk.stride, pad=1, relu=True, out_fmt=fmt, impl=impl)
                identity = cur
                if blk.downsample is not None:
                    identity = ops.conv2d([cur], pk[p + ".down"], stride=blk.stride, out_fmt=fmt, impl=impl)
                cur = ops.conv2d([out], pk[p + ".conv3"], relu=True, residual=identity, out_fmt=fmt, impl=impl)
            taps.append(cur)
            if on_tap is not None:
                on_tap(4 - li, cur)
        x2, x3, x4, x5 = taps
        return [x5, x4, x3, x2, x1]

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, ...]:
        return tuple(ops.act_to_nchw(a) for a in self.forward_act(x))
