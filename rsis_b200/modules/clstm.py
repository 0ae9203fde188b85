"""ConvLSTM cell -- counterpart of /root/reference/src/modules/clstm.py:7-62.

`ConvLSTMCell(args, input_size, hidden_size, kernel_size, padding).forward(input_, prev_state) -> [hidden, cell]`
with the reference's parameter layout (`Gates = nn.Conv2d(input_size + hidden_size, 4 * hidden_size, k, padding)`,
output-channel blocks `[in | remember | out | cell]`, input-channel order `[input_ | prev_hidden]`).  One call is
ONE CUDA kernel: gate convolution + sigmoid/tanh + state update; the gate planes and the channel concat never
exist in HBM.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from .. import ops
from ..ops import Act, PackedConv


class ConvLSTMCell(nn.Module):
    """Generate a convolutional LSTM cell (clstm.py:7)."""

    def __init__(self, args, input_size, hidden_size, kernel_size, padding):
        super().__init__()
        self.use_gpu = getattr(args, "use_gpu", True)
        self.input_size = int(input_size)
        self.hidden_size = int(hidden_size)
        self.Gates = nn.Conv2d(self.input_size + self.hidden_size, 4 * self.hidden_size, kernel_size,
                               padding=padding)
        self._packed = {}

    def packed(self, src_channels: Sequence[int], want_umma: bool) -> PackedConv:
        """Packed (gate-interleaved) weights for a given split of the input channels into sources."""
        w, b = self.Gates.weight, self.Gates.bias
        key = (tuple(src_channels), want_umma)
        ver = (w.data_ptr(), w._version, b.data_ptr(), b._version, ops.weights_epoch())
        hit = self._packed.get(key)
        if hit is None or hit[0] != ver:
            hit = (ver, PackedConv(w, b, None, gate_interleave=True, src_channels=src_channels, want_umma=want_umma))
            self._packed[key] = hit
        return hit[1]

    def packed_plain(self, want_umma: bool) -> PackedConv:
        """The gate convolution as a plain convolution pack (reference output-channel order [in|remember|out|cell],
        bias folded): the training-mode step computes the pre-activations with rsis_conv2d and keeps the activated
        gates for the backward (rsis_lstm_gates_fwd / _bwd)."""
        w, b = self.Gates.weight, self.Gates.bias
        key = ("plain", want_umma)
        ver = (w.data_ptr(), w._version, b.data_ptr(), b._version, ops.weights_epoch())
        hit = self._packed.get(key)
        if hit is None or hit[0] != ver:
            hit = (ver, PackedConv(w, b, None, want_umma=want_umma))
            self._packed[key] = hit
        return hit[1]

    def packed_hoisted(self, up_c: int, skip_c: int):
        """Packs for the hoisted form of the step (tcgen05 family): `input_ = [up(h_below) (up_c) | skip (skip_c)]`
        where the skip channels are the same at every time-step, so their share of the gate convolution (plus the
        bias) is computed once per image.  Returns (pc_skip, pc_step): a plain convolution pack over the skip
        channels (gate-interleaved output order, bias folded) and the per-step cell pack over
        `[up(h_below) | prev_hidden]` without bias."""
        w, b = self.Gates.weight, self.Gates.bias
        assert up_c + skip_c == self.input_size
        key = ("hoisted", up_c, skip_c)
        ver = (w.data_ptr(), w._version, b.data_ptr(), b._version, ops.weights_epoch())
        hit = self._packed.get(key)
        if hit is None or hit[0] != ver:
            wd = w.detach()
            w_skip = wd[:, up_c:up_c + skip_c].contiguous()
            w_step = torch.cat([wd[:, :up_c], wd[:, up_c + skip_c:]], 1).contiguous()
            pc_skip = PackedConv(w_skip, b, None, gate_interleave=True, src_channels=[skip_c], want_umma=True)
            pc_step = PackedConv(w_step, None, None, gate_interleave=True, src_channels=[up_c + self.hidden_size],
                                 want_umma=True)
            hit = (ver, (pc_skip, pc_step))
            self._packed[key] = hit
        return hit[1]

    def step_act(self, inputs: Sequence[Act], prev_h: Optional[Act], prev_c: Optional[torch.Tensor],
                 side_max: Optional[torch.Tensor], side_offset: int, impl: int = ops.IMPL_AUTO):
        """NHWC-level step used by the decoder. `inputs` are the parts of `input_` (concatenated along C)."""
        fmt = ops.activation_format(impl)
        have_state = prev_h is not None
        srcs = list(inputs) + ([prev_h] if have_state else [])
        if fmt == ops.FMT_SPLIT_BF16:
            chans = [self.input_size + self.hidden_size]  # the tcgen05 kernel reads ONE concatenated buffer
        else:
            chans = [s.c for s in inputs] + [self.hidden_size]  # weight K layout always includes the hidden block
        pc = self.packed(chans, want_umma=(fmt == ops.FMT_SPLIT_BF16))
        return ops.convlstm_cell(srcs, pc, prev_c if have_state else None, side_max, side_offset,
                                 want_split=(fmt == ops.FMT_SPLIT_BF16), impl=impl)

    def forward(self, input_: torch.Tensor, prev_state) -> List[torch.Tensor]:
        ops.require_cuda(input_, "ConvLSTMCell")
        if input_.shape[1] != self.input_size:
            raise RuntimeError(f"ConvLSTMCell: expected {self.input_size} input channels, got {input_.shape[1]}")
        impl = ops.default_impl()
        x = ops.act_from_nchw(input_, ops.activation_format(impl))
        prev_h = prev_c = None
        if prev_state is not None:
            h_t, c_t = prev_state
            prev_h = ops.act_from_nchw(h_t, ops.activation_format(impl))
            prev_c = ops.act_from_nchw(c_t, ops.FMT_F32).t
        h, c, _ = self.step_act([x], prev_h, prev_c, None, 0, impl)
        return [h.nchw(), c.nchw()]
