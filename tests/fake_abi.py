"""TEST INFRASTRUCTURE -- a CPU stand-in for librsis_b200.so, for host-logic tests only.

The product (`rsis_b200`) has no CPU path: every module raises without a CUDA device.  To test the HOST side of the
training path (which tensors are saved, in which order the backward primitives are called, how gradients are sliced,
summed and handed back to autograd) without a GPU, `install()` monkey-patches `rsis_b200._lib` so that the ctypes
entry points resolve to the small torch-CPU restatements below, which follow the CONTRACTS written in
include/rsis_b200.h (float32 NHWC tensors only -- the split-bf16 format and the tcgen05 family do not exist here).
The CUDA kernels themselves are checked on the GPU box (tests/test_gpu_backward.py).  Nothing outside tests/ imports
this module.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn.functional as F


def _buf(ptr: int, numel: int, dtype=torch.float32) -> torch.Tensor:
    size = torch.empty((), dtype=dtype).element_size()
    raw = (C.c_char * (numel * size)).from_address(ptr)
    return torch.frombuffer(raw, dtype=dtype)


def _struct(ref):
    return None if ref is None else ref._obj


def _pitch(t) -> int:
    return t.cstride if t.cstride > 0 else t.c


def _view(ref) -> torch.Tensor:
    """[N,H,W,C] float32 view (shares memory) of an rsis_tensor; pitched slices become strided views."""
    t = _struct(ref)
    assert t.fmt == 0, "fake ABI: float32 only"
    p = _pitch(t)
    flat = _buf(t.data, (t.n * t.h * t.w - 1) * p + t.c)
    return flat.as_strided((t.n, t.h, t.w, t.c), (t.h * t.w * p, t.w * p, p, 1))


def _nchw(v: torch.Tensor) -> torch.Tensor:
    return v.permute(0, 3, 1, 2)


class FakeLib:
    def __getattr__(self, name):
        raise AttributeError(f"fake ABI has no {name}")

    # ---- library ----
    def rsis_abi_version(self):
        from rsis_b200 import _lib
        return _lib.ABI_VERSION

    def rsis_strerror(self, s):
        return b"fake"

    def rsis_last_cuda_error(self):
        return b""

    def rsis_device_check(self):
        return 0

    def rsis_set_precision(self, mode):
        prev = getattr(self, "_precision", 0)
        self._precision = mode
        return prev

    def rsis_get_precision(self):
        return getattr(self, "_precision", 0)

    def rsis_set_static_weights(self, on):
        prev = getattr(self, "_static_weights", 0)
        self._static_weights = int(on)
        return prev

    def rsis_has_tcgen05(self):
        return 0

    # ---- packing: w_kc[(kh*KW+kw)*Cin + c][cout_pad], scale/shift [cout_pad] ----
    def rsis_conv_pack_bytes_simt(self, cout, cin, kh, kw):
        return kh * kw * cin * ((cout + 63) // 64 * 64) * 4

    def rsis_conv_pack_bytes_affine(self, cout):
        return ((cout + 63) // 64 * 64) * 4

    def rsis_conv_pack(self, w, bias, bn_w, bn_b, bn_m, bn_v, eps, cout, cin, kh, kw, gate_il, w_kc, scale, shift, st):
        cp = (cout + 63) // 64 * 64
        wt = _buf(w, cout * cin * kh * kw).view(cout, cin, kh, kw)
        perm = None
        if gate_il:  # packed channel j <- reference channel (j & 3) * Ch + (j >> 2)   (pack.cu::ref_cout)
            ch = cout // 4
            perm = torch.tensor([(j & 3) * ch + (j >> 2) for j in range(cout)])
            wt = wt[perm]
        if w_kc:
            dst = _buf(w_kc, kh * kw * cin * cp).view(kh * kw * cin, cp)
            dst.zero_()
            dst[:, :cout] = wt.permute(2, 3, 1, 0).reshape(kh * kw * cin, cout)
        sc = torch.ones(cout)
        sh = torch.zeros(cout) if not bias else _buf(bias, cout).clone()
        if bn_w:
            g, b, m, v = (_buf(p, cout) for p in (bn_w, bn_b, bn_m, bn_v))
            inv = g / torch.sqrt(v + eps)
            sh = (sh - m) * inv + b
            sc = inv
        if perm is not None:
            sc, sh = sc[perm], sh[perm]
        s_ = _buf(scale, cp)
        s_.zero_()
        s_[:cout] = sc
        t_ = _buf(shift, cp)
        t_.zero_()
        t_[:cout] = sh
        return 0

    def rsis_conv_pack_all(self, w, w_cout, w_cin, kh, kw, dgrad, ci0, nci, bias, bn_w, bn_b, bn_m, bn_v, eps, gate_il,
                           n_src, src_c, w_kc, scale, shift, w_umma, st):
        assert not w_umma, "fake ABI: no tcgen05 pack"
        if dgrad:
            wt = _buf(w, w_cout * w_cin * kh * kw).view(w_cout, w_cin, kh, kw)
            logical = wt[:, ci0:ci0 + nci].flip(2, 3).permute(1, 0, 2, 3).contiguous()
            self._keep = logical  # stays alive for the duration of the call below
            return self.rsis_conv_pack(logical.data_ptr(), bias, bn_w, bn_b, bn_m, bn_v, eps, nci, w_cout, kh, kw,
                                       gate_il, w_kc, scale, shift, st)
        return self.rsis_conv_pack(w, bias, bn_w, bn_b, bn_m, bn_v, eps, w_cout, w_cin, kh, kw, gate_il, w_kc, scale,
                                   shift, st)

    # ---- layout ----
    def rsis_nchw_to_nhwc(self, src, dst, st):
        d = _view(dst)
        n, h, w, c = d.shape
        d.copy_(_buf(src, n * c * h * w).view(n, c, h, w).permute(0, 2, 3, 1))
        return 0

    def rsis_convert(self, src, dst, st):
        _view(dst).copy_(_view(src))
        return 0

    def rsis_conv_workspace_bytes(self):
        return 64

    # ---- forward primitives ----
    def _weights(self, wref, cin_total):
        w = _struct(wref)
        cp = (w.cout + 63) // 64 * 64
        K = w.kh * w.kw * w.cin
        wk = _buf(w.w_kc, K * cp).view(w.kh, w.kw, w.cin, cp)[..., :w.cout]
        return wk.permute(3, 2, 0, 1).contiguous(), _buf(w.scale, cp)[:w.cout], _buf(w.shift, cp)[:w.cout], w

    def rsis_conv2d(self, srcs, n_src, wref, residual, y, y2, stride, pad, relu, impl, ws, ws_bytes, st):
        xs = [_view(C.byref(srcs[i])) for i in range(n_src)]
        x = torch.cat(xs, 3)
        wt, sc, sh, w = self._weights(wref, x.shape[3])
        assert x.shape[3] == w.cin
        out = F.conv2d(_nchw(x), wt, stride=stride, padding=pad) * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
        out = out.permute(0, 2, 3, 1)
        if residual is not None:
            out = out + _view(residual)
        if relu:
            out = out.clamp_min(0)
        _view(y).copy_(out)
        if y2 is not None:
            _view(y2).copy_(out)
        return 0

    def rsis_maxpool3x3s2(self, x, y, st):
        _view(y).copy_(F.max_pool2d(_nchw(_view(x)), 3, 2, 1).permute(0, 2, 3, 1))
        return 0

    def rsis_bn_workspace_bytes(self, c):
        return 16 * c

    def rsis_bn_train_stats(self, x, weight, bias, eps, momentum, rm, rv, nbt, ws, scale, shift, bmean, binvstd, st):
        v = _view(x)
        c = v.shape[3]
        flat = v.reshape(-1, c).double()
        m = flat.shape[0]
        mean = flat.mean(0)
        var = flat.var(0, unbiased=False)
        invstd = 1.0 / torch.sqrt(var + eps)
        g = _buf(weight, c).double() if weight else torch.ones(c, dtype=torch.float64)
        b = _buf(bias, c).double() if bias else torch.zeros(c, dtype=torch.float64)
        _buf(scale, c).copy_((g * invstd).float())
        _buf(shift, c).copy_((b - mean * g * invstd).float())
        if bmean:
            _buf(bmean, c).copy_(mean.float())
        if binvstd:
            _buf(binvstd, c).copy_(invstd.float())
        if rm and rv:
            factor = momentum
            if nbt:
                n_ = _buf(nbt, 1, torch.int64)
                n_ += 1
                if momentum < 0:
                    factor = 1.0 / float(n_.item())
            unb = var * m / (m - 1) if m > 1 else var
            r1, r2 = _buf(rm, c), _buf(rv, c)
            r1.copy_(((1 - factor) * r1.double() + factor * mean).float())
            r2.copy_(((1 - factor) * r2.double() + factor * unb).float())
        return 0

    def rsis_affine_act(self, x, scale, shift, residual, relu, y, y2, st):
        v = _view(x)
        c = v.shape[3]
        out = v * _buf(scale, c) + _buf(shift, c)
        if residual is not None:
            out = out + _view(residual)
        if relu:
            out = out.clamp_min(0)
        _view(y).copy_(out)
        if y2 is not None:
            _view(y2).copy_(out)
        return 0

    def rsis_upsample_bilinear(self, x, y, st):
        yv = _view(y)
        out = F.interpolate(_nchw(_view(x)), size=(yv.shape[1], yv.shape[2]), mode="bilinear", align_corners=True)
        yv.copy_(out.permute(0, 2, 3, 1))
        return 0

    def rsis_mask_head(self, x, w, bias, ks, logits, prob, prob_stride, st):
        v = _view(x)
        n, h, w_, c = v.shape
        wt = _buf(w, c * ks * ks).view(1, c, ks, ks)
        b = _buf(bias, 1) if bias else None
        out = F.conv2d(_nchw(v), wt, b, padding=ks // 2)
        if logits:
            _buf(logits, n * h * w_).copy_(out.reshape(-1))
        if prob:
            span = (n - 1) * prob_stride + h * w_
            _buf(prob, span).as_strided((n, h * w_), (prob_stride, 1)).copy_(torch.sigmoid(out.reshape(n, -1)))
        return 0

    def rsis_convlstm_cell(self, srcs, n_src, wref, c_prev, gate_preact, h_out, h_split, c_out, side_max, side_stride,
                           side_offset, impl, ws, ws_bytes, st):
        """Header contract of rsis_convlstm_cell, float32 sources only: gates in (hidden channel, gate) order."""
        assert not gate_preact and h_split is None
        xs = [_view(C.byref(srcs[i])) for i in range(n_src)]
        x = torch.cat(xs, 3)
        wt, sc, sh, w = self._weights(wref, x.shape[3])
        assert w.gate_interleaved
        if x.shape[3] < w.cin:   # state None: the prev_hidden block of the input is zeros
            x = torch.cat([x, torch.zeros(x.shape[:3] + (w.cin - x.shape[3],))], 3)
        g = F.conv2d(_nchw(x), wt, padding=w.kh // 2) * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
        g = g.permute(0, 2, 3, 1)
        n, h, w_, _ = g.shape
        ch = w.cout // 4
        g = g.reshape(n, h, w_, ch, 4)
        i, f, o = torch.sigmoid(g[..., 0]), torch.sigmoid(g[..., 1]), torch.sigmoid(g[..., 2])
        gg = torch.tanh(g[..., 3])
        cp = _buf(c_prev, n * h * w_ * ch).view(n, h, w_, ch) if c_prev else torch.zeros(n, h, w_, ch)
        c = f * cp + i * gg
        hh = o * torch.tanh(c)
        _view(c_out).copy_(c)
        _view(h_out).copy_(hh)
        if side_max:
            keys = _buf(side_max, n * side_stride, torch.int32).view(n, side_stride)
            keys[:, side_offset:side_offset + ch] = self._float_to_key(hh.amax(dim=(1, 2)).contiguous())
        return 0

    @staticmethod
    def _bits_to_float(bits: torch.Tensor) -> torch.Tensor:
        import numpy as np
        return torch.from_numpy(bits.numpy().astype(np.uint32).view(np.float32).copy())

    @staticmethod
    def _float_to_key(f: torch.Tensor) -> torch.Tensor:
        import numpy as np
        b = f.contiguous().numpy().view(np.uint32).astype(np.int64)
        k = np.where(b & 0x80000000, (~b) & 0xFFFFFFFF, b | 0x80000000)
        return torch.from_numpy(k.astype(np.uint32).view(np.int32).copy())

    def rsis_class_stop_heads(self, side, n, f, wc, bc, nc, ws_, bs, feat_out, probs, pstride, stop, stop_prob,
                              sstride, st):
        keys = _buf(side, n * f, torch.int32).view(n, f)
        feat = self._bits_to_float(FakeLib._decode_keys(keys)).view(n, f)
        if feat_out:
            _buf(feat_out, n * f).view(n, f).copy_(feat)
        logits = F.linear(feat, _buf(wc, nc * f).view(nc, f), _buf(bc, nc))
        p = torch.softmax(logits, 1)
        _buf(probs, (n - 1) * pstride + nc).as_strided((n, nc), (pstride, 1)).copy_(p)
        sl = F.linear(feat, _buf(ws_, f).view(1, f), _buf(bs, 1)).view(-1)
        if stop:
            _buf(stop, (n - 1) * sstride + 1).as_strided((n,), (sstride,)).copy_(sl)
        if stop_prob:
            _buf(stop_prob, (n - 1) * sstride + 1).as_strided((n,), (sstride,)).copy_(torch.sigmoid(sl))
        return 0

    @staticmethod
    def _decode_keys(keys: torch.Tensor) -> torch.Tensor:
        k = keys.to(torch.int64) & 0xFFFFFFFF
        return torch.where((k & 0x80000000) != 0, k & 0x7FFFFFFF, (~k) & 0xFFFFFFFF)

    # ---- backward primitives ----
    def rsis_conv_dgrad_weights(self, w, cout, cin, kh, kw, ci0, nci, out, st):
        wt = _buf(w, cout * cin * kh * kw).view(cout, cin, kh, kw)
        o = wt[:, ci0:ci0 + nci].flip(2, 3).permute(1, 0, 2, 3)
        _buf(out, nci * cout * kh * kw).view(nci, cout, kh, kw).copy_(o)
        return 0

    def rsis_wgrad_workspace_bytes(self):
        return 64

    def rsis_conv2d_wgrad(self, x, dy, kh, kw, stride, pad, dw, dbias, accumulate, impl, ws, ws_bytes, st):
        xv, gv = _nchw(_view(x)), _nchw(_view(dy))
        cout, cin = gv.shape[1], xv.shape[1]
        if dw:
            g = torch.nn.grad.conv2d_weight(xv.contiguous(), (cout, cin, kh, kw), gv.contiguous(), stride=stride,
                                            padding=pad)
            d = _buf(dw, cout * cin * kh * kw).view(cout, cin, kh, kw)
            d.copy_(d + g if accumulate else g)
        if dbias:
            b = _buf(dbias, cout)
            s = gv.sum((0, 2, 3))
            b.copy_(b + s if accumulate else s)
        return 0

    def rsis_dilate2x(self, x, y, st):
        yv = _view(y)
        yv.zero_()
        yv[:, ::2, ::2, :] = _view(x)
        return 0

    def rsis_bn_train_bwd(self, x_raw, y_act, dy, weight, mean, invstd, ws, dweight, dbias, dw_acc, db_acc, dx, dres,
                          st):
        xv, g = _view(x_raw), _view(dy).clone()
        c = xv.shape[3]
        if y_act is not None:
            g = g * (_view(y_act) > 0)
        mu, is_ = _buf(mean, c), _buf(invstd, c)
        xh = (xv - mu) * is_
        m = xv.numel() // c
        db = g.reshape(-1, c).double().sum(0).float()
        dw = (g * xh).reshape(-1, c).double().sum(0).float()
        w = _buf(weight, c) if weight else torch.ones(c)
        _view(dx).copy_(w * is_ * (g - db / m - xh * dw / m))
        _buf(dweight, c).copy_(dw)
        _buf(dbias, c).copy_(db)
        if dw_acc:
            _buf(dw_acc, c).add_(dw)
        if db_acc:
            _buf(db_acc, c).add_(db)
        if dres is not None:
            _view(dres).copy_(g)
        return 0

    def rsis_maxpool3x3s2_bwd(self, x, dy, dx, st):
        with torch.enable_grad():
            xv = _nchw(_view(x)).clone().requires_grad_(True)
            out = F.max_pool2d(xv, 3, 2, 1)
            out.backward(_nchw(_view(dy)).contiguous())
        _view(dx).copy_(xv.grad.permute(0, 2, 3, 1))
        return 0

    def rsis_lstm_gates_fwd(self, gates, c_prev, h_out, h_out2, c_out, st):
        g = _view(gates)
        ch = g.shape[3] // 4
        i, f, o, gg = torch.sigmoid(g[..., :ch]), torch.sigmoid(g[..., ch:2 * ch]), \
            torch.sigmoid(g[..., 2 * ch:3 * ch]), torch.tanh(g[..., 3 * ch:])
        cp = _buf(c_prev, i.numel()).view(i.shape) if c_prev else torch.zeros_like(i)
        c = f * cp + i * gg
        h = o * torch.tanh(c)
        g.copy_(torch.cat([i, f, o, gg], 3))
        _view(c_out).copy_(c)
        _view(h_out).copy_(h)
        if h_out2 is not None:
            _view(h_out2).copy_(h)
        return 0

    def rsis_lstm_gates_bwd(self, gates, c_prev, c_new, dh_a, dh_b, dc_next, dgates, dc_prev, st):
        g = _view(gates)
        ch = g.shape[3] // 4
        i, f, o, gg = g[..., :ch], g[..., ch:2 * ch], g[..., 2 * ch:3 * ch], g[..., 3 * ch:]
        cp = _buf(c_prev, i.numel()).view(i.shape) if c_prev else torch.zeros_like(i)
        cn = _buf(c_new, i.numel()).view(i.shape)
        dh = torch.zeros_like(i)
        if dh_a is not None:
            dh = dh + _view(dh_a)
        if dh_b is not None:
            dh = dh + _view(dh_b)
        dcn = _view(dc_next) if dc_next is not None else torch.zeros_like(i)
        tc = torch.tanh(cn)
        dc = dcn + dh * o * (1 - tc * tc)
        out = torch.cat([dc * gg * i * (1 - i), dc * cp * f * (1 - f), dh * tc * o * (1 - o), dc * i * (1 - gg * gg)], 3)
        _view(dgates).copy_(out)
        _buf(dc_prev, i.numel()).view(i.shape).copy_(dc * f)
        return 0

    def rsis_global_maxpool(self, h, packed, stride, off, st):
        v = _view(h)
        n, hh, ww, c = v.shape
        flat = v.reshape(n, hh * ww, c)
        best = flat.max(1)[0]
        arg = (flat == best.unsqueeze(1)).float().argmax(1)  # first maximal index
        pk = _buf(packed, n * stride, torch.int64).view(n, stride)
        key = self._float_to_key(best).to(torch.int64) & 0xFFFFFFFF
        val = (key << 32) | (0xFFFFFFFF - arg.to(torch.int64))
        # int64 storage of an unsigned 64-bit value: reinterpret through numpy
        import numpy as np
        pk[:, off:off + c] = torch.from_numpy(val.numpy().astype(np.uint64).view(np.int64).copy())
        return 0

    def rsis_global_maxpool_finish(self, packed, n, stride, keys, idx, st):
        import numpy as np
        pk = _buf(packed, n * stride, torch.int64).numpy().view(np.uint64)
        k = (pk >> np.uint64(32)).astype(np.uint32)
        ix = (np.uint64(0xFFFFFFFF) - (pk & np.uint64(0xFFFFFFFF))).astype(np.int64)
        _buf(keys, n * stride, torch.int32).copy_(torch.from_numpy(k.view(np.int32).copy()))
        _buf(idx, n * stride, torch.int32).copy_(torch.from_numpy(ix.astype(np.int32)))
        return 0

    def rsis_global_maxpool_bwd(self, dside, idx, stride, off, dh, st):
        v = _view(dh)
        n, hh, ww, c = v.shape
        d = _buf(dside, n * stride).view(n, stride)[:, off:off + c]
        ix = _buf(idx, n * stride, torch.int32).view(n, stride)[:, off:off + c].long()
        flat = v.reshape(n, hh * ww, c)  # a view: v is contiguous here
        assert flat.data_ptr() == v.data_ptr()
        for b in range(n):
            flat[b, ix[b], torch.arange(c)] += d[b]
        return 0

    def rsis_upsample_bilinear_bwd(self, dy, dx, st):
        dv = _view(dx)
        g = _nchw(_view(dy)).contiguous()
        with torch.enable_grad():
            x = torch.zeros(_nchw(dv).shape, requires_grad=True)
            out = F.interpolate(x, size=g.shape[-2:], mode="bilinear", align_corners=True)
            out.backward(g)
        dv.copy_(x.grad.permute(0, 2, 3, 1))
        return 0

    def rsis_class_stop_heads_bwd(self, feat, probs, dclass, dstop, n, f, wc, nc, ws_, scratch, dfeat, dwc, dbc, dws,
                                  dbs, st):
        ft = _buf(feat, n * f).view(n, f)
        p = _buf(probs, n * nc).view(n, nc)
        dp = _buf(dclass, n * nc).view(n, nc) if dclass else torch.zeros(n, nc)
        ds = _buf(dstop, n).view(n, 1) if dstop else torch.zeros(n, 1)
        dl = p * (dp - (p * dp).sum(1, keepdim=True))
        Wc, Ws = _buf(wc, nc * f).view(nc, f), _buf(ws_, f).view(1, f)
        _buf(dfeat, n * f).view(n, f).copy_(dl @ Wc + ds @ Ws)
        _buf(dwc, nc * f).view(nc, f).add_(dl.t() @ ft)
        _buf(dbc, nc).add_(dl.sum(0))
        _buf(dws, f).view(1, f).add_(ds.t() @ ft)
        _buf(dbs, 1).add_(ds.sum())
        return 0


def _adam(self, p, g, m, v, n, lr, b1, b2, eps, wd, step0, repeats, st):
    P, G, M, V = _buf(p, n), _buf(g, n), _buf(m, n), _buf(v, n)
    for r in range(repeats):
        t = step0 + r + 1
        gg = G + wd * P
        M.copy_(M + (gg - M) * (1 - b1))
        V.copy_(V * b2 + (1 - b2) * gg * gg)
        denom = V.sqrt() / (1 - b2 ** t) ** 0.5 + eps
        P.sub_((lr / (1 - b1 ** t)) * (M / denom))
    return 0


FakeLib.rsis_adam_step = _adam


# ---- section 8f entry points (objectives / post-processing), restated from the header contracts ----
def _gt_rows(ptr, is_u8, n):
    return _buf(ptr, n, torch.uint8).float() if is_u8 else _buf(ptr, n)


def _soft_iou_ws(self, b, g):
    return (b * g * 2 + b + 1) * 4


def _soft_iou_cost(self, logits, gt, is_u8, b, g, hw, eps, weight, ws, cost, sb, sg, num_out, den_out, st):
    s = torch.sigmoid(_buf(logits, b * hw).view(b, 1, hw))
    y = _gt_rows(gt, is_u8, b * g * hw).view(b, g, hw)
    num = (s * y).sum(-1)
    den = (s + y - s * y).sum(-1) + eps
    span = (b - 1) * sb + (g - 1) * sg + 1
    out = _buf(cost, span).as_strided((b, g), (sb, sg))
    out.copy_(weight * (1 - num / den))
    if num_out:
        _buf(num_out, b * g).copy_(num.reshape(-1))
    if den_out:
        _buf(den_out, b * g).copy_(den.reshape(-1))
    return 0


def _soft_iou_bwd(self, logits, gt, is_u8, rows, hw, num, den, dcost, weight, dlogits, st):
    s = torch.sigmoid(_buf(logits, rows * hw).view(rows, hw))
    y = _gt_rows(gt, is_u8, rows * hw).view(rows, hw)
    n, d, dc = (_buf(p_, rows).view(rows, 1) for p_ in (num, den, dcost))
    _buf(dlogits, rows * hw).view(rows, hw).copy_(-weight * dc * (y * d - n * (1 - y)) / (d * d) * s * (1 - s))
    return 0


def _hungarian(self, cost, sb, sr, sc, b, rows, cols, perm, perm_len, total, st):
    import numpy as np
    from scipy.optimize import linear_sum_assignment
    span = (b - 1) * sb + (rows - 1) * sr + (cols - 1) * sc + 1
    c = _buf(cost, span).as_strided((b, rows, cols), (sb, sr, sc)).double().numpy()
    p = _buf(perm, b * perm_len, torch.int32).view(b, perm_len)
    p.zero_()
    for i in range(b):
        r, cc = linear_sum_assignment(c[i])
        for row, col in zip(r, cc):
            if col < perm_len:
                p[i, col] = int(row)
        if total:
            _buf(total, b)[i] = float(c[i][r, cc].sum())
    return 0


def _rle_ws(self, n, h, w):
    return 16


def _rle_encode(self, masks, th, ignore, n, h, w, ws, counts, max_runs, n_runs, areas, st):
    import numpy as np
    m = _buf(masks, n * h * w).view(n, h, w).numpy()
    seg = (m > th).astype(np.uint8)
    if ignore:
        seg[_buf(ignore, n * h * w, torch.uint8).view(n, h, w).numpy() == 1] = 0
    cnt = _buf(counts, n * max_runs, torch.int32).view(n, max_runs)
    nr = _buf(n_runs, n, torch.int32)
    for i in range(n):
        t = seg[i].T.reshape(-1)
        pos = np.flatnonzero(t != np.concatenate([[0], t[:-1]]))
        runs = np.diff(np.concatenate([[0], pos, [t.size]]))
        nr[i] = len(runs)
        k = min(len(runs), max_runs)
        cnt[i, :k] = torch.from_numpy(runs[:k].astype(np.int32))
        if areas:
            _buf(areas, n, torch.int32)[i] = int(t.sum())
    return 0


def _sel(sw_ptr, n):
    return _buf(sw_ptr, n).to(torch.uint8) != 0          # the reference's `sw.byte()` mask


def _masked_nll_fwd(self, probs, target, sw, balance, rows, c, cost_rows, sum_count, st):
    p = _buf(probs, rows * c).view(rows, c)
    t = _buf(target, rows, torch.int64)
    cost = -torch.log(p.gather(1, t.view(-1, 1))).view(-1)
    if balance:
        cost = cost * _buf(balance, c)[t]
    sel = _sel(sw, rows)
    if cost_rows:
        _buf(cost_rows, rows).copy_(torch.where(sel, cost, torch.zeros_like(cost)))
    _buf(sum_count, 2).copy_(torch.stack([cost[sel].sum(), sel.float().sum()]))
    return 0


def _masked_nll_bwd(self, probs, target, sw, balance, dcost, dstride, rows, c, dprobs, st):
    p = _buf(probs, rows * c).view(rows, c)
    t = _buf(target, rows, torch.int64)
    g = _buf(dcost, (rows - 1) * dstride + 1)[::dstride] if dstride else _buf(dcost, 1).expand(rows)
    w = _buf(balance, c)[t] if balance else torch.ones(rows)
    d = torch.zeros(rows, c)
    d.scatter_(1, t.view(-1, 1), (-g * w / p.gather(1, t.view(-1, 1)).view(-1) * _sel(sw, rows).float()).view(-1, 1))
    _buf(dprobs, rows * c).view(rows, c).copy_(d)
    return 0


def _masked_bce_fwd(self, target, logits, sw, bw, n, cost_rows, out3, st):
    t, o = _buf(target, n), _buf(logits, n)
    if bw < 0:
        bw = float(t.sum() / n)
    mx = (-o).clamp(min=0)
    lv = o - o * t + mx + ((-mx).exp() + (-o - mx).exp()).log()
    cost = (1 - bw) * lv * t + bw * lv * (1 - t)
    sel = _sel(sw, n)
    if cost_rows:
        _buf(cost_rows, n).copy_(torch.where(sel, cost, torch.zeros_like(cost)))
    _buf(out3, 3).copy_(torch.stack([cost[sel].sum(), sel.float().sum(), torch.tensor(bw)]))
    return 0


def _masked_bce_bwd(self, target, logits, sw, bw_ptr, dcost, dstride, n, dlogits, st):
    t, o = _buf(target, n), _buf(logits, n)
    bw = float(_buf(bw_ptr, 1)[0])
    g = _buf(dcost, (n - 1) * dstride + 1)[::dstride] if dstride else _buf(dcost, 1).expand(n)
    _buf(dlogits, n).copy_(g * ((1 - bw) * t + bw * (1 - t)) * (torch.sigmoid(o) - t) * _sel(sw, n).float())
    return 0


FakeLib.rsis_masked_nll_fwd = _masked_nll_fwd
FakeLib.rsis_masked_nll_bwd = _masked_nll_bwd
FakeLib.rsis_masked_bce_fwd = _masked_bce_fwd
FakeLib.rsis_masked_bce_bwd = _masked_bce_bwd
FakeLib.rsis_soft_iou_workspace_bytes = _soft_iou_ws
FakeLib.rsis_soft_iou_cost = _soft_iou_cost
FakeLib.rsis_soft_iou_bwd = _soft_iou_bwd
FakeLib.rsis_hungarian_match = _hungarian
FakeLib.rsis_rle_workspace_bytes = _rle_ws
FakeLib.rsis_rle_encode = _rle_encode


def install(monkeypatch):
    """Routes rsis_b200's ABI calls to FakeLib and lifts the CUDA-only guards (host-logic tests on CPU)."""
    from rsis_b200 import _lib, ops
    fake = FakeLib()
    monkeypatch.setattr(_lib, "_lib", fake)
    monkeypatch.setattr(_lib, "load", lambda: fake)
    monkeypatch.setattr(_lib, "stream_ptr", lambda: 0)
    monkeypatch.setattr(_lib, "workspace", lambda: (0, 0))
    monkeypatch.setattr(ops, "require_cuda", lambda t, what: None)
    monkeypatch.setenv("RSIS_B200_IMPL", "simt")
    return fake
