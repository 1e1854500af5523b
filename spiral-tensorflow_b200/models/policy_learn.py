"""The learner's policy convolutions on libspiral_b200.so: forward, input gradient and weight gradient.

When the learner evaluates the A2C loss over ``[B*T]`` frames (agent.py:161-239) TensorFlow differentiates through the
trunk and the location heads of models/policy.py -- ~45 convolutions of <= 32 channels on <= 64x64 maps per frame, forward
and backward.  ``Policy(learn_native=True)`` runs all of them through ``csrc/policy_learn.cu`` (mma.sync implicit GEMMs with
a bf16 operand split -- three planes / six products, "bf16x6": fp32-exact operands, fp32 accumulation; the trunk is ~20
convolutions deep and two planes, 16 bits, leave 1.2e-4 on the logits) as ``torch.autograd.Function``s, NHWC fp32 like
the reference's graph:

    ``conv2d``            tf.layers.conv2d, stride 1 'same' (5x5, 3x3) or stride 2 'valid' (4x4)   policy.py:94-109, 349-362, 290-295
    ``conv_transpose2d``  tf.layers.conv2d_transpose 4x4 / stride 2 'same'                          policy.py:242-247, 275-281

The input gradient of each is again one of these convolutions with a re-indexed kernel -- flipped taps for stride 1, four
parity classes of 2x2 taps for the stride-2 convolution (its adjoint is a transposed convolution), a plain 4x4 / stride-2
convolution for the transposed one -- so ``spiral_lconv_pack`` (a strided gather into the MMA's [rows][taps * channels]
layout) + ``spiral_lconv_forward`` serve forward and dgrad, and ``spiral_lconv_wgrad`` is the weight gradient of both kinds
(for the transposed convolution with the roles of the two maps exchanged).  Bias gradients, ReLU, residual adds, the dense
layers and the LSTM cell are PyTorch ops (elementwise / cuBLAS GEMMs); no cuDNN kernel is left on the learner's path.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn.functional as F

from .. import _native as N

_WS = {}
# operand format of the kernels (the C ABI's `n_planes`): 4 = f16x3 (two fp16 planes, 22 operand bits, three products; the
# default), 3 = bf16x6 (three bf16 planes, 24 bits, six products), 2 = bf16x3 (16 bits)
PLANES = 4
FORMATS = {"f16x3": 4, "bf16x6": 3, "bf16x3": 2}


def _fmt(cin):
    return PLANES


def _workspace(dev, n):
    t = _WS.get(dev)
    if t is None or t.numel() < n:
        t = torch.empty((max(n, 1 << 22),), dtype=torch.float32, device=dev)
        _WS[dev] = t
    return t


def _pack(w, rows, ck, taps, s_row, s_c, s_ky, s_kx):
    """-> PLANES contiguous bf16 planes [round8(rows)][KS] (as int16) for spiral_lconv_forward."""
    n = int(N.lib().spiral_lconv_packed_elems(ck, rows, len(taps)))
    fmt = _fmt(ck)
    planes = torch.empty(((3 if fmt == 3 else 2) * n,), dtype=torch.int16, device=w.device)
    ky = (C.c_int32 * len(taps))(*[t[0] for t in taps])
    kx = (C.c_int32 * len(taps))(*[t[1] for t in taps])
    N.check(N.lib().spiral_lconv_pack(N.ptr(w), N.ptr(planes), fmt, rows, ck, len(taps), s_row, s_c, s_ky, s_kx, ky, kx,
                                      N.stream_ptr()), "spiral_lconv_pack")
    return planes


def _launch(x, packed, bias, out, cin, cout, hout, wout, kh, kw, stride, pad_t, pad_l, scatter=(1, 1, 0, 0), relu=False):
    B, hin, win, _ = x.shape
    oh, ow = out.shape[1], out.shape[2]
    N.check(N.lib().spiral_lconv_forward(N.ptr(x), N.ptr(packed), _fmt(cin), N.ptr(bias), N.ptr(out), B, hin, win,
                                         cin, hout, wout, cout, kh, kw, stride, pad_t, pad_l, oh, ow, scatter[0], scatter[1],
                                         scatter[2], scatter[3], int(relu), N.stream_ptr()), "spiral_lconv_forward")
    return out


def _wgrad(x, dy, kh, kw, stride, pad_t, pad_l):
    """-> [kh*kw*Cin, Cout] with row (ky*kw + kx)*Cin + c."""
    B, hin, win, cin = x.shape
    _, hout, wout, cout = dy.shape
    dw = torch.empty((kh * kw * cin, cout), dtype=torch.float32, device=x.device)
    ws = _workspace(x.device, int(N.lib().spiral_lconv_wgrad_workspace_floats(cin, cout, kh * kw)))
    N.check(N.lib().spiral_lconv_wgrad(N.ptr(x), N.ptr(dy), N.ptr(dw), N.ptr(ws), PLANES, B, hin, win, cin, hout, wout, cout, kh, kw,
                                       stride, pad_t, pad_l, N.stream_ptr()), "spiral_lconv_wgrad")
    return dw


def _pack_classes(w, rows, ck, class_taps, s_row, s_c, s_ky, s_kx):
    """The four parity classes' kernels, packed back to back for spiral_lconv_forward_classes."""
    n = int(N.lib().spiral_lconv_packed_elems(ck, rows, 4))
    fmt = _fmt(ck)
    per = (3 if fmt == 3 else 2) * n
    planes = torch.empty((4 * per,), dtype=torch.int16, device=w.device)
    for c, taps in enumerate(class_taps):
        ky = (C.c_int32 * 4)(*[t[0] for t in taps])
        kx = (C.c_int32 * 4)(*[t[1] for t in taps])
        N.check(N.lib().spiral_lconv_pack(N.ptr(w), N.ptr(planes[c * per:(c + 1) * per]), fmt, rows, ck, 4, s_row, s_c, s_ky, s_kx,
                                          ky, kx, N.stream_ptr()), "spiral_lconv_pack")
    return planes


def _launch_classes(x, packed, bias, out, cin, cout, hout, wout, pads_t, pads_l, relu=False):
    B, hin, win, _ = x.shape
    pt = (C.c_int32 * 4)(*pads_t)
    pl_ = (C.c_int32 * 4)(*pads_l)
    N.check(N.lib().spiral_lconv_forward_classes(N.ptr(x), N.ptr(packed), _fmt(cin), N.ptr(bias), N.ptr(out), B, hin, win, cin,
                                                 hout, wout, cout, pt, pl_, out.shape[1], out.shape[2], int(relu),
                                                 N.stream_ptr()), "spiral_lconv_forward_classes")
    return out


def _all_taps(kh, kw):
    return [(ky, kx) for ky in range(kh) for kx in range(kw)]


class _Conv2d(torch.autograd.Function):
    """y = conv(x, w) + b;  x [B,H,W,Cin] NHWC, w [Cout,Cin,KH,KW] (torch layout), stride 1 with pad = K // 2 ('same', odd K)
    or stride 2 without padding ('valid', 4x4)."""

    @staticmethod
    def forward(ctx, x, w, b, stride, pad, relu=False):
        x = x.contiguous()
        w = w.contiguous()
        cout, cin, kh, kw = w.shape
        B, H, Wd, _ = x.shape
        ho, wo = (H + 2 * pad - kh) // stride + 1, (Wd + 2 * pad - kw) // stride + 1
        packed = _pack(w, cout, cin, _all_taps(kh, kw), cin * kh * kw, kh * kw, kw, 1)
        out = torch.empty((B, ho, wo, cout), dtype=torch.float32, device=x.device)
        _launch(x, packed, b, out, cin, cout, ho, wo, kh, kw, stride, pad, pad, relu=relu)
        ctx.save_for_backward(x, w, out if relu else None)       # relu fused into the epilogue: its mask is (out > 0)
        ctx.geom = (stride, pad, b is not None)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        stride, pad, has_bias = ctx.geom
        dy = dy.contiguous()
        if y is not None:
            dy = torch.ops.aten.threshold_backward(dy, y, 0.0)
        cout, cin, kh, kw = w.shape
        B, H, Wd, _ = x.shape
        _, ho, wo, _ = dy.shape
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            if stride == 1:
                # adjoint of a stride-1 convolution: the same convolution over dy with flipped taps, rows <-> channels
                taps = [(kh - 1 - ky, kw - 1 - kx) for ky in range(kh) for kx in range(kw)]
                packed = _pack(w, cin, cout, taps, kh * kw, cin * kh * kw, kw, 1)
                dx = torch.empty((B, H, Wd, cin), dtype=torch.float32, device=x.device)
                _launch(dy, packed, None, dx, cout, cin, H, Wd, kh, kw, 1, kh - 1 - pad, kw - 1 - pad)
            else:
                # adjoint of the 4x4 / stride-2 'valid' convolution: dx[2a + py] = dy[a - 1] * w[py + 2] + dy[a] * w[py]
                assert stride == 2 and kh == 4 and kw == 4 and pad == 0
                # the four classes write rows / columns 0 .. 2 ho + 1; an odd map has one more (no output reads it: zero)
                dx = torch.empty((B, H, Wd, cin), dtype=torch.float32, device=x.device)
                if H > 2 * ho + 2:
                    dx[:, 2 * ho + 2:].zero_()
                if Wd > 2 * wo + 2:
                    dx[:, :, 2 * wo + 2:].zero_()
                class_taps = [[(ky, kx) for ky in (py + 2, py) for kx in (px + 2, px)] for py in (0, 1) for px in (0, 1)]
                packed = _pack_classes(w, cin, cout, class_taps, kh * kw, cin * kh * kw, kw, 1)
                _launch_classes(dy, packed, None, dx, cout, cin, ho + 1, wo + 1, (1, 1, 1, 1), (1, 1, 1, 1))
        if ctx.needs_input_grad[1]:
            g = _wgrad(x, dy, kh, kw, stride, pad, pad)                        # [(ky*kw + kx)*cin + ci][co]
            dw = g.view(kh, kw, cin, cout).permute(3, 2, 0, 1).contiguous()
        if has_bias and ctx.needs_input_grad[2]:
            db = dy.sum((0, 1, 2))
        return dx, dw, db, None, None, None


class _ConvT2d(torch.autograd.Function):
    """y = conv_transpose(x, w) + b, 4x4 / stride 2 / 'same' (ConvTranspose2d(k=4, s=2, p=1));  x [B,H,W,Cin] NHWC,
    w [Cin,Cout,4,4] (torch layout) -> [B,2H,2W,Cout].  Four parity classes of 2x2 taps:
    y[2a + py] = x[a - 1 + r + (py == 1)] * w[ky(py, r)],  ky(0, .) = (3, 1), ky(1, .) = (2, 0)."""

    @staticmethod
    def forward(ctx, x, w, b, relu=False):
        x = x.contiguous()
        w = w.contiguous()
        cin, cout, kh, kw = w.shape
        assert kh == 4 and kw == 4
        B, H, Wd, _ = x.shape
        out = torch.empty((B, 2 * H, 2 * Wd, cout), dtype=torch.float32, device=x.device)
        class_taps, pads_t, pads_l = [], [], []
        for py in (0, 1):
            for px in (0, 1):
                kys = (3, 1) if py == 0 else (2, 0)
                kxs = (3, 1) if px == 0 else (2, 0)
                class_taps.append([(ky, kx) for ky in kys for kx in kxs])
                pads_t.append(1 if py == 0 else 0)
                pads_l.append(1 if px == 0 else 0)
        packed = _pack_classes(w, cout, cin, class_taps, kh * kw, cout * kh * kw, kw, 1)
        _launch_classes(x, packed, b, out, cin, cout, H, Wd, pads_t, pads_l, relu=relu)
        ctx.save_for_backward(x, w, out if relu else None)
        ctx.has_bias = b is not None
        return out

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        dy = dy.contiguous()
        if y is not None:
            dy = torch.ops.aten.threshold_backward(dy, y, 0.0)
        cin, cout, kh, kw = w.shape
        B, H, Wd, _ = x.shape
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            # adjoint of the transposed convolution: the 4x4 / stride-2 / pad-1 convolution of dy with the same kernel
            packed = _pack(w, cin, cout, _all_taps(4, 4), cout * kh * kw, kh * kw, kw, 1)
            dx = torch.empty((B, H, Wd, cin), dtype=torch.float32, device=x.device)
            _launch(dy, packed, None, dx, cout, cin, H, Wd, 4, 4, 2, 1, 1)
        if ctx.needs_input_grad[1]:
            g = _wgrad(dy, x, 4, 4, 2, 1, 1)                                   # [(ky*4 + kx)*cout + co][ci]
            dw = g.view(4, 4, cout, cin).permute(3, 2, 0, 1).contiguous()
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.sum((0, 1, 2))
        return dx, dw, db, None


# the layer's ReLU in the kernel's epilogue (and its mask applied to dy from the saved output) instead of a separate
# elementwise pass; False: F.relu around the convolution (bit-identical results; the parity test that pins the activation
# pattern patches F.relu and runs this way)
FUSE_RELU = True


def conv2d(x, mod, stride=1, pad=None, relu=False):
    """NHWC fp32 -> NHWC fp32 through a torch.nn.Conv2d's parameters."""
    kh = mod.weight.shape[2]
    pad = (kh // 2) if pad is None else pad
    if relu and not FUSE_RELU:
        return F.relu(_Conv2d.apply(x, mod.weight, mod.bias, stride, pad, False))
    return _Conv2d.apply(x, mod.weight, mod.bias, stride, pad, relu)


def conv_transpose2d(x, mod, relu=False):
    if relu and not FUSE_RELU:
        return F.relu(_ConvT2d.apply(x, mod.weight, mod.bias, False))
    return _ConvT2d.apply(x, mod.weight, mod.bias, relu)


def res_block(x, blk):
    return conv2d(conv2d(x, blk.c1, relu=True), blk.c2) + x


def encode(pi, x, ac, c=None, z=None):
    """Policy.encode (policy.py:78-121) with the trunk's convolutions on the library; x [B,T,S,S,C] is already NHWC."""
    B, T = x.shape[:2]
    if pi.conditional:
        x = torch.cat([x, c], -1)
    x = x.reshape((B * T,) + tuple(x.shape[2:])).float().contiguous()
    ac = ac.reshape(B * T, -1).to(x.dtype)
    a = pi.a_enc(ac.unsqueeze(-1)).reshape(B * T, -1)
    a = F.relu(pi.a_concat_fc(a))
    add = conv2d(x, pi.x_enc, relu=True) + a[:, None, None, :]
    for conv in pi.add_enc:
        add = conv2d(add, conv, stride=2, pad=0, relu=True)
    for r in pi.encoder_res:
        add = res_block(add, r)
    out = F.relu(pi.flat_fc(add.flatten(1))).view(B, T, -1)                  # tl.flatten on NHWC
    if not pi.conditional:
        out = out + pi.z_enc(z)
    return out


def location_head(head, z_flat):
    """_LocationHead.forward (policy.py:238-295): [N, lstm] -> [N, L*L] logits."""
    x = z_flat.view(-1, 4, 4, head.c0)                                        # the reference's NHWC reshape
    x = conv_transpose2d(x, head.deconv0, relu=True)
    for r in head.res:
        x = res_block(x, r)
    for u in head.ups:
        x = conv_transpose2d(x, u, relu=True)
    return conv2d(x, head.out).flatten(1)
