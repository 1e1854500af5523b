"""Differentiable 5x5 / stride-2 / 'SAME' convolution on libspiral_b200.so (K2b).

The reference trains the discriminator with ``tf.gradients`` of ``critic + lambda * GP``
(models/discriminator.py:97-136): three shared-weight passes and, for the penalty on
``d sigmoid(D(x^)) / d x^`` (:116-120), a double backward through every conv layer.  The conv
layer's three bilinear maps -- forward, dgrad, wgrad -- are closed under differentiation, so they
are wired here as mutually recursive ``torch.autograd.Function``s: any order of derivative of the
discriminator runs on the library's kernels (tcgen05 bf16x3 where the channel counts allow, fp32
SIMT otherwise); batch-norm, leaky-ReLU and the dense head are bandwidth-bound elementwise /
reduction glue and stay PyTorch ops.

Tensors are NHWC float32 CUDA tensors, kernels HWIO, as in TensorFlow.
"""
from __future__ import annotations

import torch

from .. import _native as N

_WS = {}


def bn_name(l):
    """TensorFlow's automatic name of the unnamed tl.batch_normalization after conv<l> (discriminator.py:76-79; layer 0
    has none): the first one in the scope is 'batch_normalization', then 'batch_normalization_1', ..."""
    return "batch_normalization" if l == 1 else "batch_normalization_%d" % (l - 1)



def _workspace(dev, nbytes):
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty((max(int(nbytes), 1 << 20),), dtype=torch.uint8, device=dev)
        _WS[key] = ws
    return ws


def _check(x, name):
    if not (x.is_cuda and x.dtype == torch.float32):
        raise ValueError("%s must be a float32 CUDA tensor (there is no CPU fallback)" % name)
    return x.contiguous()


def _call(fn_name, a, b, out, B, H, Cin, Cout, precision):
    lib = N.lib()
    nbytes = lib.spiral_conv_workspace_bytes(B, H, Cin, Cout, precision)
    if nbytes < 0:
        N.check(-1, "spiral_conv_workspace_bytes")
    ws = _workspace(a.device, nbytes)
    N.check(getattr(lib, fn_name)(N.ptr(a), N.ptr(b), N.ptr(out), B, H, Cin, Cout, precision, N.ptr(ws), N.stream_ptr()),
            fn_name)
    return out


class PackScope:
    """Operand planes of the kernels used while ONE loss is evaluated and differentiated (Discriminator.wgan_gp_loss opens a
    scope per call): the three shared-weight passes and the convolutions of the penalty's double backward read the same
    kernels, so each is split into tensor-core planes once per scope (spiral_conv_pack) instead of once per convolution.
    Keyed by the kernel's storage; the weights must not change while a scope is in use (they do not: the optimiser runs
    after backward).  Scopes are not reused across steps -- under a CUDA graph the pack kernels are part of the graph."""

    def __init__(self):
        self.planes = {}

    def get(self, w, op, B, H, Cin, Cout, precision):
        key = (w.data_ptr(), op, Cin, Cout, precision)
        hit = self.planes.get(key)
        if hit is not None:
            return hit[0]
        lib = N.lib()
        nbytes = int(lib.spiral_conv_packed_bytes(op, B, H, Cin, Cout, precision))
        if nbytes <= 0:
            self.planes[key] = (None, w)
            return None
        buf = torch.empty((nbytes,), dtype=torch.uint8, device=w.device)
        N.check(lib.spiral_conv_pack(N.ptr(w), N.ptr(buf), op, B, H, Cin, Cout, precision, N.stream_ptr()), "spiral_conv_pack")
        self.planes[key] = (buf, w)                          # holds w: its storage cannot be recycled under the key
        return buf


def _call_packed(fn_name, a, packed, out, B, H, Cin, Cout, precision):
    lib = N.lib()
    nbytes = lib.spiral_conv_workspace_bytes(B, H, Cin, Cout, precision)
    if nbytes < 0:
        N.check(-1, "spiral_conv_workspace_bytes")
    ws = _workspace(a.device, nbytes)
    N.check(getattr(lib, fn_name)(N.ptr(a), N.ptr(packed), N.ptr(out), B, H, Cin, Cout, precision, N.ptr(ws), N.stream_ptr()),
            fn_name)
    return out


def conv_forward(x, w, precision=1, scope=None):
    """y[B,H/2,H/2,Cout] = conv(x[B,H,H,Cin], w[5,5,Cin,Cout]); no autograd."""
    x, w = _check(x, "x"), _check(w, "w")
    B, H, _, Cin = x.shape
    Cout = w.shape[3]
    y = torch.empty((B, H // 2, H // 2, Cout), dtype=torch.float32, device=x.device)
    packed = scope.get(w, 0, B, H, Cin, Cout, precision) if scope is not None else None
    if packed is not None:
        return _call_packed("spiral_conv_forward_packed", x, packed, y, B, H, Cin, Cout, precision)
    return _call("spiral_conv_forward", x, w, y, B, H, Cin, Cout, precision)


def conv_dgrad(dy, w, precision=1, scope=None):
    """dx[B,H,H,Cin] for dy[B,H/2,H/2,Cout]; no autograd."""
    dy, w = _check(dy, "dy"), _check(w, "w")
    B, Ho, _, Cout = dy.shape
    Cin = w.shape[2]
    dx = torch.empty((B, 2 * Ho, 2 * Ho, Cin), dtype=torch.float32, device=dy.device)
    packed = scope.get(w, 1, B, 2 * Ho, Cin, Cout, precision) if scope is not None else None
    if packed is not None:
        return _call_packed("spiral_conv_dgrad_packed", dy, packed, dx, B, 2 * Ho, Cin, Cout, precision)
    return _call("spiral_conv_dgrad", dy, w, dx, B, 2 * Ho, Cin, Cout, precision)


def conv_wgrad(x, dy, precision=1):
    """dw[5,5,Cin,Cout]; no autograd."""
    x, dy = _check(x, "x"), _check(dy, "dy")
    B, H, _, Cin = x.shape
    Cout = dy.shape[3]
    dw = torch.empty((5, 5, Cin, Cout), dtype=torch.float32, device=x.device)
    return _call("spiral_conv_wgrad", x, dy, dw, B, H, Cin, Cout, precision)


# `scope` (a PackScope or None) rides along as a non-tensor argument; only convolutions whose kernel operand is a network
# PARAMETER use it (the kernel-shaped tensors of the double backward, ggw, are different every time)
class _Conv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, precision, scope):
        ctx.save_for_backward(x, w)
        ctx.precision, ctx.scope = precision, scope
        return conv_forward(x, w, precision, scope)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gx = _Dgrad.apply(gy, w, ctx.precision, ctx.scope) if ctx.needs_input_grad[0] else None
        gw = _Wgrad.apply(x, gy, ctx.precision, ctx.scope) if ctx.needs_input_grad[1] else None
        return gx, gw, None, None


class _Dgrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gy, w, precision, scope):
        ctx.save_for_backward(gy, w)
        ctx.precision, ctx.scope = precision, scope
        return conv_dgrad(gy, w, precision, scope)

    @staticmethod
    def backward(ctx, ggx):
        gy, w = ctx.saved_tensors
        g_gy = _Conv.apply(ggx, w, ctx.precision, ctx.scope) if ctx.needs_input_grad[0] else None
        g_w = _Wgrad.apply(ggx, gy, ctx.precision, ctx.scope) if ctx.needs_input_grad[1] else None
        return g_gy, g_w, None, None


class _Wgrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gy, precision, scope):
        ctx.save_for_backward(x, gy)
        ctx.precision, ctx.scope = precision, scope
        return conv_wgrad(x, gy, precision)

    @staticmethod
    def backward(ctx, ggw):
        x, gy = ctx.saved_tensors
        g_x = _Dgrad.apply(gy, ggw, ctx.precision, None) if ctx.needs_input_grad[0] else None
        g_gy = _Conv.apply(x, ggw, ctx.precision, None) if ctx.needs_input_grad[1] else None
        return g_x, g_gy, None, None


def conv5x5s2(x, w, precision=1, scope=None):
    """Differentiable (to any order) tf.layers.conv2d(x, Cout, 5, strides=2, padding='same',
    use_bias=False) on NHWC ``x`` with HWIO ``w``."""
    return _Conv.apply(x, w, int(precision), scope)


def disc_forward_native(params, x, layer_num, batch_norm, normalize=True, precision=1, groups=1, scope=None):
    """Differentiable D(x) (models/discriminator.py:51-95) whose convolutions and all their
    gradients run on libspiral_b200.so.  x f32[N,S,S,C] NHWC (0..255 when ``normalize``).
    TF semantics: SAME padding 1/2, BN epsilon 1e-3 with the biased batch variance (training mode),
    leaky-ReLU 0.2.  Returns (probs[N], logits[N]).

    ``groups`` > 1: x holds that many equal batches back to back, each normalised with ITS OWN batch statistics, i.e.
    the result of ``groups`` separate calls of the shared-weight network (what build_optim does with real, fake and the
    interpolates, :99-117) from one launch per layer instead of one per call."""
    if normalize:
        x = (x - 127.5) / 127.5
    G = int(groups)
    if x.shape[0] % G:
        raise ValueError("batch %d does not split into %d groups" % (x.shape[0], G))
    for l in range(layer_num):
        x = conv5x5s2(x, params["conv%d/kernel" % l], precision, scope) + params["conv%d/bias" % l]
        if l > 0 and batch_norm:
            shp = x.shape
            xg = x.view(G, shp[0] // G, shp[1], shp[2], shp[3])
            mean = xg.mean(dim=(1, 2, 3), keepdim=True)
            xc = xg - mean
            var = (xc * xc).mean(dim=(1, 2, 3), keepdim=True)
            x = (xc * torch.rsqrt(var + 1e-3)).view(shp) * params[bn_name(l) + "/gamma"] + params[bn_name(l) + "/beta"]
        x = torch.where(x > 0, x, 0.2 * x)
    logits = (x.flatten(1) @ params["dense/kernel"] + params["dense/bias"]).reshape(-1)
    return torch.sigmoid(logits), logits
