"""``NativeActor`` -- policy.act with the trunk's convolutions on libspiral_b200.so (bf16 tensor cores).

The acting path of the reference (models/policy.py:196-215: one ``sess.run`` per env step on every actor) spends
its time in ~45 small convolutions (<= 32 channels, <= 64x64 maps).  Once rendering, reward and loss are native
they are >90 % of a rollout (tools/policy_act_probe.py).  This module runs them through ``spiral_pconv_forward``
(csrc/policy_conv.cu: NHWC bf16, implicit GEMM gathered from a shared-memory tile, fused bias / residual / ReLU
epilogues; transposed convolutions as four parity-class 2x2 convolutions) and keeps the rest of ``Policy`` --
the dense layers, the LSTM cell, the autoregressive head wiring, K4 sampling -- exactly as in ``models/policy.py``.

It reads the SAME parameters as the PyTorch module (call ``sync()`` after every learner update); the learner keeps
using the fp32 PyTorch module (it needs gradients and fp32 parity).  bf16 operands: logits agree with the fp32
module to ~1e-2, which only perturbs the actors' behaviour policy.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn.functional as F

from .. import _native as N
from .policy import categorical_sample


class _Conv(object):
    """One launch of spiral_pconv_forward; the kernel packed by spiral_pconv_pack from a torch Conv2d weight (or any
    [Cout, Cin, KH, KW]-indexed VIEW of one: the packer takes element strides and a tap list, so a transposed convolution's
    parity class is the permuted view of its weight plus its four taps -- no copies, one launch per layer)."""

    def __init__(self, weight_oihw, bias, stride, pad, relu, scatter=None, taps=None):
        self.Cout, self.Cin = int(weight_oihw.shape[0]), int(weight_oihw.shape[1])
        if taps is None:
            self.KH, self.KW = int(weight_oihw.shape[2]), int(weight_oihw.shape[3])
            taps = [(ky, kx) for ky in range(self.KH) for kx in range(self.KW)]
        else:
            self.KH = self.KW = 2                                                    # a 2x2 parity class of a 4x4 kernel
            assert len(taps) == 4
        self._tap_ky = (C.c_int32 * len(taps))(*[t[0] for t in taps])
        self._tap_kx = (C.c_int32 * len(taps))(*[t[1] for t in taps])
        self._n_taps = len(taps)
        self.stride, self.pad_t, self.pad_l, self.relu = stride, pad[0], pad[1], int(relu)
        K = self._n_taps * self.Cin
        cp = (self.Cout + 7) // 8 * 8
        dev = weight_oihw.device
        self.wpk = torch.empty((cp, (K + 15) // 16 * 16 + 8), dtype=torch.bfloat16, device=dev)
        self.bias = torch.zeros((cp,), dtype=torch.float32, device=dev)
        self.scatter = scatter                                                       # (sy, sx, oy, ox) for transposed-conv classes
        self.repack(weight_oihw, bias)

    def repack(self, weight_oihw, bias):
        """Refresh the packed kernel IN PLACE (a captured CUDA graph keeps pointing at these buffers)."""
        w = weight_oihw.detach()
        if w.dtype != torch.float32:
            w = w.float()
        st = w.stride()
        N.check(N.lib().spiral_pconv_pack(N.ptr(w), N.ptr(self.wpk), self.Cout, self.Cin, self._n_taps, st[0], st[1], st[2], st[3],
                                          self._tap_ky, self._tap_kx, int(self.Cout == 32), N.stream_ptr()), "spiral_pconv_pack")
        self.bias[:self.Cout].copy_(bias.detach())

    def out_hw(self, Hin, Win, same):
        if same:
            return (Hin + self.stride - 1) // self.stride, (Win + self.stride - 1) // self.stride
        return (Hin - self.KH) // self.stride + 1, (Win - self.KW) // self.stride + 1

    def __call__(self, x, out, Hout, Wout, residual=None, post_add=None, out_map=None, out_f32=False):
        B, Hin, Win, _ = x.shape
        oH, oW = out_map if out_map is not None else (Hout, Wout)
        sy, sx, oy, ox = self.scatter if self.scatter is not None else (1, 1, 0, 0)
        N.check(N.lib().spiral_pconv_forward(N.ptr(x), N.ptr(self.wpk), N.ptr(self.bias), N.ptr(residual), N.ptr(post_add),
                                             N.ptr(out), B, Hin, Win, self.Cin, Hout, Wout, self.Cout, self.KH, self.KW,
                                             self.stride, self.pad_t, self.pad_l, self.relu, oH, oW, sy, sx, oy, ox,
                                             int(out_f32), N.stream_ptr()), "spiral_pconv_forward")
        return out


class _Deconv(object):
    """ConvTranspose2d(k=4, s=2, p=1) == TF conv2d_transpose(4, strides 2, 'same') as four parity-class 2x2
    convolutions: out[2a+py, 2b+px] = sum over the two kernel rows kk with the parity of py+1 (kk = 1,3 / 0,2) of
    x[a + (py + 1 - kk) / 2] * w[kk] (same along x)."""

    def __init__(self, weight_iohw, bias, relu):
        self.classes = []
        w = weight_iohw.detach().permute(1, 0, 2, 3)                                 # [Cout, Cin, 4, 4] view
        for py in (0, 1):
            for px in (0, 1):
                kys = (3, 1) if py == 0 else (2, 0)          # tap row 0, 1 of the 2x2 window (input rows a-1,a / a,a+1)
                kxs = (3, 1) if px == 0 else (2, 0)
                conv = _Conv(w, bias, 1, (1 if py == 0 else 0, 1 if px == 0 else 0), relu, scatter=(2, 2, py, px),
                             taps=[(ky, kx) for ky in kys for kx in kxs])
                self.classes.append(conv)

    def repack(self, weight_iohw, bias):
        w = weight_iohw.detach().permute(1, 0, 2, 3)
        for conv in self.classes:
            conv.repack(w, bias)

    def __call__(self, x, out):
        B, Hin, Win, _ = x.shape
        for conv in self.classes:
            conv(x, out, Hin, Win, out_map=(2 * Hin, 2 * Win))
        return out


class NativeActor(object):
    def __init__(self, policy):
        self.pi = policy
        self.device = next(policy.parameters()).device
        self._bufs = {}
        self.sync()

    # ------------------------------------------------------------------
    def sync(self):
        """(Re)pack the convolution kernels from the PyTorch module's current parameters; after the first call the
        packed buffers are refreshed in place, so a captured CUDA graph of the rollout sees the new weights."""
        pi = self.pi
        if getattr(self, "x_enc", None) is not None:
            self.x_enc.repack(pi.x_enc.weight, pi.x_enc.bias)
            for layer, c in zip(self.add_enc, pi.add_enc):
                layer.repack(c.weight, c.bias)
            for (l1, l2), r in zip(self.enc_res, pi.encoder_res):
                l1.repack(r.c1.weight, r.c1.bias)
                l2.repack(r.c2.weight, r.c2.bias)
            for name, L in self.loc.items():
                head = pi.heads[name]
                L["deconv0"].repack(head.deconv0.weight, head.deconv0.bias)
                for (l1, l2), r in zip(L["res"], head.res):
                    l1.repack(r.c1.weight, r.c1.bias)
                    l2.repack(r.c2.weight, r.c2.bias)
                for u, m in zip(L["ups"], head.ups):
                    u.repack(m.weight, m.bias)
                L["out"].repack(head.out.weight, head.out.bias)
            for tag, (wpk, bias, blocks) in self._stacks.items():
                if bias is None:                            # fused last stage of a location head: the four class kernels
                    wpk.copy_(torch.cat([c.wpk.reshape(-1) for c in blocks.classes]))
                    continue
                wpk.copy_(torch.cat([c.wpk.reshape(-1) for blk in blocks for c in blk]))
                bias.copy_(torch.cat([c.bias[:32] for blk in blocks for c in blk]))
            return
        self._stacks = {}
        self.x_enc = _Conv(pi.x_enc.weight, pi.x_enc.bias, 1, (2, 2), True)
        self.add_enc = [_Conv(c.weight, c.bias, 2, (0, 0), True) for c in pi.add_enc]
        self.enc_res = [(_Conv(r.c1.weight, r.c1.bias, 1, (1, 1), True), _Conv(r.c2.weight, r.c2.bias, 1, (1, 1), False))
                        for r in pi.encoder_res]
        self.loc = {}
        for name in pi.acs:
            head = pi.heads[name]
            if hasattr(head, "deconv0"):
                self.loc[name] = dict(
                    deconv0=_Deconv(head.deconv0.weight, head.deconv0.bias, True),
                    res=[(_Conv(r.c1.weight, r.c1.bias, 1, (1, 1), True), _Conv(r.c2.weight, r.c2.bias, 1, (1, 1), False))
                         for r in head.res],
                    ups=[_Deconv(u.weight, u.bias, True) for u in head.ups],
                    out=_Conv(head.out.weight, head.out.bias, 1, (1, 1), False), c0=head.c0)

    def _buf(self, key, shape, dtype=torch.bfloat16):
        b = self._bufs.get(key)
        if b is None or tuple(b.shape) != tuple(shape) or b.dtype != dtype:
            b = torch.empty(shape, dtype=dtype, device=self.device)
            self._bufs[key] = b
        return b

    def _res_stack(self, x, blocks, tag):
        B, H, W, C = x.shape
        if H == W and H in (6, 8) and C == 32 and len(blocks) > 0:
            # fused: the whole stack in one launch, frames resident in shared memory (spiral_pres_stack)
            pk = self._stacks.get(tag)
            if pk is None:
                wpk = torch.cat([c.wpk.reshape(-1) for blk in blocks for c in blk]).contiguous()
                bias = torch.cat([c.bias[:32] for blk in blocks for c in blk]).contiguous()
                pk = self._stacks[tag] = (wpk, bias, blocks)
            out = self._buf((tag, "fused", H), (B, H, W, C))
            N.check(N.lib().spiral_pres_stack(N.ptr(x), N.ptr(pk[0]), N.ptr(pk[1]), N.ptr(out), B, H, 2 * len(blocks),
                                              N.stream_ptr()), "spiral_pres_stack")
            return out
        for i, (c1, c2) in enumerate(blocks):
            t = c1(x, self._buf((tag, "t", H), (B, H, W, C)), H, W)
            x = c2(t, self._buf((tag, i % 2, H), (B, H, W, C)), H, W, residual=x)
        return x

    # ------------------------------------------------------------------
    @torch.no_grad()
    def encode(self, ob, ac, condition=None, z=None):
        """Policy.encode for T = 1: ob f32 [B,S,S,C] -> [B, lstm_size]."""
        pi = self.pi
        B, S = ob.shape[0], ob.shape[1]
        x = ob if condition is None else torch.cat([ob, condition], -1)
        x = x.to(torch.bfloat16).contiguous()                                      # NHWC already
        a = pi.a_enc(ac.float().unsqueeze(-1)).reshape(B, -1)
        a = F.relu(pi.a_concat_fc(a)).float().contiguous()                         # [B,32], added after the ReLU (policy.py:101)
        h = self.x_enc(x, self._buf("e0", (B, S, S, 32)), S, S, post_add=a)
        H = S
        for i, conv in enumerate(self.add_enc):
            Ho, Wo = conv.out_hw(H, H, same=False)
            h = conv(h, self._buf(("e", i), (B, Ho, Wo, 32)), Ho, Wo)
            H = Ho
        h = self._res_stack(h, self.enc_res, "er")
        flat = h.reshape(B, -1).float()                                            # NHWC flatten (tl.flatten)
        out = F.relu(pi.flat_fc(flat))
        if not pi.conditional:
            out = out + pi.z_enc(z)
        return out

    @torch.no_grad()
    def location_logits(self, name, z_flat):
        L = self.loc[name]
        B = z_flat.shape[0]
        x = z_flat.view(B, 4, 4, L["c0"]).to(torch.bfloat16).contiguous()          # NHWC reshape of the reference
        h = L["deconv0"](x, self._buf((name, "d0"), (B, 8, 8, 32)))
        h = self._res_stack(h, L["res"], name + "r")
        H = 8
        ups = L["ups"]
        fuse_last = len(ups) > 0                         # last up-convolution + the 3x3 -> 1 logits convolution in one kernel
        for i, up in enumerate(ups[:-1] if fuse_last else ups):
            h = up(h, self._buf((name, "u", i), (B, 2 * H, 2 * H, 32)))
            H *= 2
        if fuse_last:
            pk = self._stacks.get((name, "last"))
            if pk is None:
                w_up = torch.cat([c.wpk.reshape(-1) for c in ups[-1].classes]).contiguous()
                pk = self._stacks[(name, "last")] = (w_up, None, ups[-1])
            logits = self._buf((name, "o"), (B, 2 * H, 2 * H), torch.float32)
            N.check(N.lib().spiral_phead_last(N.ptr(h), N.ptr(pk[0]), N.ptr(ups[-1].classes[0].bias), N.ptr(L["out"].wpk),
                                              N.ptr(L["out"].bias), N.ptr(logits), B, H, N.stream_ptr()), "spiral_phead_last")
            return logits.view(B, 4 * H * H)
        logits = L["out"](h, self._buf((name, "o"), (B, H, H, 1), torch.float32), H, H, out_f32=True)
        return logits.view(B, H * H)

    @torch.no_grad()
    def act(self, ob, ac, c, h, condition=None, z=None, generator=None):
        """Same contract as Policy.act (models/policy.py:196-215)."""
        pi = self.pi
        enc = self.encode(ob, ac, condition, z)
        hid, (nc, nh) = pi.lstm(enc, (c, h))
        zz = F.relu(hid)
        actions = []
        for idx, name in enumerate(pi.acs):
            logit = self.location_logits(name, zz) if name in self.loc else pi.heads[name](zz)
            s = categorical_sample(logit, generator, pi).long()
            actions.append(s)
            if idx < len(pi.acs) - 1:
                emb = pi.sample_mlp[name](s.view(-1, 1).to(zz.dtype))
                zz = F.relu(pi.concat_z_fc[name](torch.cat([zz, emb], -1)))
        vf = pi.value(hid)[:, 0]
        return torch.stack(actions, 1).to(torch.int32), vf, nc, nh
