"""``Discriminator`` -- host mirror of the reference's models/discriminator.py over K2.

``predict(images)`` (reference :157-163) returns ``sigmoid(D(norm(images)))`` for a batch of
final canvases, computed by libspiral_b200.so; it is the GAN reward the learner injects at
agent.py:336-338.  Parameters are plain float32 CUDA tensors in TensorFlow layouts (conv kernels
HWIO, dense kernel [C,1]) so that reference checkpoints map one to one by name.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np
import torch

from .. import _native as N
from . import conv_ops

PRECISIONS = {"fp32": 0, "bf16x3": 1, "bf16": 2, "bf16x6": 3}


class Discriminator(object):
    def __init__(self, args, step=None, image_shape=None, norm_fn=None, scope_name="global",
                 device=None, max_batch=None, seed=11):
        self.args = args
        self.step = step
        self.scope_name = scope_name
        self.image_shape = list(image_shape if image_shape is not None
                                else [args.screen_size, args.screen_size, args.color_channel])
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.conditional = bool(args.conditional)
        self.normalize = norm_fn is not None
        S, _, Cimg = self.image_shape
        self.in_channels = Cimg * (2 if self.conditional else 1)
        self.layer_num = int(math.log(S, 2))
        self.disc_dim = int(args.disc_dim)
        self.batch_norm = bool(args.disc_batch_norm)
        self.max_batch = int(max_batch or max(getattr(args, "num_envs", 64), args.disc_batch_size))
        self.precision = PRECISIONS[getattr(args, "disc_precision", "fp32")]
        gp = getattr(args, "disc_grad_precision", "auto")
        self.grad_precision = (3 if self.precision != 0 else 0) if gp == "auto" else PRECISIONS[gp]
        # WGAN-GP step: 1 = real / fake / interpolates as one grouped pass (wgan_gp_loss: a third of the launches, but the
        # penalty's gradient chain then runs over three times the rows, two thirds of them zeros -- measured slower at
        # batch 64: 27.2 vs 25.4 ms); default: three separate passes
        self.fused_passes = os.environ.get("SPIRAL_DISC_FUSED_PASSES", "0") != "0"

        self.params = self._init_params(seed)
        self.var_list = list(self.params.values())

        self._lib = N.lib()
        cfg = N.SpiralDiscCfg()
        cfg.abi_version = N.ABI_VERSION
        cfg.device = self.device.index if self.device.index is not None else torch.cuda.current_device()
        cfg.max_batch = self.max_batch
        cfg.in_channels = self.in_channels
        cfg.screen_size = S
        cfg.disc_dim = self.disc_dim
        cfg.batch_norm = int(self.batch_norm)
        cfg.normalize = int(self.normalize)
        cfg.precision = self.precision
        self._handle = C.c_void_p()
        N.check(self._lib.spiral_disc_create(C.byref(cfg), C.byref(self._handle)), "spiral_disc_create")
        nbytes = self._lib.spiral_disc_workspace_bytes(self._handle, self.max_batch)
        if nbytes < 0:
            raise RuntimeError("spiral_disc_workspace_bytes failed")
        self.workspace = torch.empty((int(nbytes),), dtype=torch.uint8, device=self.device)
        self.bn_stats = torch.zeros((self.layer_num, 2, self.disc_dim * 2 ** (self.layer_num - 1)),
                                    dtype=torch.float32, device=self.device)
        self.sync_weights()

    # ------------------------------------------------------------------
    def _init_params(self, seed):
        """he_normal conv kernels, zero biases, BN gamma 1 / beta 0, glorot_normal dense
        (reference :71, :87); names follow the TF variable scopes."""
        g = torch.Generator(device="cpu")
        g.manual_seed(seed)
        p = {}
        cin = self.in_channels
        for l in range(self.layer_num):
            cout = self.disc_dim * 2 ** l
            std = math.sqrt(2.0 / (25 * cin))
            p["conv%d/kernel" % l] = (torch.randn((5, 5, cin, cout), generator=g) * std)
            p["conv%d/bias" % l] = torch.zeros((cout,))
            if l > 0 and self.batch_norm:
                p[conv_ops.bn_name(l) + "/gamma"] = torch.ones((cout,))
                p[conv_ops.bn_name(l) + "/beta"] = torch.zeros((cout,))
            cin = cout
        p["dense/kernel"] = torch.randn((cin, 1), generator=g) * math.sqrt(2.0 / (cin + 1))
        p["dense/bias"] = torch.zeros((1,))
        return {k: v.to(self.device, torch.float32).contiguous() for k, v in p.items()}

    def load_numpy(self, weights):
        """weights: dict in the oracle's naming (conv%d/kernel, bn%d/gamma, dense/kernel ...)."""
        for k, v in weights.items():
            name = k
            if k.startswith("bn"):                    # oracle naming bn<l>/gamma -> TF's automatic layer names
                l, leaf = k[2:].split("/")
                name = conv_ops.bn_name(int(l)) + "/" + leaf
            self.params[name].copy_(torch.as_tensor(np.asarray(v)).to(self.params[name]))
        self.sync_weights()

    def sync_weights(self):
        """Hand the (possibly updated) parameters to the kernels; call after every optimiser step."""
        w = N.SpiralDiscWeights()
        for l in range(self.layer_num):
            w.conv_w[l] = self.params["conv%d/kernel" % l].data_ptr()
            w.conv_b[l] = self.params["conv%d/bias" % l].data_ptr()
            if l > 0 and self.batch_norm:
                w.bn_gamma[l] = self.params[conv_ops.bn_name(l) + "/gamma"].data_ptr()
                w.bn_beta[l] = self.params[conv_ops.bn_name(l) + "/beta"].data_ptr()
        w.dense_w = self.params["dense/kernel"].data_ptr()
        w.dense_b = self.params["dense/bias"].data_ptr()
        N.check(self._lib.spiral_disc_load_weights(self._handle, C.byref(w), N.ptr(self.workspace), N.stream_ptr()),
                "spiral_disc_load_weights")

    def __del__(self):
        try:
            if getattr(self, "_handle", None) is not None and self._handle.value:
                self._lib.spiral_disc_destroy(self._handle)
                self._handle = C.c_void_p()
        except Exception:
            pass

    # ------------------------------------------------------------------
    def forward(self, images, logits_out=None, probs_out=None, global_bn=None):
        """images f32[N,S,S,C] (0..255 when norm_fn is set) -> (probs[N], logits[N]).

        ``global_bn``: score with GLOBAL-batch batch-norm statistics when the batch is sharded over ranks (the reference
        normalises with the statistics of whatever batch it is fed, discriminator.py:76-79, so a sharded batch equals the
        gathered one only if every layer's per-channel sums are all-reduced): the forward runs in stages
        (``spiral_disc_forward_staged``) and ``global_bn(sums_f64 [C,2], count_f64 [1])`` is called between them to reduce
        both tensors in place over the ranks (``distributed.allreduce_bn_sums_``).  None / one rank: one call."""
        x = images
        if self.conditional:
            x = torch.cat([x, x], -1)                 # real <- concat[real, real] (reference :36-38)
        x = x.to(device=self.device, dtype=torch.float32).contiguous()
        n = x.shape[0]
        if tuple(x.shape[1:]) != (self.image_shape[0], self.image_shape[1], self.in_channels):
            raise ValueError("images must be [N,%d,%d,%d]" % (self.image_shape[0], self.image_shape[1],
                                                              self.in_channels // (2 if self.conditional else 1)))
        logits = logits_out if logits_out is not None else torch.empty((n,), dtype=torch.float32, device=self.device)
        probs = probs_out if probs_out is not None else torch.empty((n,), dtype=torch.float32, device=self.device)
        if global_bn is not None and self.batch_norm:
            if getattr(self, "_bn_sums", None) is None:
                self._bn_sums = torch.zeros((self.disc_dim * 2 ** (self.layer_num - 1), 2), dtype=torch.float64, device=self.device)
                self._bn_count = torch.zeros((1,), dtype=torch.float64, device=self.device)
            count = 0.0
            for stage in range(self.layer_num):
                N.check(self._lib.spiral_disc_forward_staged(self._handle, N.ptr(x), n, stage, C.c_double(count),
                                                             N.ptr(self._bn_sums), N.ptr(logits), N.ptr(probs),
                                                             N.ptr(self.bn_stats), N.ptr(self.workspace), N.stream_ptr()),
                        "spiral_disc_forward_staged")
                if stage + 1 < self.layer_num:
                    ho = self.image_shape[0] >> (stage + 2)                 # output side of layer stage + 1
                    cout = self.disc_dim * 2 ** (stage + 1)
                    self._bn_count.fill_(float(n * ho * ho))
                    global_bn(self._bn_sums[:cout], self._bn_count)
                    count = float(self._bn_count.item())
            return probs, logits
        N.check(self._lib.spiral_disc_forward(self._handle, N.ptr(x), n, N.ptr(logits), N.ptr(probs),
                                              N.ptr(self.bn_stats), N.ptr(self.workspace), N.stream_ptr()),
                "spiral_disc_forward")
        return probs, logits

    def predict(self, images, global_bn=None):
        """reference :157-163."""
        return self.forward(images, global_bn=global_bn)[0]

    # ------------------------------------------------------------------
    def logits_fn(self, x_normed, groups=1, scope=None):
        """Differentiable D on already-normalised NHWC input; convolutions and all their gradients
        run on libspiral_b200.so (models/conv_ops.py).  ``groups``, ``scope``: see conv_ops.disc_forward_native / PackScope."""
        return conv_ops.disc_forward_native(self.params, x_normed, self.layer_num, self.batch_norm,
                                            normalize=False, precision=self.grad_precision, groups=groups, scope=scope)

    def wgan_gp_loss(self, fake, real, alpha, wgan_lambda=None):
        """reference :97-122 (build_optim): critic + lambda * gradient penalty, the penalty on the
        gradient of the sigmoid OUTPUT w.r.t. the interpolates.  fake / real f32[N,S,S,C] in the
        env's 0..255 units (norm_fn applied here when set), alpha f32[N,1] ~ U[0,1).
        Returns a dict of differentiable scalars: d_loss, critic_loss, gradient_penalty, g_loss."""
        lam = float(self.args.wgan_lambda if wgan_lambda is None else wgan_lambda)
        f = fake.to(self.device, torch.float32)
        r = real.to(self.device, torch.float32)
        if self.normalize:
            f, r = (f - 127.5) / 127.5, (r - 127.5) / 127.5
        if self.conditional:                       # reference :36-38
            f = torch.cat([f, r], -1)
            r = torch.cat([r, r], -1)
        inter = (r.flatten(1) + alpha * (f.flatten(1) - r.flatten(1))).detach().requires_grad_(True)
        scope = conv_ops.PackScope()               # each kernel's tensor-core planes once for the three passes and their gradients
        if self.fused_passes:
            # the three shared-weight passes as ONE pass over [real; fake; interpolates], each third with its own batch
            # statistics: a third of the launches.  The penalty's gradient is taken w.r.t. the interpolates only; the
            # first two thirds of every dgrad operand are exact zeros (no statistic crosses a group), so the value is
            # that of the separate passes.
            n = r.shape[0]
            probs, logits = self.logits_fn(torch.cat([r, f, inter.view_as(f)], 0), groups=3, scope=scope)
            real_logits, fake_logits, probs = logits[:n], logits[n:2 * n], probs[2 * n:]
        else:
            _, real_logits = self.logits_fn(r, scope=scope)
            _, fake_logits = self.logits_fn(f, scope=scope)
            probs, _ = self.logits_fn(inter.view_as(f), scope=scope)
        g_loss = -fake_logits.mean()
        critic = fake_logits.mean() - real_logits.mean()
        grads = torch.autograd.grad(probs.sum(), inter, create_graph=True)[0]
        slopes = torch.sqrt(1e-8 + (grads ** 2).sum(1))
        gp = ((slopes - 1.0) ** 2).mean()
        return dict(d_loss=critic + lam * gp, critic_loss=critic, gradient_penalty=gp, g_loss=g_loss)
