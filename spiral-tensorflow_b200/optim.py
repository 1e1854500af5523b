"""``FusedTFAdam`` -- the learner's update step on libspiral_b200.so (K5).

Mirror of what the reference builds with ``tf.clip_by_global_norm(grads, 40)`` +
``tf.train.AdamOptimizer(lr).apply_gradients`` (agent.py:231-239) and
``AdamOptimizer(disc_lr, beta1=0.5, beta2=0.9)`` (models/discriminator.py:124-126): two multi-tensor
kernels per step (global norm, then clip + Adam with TensorFlow's epsilon placement), no host
synchronisation.  ``grad_scale`` folds in the explicit 1/world of the sum all-reduce.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _native as N


class FusedTFAdam(object):
    def __init__(self, params, lr, beta1=0.9, beta2=0.999, eps=1e-8):
        self.params = [p for p in params]
        for p in self.params:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                raise ValueError("FusedTFAdam needs contiguous float32 CUDA parameters (there is no CPU fallback)")
        self.lr, self.beta1, self.beta2, self.eps = float(lr), float(beta1), float(beta2), float(eps)
        self.t = 0
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        dev = self.params[0].device
        pieces = sum(max(1, -(-p.numel() // (1 << 20))) for p in self.params)      # include/spiral_b200.h: 32 partials per 2^20-element piece
        self._scratch = torch.empty((32 * pieces,), dtype=torch.float64, device=dev)
        self._norm = torch.zeros((1,), dtype=torch.float32, device=dev)
        self._lr_t = torch.zeros((1,), dtype=torch.float32, device=dev)     # this step's lr_t, device-resident (CUDA graphs)

    def advance(self):
        """Host half of a step: t += 1 and this step's ``lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t)`` into device memory.
        ``step()`` calls it itself; a step that is replayed from a CUDA graph calls it before the replay and captures
        ``step(advance=False)``, so the graph reads each replay's own value."""
        self.t += 1
        self._lr_t.fill_(self.lr * math.sqrt(1.0 - self.beta2 ** self.t) / (1.0 - self.beta1 ** self.t))

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    @torch.no_grad()
    def step(self, clip_norm=0.0, grad_scale=1.0, advance=True):
        """One update.  Returns the global norm of the (scaled) gradient as a device scalar
        (``model/grad_global_norm`` in the reference's summaries, agent.py:213)."""
        lib = N.lib()
        idx = [i for i, p in enumerate(self.params) if p.grad is not None]
        if not idx:
            return self._norm.zero_()[0]
        n = len(idx)
        grads = [self.params[i].grad.contiguous() for i in idx]
        vp = C.c_void_p * n
        pp = vp(*[self.params[i].data_ptr() for i in idx])
        gp = vp(*[g.data_ptr() for g in grads])
        mp = vp(*[self.m[i].data_ptr() for i in idx])
        vv = vp(*[self.v[i].data_ptr() for i in idx])
        sizes = (C.c_int64 * n)(*[self.params[i].numel() for i in idx])
        st = N.stream_ptr()
        N.check(lib.spiral_grad_global_norm(gp, sizes, n, N.ptr(self._scratch), N.ptr(self._norm), st),
                "spiral_grad_global_norm")
        if advance:
            self.advance()
        N.check(lib.spiral_adam_step_lr_dev(pp, gp, mp, vv, sizes, n, N.ptr(self._lr_t), self.beta1, self.beta2, self.eps,
                                            float(grad_scale), N.ptr(self._norm), float(clip_norm), st), "spiral_adam_step_lr_dev")
        return self._norm[0] * float(grad_scale)
