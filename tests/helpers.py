"""Shared helpers for the test-suite."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ENV_OF = {"simple_mnist": ["jump", "control", "end"],
          "mnist": ["jump", "pressure", "control", "end", "size"],
          "color_mnist": ["control", "end", "color", "jump", "pressure", "size"]}


def make_args(env="simple_mnist", L=32, C=1, T=5, conditional=False, jump=True, curve=True, extra=(), math="det"):
    import spiral_b200 as sp
    argv = ["--env", env, "--location_size", str(L), "--color_channel", str(C), "--episode_length", str(T),
            "--loss", "l2" if conditional else "gan", "--jump", str(jump), "--curve", str(curve),
            "--raster_math", math] + list(extra)
    return sp.get_args(argv)


def random_actions(env, B, T, seed=0, jump_prob=None):
    rng = np.random.RandomState(seed)
    a = np.stack([rng.randint(n, size=(B, T)) for n in env.head_sizes], -1).astype(np.int32)
    if jump_prob is not None and "jump" in env.ac_idx:
        a[..., env.ac_idx["jump"]] = rng.rand(B, T) < jump_prob
    return a


def oracle_for(env, B, math_mode=None):
    """The CPU oracle in the arithmetic mode the env's kernels reproduce bit for bit (det -> PO_MATH_DET, fast -> PO_MATH_FAST)."""
    from oracle import paint_oracle as po
    if math_mode is None:
        math_mode = po.PO_MATH_FAST if getattr(env, "raster_math", "det") == "fast" else po.PO_MATH_DET
    spec = po.EnvSpec(action_names=list(env.acs), location_size=env.location_size,
                      n_color=len(env.colors), colorize=env.colorize, math_mode=math_mode)
    return po.OracleBatchEnv(spec, B, channels=env.channels)


import math

import torch


class TFAdam(torch.optim.Optimizer):
    """Plain-PyTorch reference of K5 (tests only).  tf.train.AdamOptimizer: lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);  p -= lr_t * m / (sqrt(v) + eps)."""

    def __init__(self, params, lr, beta1=0.9, beta2=0.999, eps=1e-8):
        super().__init__(params, dict(lr=lr, beta1=beta1, beta2=beta2, eps=eps))

    @torch.no_grad()
    def step(self):
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["t"] = 0
                    st["m"] = torch.zeros_like(p)
                    st["v"] = torch.zeros_like(p)
                st["t"] += 1
                t, b1, b2 = st["t"], g["beta1"], g["beta2"]
                st["m"].mul_(b1).add_(p.grad, alpha=1 - b1)
                st["v"].mul_(b2).addcmul_(p.grad, p.grad, value=1 - b2)
                lr_t = g["lr"] * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
                p.addcdiv_(st["m"], st["v"].sqrt().add_(g["eps"]), value=-lr_t)


