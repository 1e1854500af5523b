"""Runs the rasteriser (both arithmetic modes, fused and two-kernel steps, small and large batches, a large-radius brush)
and the learner's convolution kernels through libspiral_b200_checked.so -- the build with the kernels' own index checks
(csrc/common.h SP_CHECK; compute-sanitizer is closed on the B200 pool) -- compares the canvases with the CPU oracle and
prints the number of index violations the device counted.

    SPIRAL_B200_LIB=spiral-tensorflow_b200/csrc/libspiral_b200_checked.so python tools/checked_run.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import spiral_b200 as sp  # noqa: E402
from helpers import make_args, oracle_for, random_actions  # noqa: E402
from importlib import import_module  # noqa: E402

N = sp._native
lib = N.lib()
assert "checked" in N.LIB_PATH, "set SPIRAL_B200_LIB to the checked library"
assert lib.spiral_debug_bounds_violations() == 0


def render(envname, L, C, B, T, math, fused_max, check_oracle=True, extra=()):
    os.environ["SPIRAL_FUSED_STEP_MAX"] = fused_max
    env = sp.create_env(make_args(envname, L=L, C=C, T=T, math=math, extra=list(extra)), num_envs=B)
    acts = random_actions(env, B, T, seed=3, jump_prob=0.2)
    env.reset()
    for t in range(T):
        env.step(acts[:, t])
    env.check_status()
    if check_oracle:
        pick = np.arange(B) if B <= 64 else np.array([0, 1, B // 2, B - 1])
        want = oracle_for(env, len(pick)).run_episodes(acts[pick])
        assert (env.canvas.cpu().numpy().view(np.uint16)[pick] == want).all(), (envname, B, math, fused_max)


for math in ("det", "fast"):
    render("color_mnist", 64, 3, 48, 4, math, "0")
    render("color_mnist", 64, 3, 48, 4, math, "100000")
    render("mnist", 32, 1, 33, 3, math, "100000")
    render("simple_mnist", 32, 1, 1, 3, math, "0")
    render("color_mnist", 64, 3, 1500, 3, math, "100000")          # second wave of the fused kernel, work-sorted CTAs
    render("color_mnist", 64, 3, 20000, 3, math, "0")              # tiered expand schedule, in-place compositing
# the learner's convolutions: every geometry, forward + dgrad + wgrad
pl = import_module(sp.__name__ + ".models.policy_learn")
g = torch.Generator(device="cuda")
g.manual_seed(0)
for planes in (4, 3, 2):
    pl.PLANES = planes
    for kind, cin, cout, k, stride, pad, H in (("conv", 3, 32, 5, 1, 2, 64), ("conv", 1, 32, 5, 1, 2, 64), ("conv", 32, 32, 4, 2, 0, 64),
                                               ("conv", 32, 32, 4, 2, 0, 31), ("conv", 32, 32, 3, 1, 1, 6), ("conv", 32, 1, 3, 1, 1, 64),
                                               ("convT", 16, 32, 4, 2, 1, 4), ("convT", 32, 32, 4, 2, 1, 32)):
        mod = (torch.nn.Conv2d(cin, cout, k, stride=stride, padding=pad) if kind == "conv"
               else torch.nn.ConvTranspose2d(cin, cout, 4, stride=2, padding=1)).cuda()
        x = torch.randn((3, H, H, cin), generator=g, device="cuda").requires_grad_(True)
        y = pl.conv2d(x, mod, stride=stride, pad=pad) if kind == "conv" else pl.conv_transpose2d(x, mod)
        y.square().sum().backward()
torch.cuda.synchronize()
print("bounds violations:", int(lib.spiral_debug_bounds_violations()))
