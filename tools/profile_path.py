"""Driver for ncu: one short pass of the whole hot path at the bench shape (used for profiles/).

    python tools/profile_path.py B [T]

reset + T env steps (expand, composite) + discriminator reward (conv0, conv_tc x5, BN, act, dense)
+ fused A2C loss (gae, row, final) for B canvases of the rgb64 workload.
"""
import os
import sys
from importlib import import_module

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import spiral_b200 as sp

rl = import_module(sp.__name__ + ".rl_utils")
models = import_module(sp.__name__ + ".models")

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
T = int(sys.argv[2]) if len(sys.argv) > 2 else 3
TA = 20                                                   # A2C always over the full [B,20] episode batch
dev = torch.device("cuda", 0)
a = sp.get_args(["--env", "color_mnist", "--location_size", "64", "--color_channel", "3", "--episode_length", "20",
                 "--loss", "gan", "--disc_dim", "64", "--disc_precision", "bf16x3"])
env = sp.create_env(a, num_envs=B, device=dev)
disc = models.Discriminator(a, None, [64, 64, 3], norm_fn=env.norm, scope_name="global", device=dev, max_batch=B)
rng = np.random.RandomState(0)
acts = torch.as_tensor(np.stack([rng.randint(n, size=(B, TA)) for n in env.head_sizes], -1).astype(np.int32)).to(dev)
gg = torch.Generator(device=dev)
gg.manual_seed(7)
logits = [torch.randn((B, TA, n), generator=gg, device=dev) for n in env.head_sizes]
values = torch.randn((B, TA), generator=gg, device=dev)
vf = torch.randn((B, TA), generator=gg, device=dev)
rewards = torch.zeros((B, TA), device=dev)
torch.cuda.synchronize()

env.reset()
for t in range(T):
    env.step(acts[:, t].contiguous())
probs, _ = disc.forward(env.state)
rewards[:, -1] = probs
res = rl.a2c_loss(logits, acts, rewards, values, vf, 0.99, 1.0, 0.01)
torch.cuda.synchronize()
env.check_status()
print("ok", float(env.state.mean()), float(probs.mean()), float(res.total))
