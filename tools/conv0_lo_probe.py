"""Probe for ncu: D forward on 8-bit images (lo plane of layer 0's input all zero) and on non-integer images."""
import os, sys
from importlib import import_module
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import spiral_b200 as sp
models = import_module(sp.__name__ + ".models")
B = 1024
a = sp.get_args(["--env", "color_mnist", "--color_channel", "3", "--loss", "gan", "--disc_dim", "64", "--disc_precision", "bf16x3"])
env = sp.create_env(a, num_envs=8)
D = models.Discriminator(a, None, [64, 64, 3], norm_fn=env.norm, scope_name="global", max_batch=B)
g = torch.Generator(device="cuda"); g.manual_seed(0)
x8 = torch.randint(0, 256, (B, 64, 64, 3), generator=g, device="cuda").float()
xf = x8 + 0.3
p1, _ = D.forward(x8)
p2, _ = D.forward(xf)
torch.cuda.synchronize()
print(float(p1.mean()), float(p2.mean()))
