"""Small driver for ncu: a few env steps of the rgb64 workload (used for profiles/)."""
import sys
import numpy as np
import torch
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import spiral_b200 as sp

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4
a = sp.get_args(["--env", "color_mnist", "--location_size", "64", "--color_channel", "3", "--episode_length", str(T)])
env = sp.create_env(a, num_envs=B)
rng = np.random.RandomState(0)
acts = np.stack([rng.randint(n, size=(B, T)) for n in env.head_sizes], -1).astype(np.int32)
env.reset()
for t in range(T):
    env.step(acts[:, t])
torch.cuda.synchronize()
env.check_status()
print("ok", float(env.state.mean()))
