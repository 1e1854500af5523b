import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from importlib import import_module
import spiral_b200 as sp
pol = import_module(sp.__name__ + ".models.policy"); pn = import_module(sp.__name__ + ".models.policy_native")
a = sp.get_args(["--env", "color_mnist", "--location_size", "64", "--color_channel", "3", "--episode_length", "20", "--loss", "gan"])
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
env = sp.create_env(a, num_envs=B); pi = pol.Policy(a, env).cuda(); actor = pn.NativeActor(pi)
state, cond, z = env.reset(); c, h = pi.get_initial_features(B)
la = torch.full((B, len(env.acs)), -1, dtype=torch.int32, device="cuda")
g = torch.Generator(device="cuda"); g.manual_seed(0)
for _ in range(2): actor.act(state, la, c, h, cond, z, g)
torch.cuda.synchronize()
print("ok")
