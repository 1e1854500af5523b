"""Probe: time of policy.act (PyTorch module) per env step at several batch sizes / dtypes."""
import os, sys, time, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from importlib import import_module
import spiral_b200 as sp
pol = import_module(sp.__name__ + ".models.policy")
pn = import_module(sp.__name__ + ".models.policy_native")
a = sp.get_args(["--env", "color_mnist", "--location_size", "64", "--color_channel", "3", "--episode_length", "20", "--loss", "gan"])
for B in (128, 1024, 8192):
    env = sp.create_env(a, num_envs=B)
    pi = pol.Policy(a, env).cuda()
    state, cond, z = env.reset()
    c, h = pi.get_initial_features(B)
    la = torch.full((B, len(env.acs)), -1, dtype=torch.int32, device="cuda")
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    actor = pn.NativeActor(pi)
    for mode in ("fp32", "tf32", "bf16", "native"):
        torch.backends.cudnn.allow_tf32 = mode != "fp32"
        torch.backends.cuda.matmul.allow_tf32 = mode != "fp32"
        m = pi
        def run():
            if mode == "native":
                return actor.act(state, la, c, h, cond, z, g)
            if mode.startswith("bf16"):
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return m.act(state, la, c, h, cond, z, g)
            return m.act(state, la, c, h, cond, z, g)
        for _ in range(3): run()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5): run()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
        print("B=%d %-8s act %.2f ms  (%.0f frames/s)" % (B, mode, dt * 1e3, B / dt))
    del env, pi; torch.cuda.empty_cache()
