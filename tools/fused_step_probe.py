"""Times a 20-step episode of env.step at small batches with the fused step kernel on and off (B200).

    python tools/fused_step_probe.py [B ...]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spiral_b200 as sp  # noqa: E402


def run(B, fused_max, T=20, reps=10):
    os.environ["SPIRAL_FUSED_STEP_MAX"] = fused_max
    a = sp.get_args(["--env", "color_mnist", "--location_size", "64", "--color_channel", "3", "--episode_length", str(T),
                     "--jump", "True", "--seed", "123"])
    env = sp.create_env(a, num_envs=B)
    g = torch.Generator(device="cpu")
    g.manual_seed(123)
    acts = torch.stack([torch.randint(0, n, (T, B), generator=g) for n in env.head_sizes], -1).to(torch.int32).cuda()

    def episode():
        env.reset()
        for t in range(T):
            env.step(acts[t])

    for _ in range(3):
        episode()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        episode()
    e.record()
    torch.cuda.synchronize()
    env.check_status()
    return s.elapsed_time(e) / reps, env.canvas.clone()


if __name__ == "__main__":
    for B in [int(x) for x in sys.argv[1:]] or [128, 1024, 1184]:
        t0, c0 = run(B, "0")
        t1, c1 = run(B, "100000")
        print("B=%d two-kernel %.2f ms/episode, fused %.2f ms/episode, identical=%s" % (B, t0, t1, bool((c0 == c1).all())), flush=True)
