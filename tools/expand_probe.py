"""Times a 20-step episode of spiral_env_expand / spiral_env_composite at large batches (B200); the tuning knobs are
the library's environment variables (SPIRAL_EXPAND_LANES, SPIRAL_EXPAND_SORT, SPIRAL_EXPAND_SOLO_DIV).

    SPIRAL_EXPAND_SOLO_DIV=64 python tools/expand_probe.py 16384 65536
"""
import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spiral_b200 as sp  # noqa: E402

N = sp._native
lib = N.lib()
T = 20
tag = " ".join("%s=%s" % (k[7:], v) for k, v in sorted(os.environ.items()) if k.startswith("SPIRAL_EXPAND"))
for B in [int(x) for x in sys.argv[1:]] or [16384, 65536]:
    a = sp.get_args(["--env", "color_mnist", "--location_size", "64", "--color_channel", "3", "--episode_length", str(T),
                     "--jump", "True", "--seed", "123"])
    env = sp.create_env(a, num_envs=B)
    g = torch.Generator(device="cpu")
    g.manual_seed(123)
    acts = torch.stack([torch.randint(0, n, (T, B), generator=g) for n in env.head_sizes], -1).to(torch.int32).cuda()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    tot = {"expand": 0.0, "composite": 0.0}
    reps = 5
    for r in range(2 + reps):
        env.reset()
        pairs = []
        for t in range(T):
            e0, e1, e2 = ev(), ev(), ev()
            e0.record()
            N.check(lib.spiral_env_expand(env._handle, N.ptr(acts[t]), N.ptr(env.brush_state), N.ptr(env.env_state),
                                          N.ptr(env.dab_scratch), N.ptr(env.dab_count), N.ptr(env.status), N.stream_ptr()))
            e1.record()
            N.check(lib.spiral_env_composite(env._handle, N.ptr(env.canvas), N.ptr(env.dab_scratch), N.ptr(env.dab_count),
                                             N.ptr(env.obs), N.stream_ptr()))
            e2.record()
            env._step += 1
            pairs.append((e0, e1, e2))
        torch.cuda.synchronize()
        if r >= 2:
            tot["expand"] += sum(p[0].elapsed_time(p[1]) for p in pairs) / reps
            tot["composite"] += sum(p[1].elapsed_time(p[2]) for p in pairs) / reps
    env.check_status()
    digest = hashlib.sha1(env.canvas.cpu().numpy().tobytes()).hexdigest()[:12]
    print("[%s] B=%d expand %.2f ms/episode, composite %.2f ms/episode, canvas sha1 %s" % (tag, B, tot["expand"], tot["composite"], digest),
          flush=True)
    del env
    torch.cuda.empty_cache()
