"""Probe: actor groups on separate CUDA streams (expand of one group overlaps composite of another).

    python tools/pipeline_probe.py B T G [prio]

Times T env steps of B canvases split into G independent groups, each on its own stream.  With
prio=1 the expand kernels go to a high-priority stream per group (event-chained to the composite).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import spiral_b200 as sp

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
T = int(sys.argv[2]) if len(sys.argv) > 2 else 20
G = int(sys.argv[3]) if len(sys.argv) > 3 else 2
PRIO = int(sys.argv[4]) if len(sys.argv) > 4 else 0
N = sp._native
lib = N.lib()
dev = torch.device("cuda", 0)
a = sp.get_args(["--env", "color_mnist", "--location_size", "64", "--color_channel", "3", "--episode_length", str(T)])
envs = [sp.create_env(a, num_envs=B // G, device=dev) for _ in range(G)]
g = torch.Generator(device="cpu")
g.manual_seed(123)
acts = torch.stack([torch.randint(0, n, (T, B), generator=g) for n in envs[0].head_sizes], -1).to(torch.int32).to(dev)
acts_g = [acts[:, i * (B // G):(i + 1) * (B // G)].contiguous() for i in range(G)]
lo, hi = -1, 0
streams = [torch.cuda.Stream(priority=0) for _ in range(G)]
hstreams = [torch.cuda.Stream(priority=-1) for _ in range(G)] if PRIO else None


def episode():
    main = torch.cuda.current_stream()
    start = torch.cuda.Event()
    start.record(main)
    for i, env in enumerate(envs):
        streams[i].wait_event(start)
        with torch.cuda.stream(streams[i]):
            env.reset()
    for t in range(T):
        for i, env in enumerate(envs):
            s = streams[i]
            if PRIO:
                h = hstreams[i]
                e = torch.cuda.Event()
                e.record(s)
                h.wait_event(e)
                N.check(lib.spiral_env_expand(env._handle, N.ptr(acts_g[i][t]), N.ptr(env.brush_state), N.ptr(env.env_state),
                                              N.ptr(env.dab_scratch), N.ptr(env.dab_count), N.ptr(env.status),
                                              N.stream_ptr(h)))
                e2 = torch.cuda.Event()
                e2.record(h)
                s.wait_event(e2)
                N.check(lib.spiral_env_composite(env._handle, N.ptr(env.canvas), N.ptr(env.dab_scratch),
                                                 N.ptr(env.dab_count), N.ptr(env.obs), N.stream_ptr(s)))
            else:
                with torch.cuda.stream(s):
                    env.step(acts_g[i][t])
    for s in streams:
        e = torch.cuda.Event()
        e.record(s)
        main.wait_event(e)


for _ in range(2):
    episode()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
R = 3
for _ in range(R):
    episode()
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / R
for env in envs:
    env.check_status()
chk = sum(float(env.state.double().sum()) for env in envs)
print("B=%d T=%d G=%d prio=%d  %.2f ms/episode  %.3f ms/step  %.2f M canvas-steps/s  checksum %.1f"
      % (B, T, G, PRIO, ms, ms / T, B * T / ms / 1e3, chk))
