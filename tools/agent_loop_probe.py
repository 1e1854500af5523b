"""Times rollout / train_policy / train_gan of the synchronous loop with and without CUDA graphs (rgb64 shape)."""
import os, sys
from importlib import import_module
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import spiral_b200 as sp
agent_mod = import_module(sp.__name__ + ".agent")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
for graphs in ("True", "False"):
    a = sp.get_args(["--env", "color_mnist", "--location_size", "64", "--color_channel", "3", "--episode_length", "20",
                     "--loss", "gan", "--disc_dim", "64", "--disc_batch_size", "64", "--cuda_graphs", graphs])
    env = sp.create_env(a, num_envs=n)
    ag = agent_mod.Agent(a, env)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    tt = [0.0, 0.0, 0.0]
    for i in range(7):
        e = [ev() for _ in range(4)]
        e[0].record(); ro = ag.rollout(); e[1].record(); ag.train_policy(ro); e[2].record(); ag.train_gan(); e[3].record()
        torch.cuda.synchronize()
        if i >= 4:
            for j in range(3):
                tt[j] += e[j].elapsed_time(e[j + 1]) / 3
    print("canvases %d graphs %s: rollout %.1f train_policy %.1f train_gan %.1f ms (learner graph: %s)"
          % (n, graphs, tt[0], tt[1], tt[2], ag._tp_graph is not None))
