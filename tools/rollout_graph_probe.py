import sys, os, time, torch
sys.path.insert(0,'/root/repo')
from importlib import import_module
import spiral_b200 as sp
agent_mod = import_module(sp.__name__ + ".agent")
for graphs in ("False","True"):
    for n in (128, 1024):
        a = sp.get_args(["--env","color_mnist","--location_size","64","--color_channel","3","--episode_length","20","--loss","gan","--disc_dim","64","--cuda_graphs",graphs])
        env = sp.create_env(a, num_envs=n); ag = agent_mod.Agent(a, env)
        for _ in range(2): ag.rollout()
        torch.cuda.synchronize(); t0=time.perf_counter()
        for _ in range(3): ag.rollout()
        torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/3
        print("graphs=%s B=%d rollout %.1f ms  (%.0f env-steps/s)" % (graphs, n, dt*1e3, n*20/dt))
        del ag, env; torch.cuda.empty_cache()
