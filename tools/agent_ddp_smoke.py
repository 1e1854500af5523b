"""torchrun smoke of the synchronous data-parallel loop on real GPUs (NCCL): every rank owns an actor partition,
gradients are all-reduced, weights must stay bit-identical across ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/agent_ddp_smoke.py
"""
import faulthandler
import os
import sys
from importlib import import_module

if os.environ.get("SMOKE_FAULT_SECS"):                      # a hang prints every thread's stack instead of nothing
    faulthandler.dump_traceback_later(int(os.environ["SMOKE_FAULT_SECS"]), exit=True)

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import spiral_b200 as sp

du = import_module(sp.__name__ + ".distributed")
agent_mod = import_module(sp.__name__ + ".agent")
os.environ.setdefault("NCCL_DEBUG", "WARN")
rank, world, local = du.init()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
torch.manual_seed(0)                                       # identical initial weights on every rank
a = sp.get_args(["--env", "color_mnist", "--location_size", "32", "--color_channel", "3", "--episode_length", "4",
                 "--loss", "gan", "--disc_dim", "64", "--disc_precision", "bf16x3", "--disc_batch_size", "16",
                 "--seed", str(123 + rank), "--disc_global_bn", os.environ.get("SMOKE_GLOBAL_BN", "True")])
start, count = du.shard(64, rank, world)
env = sp.create_env(a, num_envs=count, device=dev)
ag = agent_mod.Agent(a, env)
iters = int(os.environ.get("SMOKE_ITERS", "6"))                # the learner / discriminator graphs are captured at the third call
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for it in range(iters):
    if os.environ.get("SMOKE_VERBOSE"):
        print("rank %d iteration %d" % (rank, it), flush=True)
    ev[0].record()
    ro = ag.rollout()
    ev[1].record()
    m = ag.train_policy(ro)
    ev[2].record()
    g = ag.train_gan()
    ev[3].record()
torch.cuda.synchronize()
times = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]


def digest(params):
    return torch.stack([p.detach().double().sum() for p in params])


dp = digest(ag.policy.parameters())
dd = digest(ag.disc.params.values())
if world > 1:
    ref_p, ref_d = dp.clone(), dd.clone()
    dist.broadcast(ref_p, 0)
    dist.broadcast(ref_d, 0)
    assert torch.equal(ref_p, dp), "policy weights diverged across ranks"
    assert torch.equal(ref_d, dd), "discriminator weights diverged across ranks"
if rank == 0:
    print("ddp smoke ok: world %d, canvases/rank %d, %d iterations, loss %.4f, grad norm %.4f, d_loss %.4f; weights identical on "
          "all ranks; CUDA graphs: rollout %s, learner %s, discriminator %s; gradient exchange: %s (%.1f + %.1f MB per step); "
          "global-batch BN %s; last iteration rollout %.1f ms, train_policy %.1f ms, train_gan %.1f ms"
          % (world, count, iters, float(m["model/total_loss"]), float(m["model/grad_global_norm"]), float(g["gan/disc_loss"]),
             ag._graph is not None, ag._tp_graph is not None, ag._tg_graph is not None,
             "GradReducer (bucketed all-reduce from backward hooks)" if ag._pol_red is not None else "none (one rank)",
             (ag._pol_red.collective_bytes() if ag._pol_red else 0) / 1e6, (ag._disc_red.collective_bytes() if ag._disc_red else 0) / 1e6,
             ag._global_bn is not None, times[0], times[1], times[2]))
ag.close()                                                 # the learner graphs hold NCCL collectives: drop them first
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
