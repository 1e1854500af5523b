"""torchrun smoke of the synchronous data-parallel loop on real GPUs (NCCL): every rank owns an actor partition,
gradients are all-reduced, weights must stay bit-identical across ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/agent_ddp_smoke.py
"""
import os
import sys
from importlib import import_module

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
                 "--seed", str(123 + rank)])
start, count = du.shard(64, rank, world)
env = sp.create_env(a, num_envs=count, device=dev)
ag = agent_mod.Agent(a, env)
for it in range(2):
    ro = ag.rollout()
    m = ag.train_policy(ro)
    g = ag.train_gan()
torch.cuda.synchronize()


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
    print("ddp smoke ok: world %d, canvases/rank %d, loss %.4f, grad norm %.4f, d_loss %.4f; weights identical on all ranks"
          % (world, count, float(m["model/total_loss"]), float(m["model/grad_global_norm"]), float(g["gan/disc_loss"])))
if world > 1:
    dist.destroy_process_group()
