#!/usr/bin/env python
"""train.py -- single launcher of the synchronous data-parallel loop.

    python train.py --env simple_mnist --loss gan --num_envs 256 --max_iters 100 [reference flags ...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 train.py ...

Stands in for the reference's run.py (tmux: 1 ps + N workers), main.py (cluster spec, tf.train.Server) and the role
loop of trainer.py:107-118: there is one role here -- every rank owns an actor partition of the canvases and a
replica of the weights, and an iteration is  rollout -> train_policy -> train_gan  with NCCL all-reduced gradients
(agent.py).  The chief (rank 0) writes params.json, checkpoints (every --save_iters) and the reference's scalar
names (agent.py:207-230, discriminator.py:149-155) every --policy_log_step iterations.
"""
import json
import os
import sys
import time
from importlib import import_module

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import spiral_b200 as sp  # noqa: E402


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    extra = {"--num_envs": 256, "--max_iters": 100, "--save_iters": 50}
    opts = {}
    for k, default in extra.items():                         # launcher-only flags (not part of the reference's parser)
        opts[k] = default
        if k in argv:
            i = argv.index(k)
            opts[k] = int(argv[i + 1])
            del argv[i:i + 2]
    du = import_module(sp.__name__ + ".distributed")
    agent_mod = import_module(sp.__name__ + ".agent")
    ck = import_module(sp.__name__ + ".checkpoint")
    rank, world, local = du.init()
    torch.cuda.set_device(local)
    torch.manual_seed(0)                                     # identical initial weights on every rank
    args = sp.get_args(argv)
    if args.env == "simple":
        raise SystemExit("train.py drives the batched GPU loop; --env simple is the reference's CPU / PIL debug env "
                         "(BASELINE configs[0]) and has no device path -- step it directly (spiral_b200.create_env)")
    load_path = str(getattr(args, "load_path", "") or "")
    if load_path and os.path.isdir(load_path) and os.path.exists(os.path.join(load_path, ck.PARAM_FNAME)) \
            and ck.latest_checkpoint(load_path):
        changed = ck.load_args(args, load_path)              # utils/train.py:97-120: a resumed run keeps the saved flags
        if changed and rank == 0:
            print(json.dumps({"resumed_flags": {k: list(map(str, v)) for k, v in changed.items()}}), flush=True)
    args.seed = int(args.seed) + rank                        # main.py:21: seed + task index for the actors
    _, count = du.shard(opts["--num_envs"], rank, world)
    env = sp.create_env(args, num_envs=count, device=torch.device("cuda", local))
    agent = agent_mod.Agent(args, env)
    load_path = load_path or os.path.join("logs", "%s_%s" % (args.env, time.strftime("%Y-%m-%d_%H-%M-%S")))
    step = 0
    if os.path.isdir(load_path) and ck.latest_checkpoint(load_path):
        step = ck.load_checkpoint(agent, load_path)          # every rank restores the same file
    elif rank == 0:
        ck.save_args(args, load_path)
    t0 = time.time()
    for it in range(step, opts["--max_iters"]):
        ro = agent.rollout()
        m = agent.train_policy(ro)
        if agent.gan:
            m.update(agent.train_gan())
        log_now = (it + 1) % max(1, int(args.policy_log_step)) == 0 or it + 1 == opts["--max_iters"]
        if log_now:
            # every rank: a pending-dab overflow or an RNG-table overrun truncates strokes silently on the device; the
            # flag is read here (a synchronising read, at the logging cadence) and raises
            env.check_status()
        if rank == 0 and log_now:
            line = {k: (float(v) if torch.is_tensor(v) else v) for k, v in m.items()}
            line.update(iter=it + 1, env_steps=(it + 1 - step) * opts["--num_envs"] * env.episode_length,
                        seconds=round(time.time() - t0, 2))
            print(json.dumps(line), flush=True)
        if rank == 0 and ((it + 1) % opts["--save_iters"] == 0 or it + 1 == opts["--max_iters"]):
            ck.save_checkpoint(agent, load_path, it + 1)
    agent.close()                                            # captured graphs hold NCCL collectives: drop them before the group
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return load_path


if __name__ == "__main__":
    main()
