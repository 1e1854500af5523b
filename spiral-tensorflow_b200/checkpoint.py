"""Run directory: ``params.json`` and checkpoints of the ``global`` variables.

Mirror of the reference's bookkeeping around the hot path: ``utils/train.py:66-120`` (``save_args`` /
``load_args``: every flag as JSON next to the checkpoints, reloaded on resume except ``load_path`` / test-time
keys) and ``trainer.py:23-29,66-80`` (a Supervisor that saves every non-``local*`` variable and restores them when
the same ``--load_path`` is given).  Here one rank (the chief, rank 0) writes a single file per save:

    params.json                      flags (sorted, indent 4, like the reference)
    model.ckpt-<step>.pt             {'global/policy/...': tensor, 'global/disc/<tf name>': tensor,
                                      'optim/<policy|disc>/{t,m,v}': ..., 'step': int, 'sample_counter': int}

Discriminator tensors keep their TensorFlow names and layouts (``conv0/kernel`` HWIO, ``dense/kernel`` [C,1], ...),
so a reference checkpoint converted to numpy maps one to one through ``Discriminator.load_numpy``.
"""
from __future__ import annotations

import glob
import json
import os
import re

import torch

PARAM_FNAME = "params.json"
SKIP_ON_LOAD = ("load_path", "test_epoch", "test_dataset", "train")          # utils/train.py:97


def save_args(args, load_path):
    os.makedirs(load_path, exist_ok=True)
    info = {k: (v if isinstance(v, (int, float, bool, str, list, type(None))) else str(v)) for k, v in vars(args).items()}
    with open(os.path.join(load_path, PARAM_FNAME), "w") as f:
        json.dump(info, f, indent=4, sort_keys=True)
    return os.path.join(load_path, PARAM_FNAME)


def load_args(args, load_path, skip_list=SKIP_ON_LOAD):
    """Overwrite ``args`` with the saved flags (utils/train.py:97-120); keys the current parser does not know are
    ignored, ``skip_list`` keys keep their command-line values."""
    with open(os.path.join(load_path, PARAM_FNAME)) as f:
        saved = json.load(f)
    changed = {}
    for k, v in saved.items():
        if k in skip_list or not hasattr(args, k):
            continue
        if getattr(args, k) != v:
            changed[k] = (getattr(args, k), v)
            setattr(args, k, v)
    return changed


def _optim_state(opt):
    return {"t": opt.t, "m": [m.detach().cpu() for m in opt.m], "v": [v.detach().cpu() for v in opt.v]}


def _load_optim(opt, st):
    opt.t = int(st["t"])
    for dst, src in zip(opt.m, st["m"]):
        dst.copy_(src)
    for dst, src in zip(opt.v, st["v"]):
        dst.copy_(src)


def save_checkpoint(agent, load_path, step):
    """Chief-only save of every ``global`` variable and the optimiser slots (trainer.py:23-29: variables whose name
    starts with ``local`` are not saved -- there are none here, the synchronous design has no local copies)."""
    os.makedirs(load_path, exist_ok=True)
    state = {"step": int(step)}
    for k, v in agent.policy.state_dict().items():
        state["global/policy/" + k] = v.detach().cpu()
    state["optim/policy"] = _optim_state(agent.policy_opt)
    if agent.gan:
        for k, v in agent.disc.params.items():
            state["global/disc/" + k] = v.detach().cpu()
        state["optim/disc"] = _optim_state(agent.disc_opt)
    ctr = getattr(agent.policy, "_sample_counter_dev", None)
    state["sample_counter"] = int(ctr.item()) if ctr is not None else int(getattr(agent.policy, "_sample_offset", 0))
    # the random streams and the data a resumed run would otherwise re-draw from their seeds: env (targets, z),
    # agent (WGAN-GP interpolation), replay (sampling) generators, the replay ring and the actors' carried LSTM state
    state["rng/env"] = agent.env._rng.get_state().cpu() if hasattr(agent.env, "_rng") else None
    state["rng/agent"] = agent.rng.get_state().cpu()
    if agent.gan:
        state["rng/replay"] = agent.replay.rng.get_state().cpu()
        state["replay/data"] = agent.replay.data.detach().cpu()
        state["replay/idx"] = int(agent.replay.idx)
    if getattr(agent, "_lstm_state", None) is not None:
        state["actor/lstm_state"] = agent._lstm_state.detach().cpu()
    path = os.path.join(load_path, "model.ckpt-%d.pt" % int(step))
    tmp = path + ".tmp"
    torch.save(state, tmp)
    os.replace(tmp, path)                                    # a reader never sees a half-written checkpoint
    return path


def latest_checkpoint(load_path):
    best, best_step = None, -1
    for p in glob.glob(os.path.join(load_path, "model.ckpt-*.pt")):
        m = re.search(r"model\.ckpt-(\d+)\.pt$", p)
        if m and int(m.group(1)) > best_step:
            best, best_step = p, int(m.group(1))
    return best


def load_checkpoint(agent, path_or_dir):
    """Restore an ``Agent`` in place; returns the saved step (0 when the directory holds no checkpoint)."""
    path = latest_checkpoint(path_or_dir) if os.path.isdir(path_or_dir) else path_or_dir
    if path is None:
        return 0
    state = torch.load(path, map_location="cpu")
    pol = {k[len("global/policy/"):]: v for k, v in state.items() if k.startswith("global/policy/")}
    agent.policy.load_state_dict(pol)
    _load_optim(agent.policy_opt, state["optim/policy"])
    if agent.gan and "optim/disc" in state:
        for k, v in state.items():
            if k.startswith("global/disc/"):
                agent.disc.params[k[len("global/disc/"):]].data.copy_(v)
        agent.disc.sync_weights()
        _load_optim(agent.disc_opt, state["optim/disc"])
    ctr = getattr(agent.policy, "_sample_counter_dev", None)
    if ctr is not None:
        ctr.fill_(int(state.get("sample_counter", 0)))
    agent.policy._sample_offset = int(state.get("sample_counter", 0))    # read when the device counter is created later
    if state.get("rng/env") is not None and hasattr(agent.env, "_rng"):
        agent.env._rng.set_state(state["rng/env"])
    if state.get("rng/agent") is not None:
        agent.rng.set_state(state["rng/agent"])
    if agent.gan and state.get("rng/replay") is not None:
        agent.replay.rng.set_state(state["rng/replay"])
        if tuple(state["replay/data"].shape) == tuple(agent.replay.data.shape):
            agent.replay.data.copy_(state["replay/data"])
            agent.replay.idx = int(state["replay/idx"])
    if state.get("actor/lstm_state") is not None and getattr(agent, "_lstm_state", None) is not None \
            and tuple(state["actor/lstm_state"].shape) == tuple(agent._lstm_state.shape):
        agent._lstm_state.copy_(state["actor/lstm_state"])
    # the acting path holds its own packed (bf16) copy of the policy's convolution kernels, refreshed only after a
    # learner step: re-pack it, or the first rollout after a resume acts with the random-init network
    if getattr(agent, "actor", None) is not None:
        agent.actor.sync()
    return int(state["step"])
