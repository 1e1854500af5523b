"""Returns / advantages and the fused A2C loss.

Mirror of the reference's rl_utils.py:18-57 (``discount``, ``multiple_process_rollout``) and of the
loss built in agent.py:161-196, computed on the GPU by libspiral_b200.so (K3) for a whole
``[B, T]`` batch of episodes at once.
"""
from __future__ import annotations

import ctypes as C
from collections import namedtuple

import numpy as np
import torch

from . import _native as N

Batch = namedtuple("Batch", ["si", "a", "adv", "r", "features", "c", "z"])

_SCRATCH = {}


def _scratch(name, shape, dtype, device, zero=False):
    """Per-(name, shape, device) scratch tensor, allocated once: nothing is allocated on the step path;
    the kernels overwrite it on every call."""
    key = (name, tuple(shape), dtype, str(device))
    t = _SCRATCH.get(key)
    if t is None:
        t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=device)
        _SCRATCH[key] = t
    return t


def discount(x, gamma):
    """Reverse discounted cumulative sum along axis 1 (reference rl_utils.py:18-20), numpy."""
    x = np.asarray(x, np.float64)
    out = np.zeros_like(x)
    acc = np.zeros(x.shape[0])
    for t in range(x.shape[1] - 1, -1, -1):
        acc = x[:, t] + gamma * acc
        out[:, t] = acc
    return out


@torch.no_grad()
def batched_env_runner(env, act, lstm_state, initial_action, lstm_size, state_dtype=torch.float32,
                       target=None, z=None, generator=None):
    """One pass of the body of ``env_runner``'s ``while True`` (reference rl_utils.py:190-238) for B envs at once.

    ``act(state, last_action, c, h, condition, z, generator) -> (action [B,K], value [B], c, h)`` is ``policy.act``.
    ``lstm_state`` [B, 2, lstm] = (c, h) is the actors' ``last_features``: the reference fetches
    ``policy.get_initial_features`` once, BEFORE the loop (:186-187), so the state at the end of an episode is the state
    the next episode starts from -- it is read here and written back in place.  ``PartialRollout.add`` (:205-206) stores
    the features the step started from, so ``features[:, t]`` is the state BEFORE step t (the learner feeds
    ``features[:, 0]``, :42); ``states`` gets the final state appended (:232).  The reference resets the env after the
    episode (:222); here the episode starts with the reset, which is the same sequence one call later."""
    T, B = env.episode_length, env.num_envs
    state, condition, z = env.reset(target=target, z=z)
    dev = state.device
    c, h = lstm_state[:, 0], lstm_state[:, 1]
    last_action = initial_action
    states = torch.empty((B, T + 1) + tuple(state.shape[1:]), dtype=state_dtype, device=dev)
    actions = torch.empty((B, T, len(env.acs)), dtype=torch.int32, device=dev)
    rewards = torch.zeros((B, T), device=dev)
    values = torch.zeros((B, T, 1), device=dev)
    features = torch.empty((B, T, 2, lstm_size), device=dev)
    for t in range(T):
        states[:, t] = state
        features[:, t, 0] = c
        features[:, t, 1] = h
        action, value, c, h = act(state, last_action, c, h, condition, z, generator)
        state, reward, terminal, _ = env.step(action)
        actions[:, t] = action
        rewards[:, t] = reward
        values[:, t, 0] = value
        last_action = action
    states[:, T] = state
    lstm_state[:, 0] = c
    lstm_state[:, 1] = h
    out = dict(states=states, actions=actions, rewards=rewards, values=values, features=features)
    if condition is not None:
        out["conditions"] = condition.unsqueeze(1).expand((B, T) + tuple(condition.shape[1:]))
    else:
        out["z"] = z.unsqueeze(1).expand(B, T, z.shape[-1])
    return out


class A2CResult(object):
    __slots__ = ("pi_loss", "vf_loss", "entropy", "total", "adv", "r", "dlogits", "dvf", "scalars")


def a2c_loss(logits, actions, rewards, values, vf, gamma=0.99, lambda_=1.0, entropy_coeff=0.01,
             want_grads=True, out=None):
    """Fused GAE + A2C loss (+ gradient) on the GPU.

    logits : list of K float32 CUDA tensors [B, T, N_k] in env.acs order
    actions: int32 [B, T, K];  rewards, values, vf: float32 [B, T]
      - ``values`` are the rollout-time value predictions used for the advantage
        (rl_utils.py:29-38), ``vf`` the learner's current value output (agent.py:186)
    Returns an ``A2CResult``: scalars stay on the device (``.scalars`` f32[4] =
    pi_loss, vf_loss, entropy, total), ``adv``/``r`` f32[B,T], and d total/d logits, d total/d vf.
    """
    lib = N.lib()
    K = len(logits)
    B, T = vf.shape
    dev = vf.device
    for k, l in enumerate(logits):
        if l.dtype != torch.float32 or not l.is_cuda or not l.is_contiguous():
            raise ValueError("logits[%d] must be a contiguous float32 CUDA tensor" % k)
        if l.shape[0] != B or l.shape[1] != T:
            raise ValueError("logits[%d] must be [B,T,N]" % k)
    actions = actions.to(device=dev, dtype=torch.int32).contiguous()
    rewards = rewards.to(device=dev, dtype=torch.float32).contiguous()
    values = values.to(device=dev, dtype=torch.float32).contiguous()
    vf = vf.contiguous()
    res = out if out is not None else A2CResult()
    if out is None:
        res.adv = torch.empty((B, T), dtype=torch.float32, device=dev)
        res.r = torch.empty((B, T), dtype=torch.float32, device=dev)
        res.scalars = torch.empty((4,), dtype=torch.float32, device=dev)
        res.dlogits = [torch.empty_like(l) for l in logits] if want_grads else None
        res.dvf = torch.empty((B, T), dtype=torch.float32, device=dev) if want_grads else None
        res.pi_loss = res.vf_loss = res.entropy = res.total = None
    partial = _scratch("a2c_partial", (B * T * 4 + 2048,), torch.float32, dev)
    lp = (C.c_void_p * K)(*[l.data_ptr() for l in logits])
    sizes = (C.c_int32 * K)(*[int(l.shape[2]) for l in logits])
    dp = (C.c_void_p * K)(*[d.data_ptr() for d in res.dlogits]) if want_grads else None
    N.check(lib.spiral_a2c_loss(lp, sizes, K, N.ptr(actions), N.ptr(rewards), N.ptr(values), N.ptr(vf),
                                B, T, float(gamma), float(lambda_), float(entropy_coeff),
                                N.ptr(res.adv), N.ptr(res.r), N.ptr(res.scalars), dp,
                                N.ptr(res.dvf) if want_grads else None, N.ptr(partial), N.stream_ptr()),
            "spiral_a2c_loss")
    res.pi_loss, res.vf_loss, res.entropy, res.total = res.scalars[0], res.scalars[1], res.scalars[2], res.scalars[3]
    return res


class _A2CLossFn(torch.autograd.Function):
    """total loss as a differentiable scalar: backward hands out the kernel's dlogits / dvf."""

    @staticmethod
    def forward(ctx, actions, rewards, values, gamma, lambda_, entropy_coeff, vf, *logits):
        res = a2c_loss([l.detach().contiguous() for l in logits], actions, rewards, values,
                       vf.detach().contiguous(), gamma, lambda_, entropy_coeff, want_grads=True)
        ctx.save_for_backward(res.dvf, *res.dlogits)
        ctx.aux = res
        return res.total.clone()

    @staticmethod
    def backward(ctx, g):
        dvf, *dl = ctx.saved_tensors
        return (None, None, None, None, None, None, dvf * g) + tuple(d * g for d in dl)


def a2c_loss_autograd(logits, actions, rewards, values, vf, gamma=0.99, lambda_=1.0, entropy_coeff=0.01):
    """Differentiable total loss (agent.py:189-190) for use with a PyTorch policy."""
    return _A2CLossFn.apply(actions, rewards, values, gamma, lambda_, entropy_coeff, vf, *logits)


def multiple_process_rollout(rollout, gamma, lambda_=1.0):
    """Reference rl_utils.py:25-57 on GPU tensors: ``rollout`` is a dict with 'states' [B,T+1,...],
    'actions' [B,T,K], 'rewards' [B,T], 'values' [B,T,1], 'features' [B,T,2,lstm], and
    'conditions' or 'z'.  The bootstrap ``rollout['r']`` is the constant 0.0 as in the reference."""
    lib = N.lib()
    rewards = rollout['rewards'].to(torch.float32).contiguous()
    values = rollout['values'][:, :, 0].to(torch.float32).contiguous()
    B, T = rewards.shape
    dev = rewards.device
    adv = torch.empty((B, T), dtype=torch.float32, device=dev)       # returned to the caller: the only allocations here
    ret = torch.empty((B, T), dtype=torch.float32, device=dev)
    # reuse the fused entry point with a dummy 1-logit head to get (adv, r) only
    dummy = _scratch("mpr_dummy", (B, T, 1), torch.float32, dev, zero=True)
    acts = _scratch("mpr_acts", (B, T, 1), torch.int32, dev, zero=True)
    scal = _scratch("mpr_scal", (4,), torch.float32, dev)
    partial = _scratch("a2c_partial", (B * T * 4 + 2048,), torch.float32, dev)
    lp = (C.c_void_p * 1)(dummy.data_ptr())
    sizes = (C.c_int32 * 1)(1)
    N.check(lib.spiral_a2c_loss(lp, sizes, 1, N.ptr(acts), N.ptr(rewards), N.ptr(values), N.ptr(values), B, T,
                                float(gamma), float(lambda_), 0.0, N.ptr(adv), N.ptr(ret), N.ptr(scal), None, None,
                                N.ptr(partial), N.stream_ptr()), "spiral_a2c_loss")
    features = rollout['features'][:, 0]
    if 'conditions' in rollout and rollout['conditions'] is not None:
        batch_c, batch_z = rollout['conditions'], None
    else:
        batch_c, batch_z = None, rollout.get('z')
    return Batch(rollout['states'], rollout['actions'], adv, ret, features, batch_c, batch_z)
