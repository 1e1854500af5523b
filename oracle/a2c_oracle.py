"""oracle/a2c_oracle.py -- TEST INFRASTRUCTURE (CPU oracle), NOT product code.

numpy/scipy restatement of the reference's return/advantage computation and A2C loss:
  rl_utils.py:18-20   discount            (scipy.signal.lfilter on the reversed axis)
  rl_utils.py:25-57   multiple_process_rollout  (bootstrap rollout['r'] is the constant 0.0)
  agent.py:161-196    per-head log-softmax policy-gradient / value / entropy loss (sums)
Gradients are derived by torch autograd on the same expression (float64), standing in for
tf.gradients.  PARITY UNPINNED against TensorFlow 1.6 itself (not installable here); the loss is
elementary fp32/fp64 arithmetic, pinned by the known answers in tests/test_oracle_a2c.py.
"""
import numpy as np
import scipy.signal
import torch


def discount(x, gamma):
    return scipy.signal.lfilter([1], [1, -gamma], x[:, ::-1], axis=1)[:, ::-1]


def multiple_process_rollout(rewards, values, gamma=0.99, lambda_=1.0, r=0.0):
    """rewards, values: [B,T] -> (batch_r [B,T], batch_adv [B,T]) in float64 like numpy."""
    rewards = np.asarray(rewards, np.float64)
    values = np.asarray(values, np.float64)
    B = rewards.shape[0]
    rcol = np.full((B, 1), r)
    vpred_t = np.hstack([values, rcol])
    rewards_plus_v = np.hstack([rewards, rcol])
    batch_r = discount(rewards_plus_v, gamma)[:, :-1]
    delta_t = rewards + gamma * vpred_t[:, 1:] - vpred_t[:, :-1]
    batch_adv = discount(delta_t, gamma * lambda_)
    return batch_r, batch_adv


def a2c_loss(logits, actions, adv, r, vf, entropy_coeff=0.01, dtype=torch.float64, want_grads=True):
    """logits: list of [B,T,N_k]; actions [B,T,K] int; adv, r, vf [B,T].
    Returns dict(pi_loss, vf_loss, entropy, total, dlogits=[...], dvf) per agent.py:172-194."""
    lg = [torch.tensor(np.asarray(l), dtype=dtype, requires_grad=want_grads) for l in logits]
    vf_t = torch.tensor(np.asarray(vf), dtype=dtype, requires_grad=want_grads)
    adv_t = torch.tensor(np.asarray(adv), dtype=dtype)
    r_t = torch.tensor(np.asarray(r), dtype=dtype)
    acts = torch.tensor(np.asarray(actions), dtype=torch.int64)
    loss = 0
    pi_loss_sum, vf_loss_sum, ent_sum = 0, 0, 0
    for k, logit in enumerate(lg):
        ac = torch.nn.functional.one_hot(acts[..., k], logit.shape[-1]).to(dtype)
        log_prob = torch.log_softmax(logit, -1)
        prob = torch.softmax(logit, -1)
        pi_loss = -torch.sum(torch.sum(log_prob * ac, -1) * adv_t)
        vf_loss = 0.5 * torch.sum(torch.square(vf_t - r_t))
        entropy = -torch.sum(prob * log_prob)
        loss = loss + pi_loss + 0.5 * vf_loss - entropy * entropy_coeff
        pi_loss_sum = pi_loss_sum + pi_loss
        vf_loss_sum = vf_loss_sum + vf_loss
        ent_sum = ent_sum + entropy
    out = dict(pi_loss=float(pi_loss_sum.detach()), vf_loss=float(vf_loss_sum.detach()),
               entropy=float(ent_sum.detach()), total=float(loss.detach()))
    if want_grads:
        loss.backward()
        out["dlogits"] = [l.grad.numpy() for l in lg]
        out["dvf"] = vf_t.grad.numpy()
    return out
