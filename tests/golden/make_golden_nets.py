"""Generates tests/golden/nets_*.npz: known answers of the discriminator reward, the WGAN-GP loss and the fused
A2C loss from the float64 CPU oracles (oracle/disc_oracle.py, oracle/a2c_oracle.py).  The reference itself cannot
run here (no TensorFlow 1.6), so these are ORACLE outputs, committed so that CUDA regressions are caught on the
GPU box without evaluating the oracle, and so that the oracle itself is pinned against drift.
    python tests/golden/make_golden_nets.py
Inputs are regenerated from seeds (numpy RandomState is stable); only the answers are stored."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import a2c_oracle as ao  # noqa: E402
from oracle import disc_oracle as do  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def canvases(rng, B, C):
    imgs = np.full((B, 64, 64, C), 255.0, np.float32)
    for b in range(B):
        for _ in range(6):
            y, x = rng.randint(0, 56, 2)
            imgs[b, y:y + 8, x:x + 8] = rng.randint(0, 255)
    return imgs


def disc_case(C, dd, B, seed):
    rng = np.random.RandomState(seed)
    w = do.init_weights(C, disc_dim=dd, batch_norm=True, seed=seed + 1)
    fake, real = canvases(rng, B, C), canvases(rng, B, C)
    alpha = rng.rand(B, 1).astype(np.float32)
    probs, logits = do.predict(fake, w, dtype=torch.float64)
    loss = do.d_loss(fake, real, alpha, w, wgan_lambda=20.0)
    return dict(C=C, dd=dd, B=B, seed=seed, probs=probs, logits=logits,
                d_loss=loss["d_loss"], critic_loss=loss["critic_loss"], gradient_penalty=loss["gradient_penalty"],
                grad_conv0_kernel=loss["grads"]["conv0/kernel"], grad_dense_kernel=loss["grads"]["dense/kernel"])


def a2c_case(sizes, B, T, seed):
    rng = np.random.RandomState(seed)
    logits = [rng.randn(B, T, n).astype(np.float32) * 2 for n in sizes]
    acts = np.stack([rng.randint(n, size=(B, T)) for n in sizes], -1).astype(np.int32)
    rewards = np.zeros((B, T), np.float32)
    rewards[:, -1] = rng.rand(B)
    values = rng.randn(B, T).astype(np.float32)
    vf = rng.randn(B, T).astype(np.float32)
    r, adv = ao.multiple_process_rollout(rewards, values, 0.99, 1.0)
    want = ao.a2c_loss(logits, acts, adv.astype(np.float32), r.astype(np.float32), vf, 0.01)
    return dict(sizes=np.array(sizes), B=B, T=T, seed=seed, r=r, adv=adv,
                scalars=np.array([want[k] for k in ("pi_loss", "vf_loss", "entropy", "total")], np.float64),
                dvf=want["dvf"], dlogits0_sub=want["dlogits"][0][:, ::4, ::32], dlogits_last_sum=want["dlogits"][-1].sum(-1))


if __name__ == "__main__":
    for name, kw in (("nets_disc_c3_dd64", dict(C=3, dd=64, B=6, seed=21)), ("nets_disc_c1_dd8", dict(C=1, dd=8, B=5, seed=22))):
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **disc_case(**kw))
        print(name)
    for name, kw in (("nets_a2c_rgb64", dict(sizes=[4096, 4096, 8, 2, 2, 2], B=3, T=20, seed=31)),
                     ("nets_a2c_mnist", dict(sizes=[2, 1024, 1024], B=5, T=3, seed=32))):
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **a2c_case(**kw))
        print(name)
