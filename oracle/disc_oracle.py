"""oracle/disc_oracle.py -- TEST INFRASTRUCTURE (CPU oracle), NOT product code.

PyTorch-CPU restatement of the reference discriminator with TensorFlow-1.6 semantics
(SURVEY.md Appendix B), following models/discriminator.py:
  :32-38   norm_fn, conditional concat           :61-81  conv 5x5 s2 'same' + BN(training) + lrelu
  :83-93   flatten, dense(1), sigmoid            :97-122 WGAN-GP loss with the penalty on PROBS
  :157-163 predict (BN statistics of the scored batch)
PARITY UNPINNED against TensorFlow 1.6 itself (not installable here).  Pinned by known answers
(shape chain, SAME padding = 1 before / 2 after, BN eps 1e-3 with biased variance).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def init_weights(in_channels, disc_dim=64, screen_size=64, batch_norm=True, seed=11):
    """he_normal conv kernels (HWIO), zero biases, BN gamma=1/beta=0 perturbed so tests see them,
    glorot_normal dense; dict of float32 numpy arrays."""
    g = np.random.RandomState(seed)
    n_layers = int(math.log(screen_size, 2))
    w = {}
    cin = in_channels
    for l in range(n_layers):
        cout = disc_dim * 2 ** l
        fan_in = 25 * cin
        w["conv%d/kernel" % l] = (g.randn(5, 5, cin, cout) * math.sqrt(2.0 / fan_in)).astype(np.float32)
        w["conv%d/bias" % l] = (0.05 * g.randn(cout)).astype(np.float32)
        if l > 0 and batch_norm:
            w["bn%d/gamma" % l] = (1.0 + 0.1 * g.randn(cout)).astype(np.float32)
            w["bn%d/beta" % l] = (0.1 * g.randn(cout)).astype(np.float32)
        cin = cout
    w["dense/kernel"] = (g.randn(cin, 1) * math.sqrt(2.0 / (cin + 1))).astype(np.float32)
    w["dense/bias"] = (0.05 * g.randn(1)).astype(np.float32)
    return w


def norm_fn(x):
    return (x - 127.5) / 127.5            # envs/base.py:51-52


def build_model(x_nhwc, w, batch_norm=True, dtype=torch.float32, return_stats=False):
    """x_nhwc: [N,S,S,C] already normalised.  Returns (probs, logits)."""
    x = x_nhwc.permute(0, 3, 1, 2).to(dtype)
    n_layers = int(math.log(x.shape[2], 2))
    stats = []
    for l in range(n_layers):
        k = torch.as_tensor(w["conv%d/kernel" % l]).to(dtype).permute(3, 2, 0, 1)   # HWIO -> OIHW
        b = torch.as_tensor(w["conv%d/bias" % l]).to(dtype)
        x = F.pad(x, (1, 2, 1, 2))                      # TF 'SAME', k=5, s=2, even input
        x = F.conv2d(x, k, b, stride=2)
        if l > 0 and batch_norm:
            mean = x.mean(dim=(0, 2, 3), keepdim=True)
            var = x.var(dim=(0, 2, 3), unbiased=False, keepdim=True)
            stats.append((mean.flatten(), var.flatten()))
            gamma = torch.as_tensor(w["bn%d/gamma" % l]).to(dtype).view(1, -1, 1, 1)
            beta = torch.as_tensor(w["bn%d/beta" % l]).to(dtype).view(1, -1, 1, 1)
            x = (x - mean) * torch.rsqrt(var + 1e-3) * gamma + beta
        x = F.leaky_relu(x, 0.2)
    x = x.flatten(1)
    logits = (x @ torch.as_tensor(w["dense/kernel"]).to(dtype) + torch.as_tensor(w["dense/bias"]).to(dtype)).reshape(-1)
    probs = torch.sigmoid(logits)
    if return_stats:
        return probs, logits, stats
    return probs, logits


def predict(images_0_255, w, batch_norm=True, dtype=torch.float32):
    """discriminator.py:157-163: real_probs of the fed batch."""
    x = norm_fn(torch.as_tensor(np.asarray(images_0_255)).to(dtype))
    probs, logits = build_model(x, w, batch_norm, dtype)
    return probs.numpy(), logits.numpy()


def d_loss(fake_0_255, real_0_255, alpha, w, batch_norm=True, wgan_lambda=20.0, conditional=False,
           dtype=torch.float64):
    """discriminator.py:97-122.  alpha: [N,1] uniform numbers (tf.random_uniform in the reference)."""
    wt = {k: torch.tensor(v, dtype=dtype, requires_grad=True) for k, v in w.items()}
    fake = norm_fn(torch.as_tensor(np.asarray(fake_0_255)).to(dtype))
    real = norm_fn(torch.as_tensor(np.asarray(real_0_255)).to(dtype))
    if conditional:
        fake = torch.cat([fake, real], -1)
        real = torch.cat([real, real], -1)
    _, real_logits = build_model(real, wt, batch_norm, dtype)
    _, fake_logits = build_model(fake, wt, batch_norm, dtype)
    g_loss = -fake_logits.mean()
    critic = fake_logits.mean() - real_logits.mean()
    a = torch.as_tensor(np.asarray(alpha)).to(dtype)
    fake_data, real_data = fake.flatten(1), real.flatten(1)
    interpolates = (real_data + a * (fake_data - real_data)).detach().requires_grad_(True)
    diff_probs, _ = build_model(interpolates.reshape(fake.shape), wt, batch_norm, dtype)
    grads = torch.autograd.grad(diff_probs.sum(), interpolates, create_graph=True)[0]
    slopes = torch.sqrt(1e-8 + (grads ** 2).sum(1))
    gp = ((slopes - 1.0) ** 2).mean()
    loss = critic + wgan_lambda * gp
    loss.backward()
    return dict(d_loss=float(loss), g_loss=float(g_loss), critic_loss=float(critic), gradient_penalty=float(gp),
                grads={k: v.grad.numpy() for k, v in wt.items()})
