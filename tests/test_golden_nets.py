"""Golden vectors of the discriminator reward, the WGAN-GP loss and the fused A2C loss (tests/golden/nets_*.npz,
made by tests/golden/make_golden_nets.py from the float64 oracles).  CPU: the oracles still reproduce them (pins the
oracle).  GPU: the kernels reproduce them through the C ABI without evaluating the oracle."""
import glob
import os
import sys

import numpy as np
import pytest
import torch

from helpers import GOLDEN

sys.path.insert(0, GOLDEN)
import make_golden_nets as mk  # noqa: E402

DISC = sorted(glob.glob(os.path.join(GOLDEN, "nets_disc_*.npz")))
A2C = sorted(glob.glob(os.path.join(GOLDEN, "nets_a2c_*.npz")))


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize("path", DISC)
def test_disc_oracle_reproduces_golden(path):
    g = np.load(path)
    got = mk.disc_case(int(g["C"]), int(g["dd"]), int(g["B"]), int(g["seed"]))
    for k in ("probs", "logits", "d_loss", "critic_loss", "gradient_penalty", "grad_conv0_kernel", "grad_dense_kernel"):
        assert _rel(got[k], g[k]) < 1e-9, k


@pytest.mark.parametrize("path", A2C)
def test_a2c_oracle_reproduces_golden(path):
    g = np.load(path)
    got = mk.a2c_case([int(v) for v in g["sizes"]], int(g["B"]), int(g["T"]), int(g["seed"]))
    for k in ("r", "adv", "scalars", "dvf", "dlogits0_sub", "dlogits_last_sum"):
        assert np.allclose(got[k], g[k], rtol=1e-6, atol=1e-7), k


@pytest.mark.gpu
@pytest.mark.parametrize("path", DISC)
def test_disc_kernels_reproduce_golden(path):
    import spiral_b200 as sp
    from importlib import import_module

    from oracle import disc_oracle as do          # only for the seeded weight / image generators
    models = import_module(sp.__name__ + ".models")
    g = np.load(path)
    C, dd, B, seed = int(g["C"]), int(g["dd"]), int(g["B"]), int(g["seed"])
    rng = np.random.RandomState(seed)
    w = do.init_weights(C, disc_dim=dd, batch_norm=True, seed=seed + 1)
    fake, real = mk.canvases(rng, B, C), mk.canvases(rng, B, C)
    alpha = rng.rand(B, 1).astype(np.float32)
    prec = "bf16x3" if dd >= 64 else "fp32"
    args = sp.get_args(["--color_channel", str(C), "--disc_dim", str(dd), "--disc_precision", prec])
    D = models.Discriminator(args, None, [64, 64, C], norm_fn=lambda x: x, scope_name="global", max_batch=B)
    D.load_numpy(w)
    probs, logits = D.forward(torch.tensor(fake, device="cuda"))
    assert np.abs(logits.cpu().numpy() - g["logits"]).max() <= 1e-4 * max(1.0, np.abs(g["logits"]).max())
    assert np.abs(probs.cpu().numpy() - g["probs"]).max() <= 1e-4
    for p in D.params.values():
        p.requires_grad_(True)
    out = D.wgan_gp_loss(torch.tensor(fake, device="cuda"), torch.tensor(real, device="cuda"),
                         torch.tensor(alpha, device="cuda"), 20.0)
    out["d_loss"].backward()
    for k in ("d_loss", "critic_loss", "gradient_penalty"):
        assert abs(float(out[k].detach()) - float(g[k])) <= 1e-4 * max(1.0, abs(float(g[k]))), k
    assert _rel(D.params["conv0/kernel"].grad.cpu().numpy(), g["grad_conv0_kernel"]) < 1e-3
    assert _rel(D.params["dense/kernel"].grad.cpu().numpy(), g["grad_dense_kernel"]) < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("path", A2C)
def test_a2c_kernels_reproduce_golden(path):
    import spiral_b200 as sp
    from importlib import import_module
    rl = import_module(sp.__name__ + ".rl_utils")
    g = np.load(path)
    sizes, B, T, seed = [int(v) for v in g["sizes"]], int(g["B"]), int(g["T"]), int(g["seed"])
    rng = np.random.RandomState(seed)
    logits = [rng.randn(B, T, n).astype(np.float32) * 2 for n in sizes]
    acts = np.stack([rng.randint(n, size=(B, T)) for n in sizes], -1).astype(np.int32)
    rewards = np.zeros((B, T), np.float32)
    rewards[:, -1] = rng.rand(B)
    values = rng.randn(B, T).astype(np.float32)
    vf = rng.randn(B, T).astype(np.float32)
    res = rl.a2c_loss([torch.tensor(l, device="cuda") for l in logits], torch.tensor(acts, device="cuda"),
                      torch.tensor(rewards, device="cuda"), torch.tensor(values, device="cuda"),
                      torch.tensor(vf, device="cuda"), 0.99, 1.0, 0.01)
    assert _rel(res.r.cpu().numpy(), g["r"]) < 1e-6 and _rel(res.adv.cpu().numpy(), g["adv"]) < 1e-6
    got = res.scalars.cpu().numpy()
    for i in range(4):
        assert abs(got[i] - g["scalars"][i]) <= 2e-5 * max(1.0, abs(g["scalars"][i])), i
    assert _rel(res.dvf.cpu().numpy(), g["dvf"]) < 2e-5
    assert _rel(res.dlogits[0].cpu().numpy()[:, ::4, ::32], g["dlogits0_sub"]) < 2e-5
    assert np.allclose(res.dlogits[-1].cpu().numpy().sum(-1), g["dlogits_last_sum"], atol=2e-4)
