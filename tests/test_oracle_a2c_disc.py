"""Known answers for the A2C / discriminator oracles."""
import numpy as np
import torch

from oracle import a2c_oracle as ao
from oracle import disc_oracle as do


def test_discount_known_answer():
    out = ao.discount(np.array([[0.0, 0.0, 1.0]]), 0.99)
    assert np.allclose(out, [[0.9801, 0.99, 1.0]], rtol=0, atol=1e-15)


def test_rollout_bootstrap_zero():
    rng = np.random.RandomState(0)
    rew = np.zeros((3, 5)); rew[:, -1] = rng.rand(3)
    val = rng.randn(3, 5)
    r, adv = ao.multiple_process_rollout(rew, val, 0.99, 1.0)
    # lambda = 1: adv = discounted return - value
    assert np.allclose(adv, r - val, atol=1e-12)
    assert np.allclose(r[:, -1], rew[:, -1])
    assert np.allclose(r[:, 0], 0.99 ** 4 * rew[:, -1])


def test_loss_identity_value_term_once_per_head():
    rng = np.random.RandomState(1)
    B, T, sizes = 2, 3, [2, 7, 16]
    logits = [rng.randn(B, T, n) for n in sizes]
    acts = np.stack([rng.randint(n, size=(B, T)) for n in sizes], -1)
    adv, r, vf = rng.randn(B, T), rng.randn(B, T), rng.randn(B, T)
    out = ao.a2c_loss(logits, acts, adv, r, vf, 0.01)
    K = len(sizes)
    assert np.isclose(out["vf_loss"], K * 0.5 * np.sum((vf - r) ** 2))
    assert np.isclose(out["total"], out["pi_loss"] + 0.25 * K * np.sum((vf - r) ** 2) - 0.01 * out["entropy"])
    assert np.allclose(out["dvf"], 0.5 * K * (vf - r))
    # uniform logits: entropy = B*T*log(N)
    out = ao.a2c_loss([np.zeros((B, T, 8))], np.zeros((B, T, 1), int), adv, r, vf)
    assert np.isclose(out["entropy"], B * T * np.log(8))


def test_loss_gradient_rows_sum_to_zero_and_logit_shift_invariance():
    """Size-independent properties of agent.py:179-190 (they hold for any batch, so they are the checks that scale):
    softmax terms do not see a constant added to a row of logits, and d total / d logits sums to zero over every row:
    adv (sum p - 1) + coeff sum p (logp + entropy) = 0."""
    rng = np.random.RandomState(4)
    B, T, sizes = 3, 4, [2, 8, 33, 1024]
    logits = [rng.randn(B, T, n) * 3 for n in sizes]
    acts = np.stack([rng.randint(n, size=(B, T)) for n in sizes], -1)
    adv, r, vf = rng.randn(B, T), rng.randn(B, T), rng.randn(B, T)
    out = ao.a2c_loss(logits, acts, adv, r, vf, 0.01)
    for g in out["dlogits"]:
        assert np.abs(np.asarray(g).sum(-1)).max() < 1e-12
    shifted = [l + rng.randn(B, T, 1) * 10 for l in logits]
    out2 = ao.a2c_loss(shifted, acts, adv, r, vf, 0.01)
    for k in ("pi_loss", "vf_loss", "entropy", "total"):
        assert np.isclose(out[k], out2[k], rtol=1e-12, atol=1e-10), k
    for g, g2 in zip(out["dlogits"], out2["dlogits"]):
        assert np.allclose(g, g2, rtol=0, atol=1e-12)


def test_disc_shape_chain_and_same_padding():
    w = do.init_weights(1, disc_dim=8)
    assert [w["conv%d/kernel" % l].shape[-1] for l in range(6)] == [8, 16, 32, 64, 128, 256]
    assert w["dense/kernel"].shape == (8 * 32, 1)
    # SAME padding for k=5, s=2: output pixel (0,0) sees input rows/cols -1..3 -> kernel tap (1,1)
    # multiplies input pixel (0,0)
    x = torch.zeros(1, 64, 64, 1)
    x[0, 0, 0, 0] = 1.0
    k = np.zeros((5, 5, 1, 8), np.float32)
    k[1, 1, 0, :] = 3.0
    ww = dict(w)
    ww["conv0/kernel"] = k
    ww["conv0/bias"] = np.zeros(8, np.float32)
    xx = torch.nn.functional.pad(x.permute(0, 3, 1, 2), (1, 2, 1, 2))
    y = torch.nn.functional.conv2d(xx, torch.tensor(k).permute(3, 2, 0, 1), stride=2)
    assert y.shape == (1, 8, 32, 32) and float(y[0, 0, 0, 0]) == 3.0 and float(y.abs().sum()) == 24.0


def test_disc_bn_uses_batch_statistics_biased_variance():
    w = do.init_weights(1, disc_dim=8)
    rng = np.random.RandomState(0)
    imgs = rng.randint(0, 256, (4, 64, 64, 1)).astype(np.float32)
    x = do.norm_fn(torch.tensor(imgs))
    _, _, stats = do.build_model(x, w, return_stats=True)
    assert len(stats) == 5
    p4, _ = do.predict(imgs, w)
    p2, _ = do.predict(imgs[:2], w)
    assert not np.allclose(p4[:2], p2)        # rewards depend on batch composition (training-mode BN)
    assert ((p4 > 0) & (p4 < 1)).all()


def test_disc_loss_runs_and_penalty_is_on_probs():
    w = do.init_weights(1, disc_dim=8)
    rng = np.random.RandomState(0)
    fake = rng.randint(0, 256, (4, 64, 64, 1)).astype(np.float32)
    real = rng.randint(0, 256, (4, 64, 64, 1)).astype(np.float32)
    out = do.d_loss(fake, real, rng.rand(4, 1), w)
    assert np.isclose(out["d_loss"], out["critic_loss"] + 20.0 * out["gradient_penalty"])
    assert 0.0 < out["gradient_penalty"] < 4.0
    assert all(np.isfinite(g).all() for g in out["grads"].values())
