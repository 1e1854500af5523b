"""End-to-end smoke of the batched loop (rollout -> GAN reward -> fused A2C loss -> policy update,
WGAN-GP step) and TF-semantics checks of the PyTorch policy."""
import numpy as np
import pytest
import torch

from helpers import make_args

pytestmark = pytest.mark.gpu


def _mods():
    import spiral_b200 as sp
    from importlib import import_module
    return sp, import_module(sp.__name__ + ".agent"), import_module(sp.__name__ + ".models.policy")


def test_policy_shapes_and_teacher_forcing_is_deterministic():
    sp, agent_mod, pol = _mods()
    args = make_args("simple_mnist", L=32, C=1, T=3)
    env = sp.create_env(args, num_envs=4)
    pi = pol.Policy(args, env).cuda()
    B, T = 4, 3
    x = torch.rand(B, T + 1, 64, 64, 1, device="cuda") * 255
    ac = torch.randint(0, 2, (B, T, 3), device="cuda")
    z = torch.rand(B, T, 10, device="cuda") - 0.5
    st = pi.get_initial_features(B)
    samples = {n: ac[:, :, i] for i, n in enumerate(pi.acs)}
    l1, s1, v1, _ = pi(x, ac.float(), st, None, z, samples)
    l2, s2, v2, _ = pi(x, ac.float(), st, None, z, samples)
    assert pi.acs == ["jump", "control", "end"]
    assert l1["jump"].shape == (B, T, 2) and l1["control"].shape == (B, T, 1024) and v1.shape == (B, T)
    assert all(torch.equal(l1[k], l2[k]) for k in l1) and torch.equal(v1, v2)
    # BasicLSTMCell: forget bias 1.0 => with zero weights c' = c * sigmoid(1) + 0.5 * tanh(0)
    cell = pol._BasicLSTMCell(4, 4).cuda()
    with torch.no_grad():
        cell.kernel.weight.zero_(); cell.kernel.bias.zero_()
    c0 = torch.ones(1, 4, device="cuda")
    h, (c, _) = cell(torch.zeros(1, 4, device="cuda"), (c0, torch.zeros(1, 4, device="cuda")))
    assert torch.allclose(c, c0 * torch.sigmoid(torch.tensor(1.0))) and torch.allclose(h, torch.tanh(c) * 0.5)
    # EXTENSION: location_size 64 gets a third transposed convolution -> 4096 logits
    args64 = make_args("color_mnist", L=64, C=3, T=2)
    env64 = sp.create_env(args64, num_envs=2)
    pi64 = pol.Policy(args64, env64).cuda()
    a, v, c, h = pi64.act(env64.reset()[0], torch.full((2, 6), -1, device="cuda"), *pi64.get_initial_features(2),
                          None, torch.zeros(2, 10, device="cuda"))
    assert a.shape == (2, 6) and int(a[:, pi64.acs.index("end")].max()) < 4096


@pytest.mark.parametrize("loss", ["gan", "l2"])
def test_agent_loop_runs_and_updates_parameters(loss):
    sp, agent_mod, _ = _mods()
    args = make_args("simple_mnist", L=32, C=1, T=3, conditional=(loss == "l2"),
                     extra=["--disc_dim", "8", "--disc_batch_size", "8", "--disc_precision", "fp32"])
    env = sp.create_env(args, num_envs=8)
    ag = agent_mod.Agent(args, env)
    before = [p.detach().clone() for p in ag.policy.parameters()]
    ro = ag.rollout()
    assert ro["states"].shape == (8, 4, 64, 64, 1) and ro["actions"].shape == (8, 3, 3)
    assert float(ro["states"][:, 0].min()) == 255.0
    m = ag.train_policy(ro)
    assert torch.isfinite(m["model/total_loss"]) and torch.isfinite(m["model/grad_global_norm"])
    changed = sum(float((a - b.detach()).abs().sum()) for a, b in zip(before, ag.policy.parameters()))
    assert changed > 0
    if loss == "gan":
        assert float(ro["rewards"][:, -1].min()) > 0 and float(ro["rewards"][:, -1].max()) < 1
        dw = ag.disc.params["dense/kernel"].detach().clone()
        g = ag.train_gan()
        assert all(torch.isfinite(v) for v in g.values())
        assert float((ag.disc.params["dense/kernel"].detach() - dw).abs().sum()) > 0
        # fused reward forward and the differentiable layer-wise forward read the same parameters
        x = ro["states"][:, -1]
        with torch.no_grad():
            pt, _ = ag.disc.logits_fn((x - 127.5) / 127.5)
        assert torch.allclose(ag.disc.predict(x), pt, atol=1e-4)
    else:
        assert float(ro["rewards"][:, :-1].abs().max()) == 0.0 and float(ro["rewards"][:, -1].max()) <= 1.0
