"""GPU tests of the native acting path (csrc/policy_conv.cu, models/policy_native.py) against the fp32 PyTorch policy
(models/policy.py).  bf16 operands with fp32 accumulation: tolerances are bf16-level, the comparison is against the same
parameters rounded to bf16 where that isolates the kernel's arithmetic."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _mods():
    import spiral_b200 as sp
    from importlib import import_module
    return sp, import_module(sp.__name__ + ".models.policy"), import_module(sp.__name__ + ".models.policy_native")


def _bf(x):
    return x.to(torch.bfloat16).float()


@pytest.mark.parametrize("Cin,Cout,k,stride,pad,H", [(3, 32, 5, 1, 2, 64), (6, 32, 5, 1, 2, 64), (32, 32, 4, 2, 0, 64),
                                                    (32, 32, 4, 2, 0, 31), (32, 32, 4, 2, 0, 14), (32, 32, 3, 1, 1, 6),
                                                    (32, 32, 3, 1, 1, 8), (32, 1, 3, 1, 1, 64), (32, 1, 3, 1, 1, 32)])
def test_pconv_matches_torch_conv(Cin, Cout, k, stride, pad, H):
    _, _, pn = _mods()
    torch.manual_seed(Cin * 100 + H)
    B = 3
    conv = torch.nn.Conv2d(Cin, Cout, k, stride=stride, padding=pad).cuda()
    x = torch.randn(B, H, H, Cin, device="cuda")
    res = torch.randn(B, (H + 2 * pad - k) // stride + 1, (H + 2 * pad - k) // stride + 1, Cout, device="cuda")
    layer = pn._Conv(conv.weight, conv.bias, stride, (pad, pad), relu=True)
    Ho = res.shape[1]
    out = torch.empty((B, Ho, Ho, Cout), dtype=torch.bfloat16, device="cuda")
    post = torch.randn(B, Cout, device="cuda")
    layer(x.to(torch.bfloat16).contiguous(), out, Ho, Ho, residual=res.to(torch.bfloat16).contiguous(), post_add=post)
    with torch.no_grad():
        ref = F.conv2d(_bf(x).permute(0, 3, 1, 2), _bf(conv.weight), conv.bias, stride=stride, padding=pad).permute(0, 2, 3, 1)
        ref = F.relu(ref + _bf(res)) + post[:, None, None, :]
    err = (out.float() - ref).abs().max() / ref.abs().max()
    assert float(err) < 1e-2, float(err)              # output rounding to bf16: 2^-9


@pytest.mark.parametrize("Cin,H", [(16, 4), (32, 8), (32, 16), (32, 32)])
def test_deconv_classes_match_torch_conv_transpose(Cin, H):
    _, _, pn = _mods()
    torch.manual_seed(H)
    B = 2
    dc = torch.nn.ConvTranspose2d(Cin, 32, 4, stride=2, padding=1).cuda()
    x = torch.randn(B, H, H, Cin, device="cuda")
    layer = pn._Deconv(dc.weight, dc.bias, relu=True)
    out = torch.empty((B, 2 * H, 2 * H, 32), dtype=torch.bfloat16, device="cuda")
    layer(x.to(torch.bfloat16).contiguous(), out)
    with torch.no_grad():
        ref = F.relu(F.conv_transpose2d(_bf(x).permute(0, 3, 1, 2), _bf(dc.weight), dc.bias, stride=2, padding=1)).permute(0, 2, 3, 1)
    err = (out.float() - ref).abs().max() / ref.abs().max()
    assert float(err) < 1e-2, float(err)


@pytest.mark.parametrize("envname,L,C,cond", [("color_mnist", 64, 3, False), ("simple_mnist", 32, 1, False),
                                             ("simple_mnist", 32, 1, True)])
def test_native_actor_agrees_with_the_pytorch_policy(envname, L, C, cond):
    sp, pol, pn = _mods()
    args = sp.get_args(["--env", envname, "--location_size", str(L), "--color_channel", str(C), "--episode_length", "3",
                        "--loss", "l2" if cond else "gan", "--policy_tf32", "False"])
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    B = 6
    env = sp.create_env(args, num_envs=B)
    torch.manual_seed(0)
    pi = pol.Policy(args, env).cuda()
    actor = pn.NativeActor(pi)
    state, condition, z = env.reset()
    rng = np.random.RandomState(0)
    acts = np.stack([rng.randint(n, size=B) for n in env.head_sizes], -1).astype(np.int32)
    env.step(acts)                                    # a painted observation
    ob = env.state.clone()
    ac = torch.tensor(acts, device="cuda")
    with torch.no_grad():
        want_enc = pi.encode(ob.unsqueeze(1), ac.unsqueeze(1).float(), None if condition is None else condition.unsqueeze(1),
                             None if z is None else z.unsqueeze(1))[:, 0]
        got_enc = actor.encode(ob, ac, condition, z)
        assert float((got_enc - want_enc).abs().max() / want_enc.abs().max()) < 3e-2
        zz = F.relu(torch.randn(B, args.lstm_size, device="cuda"))
        for name in pi.acs:
            if name in actor.loc:
                want = pi.heads[name](zz)
                got = actor.location_logits(name, zz)
                assert got.shape == want.shape
                assert float((got - want).abs().max() / want.abs().max()) < 3e-2, name
    c, h = pi.get_initial_features(B)
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    a, v, nc, nh = actor.act(ob, ac, c, h, condition, z, g)
    assert a.shape == (B, len(env.acs)) and a.dtype == torch.int32 and v.shape == (B,)
    for i, n in enumerate(env.head_sizes):
        assert int(a[:, i].min()) >= 0 and int(a[:, i].max()) < n
    env.step(a)                                       # the sampled actions drive the env
    env.check_status()


def test_agent_rollout_with_native_actor_sees_weight_updates():
    """The (graph-captured) rollout acts through NativeActor; after a learner step sync() refreshes the packed kernels in
    place, so the next replay uses the new weights."""
    sp, pol, pn = _mods()
    from importlib import import_module
    agent_mod = import_module(sp.__name__ + ".agent")
    args = sp.get_args(["--env", "simple_mnist", "--color_channel", "1", "--episode_length", "3", "--loss", "gan",
                        "--disc_dim", "8", "--disc_precision", "fp32", "--policy_act_native", "True", "--cuda_graphs", "True"])
    env = sp.create_env(args, num_envs=8)
    ag = agent_mod.Agent(args, env)
    assert ag.actor is not None
    ro = ag.rollout()
    assert ag._graph is not None and not ag._graph_failed
    before = ag.actor.x_enc.wpk.clone()
    ptr = ag.actor.x_enc.wpk.data_ptr()
    ag.train_policy(ro)
    assert ag.actor.x_enc.wpk.data_ptr() == ptr                      # same buffer (the graph points at it)
    with torch.no_grad():
        for p_ in ag.policy.parameters():
            p_.add_(0.05 * torch.randn_like(p_))                    # a visible update
    ag.actor.sync()
    assert not torch.equal(ag.actor.x_enc.wpk, before)
    ro2 = ag.rollout()
    env.reset()
    for t in range(3):
        assert torch.equal(env.state, ro2["states"][:, t])
        env.step(ro2["actions"][:, t])
    assert torch.equal(env.state, ro2["states"][:, 3])
