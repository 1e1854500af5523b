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


def _learner_case(envname, L, C, T, conditional, B, seed=5):
    import spiral_b200 as sp
    from importlib import import_module
    from helpers import make_args
    pol = import_module(sp.__name__ + ".models.policy")
    args = make_args(envname, L=L, C=C, T=T, conditional=conditional, extra=["--policy_tf32", "False"])
    env = sp.create_env(args, num_envs=B)
    torch.manual_seed(3)
    pi = pol.Policy(args, env).cuda()
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    K = len(pi.acs)
    x = torch.randint(0, 256, (B, T + 1, 64, 64, C), generator=g, device="cuda").float()
    ac = torch.stack([torch.randint(0, env.head_sizes[i], (B, T), generator=g, device="cuda") for i in range(K)], -1)
    cond = torch.randint(0, 256, (B, T, 64, 64, C), generator=g, device="cuda").float() if conditional else None
    z = None if conditional else torch.rand((B, T, args.z_dim), generator=g, device="cuda") - 0.5
    st = [torch.randn((B, args.lstm_size), generator=g, device="cuda") * 0.1 for _ in range(2)]
    samples = {n: ac[:, :, i] for i, n in enumerate(pi.acs)}
    return pi, (x, ac.float(), st, cond, z, samples), g


@pytest.fixture
def strict_fp32():
    """TF32 off for the PyTorch reference path of a test, restored afterwards (other tests keep the process defaults)."""
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved


def _run_learner(pi, inp, native, dtype=torch.float32):
    x, ac, st, cond, z, samples = inp
    f = lambda t_: None if t_ is None else t_.to(dtype)
    pi.learn_native = native
    pi.zero_grad(set_to_none=True)
    logits, _, vf, _ = pi(f(x), f(ac), (f(st[0]), f(st[1])), f(cond), f(z), samples)
    gg = torch.Generator(device="cuda")
    gg.manual_seed(11)
    loss = (vf * torch.randn(vf.shape, generator=gg, device="cuda").to(dtype)).sum()
    for n in pi.acs:
        loss = loss + (logits[n] * torch.randn(logits[n].shape, generator=gg, device="cuda").to(dtype)).sum()
    loss.backward()
    return ({n: logits[n].detach().double() for n in pi.acs}, vf.detach().double(),
            {k: p.grad.detach().double() for k, p in pi.named_parameters() if p.grad is not None})


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


CASES = [("simple_mnist", 32, 1, 2, False, 3), ("color_mnist", 64, 3, 2, False, 2), ("simple_mnist", 32, 1, 2, True, 2)]


class _PinnedRelu:
    """F.relu with the activation pattern of a recorded run.  A ReLU pre-activation within rounding of zero passes the
    gradient in one arithmetic and blocks it in another; one such unit moves whole gradient tensors by 1e-3 .. 5e-2 -- for
    PyTorch's own fp32 path against float64 exactly as for the kernels -- and on 64 x 64 maps (a million units) about one
    unit per run sits that close.  Recording the masks in the float64 run and replaying them in the fp32 runs evaluates the
    same linear branch of the network everywhere, so what is compared is arithmetic.  (The forward value changes by the
    pre-activation of the flipped unit, ~1e-7.)  The module's 4-D tensors are NCHW, the kernels' NHWC."""

    def __init__(self):
        self.masks, self.i, self.record, self.nhwc = [], 0, True, False

    def replay(self, nhwc):
        self.i, self.record, self.nhwc = 0, False, nhwc

    def __call__(self, x, inplace=False):
        if self.record:
            m = x > 0
            self.masks.append(m)
        else:
            m = self.masks[self.i]
            self.i += 1
            if m.dim() == 4 and self.nhwc:
                m = m.permute(0, 2, 3, 1)
            assert m.shape == x.shape, (self.i, tuple(m.shape), tuple(x.shape))
        return x * m.to(x.dtype)


def _check_learner(monkeypatch, envname, L, C, T, conditional, B, smooth=False):
    import copy
    pi, inp, _ = _learner_case(envname, L, C, T, conditional, B)
    if smooth:
        # ReLUs replaced by softplus: no kinks, nothing pinned
        monkeypatch.setattr(torch.nn.functional, "relu", lambda x, inplace=False: torch.nn.functional.softplus(x))
    # forward: logits and value of the plain network (no pinning: the forward has no kink problem)
    with torch.no_grad():
        f64 = _forward_only(copy.deepcopy(pi).double(), inp, False, torch.float64)
        f32 = _forward_only(pi, inp, False)
        fnat = _forward_only(pi, inp, True)
    for n in pi.acs:
        assert _rel(fnat[0][n], f64[0][n]) < 1e-4 and _rel(fnat[0][n], f32[0][n]) < 1e-4, n
    assert _rel(fnat[1], f64[1]) < 1e-4 and _rel(fnat[1], f32[1]) < 1e-4
    # gradients, with the activation pattern pinned to the float64 run's (the kernels' fused ReLU epilogue is switched off so
    # that every ReLU goes through F.relu; test_learner_fused_relu_is_bit_identical ties the fused path to this one)
    import spiral_b200 as sp
    from importlib import import_module
    monkeypatch.setattr(import_module(sp.__name__ + ".models.policy_learn"), "FUSE_RELU", False)
    pinned = _PinnedRelu()
    if not smooth:
        monkeypatch.setattr(torch.nn.functional, "relu", pinned)
    ref = _run_learner(copy.deepcopy(pi).double(), inp, False, torch.float64)       # the module in float64: ground truth
    n_relu = len(pinned.masks)
    pinned.replay(nhwc=False)
    t32 = _run_learner(pi, inp, False)                                              # PyTorch fp32 (cuDNN / cuBLAS, TF32 off)
    assert pinned.i == n_relu
    pinned.replay(nhwc=True)
    got = _run_learner(pi, inp, True)                                               # csrc/policy_learn.cu
    assert pinned.i == n_relu and (smooth or n_relu > 40)
    assert set(got[2]) == set(ref[2]) and len(ref[2]) > 60
    for n in pi.acs:
        assert _rel(got[0][n], ref[0][n]) < 1e-4, n
    w64 = max((_rel(got[2][k], ref[2][k]), k) for k in ref[2])
    w32 = max((_rel(t32[2][k], ref[2][k]), k) for k in ref[2])
    print("worst parameter-gradient error vs float64: native %.2e (%s), PyTorch fp32 %.2e (%s)" % (w64 + w32))
    # EVERY gradient tensor within 1e-4 of float64 -- or, for the few small ill-conditioned ones (the action encoder's
    # first layers, at the far end of the backward pass) where PyTorch's fp32 path itself is at 0.5e-4 .. 1e-4, within 3x of
    # what that path achieves on the same tensor; and never past 3e-4
    for k in ref[2]:
        e64, e32 = _rel(got[2][k], ref[2][k]), _rel(t32[2][k], ref[2][k])
        assert e64 < max(1e-4, min(3 * e32, 3e-4)), (k, e64, e32)
    assert w64[0] < 4 * max(w32[0], 1e-5)                     # and as good as the fp32 module itself, give or take


def _forward_only(pi, inp, native, dtype=torch.float32):
    x, ac, st, cond, z, samples = inp
    f = lambda t_: None if t_ is None else t_.to(dtype)
    pi.learn_native = native
    logits, _, vf, _ = pi(f(x), f(ac), (f(st[0]), f(st[1])), f(cond), f(z), samples)
    return {n: logits[n].double() for n in pi.acs}, vf.double()


@pytest.mark.parametrize("envname,L,C,T,conditional,B", CASES)
def test_learner_forward_and_gradients_match_the_module(monkeypatch, envname, L, C, T, conditional, B, strict_fp32):
    """Policy(learn_native=True): every convolution of the trunk and the location heads, and its input and weight
    gradients, on csrc/policy_learn.cu (default operand format f16x3) against the same module evaluated in float64:
    logits of every head and the value within 1e-4 of float64 AND of PyTorch's fp32 path (cuDNN / cuBLAS, TF32 off); EVERY
    parameter gradient of a loss that touches all heads within 1e-4 (relative to the largest entry of each tensor) of
    float64 with the ReLU activation pattern pinned to the float64 run's (_PinnedRelu says why), and within a small factor
    of what PyTorch's fp32 path achieves under the same pinning."""
    _check_learner(monkeypatch, envname, L, C, T, conditional, B)


def test_learner_gradients_on_a_smooth_network(monkeypatch, strict_fp32):
    """The same comparison with the ReLUs replaced by softplus (no kinks for rounding to flip): every forward, input-
    gradient and weight-gradient kernel in its place in the graph."""
    _check_learner(monkeypatch, "color_mnist", 64, 3, 2, False, 2, smooth=True)


def test_learner_fused_relu_is_bit_identical(monkeypatch, strict_fp32):
    """ReLU in the convolution's epilogue + its mask from the saved output (policy_learn.FUSE_RELU, the default) against
    F.relu around the same kernels: logits, value and every parameter gradient bit for bit."""
    import spiral_b200 as sp
    from importlib import import_module
    pl = import_module(sp.__name__ + ".models.policy_learn")
    pi, inp, _ = _learner_case("color_mnist", 64, 3, 2, False, 2)
    monkeypatch.setattr(pl, "FUSE_RELU", True)
    fused = _run_learner(pi, inp, True)
    monkeypatch.setattr(pl, "FUSE_RELU", False)
    plain = _run_learner(pi, inp, True)
    for n in pi.acs:
        assert torch.equal(fused[0][n], plain[0][n]), n
    assert torch.equal(fused[1], plain[1])
    assert set(fused[2]) == set(plain[2])
    for k in plain[2]:
        assert torch.equal(fused[2][k], plain[2][k]), k


def test_learner_convolution_primitives_match_torch(strict_fp32):
    """Each geometry on its own (forward, dgrad, wgrad): 5x5 'same' on 1/3/6 channels, 4x4 / stride 2 'valid' on odd and even
    maps, 3x3 'same', 3x3 -> 1 channel, transposed 4x4 / stride 2 with 16 and 32 input channels; ragged tile edges."""
    import spiral_b200 as sp
    from importlib import import_module
    pl = import_module(sp.__name__ + ".models.policy_learn")
    g = torch.Generator(device="cuda")
    g.manual_seed(0)

    def check(kind, cin, cout, k, stride, pad, H, W, B=3):
        if kind == "conv":
            mod = torch.nn.Conv2d(cin, cout, k, stride=stride, padding=pad).cuda()
        else:
            mod = torch.nn.ConvTranspose2d(cin, cout, 4, stride=2, padding=1).cuda()
        x = torch.randn((B, H, W, cin), generator=g, device="cuda")
        outs = []
        for native in (False, True):
            xx = x.clone().requires_grad_(True)
            mod.zero_grad(set_to_none=True)
            if native:
                y = pl.conv2d(xx, mod, stride=stride, pad=pad) if kind == "conv" else pl.conv_transpose2d(xx, mod)
            else:
                y = mod(xx.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
            gy = torch.randn(y.shape, generator=torch.Generator(device="cuda").manual_seed(1), device="cuda")
            (y * gy).sum().backward()
            outs.append((y.detach(), xx.grad.clone(), mod.weight.grad.clone(), mod.bias.grad.clone()))
        for a, b, name in zip(outs[1], outs[0], ("y", "dx", "dw", "db")):
            assert a.shape == b.shape, (kind, name)
            err = float((a - b).abs().max() / b.abs().max())
            assert err < (5e-5 if pl.PLANES != 2 else 2e-4), (kind, cin, cout, k, stride, H, W, name, err)

    saved = pl.PLANES
    try:
        for fmt in ("f16x3", "bf16x6", "bf16x3"):
            pl.PLANES = pl.FORMATS[fmt]
            _v2_shapes(check)
    finally:
        pl.PLANES = saved


def _v2_shapes(check):
    for cin in (1, 2, 3, 6):                                  # image channels: padded to one 8-channel chunk while staged
        check("conv", cin, 32, 5, 1, 2, 64, 64, B=2)
    check("conv", 1, 32, 3, 1, 1, 33, 20, B=3)
    check("conv", 32, 32, 4, 2, 0, 64, 64, B=2)
    check("conv", 32, 32, 4, 2, 0, 31, 31)
    check("conv", 32, 32, 4, 2, 0, 14, 14)
    check("conv", 32, 32, 3, 1, 1, 6, 6, B=5)
    check("conv", 32, 32, 3, 1, 1, 8, 8)
    check("conv", 32, 1, 3, 1, 1, 32, 32)
    check("conv", 32, 1, 3, 1, 1, 64, 64, B=2)
    check("convT", 16, 32, 4, 2, 1, 4, 4)
    check("convT", 32, 32, 4, 2, 1, 8, 8)
    check("convT", 32, 32, 4, 2, 1, 32, 32, B=2)
    check("conv", 32, 32, 3, 1, 1, 4, 4, B=9)                 # eight 4 x 4 frames per CTA and a ragged last group
    check("conv", 32, 32, 3, 1, 1, 16, 16, B=2)
    check("conv", 32, 16, 4, 2, 0, 18, 18, B=3)               # 16 output channels: one n-tile pair
    check("conv", 16, 32, 3, 1, 1, 12, 9, B=3)                # 16 input channels, non-square ragged map
    check("conv", 8, 32, 3, 1, 1, 9, 20, B=3)                 # 8 input channels: two taps per k-step, odd tap count (k padding)
