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


def test_fused_clip_adam_matches_reference_formulas():
    """K5 (agent.py:231-239): clip-by-global-norm + TensorFlow-style Adam against the plain PyTorch
    restatements (tests/helpers.TFAdam, distributed.clip_by_global_norm_) on identical gradients."""
    import spiral_b200 as sp
    from importlib import import_module
    from helpers import TFAdam
    optim = import_module(sp.__name__ + ".optim")
    du = import_module(sp.__name__ + ".distributed")
    g = torch.Generator(device="cuda").manual_seed(0)
    # 61 tensors: more than one launch; one of 2.5 * 2^20 + 3 elements: three pieces of the 2^20-element table entries
    shapes = [(5, 5, 3, 64), (64,), (1152, 256), (7,), (300, 33)] * 12 + [(5 * (1 << 19) + 3,)]
    pa = [torch.randn(s, generator=g, device="cuda") for s in shapes]
    pb = [p.clone() for p in pa]
    ref = TFAdam(pb, 1e-3, beta1=0.5, beta2=0.9)
    fused = optim.FusedTFAdam(pa, 1e-3, beta1=0.5, beta2=0.9)
    world = 4
    for step in range(3):
        grads = [torch.randn(s, generator=g, device="cuda") * (10.0 if step == 1 else 0.01) for s in shapes]
        for p, q, gr in zip(pa, pb, grads):
            p.grad = gr.clone()
            q.grad = gr.clone() / world
        want_norm = du.clip_by_global_norm_([q.grad for q in pb], 40.0)
        ref.step()
        got_norm = fused.step(clip_norm=40.0, grad_scale=1.0 / world)
        assert abs(float(got_norm) - float(want_norm)) <= 1e-5 * float(want_norm)
        for p, q in zip(pa, pb):
            assert torch.allclose(p, q, rtol=1e-5, atol=1e-7)


def test_cuda_graph_rollout_matches_the_env_and_draws_fresh_actions():
    """The captured rollout (policy.act + K4 sampling + K1 env.step x T in one CUDA graph) must (1) record a
    trajectory that the env reproduces from the recorded actions alone, (2) draw new random numbers on every
    replay, (3) agree in distribution with the eager loop."""
    import spiral_b200 as sp
    from importlib import import_module
    agent_mod = import_module(sp.__name__ + ".agent")
    args = sp.get_args(["--env", "simple_mnist", "--color_channel", "1", "--episode_length", "4", "--loss", "gan",
                        "--disc_dim", "8", "--disc_precision", "fp32", "--cuda_graphs", "True"])
    B = 16
    env = sp.create_env(args, num_envs=B)
    ag = agent_mod.Agent(args, env)
    ro1 = ag.rollout()
    assert ag._graph is not None and not ag._graph_failed, "the rollout was not captured"
    ro2 = ag.rollout()
    assert not torch.equal(ro1["actions"], ro2["actions"]), "replay reused the random numbers"
    assert not torch.equal(ro1["z"], ro2["z"])
    for ro in (ro1, ro2):
        env.reset()
        for t in range(4):
            assert torch.equal(env.state, ro["states"][:, t])
            env.step(ro["actions"][:, t])
        assert torch.equal(env.state, ro["states"][:, 4])
        assert int(ro["actions"].min()) >= 0
    m = ag.train_policy(ro2)
    assert torch.isfinite(m["model/total_loss"])


def test_checkpoint_resume_restores_weights_and_optimizer(tmp_path):
    """trainer.py:23-29,66-80: the chief saves the global variables; the same --load_path resumes them."""
    import spiral_b200 as sp
    from importlib import import_module
    agent_mod = import_module(sp.__name__ + ".agent")
    ck = import_module(sp.__name__ + ".checkpoint")
    args = sp.get_args(["--env", "simple_mnist", "--color_channel", "1", "--episode_length", "3", "--loss", "gan",
                        "--disc_dim", "8", "--disc_precision", "fp32", "--cuda_graphs", "False"])
    env = sp.create_env(args, num_envs=8)
    ag = agent_mod.Agent(args, env)
    ag.train_policy(ag.rollout())
    ag.train_gan()
    path = ck.save_checkpoint(ag, str(tmp_path), step=5)
    assert ck.latest_checkpoint(str(tmp_path)) == path
    want_p = [p.detach().clone() for p in ag.policy.parameters()]
    want_d = {k: v.detach().clone() for k, v in ag.disc.params.items()}
    snap = dict(replay=ag.replay.data.clone(), idx=ag.replay.idx, lstm=ag._lstm_state.clone(),
                ctr=int(ag.policy._sample_counter_dev.item()), rng=ag.rng.get_state().clone(), env_rng=ag.env._rng.get_state().clone())
    x = ag.rollout()["states"][:, -1]
    want_reward = ag.disc.predict(x).clone()
    ag2 = agent_mod.Agent(args, sp.create_env(args, num_envs=8))
    assert ck.load_checkpoint(ag2, str(tmp_path)) == 5
    assert all(torch.equal(a, b) for a, b in zip(want_p, ag2.policy.parameters()))
    assert all(torch.equal(want_d[k], ag2.disc.params[k]) for k in want_d)
    assert ag2.policy_opt.t == ag.policy_opt.t and torch.equal(ag2.disc_opt.m[0], ag.disc_opt.m[0])
    assert torch.equal(ag2.disc.predict(x), want_reward)              # the packed kernel weights were refreshed
    # the acting path's packed (bf16) copies of the policy's convolution kernels were re-packed from the restored weights:
    # the first rollout after a resume must act with the saved network, not the fresh random one
    assert ag.actor is not None and ag2.actor is not None
    ob = x.float()
    last = torch.zeros((8, 3), dtype=torch.int32, device="cuda")
    c = torch.randn((8, args.lstm_size), device="cuda")
    h = torch.randn((8, args.lstm_size), device="cuda")
    z = torch.rand((8, args.z_dim), device="cuda") - 0.5
    e1, e2 = ag.actor.encode(ob, last, None, z), ag2.actor.encode(ob, last, None, z)
    assert torch.equal(e1, e2)
    fresh = agent_mod.Agent(args, sp.create_env(args, num_envs=8))
    assert not torch.equal(e1, fresh.actor.encode(ob, last, None, z))
    # the random streams, the replay ring and the actors' carried LSTM state continue where the saved run stopped
    assert torch.equal(ag2.replay.data, snap["replay"]) and ag2.replay.idx == snap["idx"]
    assert torch.equal(ag2._lstm_state, snap["lstm"]) and float(snap["lstm"].abs().max()) > 0
    assert int(ag2.policy._sample_counter_dev.item()) == snap["ctr"] > 0
    assert torch.equal(ag2.rng.get_state(), snap["rng"]) and torch.equal(ag2.env._rng.get_state(), snap["env_rng"])


@pytest.mark.parametrize("native", [False, True])
def test_graph_captured_learner_step_matches_eager_steps(native, monkeypatch):
    """The CUDA graph of train_policy (captured at the third call) must update the policy exactly like eager steps:
    two agents with identical weights are fed the same rollouts for five learner steps; the graph reads each step's
    bias-corrected Adam step size from device memory (FusedTFAdam.advance)."""
    sp, agent_mod, _ = _mods()
    # deterministic cuDNN algorithms and strict fp32 in both agents: what is left between the two runs is the order of
    # floating-point sums inside the kernels, not a different choice of algorithm under capture
    monkeypatch.setattr(torch.backends.cudnn, "deterministic", True)
    monkeypatch.setattr(torch.backends.cudnn, "benchmark", False)

    def make(graphs):
        args = make_args("simple_mnist", L=32, C=1, T=3, conditional=True,
                         extra=["--cuda_graphs", str(graphs), "--policy_act_native", str(native), "--policy_lr", "1e-4",
                                "--policy_tf32", "False"])
        env = sp.create_env(args, num_envs=8)
        torch.manual_seed(0)
        return agent_mod.Agent(args, env)

    a_graph, a_eager = make(True), make(False)
    a_eager.policy.load_state_dict(a_graph.policy.state_dict())
    start = [p.detach().clone() for p in a_graph.policy.parameters()]
    rollouts = []
    for i in range(5):
        ro = a_eager.rollout()
        ro["rewards"][:, -1] = torch.linspace(0, 1, 8, device="cuda") * (i + 1) / 5.0     # distinct, deterministic rewards
        rollouts.append({k: v.clone() for k, v in ro.items()})
    for i, ro in enumerate(rollouts):
        m_g = a_graph.train_policy({k: v.clone() for k, v in ro.items()})
        m_e = a_eager.train_policy({k: v.clone() for k, v in ro.items()})
        # cuDNN may pick other algorithms under capture (and its backward kernels use atomics), and Adam turns last-bit
        # differences of near-zero gradients into lr-sized steps, so the two trajectories agree closely, not bitwise;
        # the step size is kept small enough (1e-4; 1e-3 let the two runs drift apart by the fifth step on some boxes)
        # that this noise is not amplified by the loss surface
        assert torch.allclose(m_g["model/total_loss"], m_e["model/total_loss"], rtol=5e-3), i
        assert torch.allclose(m_g["model/grad_global_norm"], m_e["model/grad_global_norm"], rtol=5e-2), i
    assert a_graph._tp_graph is not None and not a_graph._tp_failed           # steps 3-5 were graph replays
    assert a_graph.policy_opt.t == a_eager.policy_opt.t == 5
    moved = sum(float((p.detach() - s0).pow(2).sum()) for p, s0 in zip(a_graph.policy.parameters(), start)) ** 0.5
    apart = sum(float((p.detach() - q.detach()).pow(2).sum())
                for p, q in zip(a_graph.policy.parameters(), a_eager.policy.parameters())) ** 0.5
    assert moved > 0 and apart < 0.05 * moved, (moved, apart)


def test_graph_captured_discriminator_step_matches_eager_steps():
    """Same check for train_gan: identical batches through a graph-replaying and an eager agent for five WGAN-GP steps."""
    sp, agent_mod, _ = _mods()

    def make(graphs):
        args = make_args("simple_mnist", L=32, C=1, T=3,
                         extra=["--cuda_graphs", str(graphs), "--disc_dim", "8", "--disc_batch_size", "8",
                                "--disc_precision", "fp32", "--disc_lr", "1e-3"])
        env = sp.create_env(args, num_envs=8)
        torch.manual_seed(0)
        return agent_mod.Agent(args, env)

    a_graph, a_eager = make(True), make(False)
    for k, v in a_graph.disc.params.items():
        a_eager.disc.params[k].data.copy_(v.data)
    a_eager.disc.sync_weights()
    start = {k: v.detach().clone() for k, v in a_graph.disc.params.items()}
    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    for i in range(5):
        fake = torch.rand((8, 64, 64, 1), generator=g, device="cuda") * 255
        real = torch.rand((8, 64, 64, 1), generator=g, device="cuda") * 255
        for ag in (a_graph, a_eager):
            ag.replay.sample = lambda n, f=fake: f          # same fakes for both
            ag.rng.manual_seed(100 + i)                     # same interpolation factors
        m_g, m_e = a_graph.train_gan(real), a_eager.train_gan(real)
        assert torch.allclose(m_g["gan/disc_loss"], m_e["gan/disc_loss"], rtol=2e-2, atol=1e-3), i
    assert a_graph._tg_graph is not None and not a_graph._tg_failed
    assert a_graph.disc_opt.t == a_eager.disc_opt.t == 5
    moved = sum(float((v.detach() - start[k]).pow(2).sum()) for k, v in a_graph.disc.params.items()) ** 0.5
    apart = sum(float((v.detach() - a_eager.disc.params[k].detach()).pow(2).sum())
                for k, v in a_graph.disc.params.items()) ** 0.5
    assert moved > 0 and apart < 0.05 * moved, (moved, apart)


@pytest.mark.parametrize("graphs", [False, True])
def test_actor_lstm_state_carries_over_and_features_are_per_step(graphs):
    """rl_utils.py:185-238: an actor fetches the initial LSTM features once, before its loop, so episode n+1 starts from
    the state episode n ended in; PartialRollout stores the features each step STARTED from.  Checked on the rollout
    (eager and CUDA-graph replay): features[:, 0] of the second episode is the final state of the first, and stepping the
    LSTM by hand from features[:, t] with the recorded frames and actions gives features[:, t + 1]."""
    sp, agent_mod, _ = _mods()
    args = make_args("simple_mnist", L=32, C=1, T=3, extra=["--cuda_graphs", str(graphs), "--policy_act_native", "False",
                                                             "--policy_tf32", "False", "--disc_dim", "8", "--disc_precision", "fp32"])
    env = sp.create_env(args, num_envs=6)
    ag = agent_mod.Agent(args, env)
    ro1 = ag.rollout()
    end1 = ag._lstm_state.clone()
    ro2 = ag.rollout()
    assert float(ro1["features"][:, 0].abs().max()) == 0.0                 # get_initial_features: zeros, once
    assert torch.equal(ro2["features"][:, 0], end1)                        # carried over, never reset
    assert float(end1.abs().max()) > 0
    pi = ag.policy
    init = torch.as_tensor(env.initial_action, device="cuda").float()
    for ro in (ro1, ro2):
        T = ro["actions"].shape[1]
        for t in range(T):
            last = init if t == 0 else ro["actions"][:, t - 1].float()
            with torch.no_grad():                          # the acting path (policy.act runs under no_grad)
                enc = pi.encode(ro["states"][:, t:t + 1].float(), last.unsqueeze(1), None, ro["z"][:, t:t + 1])
                _, (c, h) = pi.lstm(enc[:, 0], (ro["features"][:, t, 0], ro["features"][:, t, 1]))
            want = ro["features"][:, t + 1] if t + 1 < T else (end1 if ro is ro1 else ag._lstm_state)
            assert torch.allclose(c, want[:, 0], atol=1e-5) and torch.allclose(h, want[:, 1], atol=1e-5), t


def test_rgb_trajectory_states_are_uint8_and_lossless():
    """color_channel 3 observations are the read-back uint8 image: the rollout keeps `states` as uint8 (4x smaller) and the
    env reproduces exactly those frames from the recorded actions; the learner step consumes them."""
    sp, agent_mod, _ = _mods()
    args = make_args("color_mnist", L=64, C=3, T=3, extra=["--disc_dim", "8", "--disc_precision", "fp32", "--disc_batch_size", "4"])
    env = sp.create_env(args, num_envs=4)
    ag = agent_mod.Agent(args, env)
    ro = ag.rollout()
    assert ro["states"].dtype == torch.uint8 and ro["states"].shape == (4, 4, 64, 64, 3)
    env.reset()
    for t in range(3):
        assert torch.equal(env.state, ro["states"][:, t].float())
        env.step(ro["actions"][:, t])
    assert torch.equal(env.state, ro["states"][:, 3].float())
    m = ag.train_policy(ro)
    assert torch.isfinite(m["model/total_loss"])
    gray = agent_mod.Agent(make_args("simple_mnist", L=32, C=1, T=2, conditional=True), sp.create_env(make_args("simple_mnist", L=32, C=1, T=2, conditional=True), num_envs=2))
    assert gray.rollout()["states"].dtype == torch.float32                # BT.601 gray values are not integers


def test_replay_buffer_on_the_device_follows_the_reference():
    """replay.py:8-37 restated in numpy against the device-resident ring on the GPU (pushes of uint8-truncated float
    canvases, the shift-left overflow rule, sampling from [0, idx))."""
    import types
    sp, _, _ = _mods()
    from importlib import import_module
    rb = import_module(sp.__name__ + ".replay")
    args = types.SimpleNamespace(buffer_batch_num=6, disc_batch_size=8, seed=5)
    shape = [64, 64, 3]
    buf = rb.ReplayBuffer(args, shape, device="cuda")
    size = 48
    data = np.zeros([size] + shape, np.uint8)
    idx = 0
    rng = np.random.RandomState(1)
    for step in range(25):
        n = int(rng.randint(1, 20))
        batch = rng.uniform(0, 255.99, size=[n] + shape).astype(np.float32)
        if idx + n > size:
            data[:-n] = data[n:]
            data[-n:] = batch
        else:
            data[idx:idx + n] = batch
            idx += n
        buf.push(torch.from_numpy(batch).cuda())
        assert buf.idx == idx and (buf.data.cpu().numpy() == data).all(), step
    out = buf.sample(8).cpu().numpy()
    rows = {r.tobytes() for r in data[:idx].astype(np.float32)}
    assert out.shape == (8, 64, 64, 3) and all(r.tobytes() in rows for r in out)


def test_discriminator_variables_carry_tensorflow_names():
    """tl.batch_normalization is unnamed in the reference (discriminator.py:76-79): TensorFlow names the first one in the
    scope 'batch_normalization', then 'batch_normalization_1', ...; conv<l>/{kernel,bias}, dense/{kernel,bias}."""
    sp, _, _ = _mods()
    from importlib import import_module
    models = import_module(sp.__name__ + ".models")
    args = sp.get_args(["--color_channel", "3", "--disc_dim", "8", "--disc_precision", "fp32"])
    D = models.Discriminator(args, None, [64, 64, 3], norm_fn=lambda x: x, max_batch=4)
    names = set(D.params)
    want = {"conv%d/%s" % (l, k) for l in range(6) for k in ("kernel", "bias")} | {"dense/kernel", "dense/bias"}
    want |= {"batch_normalization/gamma", "batch_normalization/beta"}
    want |= {"batch_normalization_%d/%s" % (i, k) for i in range(1, 5) for k in ("gamma", "beta")}
    assert names == want


def test_grad_reducer_path_matches_the_plain_path_on_one_rank(monkeypatch):
    """distributed.GradReducer (flat gradient buffer, per-bucket all-reduce from backward hooks) is what several ranks use;
    on one rank (SPIRAL_GRAD_REDUCER=1) it must update the networks like the plain path, eagerly and under the CUDA graph."""
    sp, agent_mod, _ = _mods()

    def make(reducer):
        if reducer:
            monkeypatch.setenv("SPIRAL_GRAD_REDUCER", "1")
        else:
            monkeypatch.delenv("SPIRAL_GRAD_REDUCER", raising=False)
        args = make_args("simple_mnist", L=32, C=1, T=3, extra=["--cuda_graphs", "True", "--policy_act_native", "False",
                                                                 "--disc_dim", "8", "--disc_batch_size", "8", "--disc_precision", "fp32",
                                                                 "--policy_lr", "1e-4"])
        env = sp.create_env(args, num_envs=8)
        torch.manual_seed(0)
        return agent_mod.Agent(args, env)

    a_red, a_plain = make(True), make(False)
    assert a_red._pol_red is not None and a_red._disc_red is not None and a_plain._pol_red is None
    a_plain.policy.load_state_dict(a_red.policy.state_dict())
    for k, v in a_red.disc.params.items():
        a_plain.disc.params[k].data.copy_(v.data)
    a_plain.disc.sync_weights()
    start = [p.detach().clone() for p in a_red.policy.parameters()]
    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    for i in range(5):
        ro = a_plain.rollout()
        ro["rewards"][:, -1] = torch.linspace(0, 1, 8, device="cuda")
        a_red.replay.push(ro["states"][:, -1])
        m1 = a_red.train_policy({k: v.clone() for k, v in ro.items()})
        m2 = a_plain.train_policy({k: v.clone() for k, v in ro.items()})
        assert torch.allclose(m1["model/grad_global_norm"], m2["model/grad_global_norm"], rtol=2e-1), i
        fake = torch.rand((8, 64, 64, 1), generator=g, device="cuda") * 255
        real = torch.rand((8, 64, 64, 1), generator=g, device="cuda") * 255
        for ag in (a_red, a_plain):
            ag.replay.sample = lambda n, f=fake: f
            ag.rng.manual_seed(100 + i)
        d1, d2 = a_red.train_gan(real), a_plain.train_gan(real)
        assert torch.allclose(d1["gan/disc_loss"], d2["gan/disc_loss"], rtol=2e-2, atol=1e-3), i
    assert a_red._tp_graph is not None and a_red._tg_graph is not None          # captured with the hooks in place
    for p in a_red.policy.parameters():                                          # gradients still live in the flat buffer
        assert a_red._pol_red.flat.data_ptr() <= p.grad.data_ptr() < a_red._pol_red.flat.data_ptr() + a_red._pol_red.collective_bytes()
    moved = sum(float((p.detach() - s0).pow(2).sum()) for p, s0 in zip(a_red.policy.parameters(), start)) ** 0.5
    apart = sum(float((p.detach() - q.detach()).pow(2).sum())
                for p, q in zip(a_red.policy.parameters(), a_plain.policy.parameters())) ** 0.5
    assert moved > 0 and apart < 0.05 * moved, (moved, apart)
