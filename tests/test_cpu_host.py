"""CPU-only checks (-m "not gpu"): the C-ABI library loads and exports every symbol the header
declares, the product fails loudly without a CUDA device, and the host-side logic (brush
compilation, colours, dither table, action order, config, debug env, sharding + gloo all-reduce)
agrees with the oracle / the reference's documented behaviour."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import spiral_b200 as sp
from oracle import paint_oracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mod(name):
    from importlib import import_module
    return import_module(sp.__name__ + "." + name)


def test_library_exports_every_header_symbol():
    hdr = open(os.path.join(ROOT, "include", "spiral_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(spiral_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(sp._native.SYMBOLS), declared ^ set(sp._native.SYMBOLS)
    lib = sp._native.lib()                    # raises if the .so is missing or a symbol is absent
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.spiral_version() == sp._native.ABI_VERSION
    # struct layouts agree with the header's field lists (catches a forgotten field)
    for cname, cls in (("SpiralEnvCfg", sp._native.SpiralEnvCfg), ("SpiralDiscCfg", sp._native.SpiralDiscCfg),
                       ("SpiralBrush", sp._native.SpiralBrush), ("SpiralMapping", sp._native.SpiralMapping)):
        end = hdr.index("} %s;" % cname)
        body = hdr[hdr.rindex("typedef struct {", 0, end) + len("typedef struct {"):end]
        fields = re.findall(r"([a-z_0-9]+)(?:\[[^\]]*\])*\s*[,;]", body)
        assert fields == [f[0] for f in cls._fields_], cname


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_product_fails_loudly_without_cuda():
    args = sp.get_args(["--env", "simple_mnist", "--color_channel", "1"])
    with pytest.raises(Exception):
        sp.create_env(args, num_envs=2, device="cuda:0")


def test_no_oracle_import_in_product():
    """The oracle is test infrastructure: nothing under the product may import, link or execute it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "spiral-tensorflow_b200")):
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith(".py"):
                text = open(path).read()
                for needle in ("from oracle", "import oracle", "libpaint_oracle", "paint_oracle"):
                    assert needle not in text, (f, needle)
            elif f.endswith((".cu", ".cuh", ".h", ".sh")):
                text = open(path).read()
                assert not re.search(r'#include\s*[<"][^>"]*oracle', text), f
                assert "libpaint_oracle" not in text, f


def test_py2_action_order():
    base = _mod("envs.base")
    assert base.py2_dict_order(["jump", "control", "end"]) == ["jump", "control", "end"]
    assert base.py2_dict_order(["pressure", "jump", "size", "control", "end"]) == \
        ["jump", "pressure", "control", "end", "size"]
    assert base.py2_dict_order(["color", "shape"]) == ["color", "shape"]
    for names in (["pressure", "jump", "size", "color", "control", "end"], ["a", "bb", "ccc", "dddd", "e", "f", "g"]):
        assert base.py2_dict_order(names) == po.py2_dict_order(names)


def test_environment_bookkeeping():
    base = _mod("envs.base")

    class Dummy(base.Environment):
        action_sizes = [('pressure', [2]), ('jump', [2]), ('size', [2]), ('control', None), ('end', None)]

    args = sp.get_args(["--jump", "False", "--location_size", "8"])
    env = Dummy(args, num_envs=3)
    assert env.acs == ["pressure", "control", "end", "size"]          # 'jump' deleted, order kept
    assert env.action_sizes["control"] == [8, 8] and env.head_sizes == [2, 64, 64, 2]
    assert env.observation_shape == [64, 64, 3]
    a = env.random_action(np.random.RandomState(0))
    assert a.shape == (3, 4) and (a >= 0).all() and (a < np.array(env.head_sizes)).all()
    assert (env.initial_action == -1).all() and env.initial_action.shape == (3, 4)
    assert np.allclose(env.denorm(env.norm(np.array([0.0, 127.5, 255.0]))), [0.0, 127.5, 255.0])
    assert Dummy.action_sizes[1][0] == 'jump'                           # class attribute not mutated


def test_config_defaults_and_conditional_override():
    a = sp.get_args([])
    assert (a.screen_size, a.location_size, a.color_channel, a.episode_length) == (64, 32, 3, 5)
    assert (a.lstm_size, a.z_dim, a.disc_dim, a.wgan_lambda, a.entropy_coeff, a.grad_clip) == (256, 10, 64, 20, 0.01, 40)
    assert a.loss == "gan" and a.conditional is False                  # config.py:91-93
    assert sp.get_args(["--loss", "l2", "--conditional", "False"]).conditional is True
    assert sp.config.str2bool("True") and sp.config.str2bool("t") and not sp.config.str2bool("False")


def test_brush_compile_and_rejects_unsupported():
    info = sp.brush.BrushInfo()
    br = info.compile()
    assert abs(br.dyn[2].base - 0.6) < 1e-7 and br.dyn[2].n[2] == 2 and br.dyn[0].n[0] == 2
    assert br.dabs_per_actual_radius == 6.0 and br.slow_tracking == 2.0 and br.elliptical_dab_angle == 90.0
    assert sp.brush.BrushInfo(info.to_myb()).settings == info.settings   # text round trip
    assert po.parse_myb(info.to_myb()) == po.parse_myb(po.brush_to_myb())
    bad = sp.brush.BrushInfo()
    bad.set_base_value("smudge", 0.5)
    with pytest.raises(NotImplementedError):
        bad.compile()
    bad = sp.brush.BrushInfo(info.to_myb().replace("hardness 0.2", "hardness 0.2 | stroke (0.0 0.0), (1.0 0.5)"))
    with pytest.raises(NotImplementedError):
        bad.compile()


def test_colours_and_dither_table_match_oracle():
    assert (sp.brush.dither_table() == po.dither_table()).all()
    names = ["control", "end", "color", "jump", "pressure", "size"]
    env = po.OracleBatchEnv(po.EnvSpec(action_names=names, location_size=32, n_color=8), 1, channels=3)
    env.enable_logs()
    for ci in range(8):
        env.reset()
        env.step(np.array([[0, 100, ci, 0, 0, 0]]))
        env.pop_dabs(0)
        env.step(np.array([[300, 900, ci, 0, 0, 0]]))
        d = env.pop_dabs(0)
        assert len(d) > 10
        want = sp.brush.color_fix15(po.PALETTE[ci])
        assert (d["r"] == want[0]).all() and (d["g"] == want[1]).all() and (d["b"] == want[2]).all(), ci
    assert sp.brush.brush_color_fix15(sp.brush.BrushInfo()) == [0, 0, 0]


def test_simple_debug_env_runs_on_cpu():
    # BASELINE "debug" config: --env simple --episode_length=1 --location_size=8 --conditional=True --loss=l2
    args = sp.get_args(["--env", "simple", "--episode_length", "1", "--location_size", "8", "--loss", "l2"])
    env = sp.create_env(args, num_envs=2)
    assert env.acs == ["color", "shape"]
    state, target, z = env.reset()
    assert state.shape == (2, 64, 64, 3) and z is None and float(state.min()) == 1.0
    state, reward, terminal, _ = env.step(np.array([[0, 0], [0, 1]]))
    assert terminal and state.min() == -1.0
    # shape 0 = PIL ellipse, shape 1 = rectangle, both black, bbox (0,0,6,6) (simple.py:68,82-85)
    assert (state[1, :7, :7] == -1.0).all() and state[1, 7, 7, 0] == 1.0
    assert state[0, 3, 3, 0] == -1.0 and state[0, 0, 0, 0] == 1.0
    assert ((reward < 0) | (reward == 1.0)).all()                       # simple.py:132-133 override


def test_discount_numpy_mirror():
    rl = _mod("rl_utils")
    assert np.allclose(rl.discount(np.array([[0.0, 0.0, 1.0]]), 0.99), [[0.9801, 0.99, 1.0]])


def test_shard_partitions_cover_the_batch():
    dm = _mod("distributed")
    for total, world in ((1024, 8), (1000, 8), (5, 2), (7, 4)):
        spans = [dm.shard(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == total
        for (s0, c0), (s1, _) in zip(spans, spans[1:]):
            assert s0 + c0 == s1
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
import spiral_b200 as sp
from importlib import import_module
dm = import_module(sp.__name__ + ".distributed")
rank, world, _ = dm.init(backend="gloo")
assert world == 2
start, count = dm.shard(10, rank, world)
# per-rank "sum over the local shard" gradients of f(w) = sum_i w * x_i
x = torch.arange(10, dtype=torch.float32)[start:start + count]
grads = [x.sum().reshape(1) * torch.ones(3), torch.full((2, 2), float(rank + 1))]
dm.allreduce_grads_(grads, bucket_bytes=8)
assert torch.allclose(grads[0], torch.full((3,), 45.0 / 2)), grads[0]     # global sum / world
assert torch.allclose(grads[1], torch.full((2, 2), 1.5)), grads[1]
norm = dm.clip_by_global_norm_(grads, 10.0)
tot = torch.sqrt(sum((g ** 2).sum() for g in grads))
assert abs(float(tot) - 10.0) < 1e-4 and float(norm) > 10.0
# global-batch BN statistics from shards == statistics of the gathered batch
full = torch.arange(20, dtype=torch.float64).reshape(10, 2) ** 1.5
mine = full[start:start + count]
ss = torch.stack([mine.sum(0), (mine ** 2).sum(0)])
mean, var, n = dm.allreduce_bn_stats_(ss, mine.shape[0])
assert n == 10 and torch.allclose(mean, full.mean(0)) and torch.allclose(var, full.var(0, unbiased=False))
# GradReducer: per-bucket all-reduce launched from backward hooks into a persistent flat buffer == the plain sum
torch.manual_seed(0)
net = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Tanh(), torch.nn.Linear(8, 8), torch.nn.Tanh(), torch.nn.Linear(8, 1))
ref = [torch.zeros_like(p) for p in net.parameters()]
red = dm.GradReducer(net.parameters(), bucket_bytes=128)          # several buckets
assert len(red.buckets) > 2 and red.collective_bytes() == 4 * sum(p.numel() for p in net.parameters())
for step in range(3):
    xs = [torch.randn(5, 4, generator=torch.Generator().manual_seed(10 * step + r)) for r in range(world)]
    want = [torch.zeros_like(p) for p in net.parameters()]
    for xr in xs:                                                  # what every rank's backward contributes
        for w_, g in zip(want, torch.autograd.grad(net(xr).pow(2).sum(), list(net.parameters()))):
            w_ += g
    red.zero()
    net(xs[rank]).pow(2).sum().backward()
    red.wait()
    for p, w_ in zip(net.parameters(), want):
        assert p.grad.data_ptr() >= red.flat.data_ptr() and torch.allclose(p.grad, w_, rtol=1e-5, atol=1e-6), step
# a parameter that takes no part in the loss: its bucket is still reduced (zeros), nothing hangs
extra = torch.nn.Parameter(torch.ones(3))
red2 = dm.GradReducer(list(net.parameters()) + [extra], bucket_bytes=64)
red2.zero()
net(xs[rank]).sum().backward()
red2.wait()
assert float(extra.grad.abs().max()) == 0.0
# the global_bn hook of Discriminator.forward
sums = torch.full((4, 2), float(rank + 1), dtype=torch.float64)
cnt = torch.tensor([10.0 * (rank + 1)], dtype=torch.float64)
dm.allreduce_bn_sums_(sums, cnt)
assert float(sums[0, 0]) == 3.0 and float(cnt) == 30.0
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_gloo_world2_allreduce_and_sharding(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29517", CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                         capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


def test_philox_oracle_matches_random123_known_answers():
    """Pins oracle/sample_oracle.py's Philox4x32-10 (the RNG behind the K4 sampler) to Random123's KAT file."""
    from oracle import sample_oracle as so
    for ctr, key, want in (([0] * 4, [0] * 2, [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
                           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
                           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
                            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])):
        got = so.philox4x32_10(np.array([ctr], np.uint32), np.array([key], np.uint32))[0]
        assert [int(v) for v in got] == want


def test_sampler_oracle_distribution_is_softmax():
    from oracle import sample_oracle as so
    base = np.array([0.3, -1.2, 2.0, 0.0], np.float32)
    rows = 40000
    idx = so.categorical_sample(np.tile(base, (rows, 1)), seed=5, offset=1)
    freq = np.bincount(idx, minlength=4) / rows
    p = np.exp(base - base.max())
    p /= p.sum()
    assert rows * ((freq - p) ** 2 / p).sum() < 16.3          # chi-square, 3 dof, 99.9 %


def test_bench_reference_arm_prints_one_contract_line():
    """bench.py --impl reference (the CPU port of the reference path timed on the host cores) prints exactly
    one JSON line with the contract's keys; rank != 0 exits without work."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"].startswith("rgb64")
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=60, cwd=root, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_params_json_roundtrip(tmp_path):
    """utils/train.py:66-120: flags are saved sorted as JSON and restored on resume, except the skip list."""
    from importlib import import_module
    ck = import_module(sp.__name__ + ".checkpoint")
    a = sp.get_args(["--env", "simple_mnist", "--episode_length", "7", "--disc_dim", "8"])
    path = ck.save_args(a, str(tmp_path))
    assert os.path.basename(path) == "params.json"
    b = sp.get_args(["--env", "mnist"])
    changed = ck.load_args(b, str(tmp_path))
    assert b.env == "simple_mnist" and b.episode_length == 7 and b.disc_dim == 8
    assert "episode_length" in changed and "load_path" not in changed


def test_replay_buffer_follows_the_reference_push_semantics():
    """replay.py:8-37 restated literally in numpy (uint8 ring, `idx` only moves while a push fits) against the
    device-resident ReplayBuffer on random push sequences: same contents, same `idx`, float -> uint8 truncation."""
    import types
    from importlib import import_module

    import spiral_b200 as sp
    rb = import_module(sp.__name__ + ".replay")
    args = types.SimpleNamespace(buffer_batch_num=5, disc_batch_size=4, seed=3)
    shape = [3, 2, 1]
    buf = rb.ReplayBuffer(args, shape, device="cpu")
    size = args.buffer_batch_num * args.disc_batch_size
    data = np.zeros([size] + shape, np.uint8)
    idx = 0
    rng = np.random.RandomState(0)
    for step in range(40):
        n = int(rng.randint(1, 9))
        batch = rng.uniform(0, 255.99, size=[n] + shape).astype(np.float32)
        if idx + n > size:                       # replay.py:24-27
            data[:-n] = data[n:]
            data[-n:] = batch
        else:                                    # replay.py:28-30
            data[idx:idx + n] = batch
            idx += n
        buf.push(torch.from_numpy(batch))
        assert buf.idx == idx, step
        assert (buf.data.numpy() == data).all(), step
    out = buf.sample(64)
    assert out.dtype == torch.float32 and out.shape == (64, 3, 2, 1)
    filled = {tuple(r.ravel()) for r in data[:idx].astype(np.float32)}
    assert all(tuple(r.ravel()) in filled for r in out.numpy())          # uniform over [0, idx) only
    empty = rb.ReplayBuffer(args, shape, device="cpu")
    with pytest.raises(RuntimeError):
        empty.sample(4)


def test_batched_env_runner_follows_the_reference_env_runner():
    """rl_utils.batched_env_runner against a literal restatement of the reference's env_runner / PartialRollout
    (rl_utils.py:60-95, 185-238) run once per env: the LSTM features are fetched once before the loop and carried over
    from episode to episode, ``features[t]`` is the state BEFORE step t, ``states`` has the final state appended,
    ``last_action`` restarts from ``env.initial_action`` every episode."""
    rl = _mod("rl_utils")
    B, T, K, H = 3, 4, 2, 5

    class OneEnv(object):                       # a deterministic stand-in for envs/mnist.py: one canvas
        def __init__(self, b):
            self.b, self.t, self.ep = b, 0, 0
            self.episode_length, self.acs = T, ["a", "b"]
            self.initial_action = [-1] * K

        def reset(self):
            self.t = 0
            self.ep += 1
            self.obs = np.full((2, 2, 1), 10.0 * self.ep + self.b, np.float32)
            return self.obs, None, np.full((3,), 0.1 * self.ep + self.b, np.float32)

        def step(self, action):
            self.t += 1
            self.obs = self.obs + float(sum(action)) + 1.0
            return self.obs, 0.25 * self.t + self.b, self.t == T, {}

    def act_one(state, last_action, c, h, condition, z):      # policy.act: any deterministic function of its inputs
        s = float(np.mean(state)) + float(np.sum(last_action)) + float(np.sum(z))
        action = [int(abs(s) + float(np.sum(c))) % 3, int(abs(s) + float(np.sum(h))) % 5]
        return action, 0.5 * s + float(np.sum(h)), c * 0.5 + s * 1e-3, h * 0.25 + np.tanh(c + s * 1e-3)

    def reference_env_runner(env, episodes):                  # rl_utils.py:185-238, literally
        out = []
        last_state, condition, z = env.reset()
        last_features = (np.zeros(H, np.float32), np.zeros(H, np.float32))      # policy.get_initial_features
        for _ in range(episodes):
            ro = dict(states=[], actions=[], rewards=[], values=[], features=[], z=[])
            last_action = env.initial_action
            for _ in range(T):
                c, h = last_features
                action, value_, nc, nh = act_one(last_state, last_action, c, h, condition, z)
                state, reward, terminal, info = env.step(action)
                ro["states"] += [last_state]; ro["actions"] += [action]; ro["rewards"] += [reward]
                ro["values"] += [value_]; ro["features"] += [last_features]; ro["z"] += [z]
                last_state, last_action, last_features = state, action, (nc, nh)
            last_state, condition, z = env.reset()
            ro["states"] += [state]
            out.append(ro)
        return out

    want = [reference_env_runner(OneEnv(b), 3) for b in range(B)]

    class BatchEnv(object):                     # the same envs behind the batched API
        def __init__(self):
            self.envs = [OneEnv(b) for b in range(B)]
            self.episode_length, self.num_envs, self.acs = T, B, ["a", "b"]

        def reset(self, target=None, z=None):
            r = [e.reset() for e in self.envs]
            return torch.tensor(np.stack([x[0] for x in r])), None, torch.tensor(np.stack([x[2] for x in r]))

        def step(self, action):
            r = [e.step([int(v) for v in action[b]]) for b, e in enumerate(self.envs)]
            return torch.tensor(np.stack([x[0] for x in r])), torch.tensor([x[1] for x in r]), r[0][2], {}

    def act_batch(state, last_action, c, h, condition, z, generator):
        r = [act_one(state[b].numpy(), [int(v) for v in last_action[b]], c[b].numpy(), h[b].numpy(), None, z[b].numpy())
             for b in range(B)]
        return (torch.tensor([x[0] for x in r], dtype=torch.int32), torch.tensor([x[1] for x in r]),
                torch.tensor(np.stack([x[2] for x in r])), torch.tensor(np.stack([x[3] for x in r])))

    env = BatchEnv()
    lstm_state = torch.zeros((B, 2, H))
    init = torch.full((B, K), -1, dtype=torch.int32)
    for ep in range(3):
        got = rl.batched_env_runner(env, act_batch, lstm_state, init, H)
        for b in range(B):
            w = want[b][ep]
            assert np.allclose(got["states"][b].numpy(), np.stack(w["states"]))
            assert (got["actions"][b].numpy() == np.array(w["actions"])).all()
            assert np.allclose(got["rewards"][b].numpy(), w["rewards"]) and np.allclose(got["values"][b, :, 0].numpy(), w["values"])
            assert np.allclose(got["features"][b, :, 0].numpy(), np.stack([f[0] for f in w["features"]]), atol=1e-6)
            assert np.allclose(got["features"][b, :, 1].numpy(), np.stack([f[1] for f in w["features"]]), atol=1e-6)
            assert np.allclose(got["z"][b].numpy(), np.stack(w["z"]))
        if ep > 0:
            assert float(got["features"][:, 0].abs().max()) > 0      # the carried-over state, not zeros


def test_synthetic_targets_have_the_named_shapes():
    """SURVEY 8d / BASELINE configs: digit-shaped 28x28 -> 64x64 targets for gray canvases (dark strokes on white, like
    255 - resized MNIST, mnist.py:303-317), CelebA-shaped smooth colour fields for RGB ones; uint8, seed 7, deterministic."""
    tg = _mod("envs.targets")
    d = tg.synthetic_targets(1, count=64, seed=7)
    assert d.shape == (64, 64, 64, 1) and d.dtype == torch.uint8
    assert torch.equal(d, tg.synthetic_targets(1, count=64, seed=7))
    assert not torch.equal(d, tg.synthetic_targets(1, count=64, seed=8))
    f = d.float()
    assert float((f > 250).float().mean()) > 0.5 and float(f.min()) < 30       # white paper, dark pen
    ink = (f < 128).float().mean((1, 2, 3))
    assert float(ink.min()) > 0.005 and float(ink.max()) < 0.35                # every image has a stroke, none is a blob
    c = tg.synthetic_targets(3, count=32, seed=7)
    assert c.shape == (32, 64, 64, 3) and c.dtype == torch.uint8
    cf = c.float()
    assert float(cf.min()) == 0.0 and float(cf.max()) >= 254.0
    # smooth: neighbouring pixels differ by a few grey levels, channels are correlated but not identical
    assert float((cf[:, 1:] - cf[:, :-1]).abs().mean()) < 8.0
    assert float((cf[..., 0] - cf[..., 1]).abs().mean()) > 1.0
