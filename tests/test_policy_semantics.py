"""CPU tests of the PyTorch policy's TensorFlow-1.6 semantics (models/policy.py of the reference).
The reference graph cannot be executed here (no TensorFlow), so these pin what is derivable from its source:
the size chain of the encoder, BasicLSTMCell's gate order and forget bias, the autoregressive order of the heads
and the teacher forcing of agent.py:357-365."""
import numpy as np
import torch

import spiral_b200 as sp
from importlib import import_module

policy_mod = import_module(sp.__name__ + ".models.policy")


class _Env(object):
    """Just what Policy reads from an env (no CUDA needed)."""

    def __init__(self, args, sizes):
        self.observation_shape = [args.screen_size, args.screen_size, args.color_channel]
        self.action_sizes = sizes
        self.episode_length = args.episode_length


def _policy(L=32, C=1, T=3, conditional=False):
    args = sp.get_args(["--env", "simple_mnist", "--location_size", str(L), "--color_channel", str(C),
                        "--episode_length", str(T), "--loss", "l2" if conditional else "gan"])
    from collections import OrderedDict
    sizes = OrderedDict([("jump", [2]), ("control", [L, L]), ("end", [L, L])])
    torch.manual_seed(0)
    return policy_mod.Policy(args, _Env(args, sizes)), args


def test_encoder_size_chain_and_head_shapes():
    """policy.py:103-121: three 4x4 / stride-2 VALID convs take 64 -> 31 -> 14 -> 6, flatten 6*6*32 = 1152;
    :229-300: scalar head = dense(N), location head = L*L logits; vf [B,T]."""
    pi, args = _policy(L=32, T=3)
    assert pi.flat_fc.in_features == 6 * 6 * 32 == 1152
    B, T = 2, 3
    x = torch.rand(B, T + 1, 64, 64, 1) * 255
    ac = torch.zeros(B, T, 3)
    z = torch.rand(B, T, args.z_dim) - 0.5
    c = torch.zeros(B, args.lstm_size)
    samples = {"jump": torch.zeros(B, T, dtype=torch.long), "control": torch.zeros(B, T, dtype=torch.long),
               "end": torch.zeros(B, T, dtype=torch.long)}
    logits, smp, vf, (nc, nh) = pi(x, ac, (c, c.clone()), None, z, samples)
    assert logits["jump"].shape == (B, T, 2) and logits["control"].shape == (B, T, 1024) and logits["end"].shape == (B, T, 1024)
    assert vf.shape == (B, T) and nc.shape == (B, args.lstm_size) and nh.shape == (B, args.lstm_size)
    assert all(torch.equal(smp[k], samples[k]) for k in samples)          # teacher forcing returns the fed actions


def test_basic_lstm_cell_gate_order_and_forget_bias():
    """tf BasicLSTMCell: kernel on concat[x, h]; gates i, j, f, o; new_c = c * sigmoid(f + 1) + sigmoid(i) * tanh(j);
    new_h = tanh(new_c) * sigmoid(o)."""
    torch.manual_seed(1)
    cell = policy_mod._BasicLSTMCell(5, 4)
    x, c, h = torch.randn(3, 5), torch.randn(3, 4), torch.randn(3, 4)
    out, (nc, nh) = cell(x, (c, h))
    W, b = cell.kernel.weight.detach().numpy(), cell.kernel.bias.detach().numpy()
    g = np.concatenate([x.numpy(), h.numpy()], 1) @ W.T + b
    i, j, f, o = g[:, 0:4], g[:, 4:8], g[:, 8:12], g[:, 12:16]
    sig = lambda v: 1.0 / (1.0 + np.exp(-v))
    want_c = c.numpy() * sig(f + 1.0) + sig(i) * np.tanh(j)
    want_h = np.tanh(want_c) * sig(o)
    assert np.allclose(nc.detach().numpy(), want_c, atol=1e-6) and np.allclose(nh.detach().numpy(), want_h, atol=1e-6)
    assert torch.equal(out, nh)


def test_heads_are_autoregressive_in_env_order():
    """policy.py:224,314-323: head k sees the samples of heads < k only (through sample_mlp + concat_z_fc)."""
    pi, args = _policy(L=32, T=2)
    B, T = 2, 2
    x = torch.rand(B, T + 1, 64, 64, 1) * 255
    ac = torch.zeros(B, T, 3)
    z = torch.rand(B, T, args.z_dim) - 0.5
    c = torch.zeros(B, args.lstm_size)

    def run(jump, control, end):
        s = {"jump": torch.full((B, T), jump), "control": torch.full((B, T), control), "end": torch.full((B, T), end)}
        with torch.no_grad():
            return pi(x, ac, (c, c.clone()), None, z, s)[0]
    base = run(0, 5, 7)
    last = run(0, 5, 900)              # the last head's sample feeds nothing
    mid = run(0, 600, 7)               # the middle head's sample reaches only the last head
    first = run(1, 5, 7)               # the first head's sample reaches both later heads
    assert all(torch.equal(base[k], last[k]) for k in base)
    assert torch.equal(base["jump"], mid["jump"]) and torch.equal(base["control"], mid["control"])
    assert not torch.equal(base["end"], mid["end"])
    assert torch.equal(base["jump"], first["jump"]) and not torch.equal(base["control"], first["control"])


def test_extra_last_state_is_dropped_and_conditional_doubles_channels():
    """policy.py:30-34 slices x to T; :45-53 concatenates the condition on the channel axis."""
    pi, args = _policy(L=32, T=2)
    B, T = 1, 2
    x = torch.rand(B, T + 1, 64, 64, 1) * 255
    ac = torch.zeros(B, T, 3)
    z = torch.zeros(B, T, args.z_dim)
    c = torch.zeros(B, args.lstm_size)
    s = {k: torch.zeros(B, T, dtype=torch.long) for k in ("jump", "control", "end")}
    with torch.no_grad():
        a = pi(x, ac, (c, c.clone()), None, z, s)[0]["end"]
        x2 = x.clone()
        x2[:, T] = 0.0                 # the (T+1)-th state must not matter
        b = pi(x2, ac, (c, c.clone()), None, z, s)[0]["end"]
    assert torch.equal(a, b)
    pc, _ = _policy(L=32, T=2, conditional=True)
    assert pc.x_enc.in_channels == 2 and not hasattr(pc, "z_enc")
