"""``Policy`` -- the reference's models/policy.py graph as a PyTorch module (TF-1.6 semantics).

Scope note: the north-star path lists three hand-written kernel families (rasteriser,
discriminator, A2C loss); the policy trunk is the next row and runs here as plain
PyTorch.  What is kept exactly is the interface and the graph, so that its logits feed K3 and its
parameters map one-to-one onto a reference checkpoint:

  * inputs ``x [B,T,S,S,C]``, ``ac [B,T,K]`` (float, env.acs order), ``c`` (conditional) or
    ``z [B,T,z_dim]``, initial LSTM state ``(c0, h0)``                  (policy.py:28-66,136-147)
  * encoder: action MLP -> 32, conv5x5 same, 3 x conv4x4/s2 VALID (64->31->14->6), 8 res blocks,
    NHWC flatten -> dense 256                                           (policy.py:78-121)
  * ``BasicLSTMCell(256)``: gate order i, j, f, o on concat[x, h], forget bias 1.0   (policy.py:131-157)
  * autoregressive heads on relu(lstm_out) in env.acs order: dense for scalar heads, 4x4x16 ->
    convT -> 8 res blocks -> convT x n -> conv3x3 for location heads; each sampled (or teacher
    forced) action is embedded and mixed back into z                    (policy.py:217-325)
  * value head on the un-rectified LSTM output                          (policy.py:167-170)

EXTENSION: the reference stops after two transposed convolutions (32x32 logits); ``location_size=64``
adds a third so that the head has L*L logits.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class _MLP(nn.Module):
    """policy.py:337-347: (num_layers-1) x dense(hid, relu) then dense(dim, relu)."""

    def __init__(self, in_dim, dim, hid_dim=64, num_layers=3):
        super().__init__()
        dims = [in_dim] + [hid_dim] * (num_layers - 1) + [dim]
        self.layers = nn.ModuleList([nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])])

    def forward(self, x):
        for l in self.layers:
            x = F.relu(l(x))
        return x


class _ResBlock(nn.Module):
    """policy.py:349-362."""

    def __init__(self, ch, size=3):
        super().__init__()
        self.c1 = nn.Conv2d(ch, ch, size, padding=size // 2)
        self.c2 = nn.Conv2d(ch, ch, size, padding=size // 2)

    def forward(self, x):
        return self.c2(F.relu(self.c1(x))) + x


class _BasicLSTMCell(nn.Module):
    """tf.nn.rnn_cell.BasicLSTMCell: one kernel [in+hid, 4*hid], gates i, j, f, o, forget_bias 1."""

    def __init__(self, in_dim, hid):
        super().__init__()
        self.hid = hid
        self.kernel = nn.Linear(in_dim + hid, 4 * hid)

    def forward(self, x, state):
        c, h = state
        i, j, f, o = self.kernel(torch.cat([x, h], -1)).chunk(4, -1)
        new_c = c * torch.sigmoid(f + 1.0) + torch.sigmoid(i) * torch.tanh(j)
        new_h = torch.tanh(new_c) * torch.sigmoid(o)
        return new_h, (new_c, new_h)


class _LocationHead(nn.Module):
    def __init__(self, lstm_size, L, n_res):
        super().__init__()
        self.c0 = lstm_size // 16
        self.deconv0 = nn.ConvTranspose2d(self.c0, 32, 4, stride=2, padding=1)       # 4x4 -> 8x8, 'same'
        self.res = nn.ModuleList([_ResBlock(32) for _ in range(n_res)])
        n_up = int(round(math.log2(L / 8)))
        if 8 * 2 ** n_up != L:
            raise ValueError("location_size must be 8 * 2^n (got %d)" % L)
        self.ups = nn.ModuleList([nn.ConvTranspose2d(32, 32, 4, stride=2, padding=1) for _ in range(n_up)])
        self.out = nn.Conv2d(32, 1, 3, padding=1)                                    # named "conv_1x1" upstream

    def forward(self, z_flat):
        x = z_flat.view(-1, 4, 4, self.c0).permute(0, 3, 1, 2)                       # NHWC reshape of the reference
        x = F.relu(self.deconv0(x))
        for r in self.res:
            x = r(x)
        for u in self.ups:
            x = F.relu(u(x))
        return self.out(x).flatten(1)


def categorical_sample(logits, generator=None, owner=None):
    """policy.py:364-368 (``tf.multinomial(logits - max, 1)``) on the GPU: K4, Gumbel-max over
    Philox4x32-10 keyed by the generator's seed; every call advances a per-owner offset so a
    (seed, offset) pair is never reused.  logits f32[R,N] CUDA -> int32[R]."""
    import ctypes as C

    from .. import _native as N
    if not logits.is_cuda:
        raise RuntimeError("categorical_sample needs CUDA logits (there is no CPU fallback)")
    lg = logits.to(torch.float32).contiguous()
    seed = int(generator.initial_seed()) if generator is not None else int(torch.initial_seed())
    holder = owner if owner is not None else categorical_sample
    out = torch.empty((lg.shape[0],), dtype=torch.int32, device=lg.device)
    ctr = getattr(holder, "_sample_counter_dev", None)
    if ctr is not None:
        # device-resident call counter: graph-capturable, every replay draws fresh numbers
        N.check(N.lib().spiral_categorical_sample(N.ptr(lg), lg.shape[0], lg.shape[1], C.c_uint64(seed & (2 ** 64 - 1)),
                                                  C.c_uint64(0), N.ptr(ctr), N.ptr(out), N.stream_ptr()),
                "spiral_categorical_sample")
        ctr.add_(1)
        return out
    off = getattr(holder, "_sample_offset", 0)
    setattr(holder, "_sample_offset", off + 1)
    N.check(N.lib().spiral_categorical_sample(N.ptr(lg), lg.shape[0], lg.shape[1], C.c_uint64(seed & (2 ** 64 - 1)),
                                              C.c_uint64(off), None, N.ptr(out), N.stream_ptr()), "spiral_categorical_sample")
    return out


class Policy(nn.Module):
    def __init__(self, args, env, scope_name="global", image_shape=None, action_sizes=None, data_format="channels_last"):
        super().__init__()
        self.args = args
        self.scope_name = scope_name
        self.lstm_size = args.lstm_size
        self.image_shape = list(image_shape if image_shape is not None else env.observation_shape)
        self.action_sizes = action_sizes if action_sizes is not None else env.action_sizes
        self.acs = list(self.action_sizes.keys())
        self.conditional = bool(args.conditional)
        self.episode_length = env.episode_length
        K = len(self.acs)
        C = self.image_shape[-1] * (2 if self.conditional else 1)
        n_res = int(8 * args.scale)

        if not self.conditional:
            self.z_enc = _MLP(args.z_dim, self.lstm_size)
        self.a_enc = _MLP(1, 16)
        self.a_concat_fc = nn.Linear(16 * K, 32)
        self.x_enc = nn.Conv2d(C, 32, 5, padding=2)
        self.add_enc = nn.ModuleList([nn.Conv2d(32, 32, 4, stride=2) for _ in range(3)])
        self.encoder_res = nn.ModuleList([_ResBlock(32) for _ in range(n_res)])
        s = self.image_shape[0]
        for _ in range(3):
            s = (s - 4) // 2 + 1
        self.flat_fc = nn.Linear(s * s * 32, self.lstm_size)
        self.lstm = _BasicLSTMCell(self.lstm_size, self.lstm_size)
        self.heads = nn.ModuleDict()
        self.sample_mlp = nn.ModuleDict()
        self.concat_z_fc = nn.ModuleDict()
        for idx, name in enumerate(self.acs):
            size = self.action_sizes[name]
            if len(size) == 1:
                self.heads[name] = nn.Linear(self.lstm_size, size[0])
            else:
                self.heads[name] = _LocationHead(self.lstm_size, size[0], n_res)
            if idx < len(self.acs) - 1:
                self.sample_mlp[name] = _MLP(1, 16)
                self.concat_z_fc[name] = nn.Linear(self.lstm_size + 16, self.lstm_size)
        self.value = nn.Linear(self.lstm_size, 1)
        self.reset_parameters_tf()

    @torch.no_grad()
    def reset_parameters_tf(self):
        """The reference's initial weights: tf.layers.dense / conv2d / conv2d_transpose and BasicLSTMCell create their
        kernels with the variable scope's default initializer, glorot_uniform (limit sqrt(6 / (fan_in + fan_out)) with
        TF's fan convention: receptive field x channels), and their biases with zeros; only the scalar action heads use
        normalized_columns_initializer(0.01) (policy.py:231-236, 370-375: a normal draw rescaled so that every output
        unit's weight vector has norm 0.01).  PyTorch's defaults (kaiming-uniform(a=sqrt 5) kernels, uniform biases) give
        other initial logits and value scales."""
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Conv2d, nn.ConvTranspose2d)):
                nn.init.xavier_uniform_(m.weight)           # same fan_in + fan_out as TF for all three kernel layouts
                if m.bias is not None:
                    m.bias.zero_()
        for name in self.acs:
            head = self.heads[name]
            if isinstance(head, nn.Linear):
                w = torch.randn_like(head.weight)           # TF kernel [in, out], normalised over axis 0 = per output unit
                head.weight.copy_(w * 0.01 / w.pow(2).sum(1, keepdim=True).sqrt())

    # ------------------------------------------------------------------
    def get_initial_features(self, batch_size, flat=False, device=None):
        """policy.py:175-181: zero (c, h)."""
        dev = device or next(self.parameters()).device
        c = torch.zeros((batch_size, self.lstm_size), device=dev)
        return [c, torch.zeros_like(c)]

    def _native(self, x):
        """Learner path (gradients enabled, CUDA): the convolutions and their gradients on libspiral_b200.so
        (models/policy_learn.py).  ``act`` runs under no_grad and keeps the plain module (the actors have their own
        kernels, models/policy_native.py)."""
        return bool(getattr(self, "learn_native", False)) and x.is_cuda and torch.is_grad_enabled()

    def encode(self, x, ac, c=None, z=None):
        if self._native(x):
            from . import policy_learn
            return policy_learn.encode(self, x, ac, c, z)
        B, T = x.shape[:2]
        if self.conditional:
            x = torch.cat([x, c], -1)
        x = x.reshape((B * T,) + tuple(x.shape[2:])).permute(0, 3, 1, 2)
        ac = ac.reshape(B * T, -1).to(x.dtype)
        a = self.a_enc(ac.unsqueeze(-1)).reshape(B * T, -1)
        a = F.relu(self.a_concat_fc(a))
        add = F.relu(self.x_enc(x)) + a[:, :, None, None]
        for conv in self.add_enc:
            add = F.relu(conv(add))
        for r in self.encoder_res:
            add = r(add)
        flat = add.permute(0, 2, 3, 1).flatten(1)                      # tl.flatten on NHWC
        out = F.relu(self.flat_fc(flat)).view(B, T, -1)
        if not self.conditional:
            out = out + self.z_enc(z)
        return out

    def decode(self, z, samples=None, generator=None):
        """policy.py:217-325.  ``samples``: dict name -> int tensor [B,T] to teacher-force
        (agent.py:357-365); missing heads are sampled (tf.multinomial)."""
        B, T = z.shape[:2]
        logits, out_samples = {}, {}
        for idx, name in enumerate(self.acs):
            z_flat = z.reshape(B * T, self.lstm_size)
            if isinstance(self.heads[name], _LocationHead) and self._native(z_flat):
                from . import policy_learn
                logit = policy_learn.location_head(self.heads[name], z_flat)
            else:
                logit = self.heads[name](z_flat)
            logits[name] = logit.view(B, T, -1)
            if samples is not None and name in samples:
                s = samples[name].reshape(B * T).long()
            else:
                s = categorical_sample(logit.detach(), generator, self).long()
            out_samples[name] = s.view(B, T)
            if idx < len(self.acs) - 1:
                emb = self.sample_mlp[name](s.view(B, T, 1).to(z.dtype))
                z = F.relu(self.concat_z_fc[name](torch.cat([z, emb], -1)))
        return logits, out_samples

    def forward(self, x, ac, state_in, c=None, z=None, samples=None, generator=None):
        """x [B,T(+1),S,S,C] (the extra last state is dropped, policy.py:34).  Returns
        (logits dict [B,T,N], samples dict [B,T], vf [B,T], (c, h))."""
        x = x[:, :self.episode_length]
        T = x.shape[1]
        enc = self.encode(x, ac[:, :T], None if c is None else c[:, :T], None if z is None else z[:, :T])
        state = (state_in[0], state_in[1])
        outs = []
        for t in range(T):
            h, state = self.lstm(enc[:, t], state)
            outs.append(h)
        lstm_out = torch.stack(outs, 1)
        logits, smp = self.decode(F.relu(lstm_out), samples, generator)
        vf = self.value(lstm_out)[:, :, 0]
        return logits, smp, vf, state

    @torch.no_grad()
    def act(self, ob, ac, c, h, condition=None, z=None, generator=None):
        """policy.py:196-215 for a batch: ob [B,S,S,C], ac [B,K] (previous action, -1 at step 0).
        Returns (actions int32 [B,K], value [B], c, h)."""
        x = ob.unsqueeze(1)
        cond = None if condition is None else condition.unsqueeze(1)
        zz = None if z is None else z.unsqueeze(1)
        ep = self.episode_length
        self.episode_length = 1
        try:
            logits, smp, vf, (nc, nh) = self.forward(x, ac.unsqueeze(1).float(), (c, h), cond, zz, None, generator)
        finally:
            self.episode_length = ep
        actions = torch.stack([smp[name][:, 0] for name in self.acs], 1).to(torch.int32)
        return actions, vf[:, 0], nc, nh
