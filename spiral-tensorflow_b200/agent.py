"""``Agent`` -- synchronous, batched counterpart of the reference's agent.py / rl_utils.env_runner.

The reference runs N-2 CPU actor processes with one env each, a policy learner and a discriminator
learner talking through TF queues on a parameter server (agent.py:38-59, trainer.py:107-118).  Here
one process per GPU owns B canvases; a step of the loop is

    rollout()        rl_utils.py:185-238   policy.act -> env.step, T times, everything on the GPU
    train_policy()   agent.py:333-403      GAN reward (K2) -> fused A2C loss (K3) -> backward through
                                           the PyTorch policy -> all-reduce (overlapped, GradReducer) -> clip 40 -> Adam
    train_gan()      agent.py:409-437      WGAN-GP step on (replay fakes, real targets)

The three kernel families of the north-star path are native, including the discriminator's
backward / double-backward convolutions (models/conv_ops.py, K2b); the policy trunk is the next row
 and runs through PyTorch autograd.
"""
from __future__ import annotations

import torch

from . import distributed as dist_utils
from . import rl_utils
from .models.discriminator import Discriminator
from .models.policy import Policy
from .models.policy_native import NativeActor
from .optim import FusedTFAdam
from .replay import ReplayBuffer


class Agent(object):
    def __init__(self, args, env, device=None):
        self.args = args
        self.env = env
        self.device = env.device if device is None else torch.device(device)
        tf32 = bool(getattr(args, "policy_tf32", True))
        torch.backends.cudnn.allow_tf32 = tf32            # policy trunk only: K1-K5 do not go through cuDNN / cuBLAS
        torch.backends.cuda.matmul.allow_tf32 = tf32
        self.policy = Policy(args, env, "global", env.observation_shape, env.action_sizes).to(self.device)
        # learner: the trunk's and the location heads' convolutions, forward and backward, on libspiral_b200.so (bf16x3)
        self.policy.learn_native = bool(getattr(args, "policy_learn_native", True)) and self.device.type == "cuda"
        if self.policy.learn_native:
            from .models import policy_learn
            policy_learn.PLANES = policy_learn.FORMATS[getattr(args, "policy_learn_precision", "f16x3")]
        self.policy_opt = FusedTFAdam(self.policy.parameters(), args.policy_lr)
        # acting path: the trunk's convolutions on libspiral_b200.so (bf16 tensor cores); the learner keeps the fp32 module
        self.actor = NativeActor(self.policy) if (getattr(args, "policy_act_native", True) and self.device.type == "cuda") else None
        self.gan = args.loss == "gan"
        if self.gan:
            self.disc = Discriminator(args, None, env.observation_shape, norm_fn=env.norm, scope_name="global",
                                      device=self.device, max_batch=max(env.num_envs, args.disc_batch_size))
            for p in self.disc.params.values():
                p.requires_grad_(True)
            self.disc_opt = FusedTFAdam(list(self.disc.params.values()), args.disc_lr, beta1=0.5, beta2=0.9)
            self.replay = ReplayBuffer(args, env.observation_shape, device=self.device)
        self.rng = torch.Generator(device=self.device)
        self.rng.manual_seed(int(args.seed))
        self.world = dist_utils.dist.get_world_size() if dist_utils.dist.is_initialized() else 1
        # env_runner (rl_utils.py:186-187) fetches policy.get_initial_features ONCE, before its `while True`: an actor's
        # LSTM state carries over from the last step of one episode into the first step of the next.  [B, 2, lstm] = (c, h),
        # read at the top of every rollout and written back at its end (inside the captured graph, so replays chain).
        c0, h0 = self.policy.get_initial_features(env.num_envs)
        self._lstm_state = torch.stack([c0, h0], 1).contiguous()
        # K4's call counter lives on the device from the start (graph-capturable; a checkpoint restores it in place)
        if self.device.type == "cuda":
            self.policy._sample_counter_dev = torch.full((1,), int(getattr(self.policy, "_sample_offset", 0)),
                                                         dtype=torch.int64, device=self.device)
        # trajectory store: observations are integers 0..255 when color_channel == 3 (read-back of the uint8 image), so
        # `states` is kept as uint8 (4x smaller: 252 MB instead of 1.03 GB at 1,024 canvases x 21 frames, lossless); gray
        # observations are BT.601 dot products (non-integers) and stay float32 as in the reference's queue (agent.py:42-47)
        # several ranks: gradients accumulate into one flat buffer per network and are all-reduced bucket by bucket from
        # backward hooks, overlapping the rest of the backward (distributed.GradReducer); SPIRAL_GRAD_REDUCER=1 uses the
        # same path on one rank (tests)
        import os
        use_red = self.device.type == "cuda" and (self.world > 1 or os.environ.get("SPIRAL_GRAD_REDUCER") == "1")
        self._pol_red = dist_utils.GradReducer(self.policy.parameters(), world=self.world) if use_red else None
        self._disc_red = dist_utils.GradReducer(list(self.disc.params.values()), world=self.world) if (use_red and self.gan) else None
        self._global_bn = dist_utils.allreduce_bn_sums_ if (self.world > 1 and getattr(args, "disc_global_bn", False)) else None
        self.state_dtype = torch.uint8 if (int(getattr(env, "channels", 1)) == 3 and getattr(args, "traj_uint8", True)) else torch.float32
        self._graph, self._graph_failed, self._g_out, self._g_target, self._g_z = None, False, {}, None, None
        self._tp_graph, self._tp_failed, self._tp_calls, self._tp_in, self._tp_out = None, False, 0, None, None
        self._tg_graph, self._tg_failed, self._tg_calls, self._tg_in, self._tg_out = None, False, 0, None, None

    def close(self):
        """Drop the captured CUDA graphs.  With several ranks the learner graphs hold NCCL collectives: they must be gone
        (and the device idle) before ``torch.distributed.destroy_process_group()``, which otherwise waits forever for the
        communicator the graphs still reference."""
        if self.device.type == "cuda":
            torch.cuda.synchronize()
        self._graph = self._tp_graph = self._tg_graph = None
        self._g_out, self._tp_out, self._tg_out = {}, None, None
        import gc
        gc.collect()
        if self.device.type == "cuda":
            torch.cuda.synchronize()

    # ------------------------------------------------------------------
    @torch.no_grad()
    def rollout(self):
        """rl_utils.env_runner (rl_utils.py:185-238) for B canvases at once.  With ``--cuda_graphs`` (default) the
        whole T-step episode -- reset, T x (policy.act, K4 sampling, K1 env.step, bookkeeping) -- is captured once
        into a CUDA graph and replayed: at the reference's batch sizes the loop is launch-bound (about 150 small
        kernels per step), the graph removes that overhead and the Python interpreter from the critical path."""
        if not (getattr(self.args, "cuda_graphs", True) and self.device.type == "cuda") or self._graph_failed:
            return self._finish_rollout(self._rollout_body(None, None))
        env = self.env
        if self._graph is None:
            try:
                self._capture_rollout()
            except Exception as e:                       # capture is an optimisation: fall back to eager launches
                self._graph_failed = True
                self._graph = None
                torch.cuda.synchronize()
                import warnings
                warnings.warn("CUDA-graph capture of the rollout failed (%s); running eagerly" % (e,))
                return self._finish_rollout(self._rollout_body(None, None))
        # fresh condition / latent into the graph's static inputs (drawn as env.reset does)
        if self._g_target is not None:
            self._g_target.copy_(env.get_random_target(num=env.num_envs))
        if self._g_z is not None:
            self._g_z.copy_(torch.rand(self._g_z.shape, generator=env._rng, device=self.device) - 0.5)
        self._graph.replay()
        env._step = env.episode_length
        return self._finish_rollout(self._g_out, clone=True)

    def _capture_rollout(self):
        env, pi = self.env, self.policy
        B = env.num_envs
        self._g_target = env.get_random_target(num=B).clone() if env.conditional else None
        self._g_z = None if env.conditional else torch.zeros((B, env.z_dim), device=self.device)
        if getattr(pi, "_sample_counter_dev", None) is None:
            pi._sample_counter_dev = torch.full((1,), int(getattr(pi, "_sample_offset", 0)), dtype=torch.int64, device=self.device)
        self._g_init_action = torch.as_tensor(env.initial_action, device=self.device).to(torch.int32).expand(B, -1).contiguous()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        saved = (self._lstm_state.clone(), pi._sample_counter_dev.clone())
        with torch.cuda.stream(side):                      # warm-up: lazy initialisation (function attributes, cuBLAS) outside capture
            self._rollout_body(self._g_target, self._g_z)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._lstm_state.copy_(saved[0])                   # the warm-up episode is not part of the actor's history
        pi._sample_counter_dev.copy_(saved[1])
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._g_out = self._rollout_body(self._g_target, self._g_z)
        self._graph = g

    def _rollout_body(self, target, z_in):
        env, pi = self.env, self.policy
        last_action = self._g_init_action if getattr(self, "_g_init_action", None) is not None \
            else torch.as_tensor(env.initial_action, device=self.device)
        return rl_utils.batched_env_runner(env, (self.actor or pi).act, self._lstm_state, last_action, pi.lstm_size,
                                           state_dtype=self.state_dtype, target=target, z=z_in, generator=self.rng)

    def _finish_rollout(self, out, clone=False):
        if clone:                                         # the graph's static buffers are overwritten by the next replay
            out = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}
        if self.gan:
            self.replay.push(out["states"][:, self.env.episode_length])
        return out

    def train_policy(self, rollout):
        """agent.py:333-403.  With ``--cuda_graphs`` (default) the learner step -- policy forward over [B*T] frames with
        teacher forcing, the fused A2C loss, backward, gradient all-reduce, clip + Adam, the actors' weight pull -- is
        captured into ONE CUDA graph after two eager steps and replayed: like the rollout it is launch-bound at the
        reference's batch sizes (about 1,500 small kernels).  The step size of every replay comes from device memory
        (``FusedTFAdam.advance``); the GAN reward is injected before the graph (the discriminator re-packs its weights
        only when they changed, a host-side decision)."""
        if self.gan:
            with torch.no_grad():
                self.disc.sync_weights()
                rollout["rewards"][:, -1] = self.disc.predict(rollout["states"][:, -1], global_bn=self._global_bn)   # agent.py:336-338
        # several ranks: the NCCL all-reduce is captured into the graph with the rest of the step (every rank replays the
        # same sequence of collectives)
        use_graph = getattr(self.args, "cuda_graphs", True) and self.device.type == "cuda" and not self._tp_failed
        if not use_graph:
            return self._train_policy_body(rollout)
        self._tp_calls += 1
        if self._tp_graph is None and self._tp_calls <= 2:
            # eager steps double as the warm-up capture needs (cuDNN / cuBLAS handles, autograd's lazy state), on a side stream
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                out = self._train_policy_body(rollout)
            torch.cuda.current_stream().wait_stream(side)
            return out
        keys = [k for k in ("states", "actions", "rewards", "values", "features", "conditions", "z") if k in rollout]
        if self._tp_graph is None:
            try:
                self._tp_in = {k: rollout[k].clone() for k in keys}
                if self._pol_red is None:
                    self.policy_opt.zero_grad(set_to_none=True)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._tp_out = self._train_policy_body(self._tp_in, advance=False)
                self._tp_graph = g
            except Exception as e:                       # capture is an optimisation: fall back to eager steps
                self._tp_failed, self._tp_graph = True, None
                torch.cuda.synchronize()
                import warnings
                warnings.warn("CUDA-graph capture of the learner step failed (%s); running eagerly" % (e,))
                return self._train_policy_body(rollout)
        for k in keys:
            self._tp_in[k].copy_(rollout[k])
        self.policy_opt.advance()
        self._tp_graph.replay()
        return {k: (v.clone() if torch.is_tensor(v) else v) for k, v in self._tp_out.items()}

    def _train_policy_body(self, rollout, advance=True):
        args, pi = self.args, self.policy
        acts = rollout["actions"]
        samples = {name: acts[:, :, i] for i, name in enumerate(pi.acs)}                      # teacher forcing, agent.py:357-365
        feats = rollout["features"][:, 0]
        states = rollout["states"]
        if states.dtype != torch.float32:                 # uint8 trajectory store (color_channel 3): exact integers
            states = states[:, :pi.episode_length].float()
        logits, _, vf, _ = pi(states, acts.float(), (feats[:, 0], feats[:, 1]),
                              rollout.get("conditions"), rollout.get("z"), samples)
        loss = rl_utils.a2c_loss_autograd([logits[n].contiguous() for n in pi.acs], acts, rollout["rewards"],
                                          rollout["values"][:, :, 0], vf, 0.99, 1.0, args.entropy_coeff)
        if self._pol_red is not None:
            self._pol_red.zero()                          # gradients accumulate into the flat buffer ...
            loss.backward()                               # ... whose buckets are all-reduced from backward hooks as they fill
            self._pol_red.wait()                          # sum over ranks; 1/world is folded into K5
        else:
            self.policy_opt.zero_grad(set_to_none=True)
            loss.backward()
        gnorm = self.policy_opt.step(clip_norm=float(args.grad_clip), grad_scale=1.0 / self.world, advance=advance)   # agent.py:231-239
        if self.actor is not None:
            self.actor.sync()                             # actors pull the new weights (trainer.py:117 policy_sync)
        B = acts.shape[0]
        return {"model/total_loss": loss.detach(), "model/grad_global_norm": gnorm,
                "env/r": rollout["rewards"][:, -1].mean(), "batch": B}

    def train_gan(self, reals=None):
        """agent.py:409-437 + discriminator.py:97-136 (WGAN-GP, penalty on the sigmoid output).  The batch (replay
        sample, real images, interpolation factors) is drawn eagerly; the step itself -- three shared-weight passes, the
        penalty's double backward, Adam, the kernels' weight re-pack -- is a CUDA graph from the third call on, like
        ``train_policy``."""
        args, D = self.args, self.disc
        n = args.disc_batch_size
        fake = self.replay.sample(n)
        real = self.env.get_random_target(n) if reals is None else reals
        alpha = torch.rand((n, 1), generator=self.rng, device=self.device)
        use_graph = getattr(args, "cuda_graphs", True) and self.device.type == "cuda" and not self._tg_failed
        if not use_graph:
            return self._train_gan_body(fake, real, alpha)
        self._tg_calls += 1
        if self._tg_graph is None and self._tg_calls <= 2:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                out = self._train_gan_body(fake, real, alpha)
            torch.cuda.current_stream().wait_stream(side)
            return out
        if self._tg_graph is None:
            try:
                self._tg_in = [t.to(self.device, torch.float32).clone() for t in (fake, real, alpha)]
                if self._disc_red is None:
                    self.disc_opt.zero_grad(set_to_none=True)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._tg_out = self._train_gan_body(*self._tg_in, advance=False)
                self._tg_graph = g
            except Exception as e:
                self._tg_failed, self._tg_graph = True, None
                torch.cuda.synchronize()
                import warnings
                warnings.warn("CUDA-graph capture of the discriminator step failed (%s); running eagerly" % (e,))
                return self._train_gan_body(fake, real, alpha)
        for dst, src in zip(self._tg_in, (fake, real, alpha)):
            dst.copy_(src)
        self.disc_opt.advance()
        self._tg_graph.replay()
        return {k: v.clone() for k, v in self._tg_out.items()}

    def _train_gan_body(self, fake, real, alpha, advance=True):
        D = self.disc
        p = D.params
        out = D.wgan_gp_loss(fake, real, alpha)          # convolutions + every gradient on K2b
        d_loss, critic, gp, fake_logits_mean = out["d_loss"], out["critic_loss"], out["gradient_penalty"], -out["g_loss"]
        if self._disc_red is not None:
            self._disc_red.zero()
            d_loss.backward()
            self._disc_red.wait()
        else:
            self.disc_opt.zero_grad(set_to_none=True)
            d_loss.backward()
        self.disc_opt.step(grad_scale=1.0 / self.world, advance=advance)                      # discriminator.py:124-126
        D.sync_weights()
        return {"gan/critic_loss": critic.detach(), "gan/disc_loss": d_loss.detach(), "gan/penalty": gp.detach(),
                "gan/gen_loss": -fake_logits_mean.detach()}
