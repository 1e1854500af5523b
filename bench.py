#!/usr/bin/env python
"""bench.py -- env-steps/sec (render + D reward + A2C) on 64x64 canvases (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic input: an episode of T=20 env
steps for B canvases per GPU (reset + 20 x spiral_env_step), the discriminator reward on the
final canvases (spiral_disc_forward), and the fused A2C loss + gradient over [B,T]
(spiral_a2c_loss).  Workload = BASELINE.json configs[3] "rgb64" shape (color_channel=3,
location_size=64, episode_length=20, heads control/end 4096, color 8, jump/pressure/size 2) at the
largest single-GPU batch of configs[4]'s sweep, B = 65536 canvases per GPU; the 16384- and
1024-canvas points are reported beside it.  `disc_train` = one WGAN-GP discriminator step
(batch 64, reference disc_batch_size) through the K2b gradient kernels.

`value`  : whole-job env-steps/s, inputs resident in HBM, CUDA-event timed, max over ranks.
`e2e`    : same metric through the Python env/model API with HOST (pinned) actions/values per
           call and a device->host read of rewards + loss scalars every episode.
`roofline`: K1 (expand+composite pair and each kernel alone), algorithmic bytes per canvas-step
            / CUDA-event kernel time, against MEASURED_PEAKS.json.
`cpu_baseline` / --impl reference: the CPU oracle port of the reference path (libmypaint
           restatement + PyTorch-CPU discriminator + numpy A2C; the real libmypaint/TF-1.6 stack
           is not installable) on the host cores of this box.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json configs[3] gives the shape (64x64 RGB, L=64, T=20, colour head); configs[4] ("rendering+reward
# throughput sweep: 4096-65536 canvases x 20 strokes at 1/2/4/8 GPUs") gives the batch the throughput metric is
# quoted on.  Default: 65536 canvases per GPU (weak scaling); the 16384-canvas sweep point and the 1024-canvas point of
# configs[3] are reported in the same JSON line under "rgb64_b16384" / "rgb64_b1024".
WORKLOAD = dict(workload="rgb64-sweep (configs[3] shape at configs[4]'s largest single-GPU batch)", env="color_mnist",
                color_channel=3, location_size=64, episode_length=20, canvases_per_gpu=65536, disc_dim=64)
# DRAM bytes per canvas-step of composite_kernel from the ncu --set full capture in profiles/ (B=16384:
# dram__bytes_read 368.1 MB + dram__bytes_write 266.9 MB per launch, profiles/r1_final_k1_ncu_summary_v2.md)
COMPOSITE_TRAFFIC_PER_CANVAS_STEP = (368.139008e6 + 266.859520e6) / 16384
BYTES_PER_CANVAS_STEP = {1: 81920, 3: 114688}       # BASELINE.md section 3 (K1 algorithmic bytes)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--canvases", type=int, default=WORKLOAD["canvases_per_gpu"], help="canvases per GPU")
    p.add_argument("--episode_length", type=int, default=WORKLOAD["episode_length"])
    p.add_argument("--disc_precision", default=os.environ.get("SPIRAL_DISC_PRECISION", "bf16x3"))
    p.add_argument("--no_cpu_baseline", action="store_true")
    p.add_argument("--cpu_seconds", type=float, default=15.0, help="target seconds of the bounded CPU sample")
    p.add_argument("--jump", default="True")
    p.add_argument("--no_secondary", action="store_true", help="skip the 1024-canvas secondary measurement")
    p.add_argument("--pipeline", action="store_true", help="also time the two-stream pipelined schedule")
    return p.parse_args()


# ---------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference path on the host cores
# ---------------------------------------------------------------------------------------------
_W = {}


def _cpu_worker_init(acs, L, C):
    from oracle import paint_oracle as po
    spec = po.EnvSpec(action_names=acs, location_size=L, n_color=8, math_mode=po.PO_MATH_LIBM)
    _W["env"] = po.OracleBatchEnv(spec, 1, channels=C)


def _cpu_worker_run(actions):
    env = _W["env"]
    out = []
    for a in actions:                      # a: [T,K]
        canv, obs = env.run_episodes(a[None], want_obs=True)
        out.append(obs[0, -1])
    return np.stack(out)


def cpu_reference_run(acs, L, C, T, head_sizes, disc_dim, episodes, cores, seed=0):
    """Times `episodes` whole episodes of the CPU port: render (P worker processes, one env each,
    the reference's actor layout run.py:45-52) + torch-CPU D forward + numpy A2C.  Returns seconds."""
    import multiprocessing as mp

    import torch

    from oracle import a2c_oracle as ao
    from oracle import disc_oracle as do
    rng = np.random.RandomState(seed)
    acts = np.stack([rng.randint(n, size=(episodes, T)) for n in head_sizes], -1).astype(np.int32)
    chunks = np.array_split(acts, cores)
    w = do.init_weights(C, disc_dim=disc_dim, seed=11)
    logits = [rng.randn(min(episodes, 64), T, n).astype(np.float32) for n in head_sizes]
    values = rng.randn(episodes, T).astype(np.float32)
    torch.set_num_threads(cores)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_worker_init, initargs=(acs, L, C)) as pool:
        pool.map(_cpu_worker_run, [c[:1] for c in chunks])          # warm the workers
        t0 = time.perf_counter()
        finals = np.concatenate(pool.map(_cpu_worker_run, chunks))
        t_render = time.perf_counter() - t0
    t0 = time.perf_counter()
    probs = []
    for i in range(0, episodes, 64):                                 # disc_batch_size-like chunks
        p, _ = do.predict(finals[i:i + 64], w, dtype=torch.float32)
        probs.append(p)
    probs = np.concatenate(probs)
    t_disc = time.perf_counter() - t0
    t0 = time.perf_counter()
    rewards = np.zeros((episodes, T), np.float32)
    rewards[:, -1] = probs
    r, adv = ao.multiple_process_rollout(rewards, values, 0.99, 1.0)
    nb = logits[0].shape[0]
    reps = max(1, episodes // nb)
    for i in range(reps):                                            # logits reused: same arithmetic per episode
        ao.a2c_loss(logits, acts[:nb], adv[:nb].astype(np.float32), r[:nb].astype(np.float32), values[:nb],
                    0.01, dtype=__import__("torch").float32)
    t_a2c = (time.perf_counter() - t0) * (episodes / float(reps * nb))
    return dict(seconds=t_render + t_disc + t_a2c, render=t_render, disc=t_disc, a2c=t_a2c)


def cpu_baseline(args, acs, head_sizes, C, L, T, disc_dim):
    cores = os.cpu_count() or 1
    # probe: 2 episodes per core to size the bounded sample
    probe = cpu_reference_run(acs, L, C, T, head_sizes, disc_dim, episodes=2 * cores, cores=cores)
    per_ep = probe["seconds"] / (2 * cores)
    episodes = int(max(cores * 2, min(4096, args.cpu_seconds / max(per_ep, 1e-6))))
    episodes = (episodes // cores) * cores
    res = cpu_reference_run(acs, L, C, T, head_sizes, disc_dim, episodes=episodes, cores=cores, seed=1)
    return dict(value=episodes * T / res["seconds"], unit="env-steps/s", cores=cores, kind="port",
                sample="%d episodes x %d steps (render %.1fs on %d worker processes + torch-CPU D forward %.1fs "
                       "+ numpy/torch-CPU A2C %.1fs); C oracle of libmypaint 1.3.0 in libm mode"
                       % (episodes, T, res["render"], cores, res["disc"], res["a2c"]))


# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu = gpu
        self.rows = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([c.strip() for c in line.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(self.rows))


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fp:
            d = json.load(fp)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import spiral_b200 as sp
    from importlib import import_module
    rl = import_module(sp.__name__ + ".rl_utils")
    models = import_module(sp.__name__ + ".models")

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    def measure(B, jump=None):
        T = args.episode_length
        jump = args.jump if jump is None else jump
        C, L = WORKLOAD["color_channel"], WORKLOAD["location_size"]
        a = sp.get_args(["--env", WORKLOAD["env"], "--location_size", str(L), "--color_channel", str(C),
                         "--episode_length", str(T), "--loss", "gan", "--disc_dim", str(WORKLOAD["disc_dim"]),
                         "--disc_precision", args.disc_precision, "--jump", jump, "--seed", str(123 + rank)])
        env = sp.create_env(a, num_envs=B, device=dev)
        disc = models.Discriminator(a, None, [64, 64, C], norm_fn=env.norm, scope_name="global", device=dev, max_batch=B)
        K = len(env.acs)
        acs_list, head_list = list(env.acs), list(env.head_sizes)
        g = torch.Generator(device="cpu")
        g.manual_seed(123 + rank)                     # seed + task, as main.py:21
        acts_host = torch.stack([torch.randint(0, n, (B, T), generator=g) for n in env.head_sizes], -1).to(torch.int32)
        acts_host = acts_host.permute(1, 0, 2).contiguous().pin_memory()      # [T,B,K]
        acts_dev = acts_host.to(dev)
        gg = torch.Generator(device=dev)
        gg.manual_seed(7 + rank)
        logits = [torch.randn((B, T, n), generator=gg, device=dev) for n in env.head_sizes]
        dlogits_bytes = sum(l.numel() * 4 for l in logits)
        values_host = torch.randn((B, T), generator=g).pin_memory()
        vf_host = torch.randn((B, T), generator=g).pin_memory()
        values_dev, vf_dev = values_host.to(dev), vf_host.to(dev)
        rewards = torch.zeros((B, T), device=dev)
        a2c_out = rl.a2c_loss(logits, acts_dev.permute(1, 0, 2).contiguous(), rewards, values_dev, vf_dev)   # allocates outputs
        acts_bt = acts_dev.permute(1, 0, 2).contiguous()
        probs = torch.empty((B,), device=dev)
        lg = torch.empty((B,), device=dev)
        host_out = torch.empty((B + 4,), dtype=torch.float32).pin_memory()

        ev = lambda: torch.cuda.Event(enable_timing=True)
        kern_ev = {"expand": [], "composite": [], "disc": [], "a2c": []}
        N = sp._native
        lib = N.lib()

        def episode(resident=True, record=False):
            env.reset()
            for t in range(T):
                act = acts_dev[t] if resident else acts_host[t].to(dev, non_blocking=True)
                if record:
                    e0, e1, e2 = ev(), ev(), ev()
                    e0.record()
                    N.check(lib.spiral_env_expand(env._handle, N.ptr(act), N.ptr(env.brush_state), N.ptr(env.env_state),
                                                  N.ptr(env.dab_scratch), N.ptr(env.dab_count), N.ptr(env.status),
                                                  N.stream_ptr()))
                    e1.record()
                    N.check(lib.spiral_env_composite(env._handle, N.ptr(env.canvas), N.ptr(env.dab_scratch),
                                                     N.ptr(env.dab_count), N.ptr(env.obs), N.stream_ptr()))
                    e2.record()
                    kern_ev["expand"].append((e0, e1))
                    kern_ev["composite"].append((e1, e2))
                    env._step += 1
                else:
                    env.step(act)
            if record:
                d0, d1, d2 = ev(), ev(), ev()
                d0.record()
            disc.forward(env.state, logits_out=lg, probs_out=probs)          # agent.py:336-338: GAN reward
            rewards[:, -1] = probs
            if record:
                d1.record()
            if resident:
                v, f = values_dev, vf_dev
            else:
                v, f = values_host.to(dev, non_blocking=True), vf_host.to(dev, non_blocking=True)
            rl.a2c_loss(logits, acts_bt, rewards, v, f, 0.99, 1.0, 0.01, out=a2c_out)
            if record:
                d2.record()
                kern_ev["disc"].append((d0, d1))
                kern_ev["a2c"].append((d1, d2))
            if not resident:
                host_out[:B].copy_(probs, non_blocking=True)
                host_out[B:].copy_(a2c_out.scalars, non_blocking=True)

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        def timed_pipelined():
            """Two-stage software pipeline over consecutive batches, the reference's actor / learner split
            (actors render episode i+1 while the learner scores episode i, agent.py:336-374 vs rl_utils.py:185-238):
            stream R renders, stream S runs the D reward + A2C loss on a snapshot of the final observations.  All K
            renders and all K reward/loss passes (pipeline fill and drain included) are inside the timed region."""
            sR, sS = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
            snap = torch.empty_like(env.state)
            cur = torch.cuda.current_stream()

            def run(n):
                sR.wait_stream(cur)
                sS.wait_stream(cur)
                scored = None
                for _ in range(n):
                    with torch.cuda.stream(sR):
                        env.reset()
                        for t in range(T):
                            env.step(acts_dev[t])
                        if scored is not None:
                            sR.wait_event(scored)          # the previous snapshot has been consumed
                        snap.copy_(env.state)
                        rendered = torch.cuda.Event()
                        rendered.record(sR)
                    with torch.cuda.stream(sS):
                        sS.wait_event(rendered)
                        disc.forward(snap, logits_out=lg, probs_out=probs)
                        rewards[:, -1] = probs
                        rl.a2c_loss(logits, acts_bt, rewards, values_dev, vf_dev, 0.99, 1.0, 0.01, out=a2c_out)
                        scored = torch.cuda.Event()
                        scored.record(sS)
                cur.wait_stream(sR)
                cur.wait_stream(sS)

            run(args.warmup)
            barrier()
            s, e = ev(), ev()
            s.record()
            run(args.steps)
            e.record()
            barrier()
            ms = s.elapsed_time(e)
            if world > 1:
                t = torch.tensor([ms], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            return ms

        def timed(resident, record):
            for _ in range(args.warmup):
                episode(resident)
            barrier()
            s, e = ev(), ev()
            s.record()
            for _ in range(args.steps):
                episode(resident, record)
            e.record()
            barrier()
            ms = s.elapsed_time(e)
            if world > 1:
                t = torch.tensor([ms], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            return ms

        sampler = ClockSampler(local) if rank == 0 else None
        if sampler:
            sampler.start()
        # batches of at most one wave take spiral_env_step's fused kernel (include/spiral_b200.h): `value` is then timed
        # through spiral_env_step itself and the expand / composite split comes from a separate, untimed recorded pass
        fused = B <= int(os.environ.get("SPIRAL_FUSED_STEP_MAX", 148 * 8 * 2))
        ms = timed(resident=True, record=not fused)
        if fused:
            for _ in range(args.steps):
                episode(True, True)
            torch.cuda.synchronize()
        ms_e2e = timed(resident=False, record=False)
        ms_pipe = timed_pipelined() if args.pipeline else None
        if sampler:
            sampler.stop_flag = True
            sampler.join(timeout=2)
        env.check_status()

        steps_total = world * B * T * args.steps
        value = steps_total / (ms * 1e-3)
        e2e_value = steps_total / (ms_e2e * 1e-3)
        if rank != 0:
            return None

        def avg_ms(name):
            return float(np.mean([a_.elapsed_time(b_) for a_, b_ in kern_ev[name]]))

        t_exp, t_comp, t_disc, t_a2c = avg_ms("expand"), avg_ms("composite"), avg_ms("disc"), avg_ms("a2c")
        peak, peak_src = load_peaks()
        bytes_launch = BYTES_PER_CANVAS_STEP[C] * B
        share = {"expand": t_exp * T, "composite": t_comp * T, "disc": t_disc, "a2c": t_a2c}
        pair_gbs = bytes_launch / ((t_exp + t_comp) * 1e-3) / 1e9
        dominant = max(("expand", "composite"), key=lambda k: share[k])
        dom_ms = t_exp if dominant == "expand" else t_comp
        roof = dict(bound="hbm", kernel="spiral::%s_kernel" % dominant, achieved=bytes_launch / (dom_ms * 1e-3) / 1e9,
                    peak=peak, unit="GB/s", frac=bytes_launch / (dom_ms * 1e-3) / 1e9 / peak,
                    traffic=(COMPOSITE_TRAFFIC_PER_CANVAS_STEP * B if dominant == "composite" else 94.0e6 * B / 16384),
                    traffic_source="ncu --set full at B=16384 (profiles/r1_final_k1_ncu_summary_v2.md), scaled by B",
                    peak_source=peak_src,
                    note="algorithmic bytes = %d B/canvas-step x %d canvases per launch; K1 is instruction-issue / "
                         "latency-bound (serial brush state machine, ~270 dabs x ~20 px per painted step), its DRAM "
                         "traffic is BELOW the algorithmic bytes (sparse in-place compositing), see DESIGN.md" % (BYTES_PER_CANVAS_STEP[C], B),
                    k1_pair=dict(achieved=pair_gbs, frac=pair_gbs / peak, ms_expand=t_exp, ms_composite=t_comp),
                    a2c=dict(achieved=2.0 * dlogits_bytes / (t_a2c * 1e-3) / 1e9, unit="GB/s",
                             frac=2.0 * dlogits_bytes / (t_a2c * 1e-3) / 1e9 / peak, ms=t_a2c),
                    disc=dict(ms=t_disc, tflops=340.25e6 * B / (t_disc * 1e-3) / 1e12, precision=args.disc_precision))
        # reset + T x (expand [+ plan + scatter when the work-sorted schedule is on] + composite)
        # + D forward (conv0, 5 x conv_tc, 5 x bn_finalize, 4 x act_split, dense_sigmoid) + A2C (gae, row, final)
        launches_per_episode = 1 + ((1 if B <= 148 * 8 else 3) if fused else 4 if B >= 32768 else 2) * T + (1 + 5 + 5 + 4 + 1) + 3
        footprint = (env.canvas.numel() * 2 + env.obs.numel() * 4 + env.dab_scratch.numel() * 4 + 2 * dlogits_bytes
                     + disc.workspace.numel()) / 2 ** 20
        out = dict(metric="env-steps/sec (render+D reward+A2C) on 64x64 canvases", value=value, unit="env-steps/s",
                   n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms / args.steps,
                   higher_is_better=True, scaling="weak", vs_baseline=None, dtype="fix15/u16 render, f32 reward+loss",
                   data="synthetic",
                   config=dict(WORKLOAD, canvases_per_gpu=B, episode_length=T, heads=dict(zip(env.acs, env.head_sizes)),
                               jump=jump, disc_precision=args.disc_precision, parallelism="dp%d (actor partitions, no data-path collective)" % world,
                               l2="working set %.0f MiB per GPU > 126 MiB L2 (inputs larger than L2)" % footprint),
                   e2e=dict(value=e2e_value, unit="env-steps/s", h2d_bytes_per_step=int(B * T * K * 4 + 2 * B * T * 4),
                            d2h_bytes_per_step=int((B + 4) * 4), ms_per_step=ms_e2e / args.steps),
                   gpu_launches=launches_per_episode * args.steps,
                   kernel_ms_per_step=share, roofline=roof, clocks=sampler.summary() if sampler else None)
        if fused:
            out["env_step"] = ("fused step_small_kernel inside the timed region; kernel_ms_per_step.expand / .composite are "
                               "the two-kernel step's, recorded outside it")
        if ms_pipe is not None:
            out["pipelined"] = dict(value=steps_total / (ms_pipe * 1e-3), unit="env-steps/s", ms_per_step=ms_pipe / args.steps,
                                    note="render(i+1) on one stream overlapped with D reward + A2C(i) on another")
        del env, disc, logits, a2c_out
        torch.cuda.empty_cache()
        return out, (acs_list, head_list, C, L, T)

    def disc_train_step(args, dev, C, n=64, iters=5):
        """One WGAN-GP discriminator step (models/discriminator.py:97-136) at the reference's
        disc_batch_size: 3 shared-weight passes, gradient penalty with its double backward, all
        convolutions and their gradients on the library's kernels."""
        a = sp.get_args(["--env", WORKLOAD["env"], "--color_channel", str(C), "--loss", "gan", "--disc_dim",
                         str(WORKLOAD["disc_dim"]), "--disc_precision", args.disc_precision])
        D = models.Discriminator(a, None, [64, 64, C], norm_fn=lambda x: x, scope_name="global", device=dev, max_batch=n)
        for q in D.params.values():
            q.requires_grad_(True)
        g = torch.Generator(device=dev)
        g.manual_seed(3)
        fake = torch.rand((n, 64, 64, C), generator=g, device=dev) * 255
        real = torch.rand((n, 64, 64, C), generator=g, device=dev) * 255
        alpha = torch.rand((n, 1), generator=g, device=dev)

        def step():
            for q in D.params.values():
                q.grad = None
            out = D.wgan_gp_loss(fake, real, alpha, 20.0)
            out["d_loss"].backward()
            return out
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            out = step()
        e.record()
        torch.cuda.synchronize()
        return dict(ms_per_step=s.elapsed_time(e) / iters, batch=n, precision=args.disc_precision,
                    d_loss=float(out["d_loss"]), note="WGAN-GP step incl. double backward; conv fwd/dgrad/wgrad on K2b kernels, "
                                                      "BN/leaky-ReLU/dense glue in PyTorch")

    def agent_loop(args, dev, n=128, iters=3):
        """The whole synchronous loop at a small batch, for context: policy.act (PyTorch module + K4 sampling)
        -> env.step (K1) x T, D reward (K2) -> fused A2C (K3) -> policy backward -> K5 update, and one WGAN-GP
        step (K2b + K5).  NOT the headline metric: the learner's policy forward / backward is a PyTorch module
        (SURVEY 8f); rollout and learner step are CUDA graphs."""
        agent_mod = import_module(sp.__name__ + ".agent")
        a = sp.get_args(["--env", WORLD_ENV, "--location_size", str(WORKLOAD["location_size"]), "--color_channel",
                         str(WORKLOAD["color_channel"]), "--episode_length", str(args.episode_length), "--loss", "gan",
                         "--disc_dim", str(WORKLOAD["disc_dim"]), "--disc_precision", args.disc_precision,
                         "--disc_batch_size", "64"])
        env = sp.create_env(a, num_envs=n, device=dev)
        ag = agent_mod.Agent(a, env)
        times = {"rollout": 0.0, "train_policy": 0.0, "train_gan": 0.0}
        ev = lambda: torch.cuda.Event(enable_timing=True)
        warm = 3                                      # the rollout graph is captured at call 1, the learner graph at call 3
        for i in range(iters + warm):
            e0, e1, e2, e3 = ev(), ev(), ev(), ev()
            e0.record()
            ro = ag.rollout()
            e1.record()
            ag.train_policy(ro)
            e2.record()
            ag.train_gan()
            e3.record()
            torch.cuda.synchronize()
            if i >= warm:
                times["rollout"] += e0.elapsed_time(e1) / iters
                times["train_policy"] += e1.elapsed_time(e2) / iters
                times["train_gan"] += e2.elapsed_time(e3) / iters
        total = sum(times.values())
        return dict(canvases=n, ms=times, env_steps_per_s=n * args.episode_length / (total * 1e-3),
                    note="full loop incl. the PyTorch policy (not the headline metric)")

    WORLD_ENV = WORKLOAD["env"]
    res = measure(args.canvases)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    out, (acs_list, head_list, C, L, T) = res
    if world == 1 and not args.no_secondary:
        for nb, key, desc in ((16384, "rgb64_b16384", "configs[4] sweep point: 16384 canvases per GPU"),
                              (1024, "rgb64_b1024", "BASELINE.json configs[3]: 1024 canvases per GPU (latency-bound: "
                                                    "the longest canvas's serial brush chain)")):
            if nb == args.canvases:
                continue
            sec, _ = measure(nb)
            out[key] = {k: sec[k] for k in ("value", "unit", "ms_per_step", "e2e", "kernel_ms_per_step", "env_step") if k in sec}
            out[key]["config"] = desc
        if args.jump == "True":
            # the jump head is uniform over {0,1}, so half of the steps paint nothing; --jump=False is
            # the worst-case paint load (every step after the first draws a stroke)
            sec, _ = measure(16384, jump="False")
            out["rgb64_b16384_nojump"] = {k: sec[k] for k in ("value", "unit", "ms_per_step", "kernel_ms_per_step")}
            out["rgb64_b16384_nojump"]["config"] = "16384 canvases per GPU, --jump=False (worst-case paint load)"
        out["disc_train"] = disc_train_step(args, dev, C)
        out["agent_loop"] = agent_loop(args, dev)
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args, acs_list, head_list, C, L, T, WORKLOAD["disc_dim"])
    _emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import paint_oracle as po
    acs = po.py2_dict_order(["pressure", "jump", "size", "color", "control", "end"])
    C, L, T = WORKLOAD["color_channel"], WORKLOAD["location_size"], args.episode_length
    sizes = {"jump": 2, "pressure": 2, "size": 2, "color": 8, "control": L * L, "end": L * L}
    head_sizes = [sizes[a] for a in acs]
    cores = os.cpu_count() or 1
    episodes = 8 * cores                       # bounded sample of the workload per step
    for _ in range(min(args.warmup, 1)):
        cpu_reference_run(acs, L, C, T, head_sizes, WORKLOAD["disc_dim"], episodes=cores, cores=cores)
    t0 = time.perf_counter()
    parts = dict(render=0.0, disc=0.0, a2c=0.0)
    for s in range(args.steps):
        r = cpu_reference_run(acs, L, C, T, head_sizes, WORKLOAD["disc_dim"], episodes=episodes, cores=cores, seed=s)
        for k in parts:
            parts[k] += r[k]
    secs = sum(parts.values())
    del t0
    value = episodes * T * args.steps / secs
    sample = ("%d episodes x %d steps per bench step (bounded sample of the %d-canvas workload); CPU port of the "
              "reference path: C oracle of libmypaint 1.3.0 (libm mode) on %d worker processes + torch-CPU D forward "
              "+ numpy/torch-CPU A2C; the real libmypaint/TF-1.6 stack is not installable here"
              % (episodes, T, args.canvases, cores))
    out = dict(impl="reference", metric="env-steps/sec (render+D reward+A2C) on 64x64 canvases", value=value,
               unit="env-steps/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
               ms_per_step=secs / args.steps * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None,
               dtype="fix15/u16 render, f32 reward+loss", data="synthetic",
               config=dict(WORKLOAD, canvases_per_gpu=args.canvases, episode_length=T),
               cpu_baseline=dict(value=value, unit="env-steps/s", cores=cores, kind="port", sample=sample),
               e2e=dict(value=value, unit="env-steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
               seconds=parts)
    _emit(json.dumps(out))


def _emit(line):
    os.write(_REAL_STDOUT, (line + "\n").encode())


if __name__ == "__main__":
    # stdout carries exactly ONE JSON line: libraries that chat on fd 1 (NCCL prints its version banner there when a
    # communicator is created) are sent to stderr for the whole run, the result goes to the saved descriptor
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
