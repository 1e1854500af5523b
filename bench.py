#!/usr/bin/env python
"""bench.py -- env-steps/sec (render + D reward + A2C) on 64x64 canvases (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config rgb64|mnist-uncond|mnist-cond] [--scaling weak|strong] [--canvases B]
                    [--raster_math fast|det]

One "step" = one pass of the hot path over one batch of synthetic input: an episode of T env steps for B canvases per
GPU (reset + T x spiral_env_step), the discriminator reward on the final canvases (spiral_disc_forward), and the fused
A2C loss + gradient over [B,T] (spiral_a2c_loss).

Workloads (BASELINE.json configs):
  rgb64 (default)  configs[3] shape (color_channel=3, location_size=64, episode_length=20, heads control/end 4096,
                   color 8, jump/pressure/size 2, disc_dim=64) at the largest single-GPU batch of configs[4]'s sweep,
                   65,536 canvases per GPU; the 16,384- and 1,024-canvas points, the other arithmetic mode of the
                   rasteriser and the two MNIST configs are reported beside it in the same JSON line
  mnist-uncond     configs[1]: simple_mnist, episode_length=3, color_channel=1, location_size=32, disc_dim=8
  mnist-cond       configs[2]: simple_mnist conditional (D scores concat[canvas, target]), episode_length=5,
                   location_size=32 (BASELINE runs it on 2 GPUs: --gpus 2)

`--scaling weak` (default): --canvases is per GPU.  `--scaling strong`: --canvases is the TOTAL, split over the ranks
(B_total in {4096 .. 65536} at 1/2/4/8 GPUs).

`value`   : whole-job env-steps/s, inputs resident in HBM, CUDA-event timed, max over ranks.
`e2e`     : same metric through the Python env/model API with HOST (pinned) actions/values per call and a
            device->host read of rewards + loss scalars every episode.
`roofline`: K1 = the env step's kernel pair (expand_kernel + composite_kernel): algorithmic bytes per canvas-step
             x canvases per launch / CUDA-event time of the PAIR, against MEASURED_PEAKS.json;
            each kernel alone, K3 and K2 beside it.  `traffic` = ncu dram bytes at this batch when profiles/ holds them.
`parity`  : computed after the timed region on the tensors it left behind: sampled canvases against the CPU oracle in the
            kernels' arithmetic mode (bit-exact) and in the libm-faithful mode (north-star tolerance), the D reward of a
            256-image batch and the A2C loss of a 64-canvas slice against their float64 oracles.
`learner_exchange`: the gradient exchange of an optimiser step (discriminator 279 MB + policy) through
            distributed.GradReducer: time per step, bus GB/s against NVLink, and the K5 update that consumes the buffer.
`cpu_baseline` / --impl reference: the CPU oracle port of the reference path (libmypaint restatement + PyTorch-CPU
            discriminator + numpy A2C; the real libmypaint/TF-1.6 stack is not installable) on the host cores of this box.
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

WORKLOADS = {
    "rgb64": dict(workload="rgb64-sweep (configs[3] shape at configs[4]'s largest single-GPU batch)", env="color_mnist",
                  color_channel=3, location_size=64, episode_length=20, canvases_per_gpu=65536, disc_dim=64, conditional=False),
    "mnist-uncond": dict(workload="mnist-uncond (configs[1]: simple_mnist unconditional GAN)", env="simple_mnist",
                         color_channel=1, location_size=32, episode_length=3, canvases_per_gpu=65536, disc_dim=8, conditional=False),
    "mnist-cond": dict(workload="mnist-cond (configs[2]: simple_mnist conditional GAN, D scores concat[canvas, target])",
                       env="simple_mnist", color_channel=1, location_size=32, episode_length=5, canvases_per_gpu=65536,
                       disc_dim=64, conditional=True),
}
BYTES_PER_CANVAS_STEP = {1: 81920, 3: 114688}       # BASELINE.md section 3 (K1 algorithmic bytes)
NVLINK_BUS_GBS = 725.0                               # measured 8-rank all-reduce bus bandwidth on this pool (B200_PROFILING.md)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--config", default="rgb64", choices=sorted(WORKLOADS))
    p.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    p.add_argument("--canvases", type=int, default=None, help="canvases per GPU (weak) / in total (strong)")
    p.add_argument("--episode_length", type=int, default=None)
    p.add_argument("--raster_math", default=os.environ.get("SPIRAL_RASTER_MATH", "fast"), choices=["fast", "det"])
    p.add_argument("--disc_precision", default=os.environ.get("SPIRAL_DISC_PRECISION", None))
    p.add_argument("--no_cpu_baseline", action="store_true")
    p.add_argument("--cpu_seconds", type=float, default=15.0, help="target seconds of the bounded CPU sample")
    p.add_argument("--jump", default="True")
    p.add_argument("--no_secondary", action="store_true", help="only the headline measurement")
    p.add_argument("--no_parity", action="store_true")
    p.add_argument("--pipeline", action="store_true", help="also time the two-stream pipelined schedule")
    a = p.parse_args()
    a.W = dict(WORKLOADS[a.config])
    if a.episode_length is not None:
        a.W["episode_length"] = a.episode_length
    if a.canvases is None:
        a.canvases = a.W["canvases_per_gpu"]
    if a.disc_precision is None:
        a.disc_precision = "bf16x3" if a.W["disc_dim"] >= 64 else "fp32"
    return a


# ---------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference path on the host cores
# ---------------------------------------------------------------------------------------------
_W = {}


def _cpu_worker_init(acs, L, C):
    from oracle import paint_oracle as po
    spec = po.EnvSpec(action_names=acs, location_size=L, n_color=8 if "color" in acs else 0, math_mode=po.PO_MATH_LIBM)
    _W["env"] = po.OracleBatchEnv(spec, 1, channels=C)


def _cpu_worker_run(actions):
    env = _W["env"]
    out = []
    for a in actions:                      # a: [T,K]
        canv, obs = env.run_episodes(a[None], want_obs=True)
        out.append(obs[0, -1])
    return np.stack(out)


def cpu_reference_run(acs, L, C, T, head_sizes, disc_dim, episodes, cores, seed=0, conditional=False):
    """Times `episodes` whole episodes of the CPU port: render (P worker processes, one env each,
    the reference's actor layout run.py:45-52) + torch-CPU D forward + numpy A2C.  Returns seconds."""
    import multiprocessing as mp

    import torch

    from oracle import a2c_oracle as ao
    from oracle import disc_oracle as do
    rng = np.random.RandomState(seed)
    acts = np.stack([rng.randint(n, size=(episodes, T)) for n in head_sizes], -1).astype(np.int32)
    chunks = np.array_split(acts, cores)
    w = do.init_weights(C * (2 if conditional else 1), disc_dim=disc_dim, seed=11)
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
        x = finals[i:i + 64]
        if conditional:
            x = np.concatenate([x, x], -1)
        p, _ = do.predict(x, w, dtype=torch.float32)
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


def cpu_baseline(args, acs, head_sizes, C, L, T, disc_dim, conditional):
    cores = os.cpu_count() or 1
    # probe: 2 episodes per core to size the bounded sample
    probe = cpu_reference_run(acs, L, C, T, head_sizes, disc_dim, episodes=2 * cores, cores=cores, conditional=conditional)
    per_ep = probe["seconds"] / (2 * cores)
    episodes = int(max(cores * 2, min(4096, args.cpu_seconds / max(per_ep, 1e-6))))
    episodes = (episodes // cores) * cores
    res = cpu_reference_run(acs, L, C, T, head_sizes, disc_dim, episodes=episodes, cores=cores, seed=1, conditional=conditional)
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


def load_traffic(B, C, math):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the two K1 kernels from the ncu --set full capture kept under
    profiles/ (tools/ncu_traffic.py writes the file); None when there is no capture for this batch / mode."""
    path = os.path.join(ROOT, "profiles", "r2_k1_traffic.json")
    if not os.path.exists(path):
        return None, None
    with open(path) as fp:
        rows = json.load(fp)
    for r in rows:
        if int(r.get("canvases", -1)) == int(B) and int(r.get("channels", 3)) == int(C) and r.get("raster_math") == math:
            return float(r["expand_bytes"]) + float(r["composite_bytes"]), r.get("source")
    return None, None


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import spiral_b200 as sp
    from importlib import import_module
    rl = import_module(sp.__name__ + ".rl_utils")
    models = import_module(sp.__name__ + ".models")
    du = import_module(sp.__name__ + ".distributed")

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def measure(W, B, jump=None, math=None, parity=False, record_kernels=True):
        """B canvases on THIS rank."""
        T = W["episode_length"]
        jump = args.jump if jump is None else jump
        math = args.raster_math if math is None else math
        C, L, cond = W["color_channel"], W["location_size"], W["conditional"]
        a = sp.get_args(["--env", W["env"], "--location_size", str(L), "--color_channel", str(C),
                         "--episode_length", str(T), "--loss", "gan", "--disc_dim", str(W["disc_dim"]),
                         "--disc_precision", args.disc_precision, "--jump", jump, "--seed", str(123 + rank),
                         "--raster_math", math])
        env = sp.create_env(a, num_envs=B, device=dev)
        a.conditional = cond                            # the D graph of a conditional GAN: concat[canvas, target] (discriminator.py:36-38)
        disc = models.Discriminator(a, None, [64, 64, C], norm_fn=env.norm, scope_name="global", device=dev, max_batch=B)
        with torch.no_grad():                           # zero-initialised biases / unit BN would hide mistakes: perturb them (seeded)
            gw = torch.Generator(device=dev)
            gw.manual_seed(11)
            for k, v in disc.params.items():
                if k.endswith("/bias") or k.endswith("/beta"):
                    v.add_(0.05 * torch.randn(v.shape, generator=gw, device=dev))
                elif k.endswith("/gamma"):
                    v.add_(0.1 * torch.randn(v.shape, generator=gw, device=dev))
        disc.sync_weights()
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
        acts_bt = acts_dev.permute(1, 0, 2).contiguous()
        a2c_out = rl.a2c_loss(logits, acts_bt, rewards, values_dev, vf_dev)   # allocates outputs
        probs = torch.empty((B,), device=dev)
        lg = torch.empty((B,), device=dev)
        host_out = torch.empty((B + 4,), dtype=torch.float32).pin_memory()
        target = None
        if cond:                                      # synthetic digit-shaped targets, seed 7
            target = env.make_synthetic_targets(count=min(B, 2048))
            target = target[torch.arange(B, device=dev) % target.shape[0]].float().contiguous()
            d_in = torch.empty((B, 64, 64, 2 * C), device=dev)

        ev = lambda: torch.cuda.Event(enable_timing=True)
        kern_ev = {"expand": [], "composite": [], "disc": [], "a2c": []}
        N = sp._native
        lib = N.lib()

        def reward_forward():
            if cond:
                torch.cat([env.state, target], -1, out=d_in)
                disc.conditional = False              # the concatenation is done here; forward sees 2C channels
                disc.forward(d_in, logits_out=lg, probs_out=probs)
            else:
                disc.forward(env.state, logits_out=lg, probs_out=probs)          # agent.py:336-338: GAN reward

        if cond:
            disc.conditional = False

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
            reward_forward()
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
            return max_over_ranks(s.elapsed_time(e))

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
            return max_over_ranks(s.elapsed_time(e))

        sampler = ClockSampler(local) if rank == 0 else None
        if sampler:
            sampler.start()
        # batches of at most two waves take spiral_env_step's fused kernel (include/spiral_b200.h): `value` is then timed
        # through spiral_env_step itself and the expand / composite split comes from a separate, untimed recorded pass
        fused = B <= int(os.environ.get("SPIRAL_FUSED_STEP_MAX", 148 * 8 * (5 if math == "fast" else 2)))
        ms = timed(resident=True, record=(not fused) and record_kernels)
        if fused or not record_kernels:
            for _ in range(max(2, args.steps // 2)):
                episode(True, True)
            torch.cuda.synchronize()
        ms_e2e = timed(resident=False, record=False)
        ms_pipe = timed_pipelined() if (args.pipeline and not cond) else None
        if sampler:
            sampler.stop_flag = True
            sampler.join(timeout=2)
        env.check_status()

        B_total = B * world if args.scaling == "weak" else args.canvases_total
        steps_total = B_total * T * args.steps
        value = steps_total / (ms * 1e-3)
        e2e_value = steps_total / (ms_e2e * 1e-3)
        if rank != 0:
            del env, disc, logits, a2c_out
            torch.cuda.empty_cache()
            return None

        def avg_ms(name):
            return float(np.mean([a_.elapsed_time(b_) for a_, b_ in kern_ev[name]]))

        t_exp, t_comp, t_disc, t_a2c = avg_ms("expand"), avg_ms("composite"), avg_ms("disc"), avg_ms("a2c")
        peak, peak_src = load_peaks()
        bytes_launch = BYTES_PER_CANVAS_STEP[C] * B
        share = {"expand": t_exp * T, "composite": t_comp * T, "disc": t_disc, "a2c": t_a2c}
        pair_gbs = bytes_launch / ((t_exp + t_comp) * 1e-3) / 1e9
        traffic, traffic_src = load_traffic(B, C, math)
        flops_img = {(3, 64): 340.25e6, (1, 64): 333.9e6, (1, 8): 5.56e6}.get((C, W["disc_dim"]))
        roof = dict(bound="hbm", kernel="K1 env step = spiral::expand_kernel + spiral::composite_kernel (one canvas-step needs both)",
                    achieved=pair_gbs, peak=peak, unit="GB/s", frac=pair_gbs / peak, traffic=traffic,
                    traffic_source=traffic_src, peak_source=peak_src,
                    note="algorithmic bytes = %d B/canvas-step x %d canvases per launch / (expand + composite) time; K1 is "
                         "instruction-issue bound (serial brush state machine per canvas, ~270 dabs x ~20 px x ~100 instructions "
                         "of libmypaint's anti-aliased mask per painted step), its DRAM traffic is BELOW the algorithmic "
                         "bytes (sparse in-place compositing), see DESIGN.md" % (BYTES_PER_CANVAS_STEP[C], B),
                    expand=dict(ms=t_exp, achieved=bytes_launch / (t_exp * 1e-3) / 1e9, frac=bytes_launch / (t_exp * 1e-3) / 1e9 / peak),
                    composite=dict(ms=t_comp, achieved=bytes_launch / (t_comp * 1e-3) / 1e9,
                                   frac=bytes_launch / (t_comp * 1e-3) / 1e9 / peak),
                    a2c=dict(achieved=2.0 * dlogits_bytes / (t_a2c * 1e-3) / 1e9, unit="GB/s",
                             frac=2.0 * dlogits_bytes / (t_a2c * 1e-3) / 1e9 / peak, ms=t_a2c),
                    disc=dict(ms=t_disc, tflops=(flops_img * B / (t_disc * 1e-3) / 1e12) if flops_img else None,
                              precision=args.disc_precision))
        # reset + T x (expand [+ plan + scatter when the work-sorted schedule is on] + composite)
        # + D forward (conv0, 5 x conv_tc, 5 x bn_finalize, 4 x act_split, dense_sigmoid) + A2C (gae, row, final)
        launches_per_episode = 1 + ((1 if B <= 148 * 8 else 3) if fused else 4 if B >= 32768 else 2) * T + (1 + 5 + 5 + 4 + 1) + 3
        footprint = (env.canvas.numel() * 2 + env.obs.numel() * 4 + env.dab_scratch.numel() * 4 + 2 * dlogits_bytes
                     + disc.workspace.numel()) / 2 ** 20
        out = dict(metric="env-steps/sec (render+D reward+A2C) on 64x64 canvases", value=value, unit="env-steps/s",
                   n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms / args.steps,
                   higher_is_better=True, scaling=args.scaling, vs_baseline=None, dtype="fix15/u16 render, f32 reward+loss",
                   data="synthetic",
                   config=dict(workload=W["workload"], env=W["env"], color_channel=C, location_size=L, episode_length=T,
                               disc_dim=W["disc_dim"], conditional=cond, canvases_per_gpu=B, canvases_total=B_total,
                               heads=dict(zip(env.acs, env.head_sizes)), jump=jump, raster_math=math,
                               disc_precision=args.disc_precision,
                               parallelism="dp%d (actor partitions, no data-path collective; the gradient exchange of an optimiser "
                                           "step is timed under learner_exchange)" % world,
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
        if parity:
            out["parity"] = parity_block(env, disc, acts_host, logits, acts_bt, rewards, values_dev, vf_dev, a2c_out, probs,
                                         W, math, d_in if cond else None)
        del env, disc, logits, a2c_out
        torch.cuda.empty_cache()
        return out, (acs_list, head_list, C, L, T)

    def parity_block(env, disc, acts_host, logits, acts_bt, rewards, values_dev, vf_dev, a2c_out, probs, W, math, d_in):
        """Correctness evidence for the number above, from the tensors the last timed episode left behind (outside the timed
        region; the oracle is the checker, never the thing measured)."""
        from oracle import a2c_oracle as ao
        from oracle import disc_oracle as do
        from oracle import paint_oracle as po
        B, T = acts_bt.shape[0], acts_bt.shape[1]
        C = W["color_channel"]
        rng = np.random.RandomState(0)
        pick = np.unique(np.concatenate([[0, B - 1], rng.randint(0, B, 14)]))
        acts = acts_host.permute(1, 0, 2).numpy()[pick]                       # [n,T,K]
        got_canvas = env.canvas[torch.as_tensor(pick, device=dev)].cpu().numpy().view(np.uint16)
        got_obs = env.state[torch.as_tensor(pick, device=dev)].cpu().numpy()
        res = {}
        for name, mode in (("kernel_mode", po.PO_MATH_FAST if math == "fast" else po.PO_MATH_DET), ("libm", po.PO_MATH_LIBM)):
            spec = po.EnvSpec(action_names=list(env.acs), location_size=env.location_size, n_color=len(env.colors),
                              colorize=env.colorize, math_mode=mode)
            orc = po.OracleBatchEnv(spec, len(pick), channels=C)
            want_canvas, want_obs = orc.run_episodes(acts, want_obs=True)
            err = np.abs(got_obs.astype(np.float64) - want_obs[:, -1]).reshape(len(pick), -1) / 255.0
            res[name] = dict(canvases_bit_identical=int((got_canvas == want_canvas).reshape(len(pick), -1).all(1).sum()),
                             max_abs=float(err.max()), mean_abs=float(err.mean()),
                             canvases_within_1_255=int((err.max(1) <= 1.0 / 255 + 1e-12).sum()))
        out = dict(canvases=dict(sampled=int(len(pick)), raster_math=math,
                                 oracle_same_mode=res["kernel_mode"], oracle_libm=res["libm"],
                                 bar="bit-identical to the oracle in the kernels' mode; vs the libm-faithful mode: max |err| <= 1/255 per "
                                     "channel, mean <= 1e-3 on [0,1] observations (north star)"))
        # D reward: a 256-image batch of the final canvases through the same kernels (its own batch statistics) vs float64
        n = min(256, B)
        x = (d_in if d_in is not None else env.state)[:n].contiguous()
        p_gpu, l_gpu = disc.forward(x)
        w = {}
        for k, v in disc.params.items():               # TensorFlow variable names -> the oracle's
            if k.startswith("batch_normalization"):
                head, leaf = k.split("/")
                l = 1 if head == "batch_normalization" else int(head.rsplit("_", 1)[1]) + 1
                k = "bn%d/%s" % (l, leaf)
            w[k] = v.detach().cpu().numpy()
        want_p, want_l = do.predict(x.cpu().numpy(), w, dtype=torch.float64)
        out["reward"] = dict(images=n, max_abs_err_prob=float(np.abs(p_gpu.cpu().numpy() - want_p).max()),
                             max_rel_err_logit=float(np.abs(l_gpu.cpu().numpy() - want_l).max() / max(1.0, np.abs(want_l).max())),
                             bar="1e-4 (fp32 / bf16x3)")
        # A2C: the loss of a 64-canvas slice (row-local gradients, sums) vs the float64 oracle
        m = min(64, B)
        sl = rl.a2c_loss([l[:m].contiguous() for l in logits], acts_bt[:m].contiguous(), rewards[:m].contiguous(),
                         values_dev[:m].contiguous(), vf_dev[:m].contiguous(), 0.99, 1.0, 0.01)
        r_, adv_ = ao.multiple_process_rollout(rewards[:m].cpu().numpy().astype(np.float64), values_dev[:m].cpu().numpy().astype(np.float64), 0.99, 1.0)
        want = ao.a2c_loss([l[:m].cpu().numpy() for l in logits], acts_bt[:m].cpu().numpy(), adv_.astype(np.float32),
                           r_.astype(np.float32), vf_dev[:m].cpu().numpy(), 0.01)
        got = sl.scalars.cpu().numpy()
        out["a2c"] = dict(canvases=m, rel_err_total=float(abs(got[3] - want["total"]) / max(1.0, abs(want["total"]))),
                          rel_err_dlogits=float(max(np.abs(sl.dlogits[k].cpu().numpy() - want["dlogits"][k]).max() /
                                                    max(np.abs(want["dlogits"][k]).max(), 1e-12) for k in range(len(logits)))),
                          bar="1e-4")
        out["ok"] = bool(res["kernel_mode"]["canvases_bit_identical"] == len(pick) and res["libm"]["mean_abs"] <= 1e-3
                         and out["reward"]["max_abs_err_prob"] <= 1e-4 and out["a2c"]["rel_err_total"] <= 1e-4)
        return out

    def learner_exchange():
        """The exchange step of an optimiser step: all-reduce of the discriminator's gradients (disc_dim 64:
        ~279 MB fp32) and the policy's through distributed.GradReducer's persistent flat buffers, then the K5 update that reads
        them in place.  Collective time alone (CUDA events), algorithmic bus bandwidth 2(N-1)/N x bytes / time."""
        optim = import_module(sp.__name__ + ".optim")
        pol_mod = import_module(sp.__name__ + ".models.policy")
        W = WORKLOADS["rgb64"]
        a = sp.get_args(["--env", W["env"], "--location_size", str(W["location_size"]), "--color_channel", str(W["color_channel"]),
                         "--episode_length", "20", "--loss", "gan", "--disc_dim", "64"])
        env = sp.create_env(a, num_envs=8, device=dev)
        torch.manual_seed(0)
        pi = pol_mod.Policy(a, env).to(dev)
        D = models.Discriminator(a, None, [64, 64, 3], norm_fn=env.norm, scope_name="global", device=dev, max_batch=8)
        res = {}
        for name, params, kw in (("discriminator", list(D.params.values()), dict(beta1=0.5, beta2=0.9)),
                                 ("policy", list(pi.parameters()), {})):
            for p in params:
                p.requires_grad_(True)
            red = du.GradReducer(params, world=world)
            opt = optim.FusedTFAdam(params, 1e-5, **kw)
            evs = []
            for it in range(3 + 10):
                red.zero()
                red.flat.normal_(generator=None)
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                e0.record()
                red.wait()                                            # every bucket: all-reduce (sum), joined
                e1.record()
                opt.step(clip_norm=40.0 if name == "policy" else 0.0, grad_scale=1.0 / world)
                e2.record()
                if it >= 3:
                    evs.append((e0, e1, e2))
            barrier()
            coll = max_over_ranks(float(np.mean([a_.elapsed_time(b_) for a_, b_, _ in evs])))
            k5 = float(np.mean([b_.elapsed_time(c_) for _, b_, c_ in evs]))
            nbytes = red.collective_bytes()
            bus = (2.0 * (world - 1) / world) * nbytes / (coll * 1e-3) / 1e9 if world > 1 else None
            res[name] = dict(bytes=nbytes, buckets=len(red.buckets), collective_ms=coll if world > 1 else 0.0, bus_gbs=bus,
                             bus_frac_of_nvlink=(bus / NVLINK_BUS_GBS) if bus else None, k5_update_ms=k5)
        res["note"] = ("NCCL all-reduce over NVLink/NVSwitch, %d ranks; reference bus bandwidth %.0f GB/s (8-rank all-reduce at 1 GiB, "
                       "B200_PROFILING.md); in training the buckets are launched from backward hooks and overlap the backward pass "
                       "(agent.py train_policy / train_gan)" % (world, NVLINK_BUS_GBS))
        del env, pi, D
        torch.cuda.empty_cache()
        return res

    def disc_train_step(C, n=64, iters=5):
        """One WGAN-GP discriminator step (models/discriminator.py:97-136) at the reference's
        disc_batch_size: 3 shared-weight passes, gradient penalty with its double backward, all
        convolutions and their gradients on the library's kernels."""
        W = WORKLOADS["rgb64"]
        a = sp.get_args(["--env", W["env"], "--color_channel", str(C), "--loss", "gan", "--disc_dim",
                         str(W["disc_dim"]), "--disc_precision", "bf16x3"])
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
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # (1) as Agent.train_gan runs it: one CUDA graph replay per step (captured before any eager step touches the
        # parameters: autograd pins a leaf's gradient accumulation to the stream it first ran on)
        graph_ms = None
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    step()
            torch.cuda.current_stream().wait_stream(side)
            for q in D.params.values():
                q.grad = None
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_):
                o = D.wgan_gp_loss(fake, real, alpha, 20.0)
                o["d_loss"].backward()
            g_.replay()
            torch.cuda.synchronize()
            s.record()
            for _ in range(iters):
                g_.replay()
            e.record()
            torch.cuda.synchronize()
            graph_ms = s.elapsed_time(e) / iters
            del g_
        except Exception as ex:                              # noqa: BLE001 -- report, the eager number still stands
            graph_ms = "capture failed: %s" % (str(ex).splitlines()[0],)
        # (2) eager (one launch per kernel from Python)
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        s.record()
        for _ in range(iters):
            out = step()
        e.record()
        torch.cuda.synchronize()
        return dict(ms_per_step=graph_ms if isinstance(graph_ms, float) else s.elapsed_time(e) / iters,
                    ms_per_step_graph=graph_ms, ms_per_step_eager=s.elapsed_time(e) / iters, batch=n,
                    precision="bf16x3 forward / bf16x6 gradients", d_loss=float(out["d_loss"].detach()),
                    note="WGAN-GP step incl. double backward; conv fwd/dgrad/wgrad on K2b kernels (each kernel's operand planes "
                         "split once per step), BN/leaky-ReLU/dense glue in PyTorch; ms_per_step = the CUDA-graph replay "
                         "Agent.train_gan uses")

    def agent_loop(n=128, iters=3):
        """The whole synchronous loop at a small batch, for context: policy.act (PyTorch module + K4 sampling)
        -> env.step (K1) x T, D reward (K2) -> fused A2C (K3) -> policy backward -> K5 update, and one WGAN-GP
        step (K2b + K5).  NOT the headline metric: the learner's policy forward / backward is a PyTorch module
        (SURVEY 8f); rollout and learner step are CUDA graphs."""
        agent_mod = import_module(sp.__name__ + ".agent")
        W = WORKLOADS["rgb64"]
        a = sp.get_args(["--env", W["env"], "--location_size", str(W["location_size"]), "--color_channel",
                         str(W["color_channel"]), "--episode_length", "20", "--loss", "gan",
                         "--disc_dim", str(W["disc_dim"]), "--disc_precision", "bf16x3",
                         "--disc_batch_size", "64", "--raster_math", args.raster_math])
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
        return dict(canvases=n, ms=times, env_steps_per_s=n * 20 / (total * 1e-3),
                    acting_precision="bf16 (--policy_act_native True: the actors' convolutions on bf16 tensor cores, logits within 3e-2 "
                                     "of the fp32 module -- narrower than the reference's fp32 actors; it only perturbs the behaviour "
                                     "policy's samples, the learner's logits / loss are fp32-parity)",
                    learner_precision="f16x3 (--policy_learn_native True: 22 operand bits in two fp16 planes, gradients as close to float64 as PyTorch strict fp32)",
                    note="full loop incl. the policy (not the headline metric)")

    def learner_paths(n=128):
        """Agent.train_policy (policy forward over [n*20] frames, fused A2C loss, backward, clip + Adam; CUDA-graph replay) for
        the convolution paths of the learner: the library's kernels (f16x3: the default, fp32 parity; bf16x6: fp32 parity; bf16x3: 16 operand bits), PyTorch / cuDNN in
        strict fp32 (the parity-equivalent library path) and in TF32 (10-bit operands, what round 1 ran)."""
        agent_mod = import_module(sp.__name__ + ".agent")
        W = WORKLOADS["rgb64"]
        modes = {"native_f16x3": ["--policy_learn_native", "True", "--policy_tf32", "False", "--policy_learn_precision", "f16x3"],
                 "native_bf16x6": ["--policy_learn_native", "True", "--policy_tf32", "False", "--policy_learn_precision", "bf16x6"],
                 "native_bf16x3": ["--policy_learn_native", "True", "--policy_tf32", "False", "--policy_learn_precision", "bf16x3"],
                 "cudnn_fp32": ["--policy_learn_native", "False", "--policy_tf32", "False"],
                 "cudnn_tf32": ["--policy_learn_native", "False", "--policy_tf32", "True"]}
        res = {}
        for name, extra in modes.items():
            a = sp.get_args(["--env", W["env"], "--location_size", str(W["location_size"]), "--color_channel",
                             str(W["color_channel"]), "--episode_length", "20", "--loss", "l2", "--raster_math", args.raster_math] + extra)
            env = sp.create_env(a, num_envs=n, device=dev)
            ag = agent_mod.Agent(a, env)
            ro = ag.rollout()
            ts = []
            for i in range(7):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ag.train_policy({k: v.clone() for k, v in ro.items()})
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            res[name] = dict(train_policy_ms=float(np.mean(ts[3:])))
            ag.close()
            del ag, env
            torch.cuda.empty_cache()
        res["note"] = ("%d canvases x 20 steps = %d frames per learner step; native = csrc/policy_learn.cu (mma.sync implicit GEMMs, forward + "
                       "dgrad + wgrad; f16x3 = 22 operand bits in two fp16 planes, three products: gradients as close to float64 as cuDNN "
                       "strict fp32, tests/test_gpu_policy_native.py); TF32 is ~1e-3" % (n, n * 20))
        return res

    # ---- headline ----
    if args.scaling == "strong":
        args.canvases_total = args.canvases
        _, B = du.shard(args.canvases, rank, world)
    else:
        args.canvases_total = args.canvases * world
        B = args.canvases
    res = measure(args.W, B, parity=not args.no_parity)
    exch = learner_exchange() if (world > 1 or not args.no_secondary) else None
    if rank != 0:
        if world > 1:
            # the secondary measurements below run on rank 0 only; the other ranks are done
            dist.barrier()
            dist.destroy_process_group()
        return
    out, (acs_list, head_list, C, L, T) = res
    if exch is not None:
        out["learner_exchange"] = exch
    pick = ("value", "unit", "ms_per_step", "e2e", "kernel_ms_per_step", "env_step", "roofline", "parity")
    if world == 1 and not args.no_secondary and args.config == "rgb64":
        for nb, key, desc in ((16384, "rgb64_b16384", "configs[4] sweep point: 16384 canvases per GPU"),
                              (1024, "rgb64_b1024", "BASELINE.json configs[3]: 1024 canvases per GPU (latency-bound: "
                                                    "the longest canvas's serial brush chain)")):
            if nb == B:
                continue
            sec, _ = measure(args.W, nb)
            out[key] = {k: sec[k] for k in pick if k in sec}
            out[key]["config"] = desc
        other = "det" if args.raster_math == "fast" else "fast"
        sec, _ = measure(args.W, B, math=other, parity=not args.no_parity)
        out["raster_math_" + other] = {k: sec[k] for k in pick if k in sec}
        out["raster_math_" + other]["config"] = ("same workload with --raster_math %s (%s)" % (other, {
            "det": "double-precision correctly rounded exp / log / hypot: canvases bit-identical to the oracle's det mode, which "
                   "matches glibc on >= 99 % of canvases",
            "fast": "binary32 polynomials and a shorter dab-loop chain: bit-identical to the oracle's fast mode"}[other]))
        if args.jump == "True":
            # the jump head is uniform over {0,1}, so half of the steps paint nothing; --jump=False is
            # the worst-case paint load (every step after the first draws a stroke)
            sec, _ = measure(args.W, 16384, jump="False")
            out["rgb64_b16384_nojump"] = {k: sec[k] for k in ("value", "unit", "ms_per_step", "kernel_ms_per_step")}
            out["rgb64_b16384_nojump"]["config"] = "16384 canvases per GPU, --jump=False (worst-case paint load)"
        for cfg in ("mnist-uncond", "mnist-cond"):
            Wc = WORKLOADS[cfg]
            keep_prec = args.disc_precision
            args.disc_precision = "bf16x3" if Wc["disc_dim"] >= 64 else "fp32"
            sec, _ = measure(Wc, Wc["canvases_per_gpu"], parity=not args.no_parity)
            args.disc_precision = keep_prec
            out[cfg] = {k: sec[k] for k in pick if k in sec}
            out[cfg]["config"] = sec["config"]
        out["disc_train"] = disc_train_step(C)
        out["agent_loop"] = agent_loop()
        out["learner_paths"] = learner_paths()
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args, acs_list, head_list, C, L, T, args.W["disc_dim"], args.W["conditional"])
    _emit(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import paint_oracle as po
    W = args.W
    names = {"color_mnist": ["pressure", "jump", "size", "color", "control", "end"],
             "mnist": ["pressure", "jump", "size", "control", "end"], "simple_mnist": ["jump", "control", "end"]}[W["env"]]
    acs = po.py2_dict_order(names)
    C, L, T = W["color_channel"], W["location_size"], W["episode_length"]
    sizes = {"jump": 2, "pressure": 2, "size": 2, "color": 8, "control": L * L, "end": L * L}
    head_sizes = [sizes[a] for a in acs]
    cores = os.cpu_count() or 1
    episodes = 8 * cores                       # bounded sample of the workload per step
    for _ in range(min(args.warmup, 1)):
        cpu_reference_run(acs, L, C, T, head_sizes, W["disc_dim"], episodes=cores, cores=cores, conditional=W["conditional"])
    parts = dict(render=0.0, disc=0.0, a2c=0.0)
    for s in range(args.steps):
        r = cpu_reference_run(acs, L, C, T, head_sizes, W["disc_dim"], episodes=episodes, cores=cores, seed=s,
                              conditional=W["conditional"])
        for k in parts:
            parts[k] += r[k]
    secs = sum(parts.values())
    value = episodes * T * args.steps / secs
    sample = ("%d episodes x %d steps per bench step (bounded sample of the %d-canvas workload); CPU port of the "
              "reference path: C oracle of libmypaint 1.3.0 (libm mode) on %d worker processes + torch-CPU D forward "
              "+ numpy/torch-CPU A2C; the real libmypaint/TF-1.6 stack is not installable here"
              % (episodes, T, args.canvases, cores))
    out = dict(impl="reference", metric="env-steps/sec (render+D reward+A2C) on 64x64 canvases", value=value,
               unit="env-steps/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
               ms_per_step=secs / args.steps * 1e3, higher_is_better=True, scaling=args.scaling, vs_baseline=None,
               dtype="fix15/u16 render, f32 reward+loss", data="synthetic",
               config=dict(workload=W["workload"], env=W["env"], color_channel=C, location_size=L, episode_length=T,
                           disc_dim=W["disc_dim"], conditional=W["conditional"], canvases_per_gpu=args.canvases),
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
