"""Times Agent.train_policy (the learner step: policy forward over [B*T] frames, fused A2C loss, backward, clip + Adam) at
128 canvases x 20 steps of the rgb64 shape, for the three convolution paths of the learner:
    native   csrc/policy_learn.cu (bf16x6, fp32 parity)          --policy_learn_native True
    fp32     PyTorch / cuDNN strict fp32                          --policy_learn_native False --policy_tf32 False
    tf32     PyTorch / cuDNN TF32 (10-bit operands)               --policy_learn_native False --policy_tf32 True

    python tools/learner_probe.py [native fp32 tf32]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spiral_b200 as sp  # noqa: E402
from importlib import import_module  # noqa: E402

agent_mod = import_module(sp.__name__ + ".agent")
MODES = {"native": ["--policy_learn_native", "True", "--policy_tf32", "False"],
         "native_x6": ["--policy_learn_native", "True", "--policy_tf32", "False", "--policy_learn_precision", "bf16x6"],
         "native_x3": ["--policy_learn_native", "True", "--policy_tf32", "False", "--policy_learn_precision", "bf16x3"],
         "fp32": ["--policy_learn_native", "False", "--policy_tf32", "False"],
         "tf32": ["--policy_learn_native", "False", "--policy_tf32", "True"]}
n = int(os.environ.get("PROBE_CANVASES", "128"))
for mode in (sys.argv[1:] or ["native", "native_x3", "fp32", "tf32"]):
    a = sp.get_args(["--env", "color_mnist", "--location_size", "64", "--color_channel", "3", "--episode_length", "20",
                     "--loss", "l2", "--cuda_graphs", os.environ.get("PROBE_GRAPHS", "True")] + MODES[mode])
    env = sp.create_env(a, num_envs=n)
    ag = agent_mod.Agent(a, env)
    ro = ag.rollout()
    times = []
    for i in range(int(os.environ.get("PROBE_ITERS", "8"))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        m = ag.train_policy({k: v.clone() for k, v in ro.items()})
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    print("%-7s train_policy %.1f ms (graph replay; first eager step %.1f ms), loss %.4f" % (mode, sum(times[-4:]) / len(times[-4:]), times[0], float(m["model/total_loss"])),
          flush=True)
    ag.close()
    del ag, env
    torch.cuda.empty_cache()

if os.environ.get("PROBE_BREAKDOWN"):
    # per-call-site device time of the learner's convolution kernels (eager, CUDA events around every launch)
    pl = import_module(sp.__name__ + ".models.policy_learn")
    a = sp.get_args(["--env", "color_mnist", "--location_size", "64", "--color_channel", "3", "--episode_length", "20",
                     "--loss", "l2", "--cuda_graphs", "False"] + MODES["native"])
    env = sp.create_env(a, num_envs=n)
    ag = agent_mod.Agent(a, env)
    ro = ag.rollout()
    ag.train_policy({k: v.clone() for k, v in ro.items()})
    rec = []

    def timed(fn, label):
        def w(*args, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*args, **kw)
            e1.record()
            rec.append((label(*args, **kw), e0, e1))
            return out
        return w
    pl._launch = timed(pl._launch, lambda x, packed, bias, out, cin, cout, hout, wout, kh, kw, stride, *r, **k:
                       "conv  %2d->%2d %dx%d s%d in %dx%d" % (cin, cout, kh, kw, stride, x.shape[1], x.shape[2]))
    pl._wgrad = timed(pl._wgrad, lambda x, dy, kh, kw, stride, *r: "wgrad %2d->%2d %dx%d s%d in %dx%d" % (x.shape[3], dy.shape[3], kh, kw, stride, x.shape[1], x.shape[2]))
    pl._pack = timed(pl._pack, lambda *r, **k: "pack")
    pl._pack_classes = timed(pl._pack_classes, lambda *r, **k: "pack (4 classes)")
    pl._launch_classes = timed(pl._launch_classes, lambda x, packed, bias, out, cin, cout, hout, wout, *r, **k:
                               "conv  %2d->%2d 4 parity classes of 2x2 taps, in %dx%d" % (cin, cout, x.shape[1], x.shape[2]))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ag.train_policy({k: v.clone() for k, v in ro.items()})
    e1.record()
    torch.cuda.synchronize()
    import collections
    agg, cnt = collections.Counter(), collections.Counter()
    for label, s, e in rec:
        agg[label] += s.elapsed_time(e)
        cnt[label] += 1
    tot = sum(agg.values())
    print("eager learner step %.1f ms, of which convolution kernels %.1f ms:" % (e0.elapsed_time(e1), tot))
    for k, v in agg.most_common(30):
        print("  %6.2f ms x%3d  %s" % (v, cnt[k], k))
