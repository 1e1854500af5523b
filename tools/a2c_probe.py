"""Times K3 (spiral_a2c_loss) on the rgb64 head shape (B200).   python tools/a2c_probe.py [B ...]
SPIRAL_B200_LIB selects a tuning build of the library."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spiral_b200 as sp  # noqa: E402
from importlib import import_module  # noqa: E402

rl = import_module(sp.__name__ + ".rl_utils")
HEADS = [4096, 4096, 8, 2, 2, 2]
T = 20

for B in [int(x) for x in sys.argv[1:]] or [16384, 65536]:
    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    logits = [torch.randn((B, T, n), generator=g, device="cuda") for n in HEADS]
    acts = torch.stack([torch.randint(0, n, (B, T), generator=g, device="cuda") for n in HEADS], -1).to(torch.int32)
    rew = torch.rand((B, T), generator=g, device="cuda")
    val = torch.randn((B, T), generator=g, device="cuda")
    vf = torch.randn((B, T), generator=g, device="cuda")
    out = rl.a2c_loss(logits, acts, rew, val, vf)
    for _ in range(2):
        rl.a2c_loss(logits, acts, rew, val, vf, out=out)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        rl.a2c_loss(logits, acts, rew, val, vf, out=out)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    nbytes = 2 * 4 * sum(HEADS) * B * T
    print("lib=%s B=%d  %.3f ms  %.0f GB/s  scalars=%s" % (os.path.basename(os.environ.get("SPIRAL_B200_LIB", "default")), B, ms,
          nbytes / ms / 1e6, [round(float(v), 3) for v in out.scalars.cpu()]), flush=True)
    del logits, out
    torch.cuda.empty_cache()
