"""Driver for ncu / timing: WGAN-GP discriminator steps (batch 64, disc_dim 64, C=3) on the K2b kernels.

    python tools/profile_dstep.py [iters] [precision]
"""
import os, sys, time
from importlib import import_module
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import spiral_b200 as sp
models = import_module(sp.__name__ + ".models")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
n, C = 64, 3
dev = torch.device("cuda", 0)
a = sp.get_args(["--env", "color_mnist", "--color_channel", str(C), "--loss", "gan", "--disc_dim", "64", "--disc_precision", prec])
D = models.Discriminator(a, None, [64, 64, C], norm_fn=lambda x: x, scope_name="global", device=dev, max_batch=n)
for q in D.params.values():
    q.requires_grad_(True)
g = torch.Generator(device=dev); g.manual_seed(3)
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
if os.environ.get("DSTEP_GRAPH", "1") != "only":
  step(); torch.cuda.synchronize()
  t0 = time.perf_counter()
  s.record()
  for _ in range(iters):
    out = step()
  e.record(); torch.cuda.synchronize()
  print("ok d_loss %.5f  %.2f ms/step (events)  %.2f ms/step (wall)" % (float(out["d_loss"].detach()), s.elapsed_time(e) / iters, (time.perf_counter() - t0) * 1e3 / iters))
if os.environ.get("DSTEP_GRAPH", "1") != "0":
    # the same step as a CUDA graph replay (how Agent.train_gan runs it): no launch gaps
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            o = D.wgan_gp_loss(fake, real, alpha, 20.0)
            o["d_loss"].backward()
    torch.cuda.current_stream().wait_stream(side)
    for q in D.params.values():
        q.grad = None                                         # the gradients live in the graph's pool (as in Agent.train_gan)
    g_ = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_):
        o = D.wgan_gp_loss(fake, real, alpha, 20.0)
        o["d_loss"].backward()
    g_.replay(); torch.cuda.synchronize()
    s.record()
    for _ in range(iters):
        g_.replay()
    e.record(); torch.cuda.synchronize()
    print("graph replay: d_loss %.5f  %.2f ms/step" % (float(o["d_loss"].detach()), s.elapsed_time(e) / iters))
