"""Per-tensor gradient error of the native WGAN-GP step vs the float64 oracle, for each precision."""
import os, sys
from importlib import import_module
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import spiral_b200 as sp
from oracle import disc_oracle as do
models = import_module(sp.__name__ + ".models")
C, dd, B = 3, 64, int(sys.argv[1]) if len(sys.argv) > 1 else 8
w = do.init_weights(C, disc_dim=dd, batch_norm=True, seed=5)
rng = np.random.RandomState(2)
def canvases():
    imgs = np.full((B, 64, 64, C), 255.0, np.float32)
    for b in range(B):
        for _ in range(6):
            y, x = rng.randint(0, 56, 2)
            imgs[b, y:y + 8, x:x + 8] = rng.randint(0, 255)
    return imgs
fake, real = canvases(), canvases()
alpha = rng.rand(B, 1).astype(np.float32)
want = do.d_loss(fake, real, alpha, w, batch_norm=True, wgan_lambda=20.0)
for prec in ("fp32", "bf16x6", "bf16x3", "bf16"):
    args = sp.get_args(["--color_channel", str(C), "--disc_dim", str(dd), "--disc_precision", "bf16x3" if prec == "bf16x6" else prec,
                        "--disc_grad_precision", prec])
    D = models.Discriminator(args, None, [64, 64, C], norm_fn=lambda x: x, scope_name="global", max_batch=B)
    D.load_numpy(w)
    for p in D.params.values():
        p.requires_grad_(True)
    out = D.wgan_gp_loss(torch.tensor(fake, device="cuda"), torch.tensor(real, device="cuda"), torch.tensor(alpha, device="cuda"), 20.0)
    out["d_loss"].backward()
    print(prec, "d_loss %.6f (want %.6f) gp %.6f (want %.6f)" % (float(out["d_loss"]), want["d_loss"], float(out["gradient_penalty"]), want["gradient_penalty"]))
    for k, g in want["grads"].items():
        name = (("batch_normalization" if k[2] == "1" else "batch_normalization_%d" % (int(k[2]) - 1)) + k[3:]) if k.startswith("bn") else k
        got = D.params[name].grad.cpu().numpy().astype(np.float64)
        if np.abs(g).max() < 1e-9:
            continue
        print("   %-28s max-rel %.2e   l2-rel %.2e" % (k, np.abs(got - g).max() / np.abs(g).max(), np.linalg.norm(got - g) / np.linalg.norm(g)))
