"""GPU parity of K2b: the convolution layer's forward / dgrad / wgrad kernels and the WGAN-GP
discriminator step built from them (models/discriminator.py:97-136), against float64 PyTorch-CPU
references (oracle/disc_oracle.py).  Tolerance: 1e-4 relative (BASELINE.json north_star) for the
fp32 and bf16x3 paths."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _sp():
    import spiral_b200 as sp
    return sp


def _ops():
    from importlib import import_module
    return import_module(_sp().__name__ + ".models.conv_ops")


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def _ref_conv(x, w):
    """float64 CPU: NHWC x, HWIO w, TF 'SAME' (1 before / 2 after), stride 2."""
    y = F.conv2d(F.pad(x.permute(0, 3, 1, 2), (1, 2, 1, 2)), w.permute(3, 2, 0, 1), stride=2)
    return y.permute(0, 2, 3, 1)


SHAPES = [  # B, H, Cin, Cout, precision
    (3, 8, 3, 8, 0), (2, 16, 16, 24, 0), (5, 2, 40, 130, 0), (2, 64, 1, 8, 0), (2, 4, 64, 128, 0),
    (4, 32, 64, 128, 1), (9, 16, 128, 256, 1), (40, 4, 512, 1024, 1), (70, 2, 1024, 2048, 1),
    (3, 64, 3, 64, 1), (3, 64, 6, 64, 1), (4, 8, 256, 512, 2),
    (4, 32, 64, 128, 3), (9, 16, 128, 256, 3), (40, 4, 512, 1024, 3), (70, 2, 1024, 2048, 3), (3, 64, 3, 64, 3),
]


@pytest.mark.parametrize("B,H,Cin,Cout,prec", SHAPES)
def test_conv_forward_dgrad_wgrad_match_float64(B, H, Cin, Cout, prec):
    ops = _ops()
    g = torch.Generator().manual_seed(B * 1000 + H)
    x = torch.randn((B, H, H, Cin), generator=g, dtype=torch.float64)
    w = torch.randn((5, 5, Cin, Cout), generator=g, dtype=torch.float64) / np.sqrt(25 * Cin)
    dy = torch.randn((B, H // 2, H // 2, Cout), generator=g, dtype=torch.float64)
    xr, wr = x.clone().requires_grad_(), w.clone().requires_grad_()
    y_ref = _ref_conv(xr, wr)
    dx_ref, dw_ref = torch.autograd.grad((y_ref * dy).sum(), (xr, wr))
    xc, wc, dyc = x.float().cuda(), w.float().cuda(), dy.float().cuda()
    # bf16x6: fp32-exact products, four TMEM accumulators against the tensor core's truncating accumulation
    # (measured <= 5e-6, better than the fp32 SIMT kernels)
    tol = {0: 1e-4, 1: 1e-4, 2: 2e-2, 3: 1e-5}[prec]
    assert _rel(ops.conv_forward(xc, wc, prec).cpu(), y_ref.detach()) < tol
    assert _rel(ops.conv_dgrad(dyc, wc, prec).cpu(), dx_ref) < tol
    assert _rel(ops.conv_wgrad(xc, dyc, prec).cpu(), dw_ref) < tol


def test_conv_autograd_first_and_second_order():
    """The mutually recursive autograd functions give the right first and second derivatives."""
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    B, H, Cin, Cout = 3, 8, 4, 6
    x = torch.randn((B, H, H, Cin), generator=g, dtype=torch.float64)
    w = torch.randn((5, 5, Cin, Cout), generator=g, dtype=torch.float64) / 10
    v = torch.randn((B, H // 2, H // 2, Cout), generator=g, dtype=torch.float64)

    def scalar(conv, x, w, v):
        y = conv(x, w)
        gx, = torch.autograd.grad((torch.tanh(y) * v).sum(), x, create_graph=True)
        return (gx ** 2).sum() + (y ** 3).sum()

    xr, wr = x.clone().requires_grad_(), w.clone().requires_grad_()
    ref = scalar(_ref_conv, xr, wr, v)
    gxr, gwr = torch.autograd.grad(ref, (xr, wr))
    xc, wc = x.float().cuda().requires_grad_(), w.float().cuda().requires_grad_()
    got = scalar(lambda a, b: ops.conv5x5s2(a, b, 0), xc, wc, v.float().cuda())
    gx, gw = torch.autograd.grad(got, (xc, wc))
    assert abs(float(got) - float(ref)) < 1e-4 * abs(float(ref))
    assert _rel(gx.cpu(), gxr) < 1e-4
    assert _rel(gw.cpu(), gwr) < 1e-4


@pytest.mark.parametrize("C,dd,bn,B,prec,cond", [(1, 8, True, 6, "fp32", False), (3, 8, False, 5, "fp32", False),
                                                 (1, 8, True, 4, "fp32", True), (3, 64, True, 8, "bf16x3", False),
                                                 (3, 64, True, 8, "bf16x6", False)])
def test_wgan_gp_step_matches_oracle(C, dd, bn, B, prec, cond):
    """d_loss, critic, penalty and every parameter gradient of the WGAN-GP step vs the float64 oracle."""
    sp = _sp()
    from importlib import import_module

    from oracle import disc_oracle as do
    models = import_module(sp.__name__ + ".models")
    argv = ["--color_channel", str(C), "--disc_dim", str(dd), "--disc_batch_norm", str(bn), "--disc_precision",
            "bf16x3" if prec == "bf16x6" else prec, "--disc_grad_precision", prec, "--loss", "l2" if cond else "gan"]
    args = sp.get_args(argv)
    w = do.init_weights(C * (2 if cond else 1), disc_dim=dd, batch_norm=bn, seed=5)
    D = models.Discriminator(args, None, [64, 64, C], norm_fn=lambda x: x, scope_name="global", max_batch=B)
    assert D.conditional == cond
    D.load_numpy(w)
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
    want = do.d_loss(fake, real, alpha, w, batch_norm=bn, wgan_lambda=20.0, conditional=cond)
    for p in D.params.values():
        p.requires_grad_(True)
    out = D.wgan_gp_loss(torch.tensor(fake, device="cuda"), torch.tensor(real, device="cuda"),
                         torch.tensor(alpha, device="cuda"), 20.0)
    out["d_loss"].backward()
    for k in ("d_loss", "critic_loss", "gradient_penalty", "g_loss"):
        assert abs(float(out[k]) - want[k]) <= 1e-4 * max(1.0, abs(want[k])), (k, float(out[k]), want[k])
    for k, gref in want["grads"].items():
        name = (("batch_normalization" if k[2] == "1" else "batch_normalization_%d" % (int(k[2]) - 1)) + k[3:]) if k.startswith("bn") else k
        got = D.params[name].grad.cpu().numpy()
        if np.abs(gref).max() < 1e-9:        # a conv bias in front of a batch norm: exactly zero gradient
            assert np.abs(got).max() < 1e-5, k
        else:
            # The losses hold 1e-4 in both modes.  The gradient of the penalty through training-mode BN
            # at batch 4-8 is ill-conditioned (it amplifies operand round-off ~100x: plain bf16 is 20 % off):
            # fp32 kernels and the bf16x6 tensor-core split (24 operand bits) stay within 1e-3 of the float64
            # oracle, the bf16x3 split (16 bits) within 2e-2 in the l2 sense (measured 1e-3..1e-2,
            # tools/grad_err_probe.py)
            if prec == "fp32":
                assert _rel(got, gref) < 1e-3, (k, _rel(got, gref))
            elif prec == "bf16x6":
                assert _rel(got, gref) < 2e-4, (k, _rel(got, gref))       # measured 1e-5 .. 2e-5
            else:
                l2 = np.linalg.norm(got.astype(np.float64) - gref) / np.linalg.norm(gref)
                assert l2 < 2e-2, (k, l2)
