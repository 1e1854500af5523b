"""Synthetic targets of the shapes SURVEY 8d / BASELINE.json name (there is no dataset access here).

Stand-in for the reference's ``prepare_mnist`` (envs/mnist.py:280-317): MNIST digits as 28x28 floats in [0, 1], resized
to the canvas with ``imresize(image, [64, 64], interp='cubic')`` (:303-305, a uint8 image) and inverted,
``real_data = 255 - image`` (:317) -- dark strokes on a white canvas, ``[N, 64, 64, 1]``.
"""
import numpy as np
import torch


def synthetic_targets(channels, count=2048, seed=7, size=64):
    """uint8 CPU tensor ``[count, size, size, channels]``, deterministic in ``seed``.

    * ``channels == 1``: "digit-shaped" -- 28x28 images of a few pen strokes (and, half of the time, a loop) of MNIST's
      stroke width around the centre of the frame, bicubically resized to ``size`` and inverted.
    * ``channels == 3``: "CelebA-shaped" -- smooth random colour fields (three octaves of bicubically upsampled noise with
      correlated channels), values 0..255."""
    g = np.random.RandomState(seed)
    if channels == 1:
        yy, xx = np.mgrid[0:28, 0:28].astype(np.float32)
        imgs = np.zeros((count, 28, 28), np.float32)
        for i in range(count):
            pts = 14 + g.uniform(-8, 8, (g.randint(3, 6), 2))              # a 2..4-segment polyline through the centre region
            acc = np.zeros((28, 28), np.float32)
            for p0, p1 in zip(pts[:-1], pts[1:]):
                d = p1 - p0
                t = np.clip(((xx - p0[0]) * d[0] + (yy - p0[1]) * d[1]) / max(float(d @ d), 1e-6), 0, 1)
                dist2 = (xx - (p0[0] + t * d[0])) ** 2 + (yy - (p0[1] + t * d[1])) ** 2
                acc = np.maximum(acc, np.exp(-dist2 / (2 * 1.1 ** 2)))          # pen of ~2.5 px like MNIST's strokes
            if g.rand() < 0.5:                                                  # a loop, as in 0 6 8 9
                c, r = 14 + g.uniform(-4, 4, 2), g.uniform(3, 6)
                ring = np.abs(np.sqrt((xx - c[0]) ** 2 + (yy - c[1]) ** 2) - r)
                acc = np.maximum(acc, np.exp(-ring ** 2 / (2 * 1.1 ** 2)))
            imgs[i] = np.clip(acc * 1.2, 0, 1)
        t = torch.from_numpy(imgs)[:, None]
        big = torch.nn.functional.interpolate(t, size=(size, size), mode="bicubic", align_corners=False)
        big = (big.clamp(0, 1) * 255).to(torch.uint8)[:, 0, :, :, None]         # imresize returns uint8
        return (255 - big).contiguous()
    if channels != 3:
        raise ValueError("channels must be 1 or 3")
    fields = torch.zeros((count, 3, size, size))
    for res, amp in ((4, 1.0), (8, 0.5), (16, 0.25)):
        base = torch.from_numpy(g.randn(count, 1, res, res).astype(np.float32))
        tint = torch.from_numpy(g.randn(count, 3, res, res).astype(np.float32))
        fields += amp * torch.nn.functional.interpolate(base + 0.5 * tint, size=(size, size), mode="bicubic", align_corners=False)
    fields = fields - fields.amin(dim=(1, 2, 3), keepdim=True)
    fields = fields / fields.amax(dim=(1, 2, 3), keepdim=True).clamp_min(1e-6)
    return (fields * 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
