"""How far are the rasteriser's arithmetic modes from the libm-faithful oracle?  (CPU only; test infrastructure.)

Runs N random episodes through oracle/paint_oracle.c in PO_MATH_LIBM (glibc functions exactly where libmypaint calls
them -- the upstream-faithful restatement) and in PO_MATH_DET / PO_MATH_FAST (the two modes the CUDA rasteriser
reproduces bit for bit) and reports, per canvas, the north star's two figures on the final observation scaled to
[0, 1]: max abs error per channel (bar: <= 1/255) and mean abs error (bar: <= 1e-3).

    python tools/raster_tolerance.py --episodes 10000 --env color_mnist --T 20 --workers 8 > profiles/r2_raster_tolerance.json
"""
import argparse
import json
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ENVS = {"simple_mnist": (["jump", "control", "end"], 32, 1),
        "mnist": (["jump", "pressure", "control", "end", "size"], 32, 1),
        "color_mnist": (["control", "end", "color", "jump", "pressure", "size"], 64, 3)}


def _heads(names, L, n_color):
    size = {"jump": 2, "pressure": 2, "size": 2, "color": n_color, "control": L * L, "end": L * L}
    return [size[n] for n in names]


def _work(job):
    from oracle import paint_oracle as po
    env, T, seed, n, modes = job
    names, L, C = ENVS[env]
    rng = np.random.RandomState(seed)
    acts = np.stack([rng.randint(0, h, (n, T)) for h in _heads(names, L, 8)], -1).astype(np.int32)
    out = {}
    for m in modes:
        spec = po.EnvSpec(action_names=names, location_size=L, n_color=8 if "color" in names else 0, math_mode=m)
        e = po.OracleBatchEnv(spec, n, channels=C)
        canv, obs = e.run_episodes(acts, want_obs=True)
        out[m] = (canv, obs[:, -1])
    ref_c, ref_o = out[po.PO_MATH_LIBM]
    res = {}
    for m in modes:
        if m == po.PO_MATH_LIBM:
            continue
        c, o = out[m]
        err = np.abs(o.astype(np.float64) - ref_o.astype(np.float64)).reshape(n, -1) / 255.0
        res[m] = dict(identical=(c == ref_c).reshape(n, -1).all(1), max_abs=err.max(1), mean_abs=err.mean(1),
                      n_bad_px=(err > 1.0 / 255 + 1e-12).sum(1))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--episodes", type=int, default=2000)
    ap.add_argument("--env", default="color_mnist", choices=sorted(ENVS))
    ap.add_argument("--T", type=int, default=20)
    ap.add_argument("--workers", type=int, default=os.cpu_count())
    ap.add_argument("--chunk", type=int, default=50)
    a = ap.parse_args()
    from oracle import paint_oracle as po
    modes = (po.PO_MATH_LIBM, po.PO_MATH_DET, po.PO_MATH_FAST)
    jobs = [(a.env, a.T, 1000 + i, min(a.chunk, a.episodes - i * a.chunk), modes)
            for i in range((a.episodes + a.chunk - 1) // a.chunk)]
    with mp.Pool(a.workers) as pool:
        parts = pool.map(_work, jobs)
    report = dict(env=a.env, T=a.T, episodes=a.episodes, reference="oracle PO_MATH_LIBM (glibc expf/logf/hypotf)",
                  bar=dict(max_abs=1.0 / 255, mean_abs=1e-3))
    for m, name in ((po.PO_MATH_DET, "det"), (po.PO_MATH_FAST, "fast")):
        cat = {k: np.concatenate([p[m][k] for p in parts]) for k in parts[0][m]}
        ok_max = cat["max_abs"] <= 1.0 / 255 + 1e-12
        report[name] = dict(
            bit_identical_frac=float(cat["identical"].mean()),
            within_max_abs_frac=float(ok_max.mean()),
            within_mean_abs_frac=float((cat["mean_abs"] <= 1e-3).mean()),
            mean_abs_mean=float(cat["mean_abs"].mean()), mean_abs_worst=float(cat["mean_abs"].max()),
            max_abs_median=float(np.median(cat["max_abs"])), max_abs_worst=float(cat["max_abs"].max()),
            bad_pixels_per_failing_canvas_median=float(np.median(cat["n_bad_px"][~ok_max])) if (~ok_max).any() else 0.0,
            bad_pixels_per_failing_canvas_max=int(cat["n_bad_px"].max()))
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
