"""Generates tests/golden/*.npz from the CPU oracle (oracle/paint_oracle.c; det-math mode, and the fast mode for *_fast).
The reference itself cannot run here (no libmypaint / Python 2 / TensorFlow), so these are
ORACLE outputs, committed so CUDA regressions are caught without re-running the oracle and so the
GPU box (which has no /root/reference) can check them.   python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import paint_oracle as po  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SIZES = {"jump": 2, "pressure": 2, "size": 2, "color": 8}

CASES = {
    # name: (acs order, L, channels, B, T)
    "simple_mnist_L32_c1": (["jump", "control", "end"], 32, 1, 8, 5),
    "mnist_L32_c1": (["jump", "pressure", "control", "end", "size"], 32, 1, 8, 5),
    "color_mnist_L64_c3": (["control", "end", "color", "jump", "pressure", "size"], 64, 3, 8, 5),
    "straight_only_L32_c1": (["end"], 32, 1, 4, 4),
    # the rasteriser's FAST arithmetic mode (--raster_math fast, PO_MATH_FAST)
    "color_mnist_L64_c3_fast": (["control", "end", "color", "jump", "pressure", "size"], 64, 3, 8, 5),
    "simple_mnist_L32_c1_fast": (["jump", "control", "end"], 32, 1, 8, 5),
}


def make(name):
    acs, L, C, B, T = CASES[name]
    rng = np.random.RandomState(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    acts = np.stack([rng.randint(SIZES.get(a, L * L), size=(B, T)) for a in acs], -1).astype(np.int32)
    if "jump" in acs:
        acts[:, 1, acs.index("jump")] = 0        # make sure step 1 paints
    fast = name.endswith("_fast")
    spec = po.EnvSpec(action_names=acs, location_size=L, n_color=8, math_mode=po.PO_MATH_FAST if fast else po.PO_MATH_DET)
    env = po.OracleBatchEnv(spec, B, channels=C)
    canv, obs = env.run_episodes(acts, want_obs=True)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), acs=np.array(acs), L=L, C=C, math=np.array("fast" if fast else "det"),
                        actions=acts, canvas=canv, obs_final=obs[:, -1].astype(np.float32),
                        obs_mean=obs.astype(np.float64).reshape(B, T + 1, -1).mean(-1))
    print(name, canv.shape, float(obs[:, -1].mean()))


if __name__ == "__main__":
    for n in (sys.argv[1:] or CASES):
        make(n)
