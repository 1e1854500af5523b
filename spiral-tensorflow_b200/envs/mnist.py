"""Batched MyPaint painting environments: ``MNIST`` and ``SimpleMNIST``.

Same semantics per canvas as the reference's envs/mnist.py (reset :81-105, draw :107-148,
_draw :150-167, curve :173-214, step :231-245, state :251-260), but B canvases live on one GPU
and one ``step`` call advances all of them through libspiral_b200.so (no libmypaint, no SWIG):

    env = SimpleMNIST(args, num_envs=1024)
    state, target, z = env.reset()          # state  f32[B,64,64,C] on cuda, values 0..255
    state, reward, terminal, info = env.step(actions)   # actions int[B,K] in env.acs order

Differences from the reference that are part of the contract:
  * arrays are torch CUDA tensors with a leading batch axis; ``terminal`` is one bool
    (all canvases share the episode length, mnist.py:234); ``reward`` is f32[B]
  * the returned ``state`` tensor is the env's observation buffer (overwritten by the next
    step) -- copy it if you keep it (rl_utils does)
  * targets are synthetic (BASELINE configs; no dataset download): digit-shaped 28x28 -> 64x64 for gray canvases,
    CelebA-shaped smooth colour fields for RGB ones, seed 7 (``make_synthetic_targets``)
  * EXTENSIONS beyond the reference, flagged in SURVEY.md: ``color_channel=3`` returns
    ``image[..., :3]`` instead of failing; a ``color`` action head; ``location_size=64``
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _native as N
from .. import brush as brushlib
from . import targets, utils
from .base import Environment


class MNIST(Environment):
    head = 0.25
    tail = 0.75

    # source order of the reference's dict literal (envs/mnist.py:26-32)
    action_sizes = [
        ('pressure', [2]),
        ('jump', [2]),
        ('size', [2]),
        ('control', None),
        ('end', None),
    ]

    size = 0.2
    pressure = 0.3

    def __init__(self, args, num_envs=None, device=None, max_dabs=None):
        super(MNIST, self).__init__(args, num_envs=num_envs)
        if self.screen_size != N.TILE:
            raise NotImplementedError("screen_size must be 64 (one MyPaint tile), got %d" % self.screen_size)
        self.mnist_nums = getattr(args, "mnist_nums", list(range(10)))
        self.colorize = not getattr(args, "train", True)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.channels = int(args.color_channel)
        self.z_dim = getattr(args, "z_dim", 10)

        # jump
        self.jumps = [0, 1]
        # size (mnist.py:48-52)
        self.sizes = np.arange(0.2, 2.0, 0.5)
        self.sizes = self.sizes * 1
        if 'size' in self.action_sizes:
            self.sizes = self.sizes[:self.action_sizes['size'][0]]
        # pressure (mnist.py:55-58)
        self.pressures = np.arange(0.8, 0, -0.3)
        if 'pressure' in self.action_sizes:
            self.pressures = self.pressures[:self.action_sizes['pressure'][0]]
        # palette (mnist.py:60-73)
        self.colors = np.array([
            (0., 0., 0.),        # black
            (102., 217., 232.),  # cyan 3
            (173., 181., 189.),  # gray 5
            (255., 224., 102.),  # yellow 3
            (229., 153., 247.),  # grape 3
            (99., 230., 190.),   # teal 3
            (255., 192., 120.),  # orange 3
            (255., 168., 168.),  # red 3
        ]) / 255.
        if 'color' in self.action_sizes:
            self.colors = self.colors[:self.action_sizes['color'][0]]
        self.controls = utils.uniform_locations(self.screen_size, self.location_size, 0)
        self.ends = utils.uniform_locations(self.screen_size, self.location_size, 0)

        # brush (mnist.py:93-95 reads args.brush_path at every reset; the file is parsed once here)
        brush_path = getattr(args, "brush_path", None)
        if brush_path and str(brush_path) not in ("assets/brushes/dry_brush.myb",):
            with open(str(brush_path)) as fp:
                self.brushinfo = brushlib.BrushInfo(fp.read())
        else:
            self.brushinfo = brushlib.BrushInfo()
        self.max_dabs = int(max_dabs or getattr(args, "max_dabs", 2560))
        # arithmetic mode of the brush state machine ('det' | 'fast'; SPIRAL_RASTER_MATH overrides the flag: profiling)
        import os
        self.raster_math = os.environ.get("SPIRAL_RASTER_MATH") or getattr(args, "raster_math", "fast")
        if self.raster_math not in ("det", "fast"):
            raise ValueError("raster_math must be 'det' or 'fast', got %r" % (self.raster_math,))
        self.math_mode = N.MATH_FAST if self.raster_math == "fast" else N.MATH_DET

        self._lib = N.lib()
        self._handle = C.c_void_p()
        cfg = self._make_cfg()
        N.check(self._lib.spiral_env_create(C.byref(cfg), C.byref(self._handle)), "spiral_env_create")

        B, dev = self.num_envs, self.device
        self.canvas = torch.empty((B, N.TILE, N.TILE, 4), dtype=torch.int16, device=dev)   # fix15 RGBA (u16 bits)
        self.brush_state = torch.empty((B, N.BRUSH_STATE_FLOATS), dtype=torch.float32, device=dev)
        self.env_state = torch.empty((B, N.ENV_STATE_DOUBLES), dtype=torch.float64, device=dev)
        self.dab_scratch = torch.empty((B, self.max_dabs, N.DAB_FLOATS), dtype=torch.float32, device=dev)
        self.dab_count = torch.zeros((B, 2), dtype=torch.int32, device=dev)
        self.obs = torch.empty((B, N.TILE, N.TILE, self.channels), dtype=torch.float32, device=dev)
        self.status = torch.zeros((1,), dtype=torch.int32, device=dev)
        self._actions = torch.empty((B, len(self.acs)), dtype=torch.int32, device=dev)
        self._reward = torch.zeros((B,), dtype=torch.float32, device=dev)
        self._zero_reward = torch.zeros((B,), dtype=torch.float32, device=dev)
        self._step = 0
        self.random_target = None
        self.z = None
        self.real_data = None
        self._rng = torch.Generator(device=dev)
        self._rng.manual_seed(int(getattr(args, "seed", 123)))

    # ------------------------------------------------------------------
    def _make_cfg(self) -> N.SpiralEnvCfg:
        cfg = N.SpiralEnvCfg()
        cfg.abi_version = N.ABI_VERSION
        cfg.device = self.device.index if self.device.index is not None else torch.cuda.current_device()
        cfg.batch = self.num_envs
        cfg.channels = self.channels
        cfg.location_size = self.location_size
        cfg.n_heads = len(self.acs)
        for name in ("jump", "control", "end", "pressure", "size", "color"):
            setattr(cfg, "idx_" + name, self.ac_idx.get(name, -1))
        cfg.n_pressures, cfg.n_sizes, cfg.n_colors = len(self.pressures), len(self.sizes), len(self.colors)
        cfg.colorize = int(self.colorize)
        cfg.dither = 1 if getattr(self.args, "dither", True) else 0
        cfg.max_dabs = self.max_dabs
        # every Brush walks the same seed-1000 stream; one episode draws at most ~64k values per step
        cfg.rng_table_len = int(min(1 << 24, max(1 << 16, (self.episode_length + 1) * 65536)))
        lin = np.linspace(0, self.screen_size - 0, self.location_size)
        for i, v in enumerate(lin):
            cfg.lin[i] = v
        for i, v in enumerate(self.pressures):
            cfg.pressures[i] = v
        for i, v in enumerate(self.sizes):
            cfg.sizes[i] = v
        if 'color' in self.action_sizes or self.colorize:
            for i, rgb in enumerate(self.colors):
                for c, v in enumerate(brushlib.color_fix15(rgb)):
                    cfg.colors_fix15[i][c] = v
        else:   # colour never set by the env: the brush file's color_h/s/v (black for dry_brush)
            for c, v in enumerate(brushlib.brush_color_fix15(self.brushinfo)):
                cfg.colors_fix15[0][c] = v
        cfg.default_pressure, cfg.default_size = self.pressure, self.size
        cfg.head, cfg.tail = self.head, self.tail
        cfg.entry_pressure0 = float(np.min(self.pressures))
        cfg.brush = self.brushinfo.compile()
        table = brushlib.dither_table()
        C.memmove(cfg.dither_table, table.ctypes.data, table.nbytes)
        cfg.math_mode = self.math_mode
        return cfg

    def __del__(self):
        try:
            if getattr(self, "_handle", None) is not None and self._handle.value:
                self._lib.spiral_env_destroy(self._handle)
                self._handle = C.c_void_p()
        except Exception:
            pass

    # ------------------------------------------------------------------
    def reset(self, target=None, z=None):
        """mnist.py:81-105 for every canvas: white surface, fresh Brush (RNG re-seeded), step 0.
        ``target`` / ``z`` may be handed in (static buffers of a captured CUDA graph); otherwise they are
        drawn here as the reference does (:84-87, :100-103)."""
        if self.conditional:
            self.random_target = self.get_random_target(num=self.num_envs) if target is None else target
        else:
            self.random_target = None
        N.check(self._lib.spiral_env_reset(self._handle, N.ptr(self.canvas), N.ptr(self.brush_state),
                                           N.ptr(self.env_state), N.ptr(self.obs), N.stream_ptr()),
                "spiral_env_reset")
        self._step = 0
        if self.args.conditional:
            self.z = None
        else:
            self.z = (torch.rand((self.num_envs, self.z_dim), generator=self._rng, device=self.device) - 0.5) if z is None else z
        return self.state, self.random_target, self.z

    def _to_actions(self, acs):
        if isinstance(acs, torch.Tensor):
            a = acs.to(device=self.device, dtype=torch.int32, non_blocking=True)
        else:
            a = torch.as_tensor(np.asarray(acs), dtype=torch.int32).to(self.device, non_blocking=True)
        if a.shape != self._actions.shape:
            raise ValueError("actions must have shape %s (env.acs = %s), got %s"
                             % (tuple(self._actions.shape), self.acs, tuple(a.shape)))
        return a.contiguous()

    def draw(self, acs):
        """mnist.py:107-167 for every canvas (stroke expansion + compositing + read-back)."""
        a = self._to_actions(acs)
        N.check(self._lib.spiral_env_step(self._handle, N.ptr(a), N.ptr(self.canvas), N.ptr(self.brush_state),
                                          N.ptr(self.env_state), N.ptr(self.dab_scratch), N.ptr(self.dab_count),
                                          N.ptr(self.obs), N.ptr(self.status), N.stream_ptr()),
                "spiral_env_step")

    def step(self, acs):
        """mnist.py:231-245.  reward: f32[B] (zeros unless terminal and conditional)."""
        self.draw(acs)
        self._step += 1
        terminal = (self._step == self.episode_length)
        if terminal and self.conditional:
            n = int(np.prod(self.observation_shape))
            N.check(self._lib.spiral_l2_reward(N.ptr(self.obs), N.ptr(self.random_target), N.ptr(self._reward),
                                               self.num_envs, n, N.stream_ptr()), "spiral_l2_reward")
            reward = self._reward
        else:
            reward = self._zero_reward
        return self.state, reward, terminal, {}

    def check_status(self):
        """Synchronising check of the device-side overflow flag (pending-dab list too small)."""
        st = int(self.status.item())
        if st & 1:
            raise RuntimeError("a canvas produced more than max_dabs=%d dabs in one step; "
                               "re-create the env with a larger max_dabs" % self.max_dabs)
        if st & 2:
            raise RuntimeError("a canvas consumed more random numbers than the tabulated stream holds "
                               "(episode longer than episode_length=%d?)" % self.episode_length)

    @property
    def image(self):
        """uint8 RGBA strip like surface.scanline_strips_iter (mnist.py:251-256); debugging aid."""
        if self.channels == 3:
            rgb = self.obs.to(torch.uint8)
        else:
            raise NotImplementedError("image is only materialised for color_channel=3; use .state")
        alpha = torch.full_like(rgb[..., :1], 255)
        return torch.cat([rgb, alpha], -1)

    @property
    def state(self):
        return self.obs

    @property
    def canvas_fix15(self):
        """u16 view of the fix15 RGBA tiles [B,64,64,4] (as int32 so values up to 32768 print right)."""
        return self.canvas.to(torch.int32) & 0xFFFF

    # ------------------------------------------------------------------
    def get_random_target(self, num=1, squeeze=False):
        """mnist.py:224-229: ``num`` distinct random rows of ``real_data`` (float32, 0..255)."""
        if self.real_data is None:
            self.real_data = self.make_synthetic_targets()
        n = self.real_data.shape[0]
        if num <= n:
            idx = torch.randperm(n, generator=self._rng, device=self.device)[:num]
        else:
            idx = torch.randint(0, n, (num,), generator=self._rng, device=self.device)
        out = self.real_data[idx].to(torch.float32)
        if squeeze:
            out = out.squeeze(0)
        return out

    def make_synthetic_targets(self, count=2048, seed=7):
        """Synthetic stand-in for ``prepare_mnist`` (mnist.py:280-317): see ``envs/targets.py``."""
        return targets.synthetic_targets(self.channels, count=count, seed=seed, size=N.TILE).to(self.device)

    def make_stroke_targets(self, count=2048, strokes=4, seed=7):
        """Alternative targets that are reachable by construction: ``count`` stroke drawings rendered with this very
        renderer -- dark strokes on white, uint8 [count,64,64,C]."""
        B = self.num_envs
        g = np.random.RandomState(seed)
        out = []
        saved = (self.canvas.clone(), self.brush_state.clone(), self.env_state.clone(), self.obs.clone(), self._step)
        done = 0
        while done < count:
            N.check(self._lib.spiral_env_reset(self._handle, N.ptr(self.canvas), N.ptr(self.brush_state),
                                               N.ptr(self.env_state), N.ptr(self.obs), N.stream_ptr()))
            for t in range(strokes + 1):
                a = np.stack([g.randint(n, size=B) for n in self.head_sizes], 1).astype(np.int32)
                if 'jump' in self.ac_idx:
                    a[:, self.ac_idx['jump']] = (g.rand(B) < 0.2)
                self.draw(a)
            out.append(self.obs.to(torch.uint8).clone())
            done += B
        self.canvas.copy_(saved[0]); self.brush_state.copy_(saved[1]); self.env_state.copy_(saved[2])
        self.obs.copy_(saved[3]); self._step = saved[4]
        return torch.cat(out, 0)[:count]

    def get_action_desc(self, ac):
        desc = []
        for name in self.action_sizes:
            named_ac = int(ac[self.ac_idx[name]])
            actual_ac = getattr(self, name + "s")[named_ac]
            desc.append("{}: {} ({})".format(name, actual_ac, named_ac))
        return "\n".join(desc)


class SimpleMNIST(MNIST):

    # source order of the reference's dict literal (envs/mnist.py:322-329)
    action_sizes = [
        ('jump', [2]),
        ('control', None),
        ('end', None),
    ]

    def __init__(self, args, num_envs=None, device=None, max_dabs=None):
        super(SimpleMNIST, self).__init__(args, num_envs=num_envs, device=device, max_dabs=max_dabs)


class ColorMNIST(MNIST):
    """EXTENSION (BASELINE.json rgb64 config): MNIST heads plus an 8-colour palette head."""

    action_sizes = [
        ('pressure', [2]),
        ('jump', [2]),
        ('size', [2]),
        ('color', [8]),
        ('control', None),
        ('end', None),
    ]
