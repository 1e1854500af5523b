"""oracle/paint_oracle.py -- TEST INFRASTRUCTURE (CPU oracle), NOT product code.

ctypes binding of oracle/libpaint_oracle.so plus two restatements of the
reference's MNIST painting env (envs/mnist.py):

* ``LiteralMNISTEnv``  -- follows envs/mnist.py line by line in Python (float64
  arithmetic, one ``stroke_to`` ctypes call per event exactly like the reference
  crosses SWIG per event, envs/mnist.py:216-222).  Slow; used to validate ...
* ``OracleBatchEnv``   -- the C env (po_env_*), batched by a Python loop; used
  by the parity tests and as bench.py's cpu_baseline.

PARITY UNPINNED: see paint_oracle.c.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import colorsys
import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpaint_oracle.so")

PO_MATH_LIBM = 0
PO_MATH_DET = 1
PO_MATH_FAST = 2      # the CUDA kernel's binary32 mode (oracle/fast_math.h + stroke_to_fast)
PO_SETTING_COUNT = 45
PO_INPUT_COUNT = 9
PO_MAX_POINTS = 8
PO_MAX_TABLE = 8
PO_MAX_L = 64


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (gcc only)."""
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("paint_oracle.c", "paint_oracle.h", "det_math.h")
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


class PODab(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("radius", C.c_float), ("hardness", C.c_float),
                ("opacity", C.c_uint16), ("r", C.c_uint16), ("g", C.c_uint16), ("b", C.c_uint16)]


class POEvent(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("pressure", C.c_float), ("dtime", C.c_double)]


class POEnvCfg(C.Structure):
    _fields_ = [
        ("math_mode", C.c_int),
        ("base", C.c_float * PO_SETTING_COUNT),
        ("map_n", (C.c_int * PO_INPUT_COUNT) * PO_SETTING_COUNT),
        ("map_x", ((C.c_float * PO_MAX_POINTS) * PO_INPUT_COUNT) * PO_SETTING_COUNT),
        ("map_y", ((C.c_float * PO_MAX_POINTS) * PO_INPUT_COUNT) * PO_SETTING_COUNT),
        ("L", C.c_int),
        ("lin", C.c_double * PO_MAX_L),
        ("has_jump", C.c_int), ("has_control", C.c_int), ("has_size", C.c_int),
        ("idx_jump", C.c_int), ("idx_control", C.c_int), ("idx_end", C.c_int),
        ("idx_pressure", C.c_int), ("idx_size", C.c_int), ("idx_color", C.c_int),
        ("pressures", C.c_double * PO_MAX_TABLE),
        ("sizes", C.c_double * PO_MAX_TABLE),
        ("n_colors", C.c_int),
        ("colors_hsv", (C.c_double * 3) * PO_MAX_TABLE),
        ("colorize", C.c_int),
        ("default_pressure", C.c_double), ("default_size", C.c_double),
        ("head", C.c_double), ("tail", C.c_double),
        ("entry_pressure0", C.c_double),
    ]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    vp = C.c_void_p
    L.po_surface_new.restype = vp
    L.po_surface_free.argtypes = [vp]
    L.po_surface_clear.argtypes = [vp]
    L.po_surface_fill_white.argtypes = [vp]
    L.po_surface_data.restype = C.POINTER(C.c_uint16)
    L.po_surface_data.argtypes = [vp]
    L.po_surface_enable_dablog.argtypes = [vp, C.c_int]
    L.po_surface_dablog_count.argtypes = [vp]
    L.po_surface_dablog.restype = C.POINTER(PODab)
    L.po_surface_dablog.argtypes = [vp]
    L.po_surface_dablog_reset.argtypes = [vp]
    L.po_surface_counters.argtypes = [vp, C.POINTER(C.c_long), C.POINTER(C.c_long), C.POINTER(C.c_long)]
    L.po_surface_read_rgbu8.argtypes = [vp, vp, vp]
    L.po_glibc_rand_stream.argtypes = [C.c_uint, C.c_int, vp]
    L.po_dither_table.argtypes = [C.c_uint, C.c_int, vp]
    L.po_setting_name.restype = C.c_char_p
    L.po_setting_name.argtypes = [C.c_int]
    L.po_input_name.restype = C.c_char_p
    L.po_input_name.argtypes = [C.c_int]
    L.po_setting_default.restype = C.c_float
    L.po_setting_default.argtypes = [C.c_int]
    L.po_brush_new.restype = vp
    L.po_brush_new.argtypes = [C.c_int]
    L.po_brush_free.argtypes = [vp]
    L.po_brush_set_base_value.argtypes = [vp, C.c_int, C.c_float]
    L.po_brush_set_mapping_n.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.po_brush_set_mapping_point.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float]
    L.po_brush_get_state.restype = C.c_float
    L.po_brush_get_state.argtypes = [vp, C.c_int]
    L.po_brush_get_rng.argtypes = [vp, vp, vp, C.POINTER(C.c_int)]
    L.po_brush_rng_next.restype = C.c_double
    L.po_brush_rng_next.argtypes = [vp]
    L.po_brush_stroke_to.argtypes = [vp, vp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_double]
    L.po_hsv_to_rgb_float.argtypes = [C.POINTER(C.c_float)] * 3
    L.po_env_new.restype = vp
    L.po_env_new.argtypes = [C.POINTER(POEnvCfg)]
    L.po_env_free.argtypes = [vp]
    L.po_env_surface.restype = vp
    L.po_env_surface.argtypes = [vp]
    L.po_env_brush.restype = vp
    L.po_env_brush.argtypes = [vp]
    L.po_env_enable_eventlog.argtypes = [vp, C.c_int]
    L.po_env_eventlog_count.argtypes = [vp]
    L.po_env_eventlog.restype = C.POINTER(POEvent)
    L.po_env_eventlog.argtypes = [vp]
    L.po_env_eventlog_reset.argtypes = [vp]
    L.po_env_reset.argtypes = [vp]
    L.po_env_draw.argtypes = [vp] + [C.c_double] * 6 + [C.c_int, C.c_int] + [C.c_double] * 3
    L.po_env_step_actions.argtypes = [vp, vp]
    L.po_env_step_index.argtypes = [vp]
    L.po_env_entry_pressure.restype = C.c_double
    L.po_env_entry_pressure.argtypes = [vp]
    L.po_env_state.argtypes = [vp, vp, C.c_int, vp]
    L.po_env_run_episode.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int, vp, vp]
    L.po_det_exp.restype = C.c_double
    L.po_det_exp.argtypes = [C.c_double]
    L.po_det_log.restype = C.c_double
    L.po_det_log.argtypes = [C.c_double]
    L.po_det_hypotf.restype = C.c_float
    L.po_det_hypotf.argtypes = [C.c_float, C.c_float]
    for fn in ("po_fast_expf", "po_fast_exp_neg", "po_fast_logf"):
        getattr(L, fn).restype = C.c_float
        getattr(L, fn).argtypes = [C.c_float]
    L.po_fast_hypotf.restype = C.c_float
    L.po_fast_hypotf.argtypes = [C.c_float, C.c_float]
    L.po_brush_get_inv_radius.restype = C.c_float
    L.po_brush_get_inv_radius.argtypes = [vp]
    _lib = L
    return L


def setting_names():
    L = lib()
    return [L.po_setting_name(i).decode() for i in range(PO_SETTING_COUNT)]


def input_names():
    L = lib()
    return [L.po_input_name(i).decode() for i in range(PO_INPUT_COUNT)]


# --------------------------------------------------------------------------
# .myb (v2 text) parser -- mypaint 1.2.1 lib/brush.py BrushInfo.load_from_string
# (the reference reads the file at envs/mnist.py:93-95)
# --------------------------------------------------------------------------
def parse_myb(text: str):
    """-> {setting: (base, {input: [(x, y), ...]})} for every libmypaint setting."""
    L = lib()
    names = setting_names()
    out = {n: (float(L.po_setting_default(i)), {}) for i, n in enumerate(names)}
    version = 1
    for raw in text.splitlines():
        line = raw.strip()
        if not line or line.startswith("#"):
            continue
        cname, rawvalue = line.split(" ", 1)
        if cname == "version":
            version = int(rawvalue)
            if version != 2:
                raise NotImplementedError("only version-2 .myb files are restated (got %d)" % version)
            continue
        if cname == "stroke_treshold":          # historic spelling in old brush files
            cname = "stroke_threshold"
        if cname == "speed":
            cname = "speed1"
        if cname not in out:
            continue                             # upstream: "ignoring unknown setting"
        parts = rawvalue.split("|")
        base = float(parts[0])
        pts = {}
        for part in parts[1:]:
            inputname, rawpoints = part.strip().split(" ", 1)
            plist = []
            for s in rawpoints.split(","):
                s = s.strip()
                assert s.startswith("(") and s.endswith(")") and " " in s, s
                x, y = [float(ss) for ss in s[1:-1].split(" ")]
                plist.append((x, y))
            pts[inputname] = plist
        out[cname] = (base, pts)
    return out


# Brush constants of the reference's default brush (assets/brushes/dry_brush.myb:1-37,
# config.py:38), restated as data: (setting, base value, {input: [(x, y), ...]}).
DRY_BRUSH = [
    ("opaque", 0.8, {"pressure": [(0.0, 0.0), (1.0, 0.2)]}),
    ("opaque_multiply", 0.0, {"pressure": [(0.0, 0.0), (1.0, 1.0)]}),
    ("opaque_linearize", 0.0, {}),
    ("radius_logarithmic", 0.6, {"speed2": [(0.0, 0.042857), (4.0, -0.3)]}),
    ("hardness", 0.2, {}),
    ("dabs_per_basic_radius", 6.0, {}),
    ("dabs_per_actual_radius", 6.0, {}),
    ("dabs_per_second", 0.0, {}),
    ("radius_by_random", 0.1, {}),
    ("speed1_slowness", 0.04, {}),
    ("speed2_slowness", 0.8, {}),
    ("speed1_gamma", 4.0, {}),
    ("speed2_gamma", 4.0, {}),
    ("offset_by_random", 0.0, {"pressure": [(0.0, 0.0), (1.0, 1.4)]}),
    ("offset_by_speed", 0.0, {}),
    ("offset_by_speed_slowness", 1.0, {}),
    ("slow_tracking", 2.0, {}),
    ("slow_tracking_per_dab", 0.0, {}),
    ("tracking_noise", 0.0, {}),
    ("color_h", 0.0, {}), ("color_s", 0.0, {}), ("color_v", 0.0, {}),
    ("change_color_h", 0.0, {}), ("change_color_l", 0.0, {}), ("change_color_hsl_s", 0.0, {}),
    ("change_color_v", 0.0, {}), ("change_color_hsv_s", 0.0, {}),
    ("smudge", 0.0, {}), ("smudge_length", 0.5, {}),
    ("stroke_treshold", 0.0, {}),
    ("stroke_duration_logarithmic", 4.0, {}),
    ("stroke_holdtime", 0.0, {}),
    ("custom_input", 0.0, {}), ("custom_input_slowness", 0.0, {}),
]


def brush_to_myb(brush=DRY_BRUSH) -> str:
    """Serialise to the version-2 text format parse_myb() reads."""
    lines = ["# mypaint brush file", "version 2"]
    for name, base, pts in brush:
        line = "%s %s" % (name, repr(float(base)))
        for iname, plist in pts.items():
            line += " | %s %s" % (iname, ", ".join("(%f %f)" % (x, y) for x, y in plist))
        lines.append(line)
    return "\n".join(lines) + "\n"

# 8-colour palette of envs/mnist.py:60-71
PALETTE = np.array([
    (0., 0., 0.), (102., 217., 232.), (173., 181., 189.), (255., 224., 102.),
    (229., 153., 247.), (99., 230., 190.), (255., 192., 120.), (255., 168., 168.)]) / 255.


def rgb_to_hsv(rgb):
    """mypaint lib/helpers.py rgb_to_hsv (clamped colorsys), used by set_color_rgb."""
    r, g, b = [min(max(float(c), 0.0), 1.0) for c in rgb]
    return colorsys.rgb_to_hsv(r, g, b)


def py2_dict_order(names):
    """Iteration order of a CPython-2.7 (64-bit, no hash randomisation) dict literal
    with these string keys in this source order (envs/base.py:32 relies on it)."""
    def strhash(s):
        if not s:
            return 0
        x = (ord(s[0]) << 7) & 0xFFFFFFFFFFFFFFFF
        for ch in s:
            x = ((1000003 * x) ^ ord(ch)) & 0xFFFFFFFFFFFFFFFF
        x ^= len(s)
        if x >= 1 << 63:
            x -= 1 << 64
        if x == -1:
            x = -2
        return x

    def insert(table, mask, key):
        h = strhash(key)
        i = h & mask
        perturb = h & 0xFFFFFFFFFFFFFFFF          # size_t perturb = hash
        while table[i & mask] is not None:
            i = (i << 2) + i + perturb + 1
            perturb >>= 5
            i &= 0xFFFFFFFFFFFFFFFF
        table[i & mask] = key

    size = 8
    if len(names) > 5:                            # BUILD_MAP n -> _PyDict_NewPresized(n)
        while size <= len(names):
            size <<= 1
    table = [None] * size
    fill = 0
    for k in names:
        insert(table, size - 1, k)
        fill += 1
        if fill * 3 >= size * 2:                  # dictresize(used * 4)
            want = fill * 4
            new = 8
            while new <= want:
                new <<= 1
            old = [t for t in table if t is not None]
            table = [None] * new
            size = new
            for kk in old:
                insert(table, size - 1, kk)
    return [t for t in table if t is not None]


@dataclass
class EnvSpec:
    """Host-side description of one MNIST-family env (envs/mnist.py:22-79)."""
    action_names: list            # env.acs order
    location_size: int = 32
    screen_size: int = 64
    n_pressure: int = 2
    n_size: int = 2
    n_color: int = 0
    colorize: bool = False
    brush_text: str | None = None
    math_mode: int = PO_MATH_DET
    extras: dict = field(default_factory=dict)

    def make_cfg(self) -> POEnvCfg:
        cfg = POEnvCfg()
        cfg.math_mode = self.math_mode
        text = self.brush_text
        if text is None:
            text = brush_to_myb()
        brush = parse_myb(text)
        names = setting_names()
        inames = input_names()
        for i, n in enumerate(names):
            base, pts = brush[n]
            cfg.base[i] = base
            for iname, plist in pts.items():
                j = inames.index(iname)
                cfg.map_n[i][j] = len(plist)
                for k, (x, y) in enumerate(plist):
                    cfg.map_x[i][j][k] = x
                    cfg.map_y[i][j][k] = y
        Lsz = self.location_size
        cfg.L = Lsz
        lin = np.linspace(0, self.screen_size - 0, Lsz)      # envs/utils.py:12 with radius 0
        for i in range(Lsz):
            cfg.lin[i] = lin[i]
        acs = list(self.action_names)
        idx = {n: (acs.index(n) if n in acs else -1) for n in
               ("jump", "control", "end", "pressure", "size", "color")}
        cfg.idx_jump, cfg.idx_control, cfg.idx_end = idx["jump"], idx["control"], idx["end"]
        cfg.idx_pressure, cfg.idx_size, cfg.idx_color = idx["pressure"], idx["size"], idx["color"]
        cfg.has_jump = int("jump" in acs)
        cfg.has_control = int("control" in acs)
        cfg.has_size = int("size" in acs)
        sizes = np.arange(0.2, 2.0, 0.5) * 1                  # mnist.py:48-52
        if "size" in acs:
            sizes = sizes[:self.n_size]
        pressures = np.arange(0.8, 0, -0.3)                   # mnist.py:55-58
        if "pressure" in acs:
            pressures = pressures[:self.n_pressure]
        for i, v in enumerate(sizes):
            cfg.sizes[i] = v
        for i, v in enumerate(pressures):
            cfg.pressures[i] = v
        colors = PALETTE
        if "color" in acs:
            colors = colors[:self.n_color]
        cfg.n_colors = len(colors)
        for i, c in enumerate(colors):
            h, s, v = rgb_to_hsv(c)
            cfg.colors_hsv[i][0], cfg.colors_hsv[i][1], cfg.colors_hsv[i][2] = h, s, v
        cfg.colorize = int(self.colorize)
        cfg.default_pressure = 0.3
        cfg.default_size = 0.2
        cfg.head, cfg.tail = 0.25, 0.75
        cfg.entry_pressure0 = float(np.min(pressures))
        return cfg


_DITHER = {}


def dither_table(seed: int = 1, n: int = 4096) -> np.ndarray:
    key = (seed, n)
    if key not in _DITHER:
        out = np.zeros(n, np.uint16)
        lib().po_dither_table(seed, n, out.ctypes.data)
        _DITHER[key] = out
    return _DITHER[key]


class OracleBatchEnv:
    """B independent C envs (po_env_*).  Mirrors the batched product API:
    reset() -> obs [B,64,64,C] float32;  step(actions[B,K]) -> obs."""

    def __init__(self, spec: EnvSpec, batch: int, channels: int = 1, dither="glibc"):
        self.L = lib()
        self.spec = spec
        self.cfg = spec.make_cfg()
        self.B = batch
        self.channels = channels
        self.envs = [self.L.po_env_new(C.byref(self.cfg)) for _ in range(batch)]
        self.noise = dither_table() if dither == "glibc" else None
        self.K = len(spec.action_names)

    def __del__(self):
        try:
            for e in self.envs:
                self.L.po_env_free(e)
        except Exception:
            pass

    def _noise_ptr(self):
        return self.noise.ctypes.data if self.noise is not None else None

    def state(self) -> np.ndarray:
        n = 64 * 64 * self.channels
        out = np.zeros((self.B, n), np.float64)
        for i, e in enumerate(self.envs):
            self.L.po_env_state(e, self._noise_ptr(), self.channels, out[i].ctypes.data)
        return out.reshape(self.B, 64, 64, self.channels)

    def canvas(self) -> np.ndarray:
        out = np.zeros((self.B, 64, 64, 4), np.uint16)
        for i, e in enumerate(self.envs):
            p = self.L.po_surface_data(self.L.po_env_surface(e))
            out[i] = np.ctypeslib.as_array(p, shape=(64, 64, 4))
        return out

    def reset(self) -> np.ndarray:
        for e in self.envs:
            self.L.po_env_reset(e)
        return self.state()

    def step(self, actions) -> np.ndarray:
        a = np.ascontiguousarray(actions, dtype=np.int32).reshape(self.B, self.K)
        for i, e in enumerate(self.envs):
            self.L.po_env_step_actions(e, a[i].ctypes.data)
        return self.state()

    def enable_logs(self, dab_cap=8192, event_cap=256):
        for e in self.envs:
            self.L.po_surface_enable_dablog(self.L.po_env_surface(e), dab_cap)
            self.L.po_env_enable_eventlog(e, event_cap)

    def pop_dabs(self, i):
        s = self.L.po_env_surface(self.envs[i])
        n = self.L.po_surface_dablog_count(s)
        p = self.L.po_surface_dablog(s)
        arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n, C.sizeof(PODab))).copy()
        self.L.po_surface_dablog_reset(s)
        dt = np.dtype([("x", "<f4"), ("y", "<f4"), ("radius", "<f4"), ("hardness", "<f4"),
                       ("opacity", "<u2"), ("r", "<u2"), ("g", "<u2"), ("b", "<u2")])
        return arr.view(dt).reshape(n)

    def pop_events(self, i):
        e = self.envs[i]
        n = self.L.po_env_eventlog_count(e)
        p = self.L.po_env_eventlog(e)
        ev = [(p[k].x, p[k].y, p[k].pressure, p[k].dtime) for k in range(n)]
        self.L.po_env_eventlog_reset(e)
        return ev

    def brush_states(self, i):
        b = self.L.po_env_brush(self.envs[i])
        return np.array([self.L.po_brush_get_state(b, k) for k in range(30)], np.float32)

    def counters(self, i):
        a, b, c = C.c_long(), C.c_long(), C.c_long()
        self.L.po_surface_counters(self.L.po_env_surface(self.envs[i]), C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def run_episodes(self, actions, want_obs=False):
        """actions [B,T,K] -> final canvases [B,64,64,4] u16 (and obs [B,T+1,...] f32)."""
        a = np.ascontiguousarray(actions, dtype=np.int32)
        B, T, K = a.shape
        canv = np.zeros((B, 64, 64, 4), np.uint16)
        obs = np.zeros((B, T + 1, 64, 64, self.channels), np.float32) if want_obs else None
        for i, e in enumerate(self.envs):
            self.L.po_env_run_episode(e, a[i].ctypes.data, T, K, self._noise_ptr(), self.channels,
                                      obs[i].ctypes.data if want_obs else None, canv[i].ctypes.data)
        return (canv, obs) if want_obs else canv


# --------------------------------------------------------------------------
# Literal Python restatement of envs/mnist.py (single canvas)
# --------------------------------------------------------------------------
def _multiply(x, y, d):
    return x * d, y * d


def _add(x1, y1, x2, y2):
    return x1 + x2, y1 + y2


def _multiply_add(x1, y1, x2, y2, d):
    x3, y3 = _multiply(x2, y2, d)
    return _add(x1, y1, x3, y3)


def _difference(x1, y1, x2, y2):
    return x2 - x1, y2 - y1


def _midpoint(x1, y1, x2, y2):
    return (x1 + x2) / 2.0, (y1 + y2) / 2.0


def _length_and_normal(x1, y1, x2, y2):
    import math
    x, y = _difference(x1, y1, x2, y2)
    length = math.sqrt(x * x + y * y)
    if length == 0.0:
        x, y = 0.0, 0.0
    else:
        x, y = x / length, y / length
    return length, x, y


def _point_on_curve_1(t, cx, cy, sx, sy, x1, y1, x2, y2):
    ratio = t / 100.0
    x3, y3 = _multiply_add(sx, sy, x1, y1, ratio)
    x4, y4 = _multiply_add(cx, cy, x2, y2, ratio)
    x5, y5 = _difference(x3, y3, x4, y4)
    return _multiply_add(x3, y3, x5, y5, ratio)


class LiteralMNISTEnv:
    """envs/mnist.py MNIST/SimpleMNIST for ONE canvas, Python control flow as in the
    reference, brush/surface = the C restatement reached per event via ctypes."""
    head = 0.25
    tail = 0.75
    size = 0.2
    pressure = 0.3

    def __init__(self, spec: EnvSpec):
        self.L = lib()
        self.spec = spec
        self.acs = list(spec.action_names)
        self.ac_idx = {a: i for i, a in enumerate(self.acs)}
        self.cfg = spec.make_cfg()
        self.brush = parse_myb(brush_to_myb() if spec.brush_text is None else spec.brush_text)
        self.names = setting_names()
        self.inames = input_names()
        Lsz = spec.location_size
        x = np.linspace(0, spec.screen_size, Lsz)
        grid = np.meshgrid(x, x)
        self.ends = np.array(list(zip(*np.vstack(list(map(np.ravel, grid))))))   # utils.py:12-14
        self.controls = self.ends
        self.sizes = np.arange(0.2, 2.0, 0.5) * 1
        if "size" in self.acs:
            self.sizes = self.sizes[:spec.n_size]
        self.pressures = np.arange(0.8, 0, -0.3)
        if "pressure" in self.acs:
            self.pressures = self.pressures[:spec.n_pressure]
        self.jumps = [0, 1]
        self.colors = PALETTE[:spec.n_color] if "color" in self.acs else PALETTE
        self.colorize = spec.colorize
        self.s = self.L.po_surface_new()
        self.b = None
        self.events = []

    def reset(self):
        self.entry_pressure = np.min(self.pressures)
        self.L.po_surface_clear(self.s)
        self.L.po_surface_fill_white(self.s)
        if self.b is not None:
            self.L.po_brush_free(self.b)
        self.b = self.L.po_brush_new(self.spec.math_mode)
        for i, n in enumerate(self.names):
            base, pts = self.brush[n]
            self.L.po_brush_set_base_value(self.b, i, base)
            for iname, plist in pts.items():
                j = self.inames.index(iname)
                self.L.po_brush_set_mapping_n(self.b, i, j, len(plist))
                for k, (x, y) in enumerate(plist):
                    self.L.po_brush_set_mapping_point(self.b, i, j, k, x, y)
        self._step = 0
        self.s_x, self.s_y = None, None

    def _stroke_to(self, x, y, pressure, duration=0.1):
        self.events.append((np.float32(x), np.float32(y), np.float32(pressure), duration))
        self.L.po_brush_stroke_to(self.b, self.s, x, y, pressure, 0.0, 0.0, duration)

    def draw(self, ac, dtime=1):
        jump = 0
        x, y = self.ends[0]
        c_x, c_y = self.controls[0]
        pressure, size = self.pressure, self.size
        color = None
        for name in self.acs:
            named_ac = ac[self.ac_idx[name]]
            value = getattr(self, name + "s")[named_ac]
            if name == 'end':
                x, y = value
            if name == 'control':
                c_x, c_y = value
            elif name == 'pressure':
                pressure = value
            elif name == 'size':
                size = value
            elif name == 'jump':
                jump = value
            elif name == 'color':
                color = value
        sn = self.names
        if color is not None or self.colorize:
            rgb = color if color is not None else self.colors[self._step % len(self.colors)]
            h, s, v = rgb_to_hsv(rgb)
            self.L.po_brush_set_base_value(self.b, sn.index("color_h"), h)
            self.L.po_brush_set_base_value(self.b, sn.index("color_s"), s)
            self.L.po_brush_set_base_value(self.b, sn.index("color_v"), v)
        if 'size' in self.acs:
            self.L.po_brush_set_base_value(self.b, sn.index("radius_logarithmic"), size)

        if self.s_x is None and self.s_y is None:
            pressure = 0
            self.s_x, self.s_y = 0, 0
            self._stroke_to(self.s_x, self.s_y, pressure)
        elif 'jump' in self.acs and jump:
            pressure = 0
            self._stroke_to(self.s_x, self.s_y, pressure)
        else:
            self._stroke_to(self.s_x, self.s_y, pressure)
        self._draw(x, y, c_x, c_y, pressure, size, dtime)

    def _draw(self, x, y, c_x, c_y, pressure, size, dtime):
        end_pressure = pressure
        if 'control' not in self.acs or pressure == 0:
            self.events.append((np.float32(x), np.float32(y), np.float32(pressure), dtime))
            self.L.po_brush_stroke_to(self.b, self.s, x, y, pressure, 0, 0, dtime)
        else:
            end_pressure = self.curve(c_x, c_y, self.s_x, self.s_y, x, y, pressure)
        self.entry_pressure = end_pressure
        self.s_x, self.s_y = x, y

    def _line_settings(self, pressure):
        p1 = self.entry_pressure
        p2 = (self.entry_pressure + pressure) / 2
        p3 = pressure
        if self.head == 0.0001:
            p1 = p2
        prange1 = p2 - p1
        prange2 = p3 - p2
        return p1, p2, prange1, prange2, self.head, self.tail

    def curve(self, cx, cy, sx, sy, ex, ey, pressure):
        entry_p, midpoint_p, prange1, prange2, h, t = self._line_settings(pressure)
        points_in_curve = 100
        mx, my = _midpoint(sx, sy, ex, ey)
        length, nx, ny = _length_and_normal(mx, my, cx, cy)
        cx, cy = _multiply_add(mx, my, nx, ny, length * 2)
        x1, y1 = _difference(sx, sy, cx, cy)
        x2, y2 = _difference(cx, cy, ex, ey)
        head = points_in_curve * h
        head_range = int(head) + 1
        tail = points_in_curve * t
        tail_range = int(tail) + 1
        tail_length = points_in_curve - tail

        px, py = _point_on_curve_1(1, cx, cy, sx, sy, x1, y1, x2, y2)
        length, nx, ny = _length_and_normal(sx, sy, px, py)
        bx, by = _multiply_add(sx, sy, nx, ny, 0.25)
        self._stroke_to(bx, by, entry_p)
        pressure = abs(1 / head * prange1 + entry_p)
        self._stroke_to(px, py, pressure)
        for i in range(2, head_range):
            px, py = _point_on_curve_1(i, cx, cy, sx, sy, x1, y1, x2, y2)
            pressure = abs(i / head * prange1 + entry_p)
            self._stroke_to(px, py, pressure)
        for i in range(head_range, tail_range):
            px, py = _point_on_curve_1(i, cx, cy, sx, sy, x1, y1, x2, y2)
            self._stroke_to(px, py, midpoint_p)
        for i in range(tail_range, points_in_curve + 1):
            px, py = _point_on_curve_1(i, cx, cy, sx, sy, x1, y1, x2, y2)
            pressure = abs((i - tail) / tail_length * prange2 + midpoint_p)
            self._stroke_to(px, py, pressure)
        return pressure

    def step(self, acs):
        self.draw(acs)
        self._step += 1

    def canvas(self):
        p = self.L.po_surface_data(self.s)
        return np.ctypeslib.as_array(p, shape=(64, 64, 4)).copy()

    def image(self, noise=None):
        out = np.zeros((64, 64, 4), np.uint8)
        self.L.po_surface_read_rgbu8(self.s, noise.ctypes.data if noise is not None else None,
                                     out.ctypes.data)
        return out

    def state(self, noise=None):
        rgb = self.image(noise)
        return np.expand_dims(np.dot(rgb[..., :3], [0.299, 0.587, 0.114]), -1)   # utils.py:3-5
