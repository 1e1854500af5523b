"""MyPaint brush definitions for the batched rasteriser.

Host-side mirror of what the reference does with mypaint 1.2.1's ``lib/brush.py`` at
``envs/mnist.py:93-95`` (``BrushInfo(fp.read())``) and ``:132-135`` (``set_color_rgb``,
``set_base_value('radius_logarithmic', size)``): parse a version-2 ``.myb`` text file into base
values + input mappings, then compile the subset the CUDA kernels implement into the
``SpiralBrush`` struct of include/spiral_b200.h.  Brushes that need anything else (smudge,
eraser, elliptical dabs, mappings on random/stroke/direction/tilt/custom inputs, ...) are
rejected loudly -- there is no slow path.
"""
from __future__ import annotations

import colorsys
import math

import numpy as np

from . import _native as N

# Default brush of the reference (assets/brushes/dry_brush.myb:1-37, selected by config.py:38),
# restated as data: (setting, base value, {input: [(x, y), ...]}).
DRY_BRUSH = (
    ("opaque", 0.8, {"pressure": ((0.0, 0.0), (1.0, 0.2))}),
    ("opaque_multiply", 0.0, {"pressure": ((0.0, 0.0), (1.0, 1.0))}),
    ("opaque_linearize", 0.0, {}),
    ("radius_logarithmic", 0.6, {"speed2": ((0.0, 0.042857), (4.0, -0.3))}),
    ("hardness", 0.2, {}),
    ("dabs_per_basic_radius", 6.0, {}),
    ("dabs_per_actual_radius", 6.0, {}),
    ("dabs_per_second", 0.0, {}),
    ("radius_by_random", 0.1, {}),
    ("speed1_slowness", 0.04, {}),
    ("speed2_slowness", 0.8, {}),
    ("speed1_gamma", 4.0, {}),
    ("speed2_gamma", 4.0, {}),
    ("offset_by_random", 0.0, {"pressure": ((0.0, 0.0), (1.0, 1.4))}),
    ("offset_by_speed", 0.0, {}),
    ("offset_by_speed_slowness", 1.0, {}),
    ("slow_tracking", 2.0, {}),
    ("slow_tracking_per_dab", 0.0, {}),
    ("tracking_noise", 0.0, {}),
    ("color_h", 0.0, {}), ("color_s", 0.0, {}), ("color_v", 0.0, {}),
    ("change_color_h", 0.0, {}), ("change_color_l", 0.0, {}), ("change_color_hsl_s", 0.0, {}),
    ("change_color_v", 0.0, {}), ("change_color_hsv_s", 0.0, {}),
    ("smudge", 0.0, {}), ("smudge_length", 0.5, {}),
    ("stroke_threshold", 0.0, {}),
    ("stroke_duration_logarithmic", 4.0, {}),
    ("stroke_holdtime", 0.0, {}),
    ("custom_input", 0.0, {}), ("custom_input_slowness", 0.0, {}),
)

# libmypaint brushsettings.json defaults for settings a file may omit
SETTING_DEFAULTS = {
    "opaque": 1.0, "opaque_multiply": 0.0, "opaque_linearize": 0.9, "radius_logarithmic": 2.0,
    "hardness": 0.8, "anti_aliasing": 1.0, "dabs_per_basic_radius": 0.0,
    "dabs_per_actual_radius": 2.0, "dabs_per_second": 0.0, "radius_by_random": 0.0,
    "speed1_slowness": 0.04, "speed2_slowness": 0.8, "speed1_gamma": 4.0, "speed2_gamma": 4.0,
    "offset_by_random": 0.0, "offset_by_speed": 0.0, "offset_by_speed_slowness": 1.0,
    "slow_tracking": 0.0, "slow_tracking_per_dab": 0.0, "tracking_noise": 0.0,
    "color_h": 0.0, "color_s": 0.0, "color_v": 0.0, "restore_color": 0.0,
    "change_color_h": 0.0, "change_color_l": 0.0, "change_color_hsl_s": 0.0,
    "change_color_v": 0.0, "change_color_hsv_s": 0.0, "smudge": 0.0, "smudge_length": 0.5,
    "smudge_radius_log": 0.0, "eraser": 0.0, "stroke_threshold": 0.0,
    "stroke_duration_logarithmic": 4.0, "stroke_holdtime": 0.0, "custom_input": 0.0,
    "custom_input_slowness": 0.0, "elliptical_dab_ratio": 1.0, "elliptical_dab_angle": 90.0,
    "direction_filter": 2.0, "lock_alpha": 0.0, "colorize": 0.0, "snap_to_pixel": 0.0,
    "pressure_gain_log": 0.0,
}

# settings that must keep this constant value for the kernels to be exact
_MUST_BE = {
    "opaque_linearize": 0.0, "offset_by_speed": 0.0, "slow_tracking_per_dab": 0.0,
    "tracking_noise": 0.0, "smudge": 0.0, "eraser": 0.0, "change_color_h": 0.0,
    "change_color_l": 0.0, "change_color_hsl_s": 0.0, "change_color_v": 0.0,
    "change_color_hsv_s": 0.0, "elliptical_dab_ratio": 1.0, "lock_alpha": 0.0, "colorize": 0.0,
    "snap_to_pixel": 0.0, "pressure_gain_log": 0.0, "stroke_threshold": 0.0,
}
# constant-only settings the kernels read
_CONST = ("slow_tracking", "speed1_gamma", "speed2_gamma", "dabs_per_basic_radius",
          "dabs_per_actual_radius", "dabs_per_second", "elliptical_dab_angle")
# settings that only drive inputs no supported mapping can read (dead state)
_IGNORED = ("offset_by_speed_slowness", "color_h", "color_s", "color_v", "restore_color",
            "smudge_length", "smudge_radius_log", "stroke_duration_logarithmic", "stroke_holdtime",
            "custom_input", "custom_input_slowness", "direction_filter")


class BrushInfo(object):
    """Subset of mypaint ``lib/brush.py:BrushInfo``: ``settings[name] = [base, {input: points}]``."""

    def __init__(self, string=None):
        self.settings = {k: [float(v), {}] for k, v in SETTING_DEFAULTS.items()}
        if string is not None:
            self.load_from_string(string)
        else:
            for name, base, pts in DRY_BRUSH:
                self.settings[name] = [float(base), {k: list(v) for k, v in pts.items()}]

    def load_from_string(self, text):
        version = 1
        for raw in text.splitlines():
            line = raw.strip()
            if not line or line.startswith("#"):
                continue
            cname, rawvalue = line.split(" ", 1)
            if cname == "version":
                version = int(rawvalue)
                if version != 2:
                    raise NotImplementedError("only version-2 .myb brush files are supported, got %d" % version)
                continue
            if cname == "stroke_treshold":
                cname = "stroke_threshold"
            if cname == "speed":
                cname = "speed1"
            if cname not in self.settings:
                continue
            parts = rawvalue.split("|")
            points = {}
            for part in parts[1:]:
                inputname, rawpoints = part.strip().split(" ", 1)
                plist = []
                for s in rawpoints.split(","):
                    s = s.strip()
                    if not (s.startswith("(") and s.endswith(")") and " " in s):
                        raise ValueError('(x y) expected, got "%s"' % s)
                    x, y = [float(v) for v in s[1:-1].split(" ")]
                    plist.append((x, y))
                points[inputname] = plist
            self.settings[cname] = [float(parts[0]), points]
        return self

    def get_base_value(self, cname):
        return self.settings[cname][0]

    def set_base_value(self, cname, value):
        self.settings[cname][0] = float(value)

    def to_myb(self):
        lines = ["# mypaint brush file", "version 2"]
        for name, (base, pts) in self.settings.items():
            line = "%s %r" % (name, float(base))
            for iname, plist in pts.items():
                line += " | %s %s" % (iname, ", ".join("(%f %f)" % (x, y) for x, y in plist))
            lines.append(line)
        return "\n".join(lines) + "\n"

    # ------------------------------------------------------------------
    def compile(self) -> N.SpiralBrush:
        """Validate and pack into the C struct.  Raises NotImplementedError for unsupported brushes."""
        for name, want in _MUST_BE.items():
            base, pts = self.settings[name]
            if pts or np.float32(base) != np.float32(want):
                raise NotImplementedError(
                    "brush setting %s=%r%s is outside the subset the CUDA rasteriser implements "
                    "(needs constant %r)" % (name, base, " with input mappings" if pts else "", want))
        for name in _CONST:
            if self.settings[name][1]:
                raise NotImplementedError("brush setting %s must be constant (no input mapping)" % name)
        br = N.SpiralBrush()
        for s, name in enumerate(N.DYN_NAMES):
            base, pts = self.settings[name]
            m = br.dyn[s]
            m.base = base
            for iname, plist in pts.items():
                if iname not in N.IN_NAMES:
                    raise NotImplementedError(
                        "mapping %s <- %s: only the inputs %s are implemented" % (name, iname, N.IN_NAMES))
                j = N.IN_NAMES.index(iname)
                if not (2 <= len(plist) <= N.MAX_POINTS):
                    raise NotImplementedError("mapping %s <- %s has %d points (2..%d supported)"
                                              % (name, iname, len(plist), N.MAX_POINTS))
                m.n[j] = len(plist)
                for k, (x, y) in enumerate(plist):
                    m.x[j][k] = x
                    m.y[j][k] = y
        for name in self.settings:
            if name in N.DYN_NAMES or name in _MUST_BE or name in _CONST or name in _IGNORED:
                continue
            raise NotImplementedError("unknown brush setting %s" % name)
        for name in _IGNORED:
            if self.settings[name][1] and name.startswith("color_"):
                raise NotImplementedError("mappings on %s are not supported" % name)
        br.slow_tracking = self.settings["slow_tracking"][0]
        br.speed1_gamma = self.settings["speed1_gamma"][0]
        br.speed2_gamma = self.settings["speed2_gamma"][0]
        br.dabs_per_basic_radius = self.settings["dabs_per_basic_radius"][0]
        br.dabs_per_actual_radius = self.settings["dabs_per_actual_radius"][0]
        br.dabs_per_second = self.settings["dabs_per_second"][0]
        br.elliptical_dab_angle = self.settings["elliptical_dab_angle"][0]
        return br


# ----------------------------------------------------------------------
# colour: set_color_rgb (python, float64) -> color_h/s/v base values (float32) ->
# hsv_to_rgb_float per dab (libmypaint helpers.c) -> fix15
# ----------------------------------------------------------------------
def rgb_to_hsv(rgb):
    """mypaint lib/helpers.py:rgb_to_hsv (clamp + colorsys)."""
    r, g, b = [min(max(float(c), 0.0), 1.0) for c in rgb]
    return colorsys.rgb_to_hsv(r, g, b)


def hsv_to_rgb_float(h, s, v):
    """libmypaint helpers.c:hsv_to_rgb_float with its float/double mix."""
    f32 = np.float32
    h, s, v = f32(h), f32(s), f32(v)
    h = f32(float(h) - math.floor(float(h)))
    s = f32(min(max(float(s), 0.0), 1.0))
    v = f32(min(max(float(v), 0.0), 1.0))
    if float(s) == 0.0:
        return v, v, v
    hue = float(h)
    if hue == 1.0:
        hue = 0.0
    hue *= 6.0
    i = int(hue)
    f = hue - i
    vd, sd = float(v), float(s)
    w = f32(vd * (1.0 - sd))
    q = f32(vd * (1.0 - (sd * f)))
    t = f32(vd * (1.0 - (sd * (1.0 - f))))
    return [(v, t, w), (q, v, w), (w, v, t), (w, q, v), (t, w, v), (v, w, q)][i]


def color_fix15(rgb):
    """Dab colour as draw_dab stores it: uint16(clamp(c) * (1 << 15)) after the HSV round trip."""
    h, s, v = rgb_to_hsv(rgb)
    out = []
    for c in hsv_to_rgb_float(h, s, v):
        c = np.float32(min(max(float(c), 0.0), 1.0))
        out.append(int(np.float32(c * np.float32(32768.0))))
    return out


def brush_color_fix15(info: BrushInfo):
    """Colour of a brush whose colour is never set by the env (color_h/s/v from the file)."""
    out = []
    for c in hsv_to_rgb_float(info.get_base_value("color_h"), info.get_base_value("color_s"),
                              info.get_base_value("color_v")):
        c = np.float32(min(max(float(c), 0.0), 1.0))
        out.append(int(np.float32(c * np.float32(32768.0))))
    return out


# ----------------------------------------------------------------------
# read-back dither: mypaint lib/pixops.cpp fills dithering_noise[] from the unseeded libc rand()
# ----------------------------------------------------------------------
def glibc_rand_stream(n, seed=1):
    """glibc rand(): TYPE_3 additive-feedback generator r[i] = r[i-3] + r[i-31] (mod 2^32) >> 1."""
    r = [0] * 34
    r[0] = seed if seed else 1
    for i in range(1, 31):
        hi, lo = divmod(r[i - 1], 127773) if r[i - 1] >= 0 else (-(-r[i - 1] // 127773), -(-r[i - 1] % 127773))
        word = 16807 * lo - 2836 * hi
        if word < 0:
            word += 2147483647
        r[i] = word
    ring = [x & 0xFFFFFFFF for x in r[:31]]
    f, b = 3, 0
    out = np.zeros(n, np.int64)
    for i in range(310 + n):
        ring[f] = (ring[f] + ring[b]) & 0xFFFFFFFF
        if i >= 310:
            out[i - 310] = ring[f] >> 1
        f = (f + 1) % 31
        b = (b + 1) % 31
    return out


_DITHER = None


def dither_table():
    """dithering_noise[i] = (rand() % (1<<15)) * 240/256 + (1<<15) * 8/256, first 64*64 entries."""
    global _DITHER
    if _DITHER is None:
        v = glibc_rand_stream(N.TILE * N.TILE)
        _DITHER = ((v % (1 << 15)) * 240 // 256 + (1 << 15) * 8 // 256).astype(np.uint16)
    return _DITHER
