"""ctypes binding of libspiral_b200.so (the C-ABI declared in include/spiral_b200.h).

The CUDA extension is the product: there is NO CPU fallback.  Importing this module never
touches the GPU, but every entry point raises if the shared library is missing, and the
library itself returns SPIRAL_ECUDA when no sm_100a device is present.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("SPIRAL_B200_LIB") or os.path.join(CSRC, "libspiral_b200.so")   # override: tuning builds

ABI_VERSION = 3
MATH_DET, MATH_FAST = 1, 2       # SpiralEnvCfg.math_mode
MAX_POINTS, MAX_TABLE, MAX_L, TILE = 4, 8, 64, 64
BRUSH_STATE_FLOATS, ENV_STATE_DOUBLES, DAB_FLOATS = 16, 16, 8
DYN_NAMES = ["opaque", "opaque_multiply", "radius_logarithmic", "hardness", "radius_by_random",
             "offset_by_random", "speed1_slowness", "speed2_slowness", "anti_aliasing"]
IN_NAMES = ["pressure", "speed1", "speed2"]
DISC_MAX_LAYERS = 8
MAX_HEADS = 8


class SpiralMapping(C.Structure):
    _fields_ = [("base", C.c_float),
                ("n", C.c_int32 * len(IN_NAMES)),
                ("x", (C.c_float * MAX_POINTS) * len(IN_NAMES)),
                ("y", (C.c_float * MAX_POINTS) * len(IN_NAMES))]


class SpiralBrush(C.Structure):
    _fields_ = [("dyn", SpiralMapping * len(DYN_NAMES)),
                ("slow_tracking", C.c_float),
                ("speed1_gamma", C.c_float), ("speed2_gamma", C.c_float),
                ("dabs_per_basic_radius", C.c_float), ("dabs_per_actual_radius", C.c_float),
                ("dabs_per_second", C.c_float),
                ("elliptical_dab_angle", C.c_float)]


class SpiralEnvCfg(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("batch", C.c_int32),
                ("channels", C.c_int32), ("location_size", C.c_int32), ("n_heads", C.c_int32),
                ("idx_jump", C.c_int32), ("idx_control", C.c_int32), ("idx_end", C.c_int32),
                ("idx_pressure", C.c_int32), ("idx_size", C.c_int32), ("idx_color", C.c_int32),
                ("n_pressures", C.c_int32), ("n_sizes", C.c_int32), ("n_colors", C.c_int32),
                ("colorize", C.c_int32), ("dither", C.c_int32), ("max_dabs", C.c_int32),
                ("rng_table_len", C.c_int32),
                ("lin", C.c_double * MAX_L),
                ("pressures", C.c_double * MAX_TABLE),
                ("sizes", C.c_double * MAX_TABLE),
                ("colors_fix15", (C.c_uint16 * 4) * MAX_TABLE),
                ("default_pressure", C.c_double), ("default_size", C.c_double),
                ("head", C.c_double), ("tail", C.c_double),
                ("entry_pressure0", C.c_double),
                ("brush", SpiralBrush),
                ("dither_table", C.c_uint16 * (TILE * TILE)),
                ("math_mode", C.c_int32)]


class SpiralDiscCfg(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("max_batch", C.c_int32),
                ("in_channels", C.c_int32), ("screen_size", C.c_int32), ("disc_dim", C.c_int32),
                ("batch_norm", C.c_int32), ("normalize", C.c_int32), ("precision", C.c_int32)]


class SpiralDiscWeights(C.Structure):
    _fields_ = [("conv_w", C.c_void_p * DISC_MAX_LAYERS), ("conv_b", C.c_void_p * DISC_MAX_LAYERS),
                ("bn_gamma", C.c_void_p * DISC_MAX_LAYERS), ("bn_beta", C.c_void_p * DISC_MAX_LAYERS),
                ("dense_w", C.c_void_p), ("dense_b", C.c_void_p)]


# every symbol include/spiral_b200.h declares: name -> (restype, argtypes)
_vp = C.c_void_p
SYMBOLS = {
    "spiral_version": (C.c_int, []),
    "spiral_last_error": (C.c_char_p, []),
    "spiral_env_create": (C.c_int, [C.POINTER(SpiralEnvCfg), C.POINTER(_vp)]),
    "spiral_env_destroy": (C.c_int, [_vp]),
    "spiral_env_reset": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "spiral_env_step": (C.c_int, [_vp] * 10),
    "spiral_env_expand": (C.c_int, [_vp] * 8),
    "spiral_env_composite": (C.c_int, [_vp] * 6),
    "spiral_env_debug_dabs": (C.c_int, [_vp] * 6),
    "spiral_l2_reward": (C.c_int, [_vp, _vp, _vp, C.c_int32, C.c_int32, _vp]),
    "spiral_debug_detmath": (C.c_int, [_vp, _vp, _vp, C.c_int32, _vp]),
    "spiral_debug_bounds_violations": (C.c_int64, []),
    "spiral_debug_divcheck": (C.c_int, [C.c_uint64, C.c_int64, _vp, _vp]),
    "spiral_debug_fastmath": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int32, _vp]),
    "spiral_a2c_loss": (C.c_int, [C.POINTER(_vp), C.POINTER(C.c_int32), C.c_int32, _vp, _vp, _vp, _vp,
                                  C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_float,
                                  _vp, _vp, _vp, C.POINTER(_vp), _vp, _vp, _vp]),
    "spiral_disc_create": (C.c_int, [C.POINTER(SpiralDiscCfg), C.POINTER(_vp)]),
    "spiral_disc_destroy": (C.c_int, [_vp]),
    "spiral_disc_workspace_bytes": (C.c_int64, [_vp, C.c_int32]),
    "spiral_disc_load_weights": (C.c_int, [_vp, C.POINTER(SpiralDiscWeights), _vp, _vp]),
    "spiral_disc_forward": (C.c_int, [_vp, _vp, C.c_int32, _vp, _vp, _vp, _vp, _vp]),
    "spiral_disc_forward_staged": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, C.c_double, _vp, _vp, _vp, _vp, _vp, _vp]),
    "spiral_categorical_sample": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_uint64, C.c_uint64, _vp, _vp, _vp]),
    "spiral_debug_philox": (C.c_int, [C.POINTER(C.c_uint32), _vp, _vp]),
    "spiral_grad_global_norm": (C.c_int, [C.POINTER(_vp), C.POINTER(C.c_int64), C.c_int32, _vp, _vp, _vp]),
    "spiral_adam_step": (C.c_int, [C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(C.c_int64),
                                   C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _vp, C.c_float, _vp]),
    "spiral_adam_step_lr_dev": (C.c_int, [C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(C.c_int64),
                                          C.c_int32, _vp, C.c_float, C.c_float, C.c_float, C.c_float, _vp, C.c_float, _vp]),
    "spiral_pconv_forward": (C.c_int, [_vp] * 6 + [C.c_int32] * 20 + [_vp]),
    "spiral_pconv_pack": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _vp, _vp,
                                    C.c_int32, _vp]),
    "spiral_pres_stack": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, _vp]),
    "spiral_phead_last": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int32, C.c_int32, _vp]),
    "spiral_lconv_packed_elems": (C.c_int64, [C.c_int32] * 3),
    "spiral_lconv_pack": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                    C.POINTER(C.c_int32), C.POINTER(C.c_int32), _vp]),
    "spiral_lconv_forward": (C.c_int, [_vp, _vp, C.c_int32, _vp, _vp] + [C.c_int32] * 19 + [_vp]),
    "spiral_lconv_forward_classes": (C.c_int, [_vp, _vp, C.c_int32, _vp, _vp] + [C.c_int32] * 7 + [_vp, _vp] + [C.c_int32] * 3 + [_vp]),
    "spiral_lconv_wgrad_workspace_floats": (C.c_int64, [C.c_int32] * 3),
    "spiral_lconv_wgrad": (C.c_int, [_vp] * 4 + [C.c_int32] * 13 + [_vp]),
    "spiral_conv_workspace_bytes": (C.c_int64, [C.c_int32] * 5),
    "spiral_conv_forward": (C.c_int, [_vp, _vp, _vp] + [C.c_int32] * 5 + [_vp, _vp]),
    "spiral_conv_dgrad": (C.c_int, [_vp, _vp, _vp] + [C.c_int32] * 5 + [_vp, _vp]),
    "spiral_conv_wgrad": (C.c_int, [_vp, _vp, _vp] + [C.c_int32] * 5 + [_vp, _vp]),
    "spiral_conv_packed_bytes": (C.c_int64, [C.c_int32] * 6),
    "spiral_conv_pack": (C.c_int, [_vp, _vp] + [C.c_int32] * 6 + [_vp]),
    "spiral_conv_forward_packed": (C.c_int, [_vp, _vp, _vp] + [C.c_int32] * 5 + [_vp, _vp]),
    "spiral_conv_dgrad_packed": (C.c_int, [_vp, _vp, _vp] + [C.c_int32] * 5 + [_vp, _vp]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile every CUDA translation unit for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["bash", os.path.join(CSRC, "build.sh")], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building libspiral_b200.so failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout + out.stderr)
    return LIB_PATH


def lib():
    """Load the shared library; raise (never fall back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH) and not os.environ.get("SPIRAL_B200_LIB") and shutil.which(
            os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")):
        build()                      # in-tree build (csrc/build.sh, sm_100a); raises loudly when nvcc fails
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libspiral_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; "
            "g.build()'` or spiral-tensorflow_b200/csrc/build.sh; there is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(L, name)        # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if L.spiral_version() != ABI_VERSION:
        raise RuntimeError("libspiral_b200.so ABI %d != binding ABI %d" % (L.spiral_version(), ABI_VERSION))
    _lib = L
    return L


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().spiral_last_error().decode(errors="replace")
        raise RuntimeError("%s failed (%d): %s" % (what or "libspiral_b200", rc, msg))


def ptr(t):
    """Device (or host) address of a torch tensor / None."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(stream=None):
    import torch
    s = torch.cuda.current_stream() if stream is None else stream
    return C.c_void_p(s.cuda_stream)
