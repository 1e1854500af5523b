// raster.cuh -- device-side parameter block shared by the K1 kernels and the C-ABI glue.
#pragma once
#include <cstdint>

namespace spiral {

constexpr int kTile = 64;
constexpr int kBrushFloats = 16;
constexpr int kEnvDoubles = 16;
constexpr int kDabFloats = 8;

// brush_state float[B,16]
enum {
    BS_X = 0, BS_Y, BS_PRESSURE, BS_PARTIAL_DABS, BS_ACTUAL_RADIUS, BS_ACTUAL_X, BS_ACTUAL_Y,
    BS_SPEED1_SLOW, BS_SPEED2_SLOW,
    BS_FLAGS,      // int bits: 1 = reset_requested, 2 = settings_value[] initialised
    BS_STEP,       // int: env._step
    BS_RNG_POS,    // int: rng-double stream position (values consumed since the Brush was created)
    BS_INV_RADIUS, // FAST arithmetic mode: exp(-radius_logarithmic), kept beside BS_ACTUAL_RADIUS
};
// env_state double[B,16]
enum {
    ES_RNG0 = 0,             // [0..9] first RNG output block after seeding (informational)
    ES_SX = 10, ES_SY = 11,  // env.s_x, env.s_y
    ES_ENTRY_PRESSURE = 12,
    ES_HAVE_START = 13,      // 0.0 while s_x is None
};
// pending dab record float[8]
enum { DAB_X = 0, DAB_Y, DAB_RLOG, DAB_ACTUAL_RADIUS, DAB_OPAQUE, DAB_HARDNESS, DAB_AA, DAB_HAS_RAND };

// arithmetic of the brush state machine (SpiralEnvCfg.math_mode)
constexpr int kMathDet = 1;    // detmath.cuh: double-precision exp / log / hypot, bit-exact vs the oracle's PO_MATH_DET
constexpr int kMathFast = 2;   // fastmath.cuh + restructured binary32 dab loop, bit-exact vs the oracle's PO_MATH_FAST

constexpr int kDyn = 9;
constexpr int kIn = 3;
constexpr int kPts = 4;

struct MapDev {
    float base;
    int n[kIn];
    float x[kIn][kPts];
    float y[kIn][kPts];
    float inv[kIn][kPts];   // 1/(x[i+1]-x[i]) when that is a power of two (division == multiplication), else 0
    float slope[kIn][kPts]; // FAST arithmetic mode: segment i is fmaf(slope[i], x, icpt[i])
    float icpt[kIn][kPts];
};

struct EnvParams {
    int B, C, L, K;
    int idx_jump, idx_control, idx_end, idx_pressure, idx_size, idx_color;
    int n_colors, colorize, dither, max_dabs;
    int uses_speed1, uses_speed2;
    int mapped_mask;             // bit s set: dyn[s] has at least one input mapping
    int dry_structure;           // 1: mapping structure of dry_brush.myb (fast path in update_states)
    int rng_table_len;           // entries of the rand_gauss table (stream positions)
    int expand_lanes;            // canvases per warp in expand_kernel (0 = choose from the batch size)
    int expand_sort;             // 1: sort canvases by predicted chain length before expansion
    int expand_solo_div;         // the B / expand_solo_div longest canvases get a warp each
    double lin[64];
    double pressures[8];
    double sizes[8];
    float base_rlog[9];          // radius_logarithmic base per size entry; [8] = brush file value
    float base_radius[9];        // expf(base_rlog)
    unsigned short colors[8][4];
    double default_pressure, default_size, head, tail, entry_pressure0;
    MapDev dyn[kDyn];
    float slow_tracking;
    float track_fac[2];          // 1 - exp_decay(slow_tracking, 100*dtime) for dtime = 0.1, 1.0
    float sp_gamma[2], sp_m[2], sp_q[2];
    float dabs_basic, dabs_actual, dabs_second;
    float cs, sn, l2, rad_area_1;
    int fast_spans;              // round upright dabs (ratio 1, sn == 1, |cs| ~ 0): composite may list candidate pixels
    float cull_factor;           // upper bound of radius / ACTUAL_RADIUS (radius_by_random)
    double rng_seed_block[10];   // first output block of rng_double after set_seed(1000)
    // ---- FAST arithmetic mode (kMathFast) ----
    int math_mode;               // kMathDet | kMathFast
    float f_inv_norm01[9];       // 1 / ((float)0.1 * base_radius[i]): events of 0.1 s
    float f_inv_norm1[9];        // 1 / (1.0f * base_radius[i]): the straight stroke's 1 s event
    float f_inv_norm001[9];      // 1 / (0.001f * base_radius[i]): step_dtime clamped
    float f_k_basic[9];          // dabs_per_basic_radius / clamp(base_radius[i])
    float f_neg_inv_slow[2];     // -1 / speed{1,2}_slowness when that setting has no mapping (else 0)
    // ---- raw dab hand-off (raw_dab_may_paint) ----
    float cull_pad[9];           // per size entry: largest position jitter + anti-aliasing fringe + 3 px
    int zero_pressure_blank;     // 1: a dab of zero pressure has zero opacity (opaque_multiply = 0 + y(pressure), y(0) = 0)
};

}  // namespace spiral
