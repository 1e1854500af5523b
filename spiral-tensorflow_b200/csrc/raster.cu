// raster.cu -- K1: batched brush-stroke rasteriser for sm_100a.
//
// Replaces, for B canvases at once, what the reference does one canvas per process through
// SWIG: envs/mnist.py:107-222 (action decode, Bezier stroke expansion into stroke_to events) and
// beneath it libmypaint 1.3.0 (mypaint_brush_stroke_to state machine, rng-double, draw_dab,
// render_dab_mask, BlendMode_Normal) and mypaint 1.2.1's rgbu16 -> rgbu8 read-back.
//
// Two kernels per env step:
//   expand_kernel    one THREAD per canvas: the brush state machine is a strictly serial chain per
//                    canvas (dab i's position/radius depend on the filtered speed, the partial-dab
//                    carry and the RNG position after dab i-1), so parallelism comes from the
//                    batch.  Emits "pending dabs" (8 floats) into a caller-owned list.
//   composite_kernel one WARP per canvas walks the dab list in order, one lane per candidate pixel,
//                    read-modify-writing the fix15 tile in place (64-bit RGBA words through L1/L2);
//                    only the rows a step dirtied are converted (fix15 -> uint8 with dither ->
//                    observation).  32 independent canvases per SM hide the per-dab latency.
//
// Arithmetic is bit-exact against oracle/paint_oracle.c in PO_MATH_DET mode: compile with
// --fmad=false, transcendental functions from detmath.cuh.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "common.h"
#include "detmath.cuh"
#include "fastmath.cuh"
#include "raster.cuh"

namespace spiral {

using detmath::det_exp;
using detmath::det_expf;
using detmath::det_hypotf;
using detmath::det_log;

// ------------------------------------------------------------------------------------------
// libmypaint mapping (mypaint-mapping.c: mypaint_mapping_calculate), inputs pressure/speed1/speed2.
// Settings without any mapping are the common case: `mapped` (warp-uniform, from the parameter
// block) skips the evaluation and returns the base value.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float map_eval(const MapDev &m, bool mapped, const float (&in)[kIn], float base)
{
    float result = base;
    if (!mapped) return result;
#pragma unroll
    for (int j = 0; j < kIn; j++) {
        const int n = m.n[j];
        if (n) {
            const float x = in[j];
            float x0 = m.x[j][0], y0 = m.y[j][0];
            float x1 = m.x[j][1], y1 = m.y[j][1];
            float inv = m.inv[j][0];
            if (n > 2) {
                for (int i = 2; i < n && x > x1; i++) {
                    x0 = x1;
                    y0 = y1;
                    x1 = m.x[j][i];
                    y1 = m.y[j][i];
                    inv = m.inv[j][i - 1];
                }
            }
            float y;
            if (x0 == x1 || y0 == y1) {
                y = y0;
            } else {
                const float num = y1 * (x - x0) + y0 * (x1 - x);
                // dividing by a power of two == multiplying by its reciprocal (both exact scalings)
                y = (inv != 0.0f) ? num * inv : num / (x1 - x0);
            }
            result += y;
        }
    }
    return result;
}

// one 2-point segment of a mapping (the n == 2 case of mypaint_mapping_calculate)
__device__ __forceinline__ float map_seg(const MapDev &m, int j, float x)
{
    const float x0 = m.x[j][0], y0 = m.y[j][0], x1 = m.x[j][1], y1 = m.y[j][1], inv = m.inv[j][0];
    if (x0 == x1 || y0 == y1) return y0;
    const float num = y1 * (x - x0) + y0 * (x1 - x);
    return (inv != 0.0f) ? num * inv : num / (x1 - x0);
}

// FAST arithmetic mode: a segment is fmaf(slope, x, intercept) with slope = (y1 - y0) / (x1 - x0) and intercept =
// y0 - slope * x0 tabulated at create time (oracle: mapping_calculate_fast evaluates the same expressions in place)
__device__ __forceinline__ float map_eval_fast(const MapDev &m, bool mapped, const float (&in)[kIn], float base)
{
    float result = base;
    if (!mapped) return result;
#pragma unroll
    for (int j = 0; j < kIn; j++) {
        const int n = m.n[j];
        if (n) {
            const float x = in[j];
            int seg = 0;
            for (int i = 2; i < n && x > m.x[j][i - 1]; i++) seg = i - 1;
            result += fmaf(m.slope[j][seg], x, m.icpt[j][seg]);
        }
    }
    return result;
}
__device__ __forceinline__ float map_seg_fast(const MapDev &m, int j, float x) { return fmaf(m.slope[j][0], x, m.icpt[j][0]); }

template <bool FASTM>
__device__ __forceinline__ float map_any(const MapDev &m, bool mapped, const float (&in)[kIn], float base)
{
    return FASTM ? map_eval_fast(m, mapped, in, base) : map_eval(m, mapped, in, base);
}
template <bool FASTM>
__device__ __forceinline__ float seg_any(const MapDev &m, int j, float x)
{
    return FASTM ? map_seg_fast(m, j, x) : map_seg(m, j, x);
}

// The settings a dab needs beyond its position and radius_logarithmic (prepare_and_draw_dab): evaluated from the brush
// inputs of the dab's update_states call -- by the compositing side (finalize_raw), lane-parallel, not by the serial chain.
struct DabSettings { float opaque, opaque_mul, hard, rbr, obr, aa; };
template <bool FASTM>
__device__ __forceinline__ DabSettings dab_settings(const EnvParams &P, const float (&in)[kIn])
{
    DabSettings v;
    if (P.dry_structure) {
        v.opaque = P.dyn[0].base + seg_any<FASTM>(P.dyn[0], 0, in[0]);
        v.opaque_mul = P.dyn[1].base + seg_any<FASTM>(P.dyn[1], 0, in[0]);
        v.hard = P.dyn[3].base;
        v.rbr = P.dyn[4].base;
        v.obr = P.dyn[5].base + seg_any<FASTM>(P.dyn[5], 0, in[0]);
        v.aa = P.dyn[8].base;
    } else {
        const int mm = P.mapped_mask;
        v.opaque = map_any<FASTM>(P.dyn[0], mm & 1, in, P.dyn[0].base);
        v.opaque_mul = map_any<FASTM>(P.dyn[1], mm & 2, in, P.dyn[1].base);
        v.hard = map_any<FASTM>(P.dyn[3], mm & 8, in, P.dyn[3].base);
        v.rbr = map_any<FASTM>(P.dyn[4], mm & 16, in, P.dyn[4].base);
        v.obr = map_any<FASTM>(P.dyn[5], mm & 32, in, P.dyn[5].base);
        v.aa = map_any<FASTM>(P.dyn[8], mm & 256, in, P.dyn[8].base);
    }
    return v;
}
// the two settings the chain itself needs (they decide how many random numbers a dab consumes)
template <bool FASTM>
__device__ __forceinline__ void jitter_settings(const EnvParams &P, const float (&in)[kIn], float &obr, float &rbr)
{
    if (P.dry_structure) {
        rbr = P.dyn[4].base;
        obr = P.dyn[5].base + seg_any<FASTM>(P.dyn[5], 0, in[0]);
    } else {
        const int mm = P.mapped_mask;
        rbr = map_any<FASTM>(P.dyn[4], mm & 16, in, P.dyn[4].base);
        obr = map_any<FASTM>(P.dyn[5], mm & 32, in, P.dyn[5].base);
    }
}

// exp_decay (mypaint-brush.c): returns float
__device__ __forceinline__ float exp_decay(float T_const, float t)
{
    if (T_const < 0.001f) return 0.0f;      // (double)T <= 0.001  <=>  T < 0.001f  ((float)0.001 > 0.001)
    return (float)det_exp((double)fastmath::div_exact(-t, T_const));
}
__device__ __forceinline__ float one_minus_decay(float T_const, float t)
{
    return (float)(1.0 - (double)exp_decay(T_const, t));
}

// ------------------------------------------------------------------------------------------
// K1a: stroke expansion
//
// RNG: every Brush is seeded with 1000 (mypaint_brush_new), so ALL canvases walk the same
// rng-double stream u[0], u[1], ... and differ only in how far they have got.  The library
// therefore tabulates, once per env handle, g[i] = rand_gauss over u[i..i+3] (helpers.c) for the
// whole stream an episode can consume; a canvas keeps just its cursor.  Per dab the engine draws
// u[c] (the RANDOM input, unused by the supported mappings), then two gaussians for the position
// jitter (if offset_by_random != 0) and one for the radius jitter (if radius_by_random != 0).
// ------------------------------------------------------------------------------------------
struct Brush {
    float X, Y, P, partial, AR, AX, AY, s1, s2;
    int flags;
    float v_rlog;      // settings_value[radius_logarithmic] of the last update_states call
    float in[kIn];     // brush inputs of the last update_states call (pressure, speed1, speed2)
};

constexpr int kExpandThreads = 32;
#ifndef EXPAND_MIN_BLOCKS
#define EXPAND_MIN_BLOCKS 16
#endif

using fastmath::div_exact;   // == IEEE `/` for the operands of the dab loop, branch-free (fastmath.cuh)

__device__ __forceinline__ float clamp_radius(float r)
{
    // (double)r < 0.2  <=>  r < 0.2f : no float lies in [0.2, (float)0.2)
    if (r < 0.2f) r = 0.2f;
    if (r > 1000.0f) r = 1000.0f;
    return r;
}

// mypaint-brush.c count_dabs_to
__device__ __forceinline__ float count_dabs_to(const EnvParams &P, Brush &S, float base_radius_raw,
                                               float x, float y, float dt)
{
    if (S.AR == 0.0f) S.AR = base_radius_raw;
    S.AR = clamp_radius(S.AR);
    const float base_radius = clamp_radius(base_radius_raw);
    const float xx = x - S.X;
    const float yy = y - S.Y;
    const float dist = det_hypotf(xx, yy);
    const bool init = (S.flags & 2) != 0;
    const float res1 = div_exact(dist, S.AR) * (init ? P.dabs_actual : 0.0f);
    const float res2 = div_exact(dist, base_radius) * (init ? P.dabs_basic : 0.0f);
    const float res3 = dt * (init ? P.dabs_second : 0.0f);
    return res1 + res2 + res3;
}

// mypaint-brush.c update_states_and_setting_values, live state only (the stroke / direction /
// custom / norm_d{x,y}_slow states feed inputs no supported mapping reads).  Of the settings only
// radius_logarithmic is evaluated here: the others are needed per dab, not by the chain (dab_settings).
__device__ __forceinline__ void update_states(const EnvParams &P, Brush &S, float base_radius,
                                              float rlog_base, float step_dx, float step_dy,
                                              float step_dpressure, float step_dtime)
{
    if (step_dtime < 0.0f) step_dtime = 0.001f;
    else if (step_dtime == 0.0f) step_dtime = 0.001f;

    S.X += step_dx;
    S.Y += step_dy;
    S.P += step_dpressure;
    if (S.P <= 0.0f) S.P = 0.0f;

    S.in[0] = S.P * 1.0f;   // * expf(pressure_gain_log = 0)
    S.in[1] = 0.0f;
    S.in[2] = 0.0f;
    if (P.uses_speed1)
        S.in[1] = (float)(det_log((double)(P.sp_gamma[0] + S.s1)) * (double)P.sp_m[0] + (double)P.sp_q[0]);
    if (P.uses_speed2)
        S.in[2] = (float)(det_log((double)(P.sp_gamma[1] + S.s2)) * (double)P.sp_m[1] + (double)P.sp_q[1]);

    const int mm = P.mapped_mask;
    // the size action rewrites the base value of radius_logarithmic every step (mnist.py:134-135)
    if (P.dry_structure) S.v_rlog = rlog_base + map_seg(P.dyn[2], 2, S.in[2]);
    else S.v_rlog = map_eval(P.dyn[2], mm & 4, S.in, rlog_base);

    // slow_tracking_per_dab == 0  =>  fac = 1.0 - 0.0
    {
        const float fac = 1.0f;
        S.AX += (S.X - S.AX) * fac;
        S.AY += (S.Y - S.AY) * fac;
    }
    if (P.uses_speed1 | P.uses_speed2) {
        const float norm_dx = div_exact(div_exact(step_dx, step_dtime), base_radius);
        const float norm_dy = div_exact(div_exact(step_dy, step_dtime), base_radius);
        const float norm_speed = det_hypotf(norm_dx, norm_dy);
        if (P.uses_speed1) {
            const float v_s1slow = map_eval(P.dyn[6], mm & 64, S.in, P.dyn[6].base);
            const float fac = one_minus_decay(v_s1slow, step_dtime);
            S.s1 += (norm_speed - S.s1) * fac;
        }
        if (P.uses_speed2) {
            const float v_s2slow = map_eval(P.dyn[7], mm & 128, S.in, P.dyn[7].base);
            const float fac = one_minus_decay(v_s2slow, step_dtime);
            S.s2 += (norm_speed - S.s2) * fac;
        }
    }
    S.AR = clamp_radius(det_expf(S.v_rlog));
    S.flags |= 2;
}

// A dab as the chain hands it to the compositing side: the raw state of its update_states call.
//   r0 = (ACTUAL_X, ACTUAL_Y, pressure input, settings_value[radius_logarithmic])
//   r1 = (ACTUAL_RADIUS, RNG cursor at the start of the call [int bits], speed1 input, speed2 input)
// finalize_raw turns it into the dab libmypaint would draw (opacity and jitter settings, the three gaussians of the
// stream window, radius jitter, anti-aliasing): ~60 instructions that nothing in the chain depends on, so they run
// lane-parallel over 32 dabs there instead of serially here.  Dropped here (conservatively): dabs that cannot
// touch tile (0,0) whatever their jitter, and -- when the brush maps zero pressure to zero opacity -- dabs of
// pressure-free moves (jump steps).
__device__ __forceinline__ bool raw_dab_may_paint(const EnvParams &P, float ax, float ay, float ar, float in0, float pad)
{
    if (P.zero_pressure_blank && in0 == 0.0f) return false;
    const float R = fmaf(ar, P.cull_factor, pad);
    return ax + R >= 0.0f && ax - R < 64.0f && ay + R >= 0.0f && ay - R < 64.0f;
}

struct StepSetup {
    // decoded action (mnist.py:107-130)
    double x, y, c_x, c_y;
    int jump;
    // curve() locals (mnist.py:173-190)
    double cx, cy, sx, sy, x1, y1, x2, y2, bx, by;
    double entry_p, midpoint_p, prange1, prange2, head, tail, tail_length;
    int head_range, tail_range;
    int n_events;
    bool curved;
    // env state at the start of the step and the decoded tables
    double s_x, s_y, pressure;
    int size_idx, color_idx;
};

__device__ __forceinline__ void curve_point(const StepSetup &st, int i, double &px, double &py)
{
    // mypaint_utils.py:5-11 point_on_curve_1
    const double ratio = (double)i / 100.0;
    const double x3 = st.sx + st.x1 * ratio, y3 = st.sy + st.y1 * ratio;
    const double x4 = st.cx + st.x2 * ratio, y4 = st.cy + st.y2 * ratio;
    const double x5 = x4 - x3, y5 = y4 - y3;
    px = x3 + x5 * ratio;
    py = y3 + y5 * ratio;
}

// action decode (mnist.py:107-135), pre-positioning event and stroke geometry (mnist.py:137-160, 173-190)
__device__ __forceinline__ void decode_step(const EnvParams &P, const int32_t *__restrict__ ac,
                                            const double *__restrict__ es, const int step, StepSetup &st)
{
    const int L = P.L;
    st.x = P.lin[0]; st.y = P.lin[0]; st.c_x = P.lin[0]; st.c_y = P.lin[0];
    double pressure = P.default_pressure;
    st.jump = 0;
    st.size_idx = 8;
    st.color_idx = 0;
    if (P.idx_end >= 0) { const int a = ac[P.idx_end]; st.x = P.lin[a % L]; st.y = P.lin[a / L]; }
    if (P.idx_control >= 0) { const int a = ac[P.idx_control]; st.c_x = P.lin[a % L]; st.c_y = P.lin[a / L]; }
    if (P.idx_pressure >= 0) pressure = P.pressures[ac[P.idx_pressure]];
    if (P.idx_size >= 0) st.size_idx = ac[P.idx_size];
    if (P.idx_jump >= 0) st.jump = ac[P.idx_jump];
    if (P.idx_color >= 0) st.color_idx = ac[P.idx_color];
    else if (P.colorize) st.color_idx = step % P.n_colors;

    const bool have_start = es[ES_HAVE_START] != 0.0;
    double s_x = es[ES_SX], s_y = es[ES_SY];
    const double entry_pressure = es[ES_ENTRY_PRESSURE];
    if (!have_start) { pressure = 0; s_x = 0; s_y = 0; }
    else if (P.idx_jump >= 0 && st.jump) pressure = 0;
    st.curved = !(P.idx_control < 0 || pressure == 0);
    st.sx = s_x; st.sy = s_y;
    st.s_x = s_x; st.s_y = s_y; st.pressure = pressure;
    if (st.curved) {
        // _line_settings (mnist.py:270-278)
        double p1 = entry_pressure;
        const double p2 = (entry_pressure + pressure) / 2;
        const double p3 = pressure;
        if (P.head == 0.0001) p1 = p2;
        st.prange1 = p2 - p1;
        st.prange2 = p3 - p2;
        st.entry_p = p1;
        st.midpoint_p = p2;
        const double mx = (s_x + st.x) / 2.0, my = (s_y + st.y) / 2.0;
        double dx = st.c_x - mx, dy = st.c_y - my;
        double length = sqrt(dx * dx + dy * dy);
        double nx, ny;
        if (length == 0.0) { nx = 0.0; ny = 0.0; } else { nx = dx / length; ny = dy / length; }
        const double l2x = length * 2;
        st.cx = mx + nx * l2x;
        st.cy = my + ny * l2x;
        st.x1 = st.cx - s_x; st.y1 = st.cy - s_y;
        st.x2 = st.x - st.cx; st.y2 = st.y - st.cy;
        st.head = 100 * P.head;
        st.head_range = (int)st.head + 1;
        st.tail = 100 * P.tail;
        st.tail_range = (int)st.tail + 1;
        st.tail_length = 100 - st.tail;
        double px, py;
        curve_point(st, 1, px, py);
        dx = px - s_x; dy = py - s_y;
        length = sqrt(dx * dx + dy * dy);
        if (length == 0.0) { nx = 0.0; ny = 0.0; } else { nx = dx / length; ny = dy / length; }
        st.bx = s_x + nx * 0.25;
        st.by = s_y + ny * 0.25;
        st.n_events = 102;
    } else {
        st.n_events = 2;
    }
}

// stroke_to event `ev` of the step as the env hands it to the brush (mnist.py:137-160, 191-214): position and
// pressure in double (Python floats), duration.  cur_pressure_var tracks curve()'s `pressure` variable.
__device__ __forceinline__ void event_raw(const StepSetup &st, const int ev, double &px, double &py, double &pp,
                                          double &edt, double &cur_pressure_var)
{
    if (ev == 0) {
        px = st.s_x; py = st.s_y; pp = st.pressure; edt = 0.1;
    } else if (!st.curved) {
        px = st.x; py = st.y; pp = st.pressure; edt = 1.0;
    } else if (ev == 1) {
        px = st.bx; py = st.by; pp = st.entry_p; edt = 0.1;
    } else {
        const int i = ev - 1;
        curve_point(st, i, px, py);
        if (i < st.head_range) {
            cur_pressure_var = fabs((double)i / st.head * st.prange1 + st.entry_p);
            pp = cur_pressure_var;
        } else if (i < st.tail_range) {
            pp = st.midpoint_p;
        } else {
            cur_pressure_var = fabs(((double)i - st.tail) / st.tail_length * st.prange2 + st.midpoint_p);
            pp = cur_pressure_var;
        }
        edt = 0.1;
    }
}

// One canvas's env step of the brush state machine (the body of expand_kernel).  STREAM: the fused small-batch step
// (step_small_kernel) runs this in one thread while a second warp of the CTA composites the dabs as they appear;
// prog = {dabs published, done, colour index, started} in shared memory.
constexpr int kPublishEvery = 8;        // STREAM: dabs between two publications of the count (power of two)
template <bool STREAM>
__device__ __forceinline__ void expand_one(const EnvParams &P, const int b, const int32_t *__restrict__ actions,
                                           float *__restrict__ brush_state, double *__restrict__ env_state,
                                           float *__restrict__ dabs, int32_t *__restrict__ dab_count,
                                           int32_t *__restrict__ status, const float *__restrict__ gauss_tab,
                                           volatile int *prog)
{
    float *bs = brush_state + (size_t)b * kBrushFloats;
    double *es = env_state + (size_t)b * kEnvDoubles;
    Brush S;
    S.X = bs[BS_X]; S.Y = bs[BS_Y]; S.P = bs[BS_PRESSURE]; S.partial = bs[BS_PARTIAL_DABS];
    S.AR = bs[BS_ACTUAL_RADIUS]; S.AX = bs[BS_ACTUAL_X]; S.AY = bs[BS_ACTUAL_Y];
    S.s1 = bs[BS_SPEED1_SLOW]; S.s2 = bs[BS_SPEED2_SLOW];
    S.flags = __float_as_int(bs[BS_FLAGS]);
    const int step = __float_as_int(bs[BS_STEP]);
    int cursor = __float_as_int(bs[BS_RNG_POS]);
    const int cursor_max = P.rng_table_len - 16;
    S.v_rlog = 0.0f;
    S.in[0] = S.in[1] = S.in[2] = 0.0f;

    StepSetup st;
    decode_step(P, actions + (size_t)b * P.K, es, step, st);
    const double pressure = st.pressure;
    const int color_idx = st.color_idx | (st.size_idx << 8);  // what the compositing side needs to know about the step
    const float base_radius = P.base_radius[st.size_idx];     // expf(base radius_logarithmic), unclamped
    const float rlog_base = P.base_rlog[st.size_idx];         // base value of radius_logarithmic this step
    const float cull_pad = P.cull_pad[st.size_idx];
    double cur_pressure_var = pressure;     // the `pressure` variable inside curve()

    if (STREAM) {                                        // the consumer needs the colour before the first dab
        prog[2] = color_idx;
        __threadfence_block();
        prog[3] = 1;
    }
    float *my_dabs = dabs + (size_t)b * P.max_dabs * kDabFloats;
    int n_dabs = 0;
    int err = 0;

    // ---- event cursor ----
    int ev = 0;
    int split = 0;                 // stroke_to's "pressure after zero pressure" recursion
    float ex = 0, ey = 0, ep = 0;  // raw event (float32 at the SWIG boundary)
    double edt = 0;
    float tx = 0, ty = 0, tp = 0;  // target after slow tracking
    double dtime_left = 0;
    float dabs_moved = 0, dabs_todo = 0;

    // begin event `ev` (mypaint_brush_stroke_to up to the dab loop); false when the step is over
    auto begin_event = [&]() -> bool {
        for (;;) {
            if (ev >= st.n_events) return false;
            if (split == 2) {
                edt = 0.0001;      // second half of the recursion: same x, y, pressure
                split = 0;
            } else {
                double px, py, pp;
                event_raw(st, ev, px, py, pp, edt, cur_pressure_var);
                ex = (float)px; ey = (float)py; ep = (float)pp;
                if (ep <= 0.0f) ep = 0.0f;
                if (edt > 0.100 && ep != 0.0f && S.P == 0.0f) split = 1;
            }
            float evp = ep;
            double dt = edt;
            if (split == 1) { evp = 0.0f; dt = edt - 0.0001; }
            // slow position tracking: fac = 1 - exp_decay(slow_tracking, 100 * dtime)
            float fac;
            if (dt == 0.1) fac = P.track_fac[0];
            else if (dt == 1.0) fac = P.track_fac[1];
            else fac = one_minus_decay(P.slow_tracking, (float)(100.0 * dt));
            tx = S.X + (ex - S.X) * fac;
            ty = S.Y + (ey - S.Y) * fac;
            tp = evp;
            dabs_moved = S.partial;
            dtime_left = dt;
            if (!(S.flags & 1)) return true;
            // first event after reset: initialise only (mypaint-brush.c reset_requested)
            S.flags &= ~1;
            S.X = tx; S.Y = ty; S.P = tp;
            S.partial = 0; S.AR = 0; S.s1 = 0; S.s2 = 0;
            S.AX = S.X; S.AY = S.Y;
            if (split == 1) split = 2; else ev++;
        }
    };

    bool live = begin_event();
    if (live) dabs_todo = count_dabs_to(P, S, base_radius, tx, ty, (float)dtime_left);

    // ---- one update_states_and_setting_values per iteration (dab, or end of event) ----
    while (live) {
        const int c0 = cursor;
        const bool is_dab = (dabs_moved + dabs_todo >= 1.0f);
        float step_dx, step_dy, step_dp, step_dtime;
        if (is_dab) {
            float step_ddab;
            if (dabs_moved > 0.0f) {
                step_ddab = 1.0f - dabs_moved;           // == (float)(1.0 - (double)dabs_moved): one correctly rounded subtraction
                dabs_moved = 0.0f;
            } else {
                step_ddab = 1.0f;
            }
            const float frac = div_exact(step_ddab, dabs_todo);
            step_dx = frac * (tx - S.X);
            step_dy = frac * (ty - S.Y);
            step_dp = frac * (tp - S.P);
            step_dtime = (float)((double)frac * (dtime_left - 0.0));
        } else {
            step_dx = tx - S.X;
            step_dy = ty - S.Y;
            step_dp = tp - S.P;
            step_dtime = (float)dtime_left;
        }

        update_states(P, S, base_radius, rlog_base, step_dx, step_dy, step_dp, step_dtime);
        cursor += 1;                     // inputs[RANDOM] = rng_double_next(): consumed on every call

        if (is_dab) {
            // prepare_and_draw_dab: the chain only keeps count of the random numbers the dab draws; the dab itself is
            // handed over raw (finalize_raw)
            float v_obr, v_rbr;
            jitter_settings<false>(P, S.in, v_obr, v_rbr);
            if (v_obr != 0.0f) cursor += 8;
            if (v_rbr != 0.0f) cursor += 4;
            if (raw_dab_may_paint(P, S.AX, S.AY, S.AR, S.in[0], cull_pad)) {
                if (n_dabs < P.max_dabs) {
                    SP_CHECK(n_dabs >= 0 && n_dabs < P.max_dabs && b >= 0 && b < P.B);
                    float4 *d = reinterpret_cast<float4 *>(my_dabs + (size_t)n_dabs * kDabFloats);
                    d[0] = make_float4(S.AX, S.AY, S.in[0], S.v_rlog);
                    d[1] = make_float4(S.AR, __int_as_float(c0), S.in[1], S.in[2]);
                    n_dabs++;
                    if (STREAM && (n_dabs & (kPublishEvery - 1)) == 0) {
                        // publish to the compositing warp of this CTA -- every 8th dab only: the fence waits for the
                        // dab stores above, and the consumer works in rounds of 32 anyway
                        __threadfence_block();
                        prog[0] = n_dabs;
                    }
                } else {
                    err |= 1;
                }
            }
            dtime_left -= (double)step_dtime;
        } else {
            S.partial = dabs_moved + dabs_todo;
            if (split == 1) split = 2; else ev++;
            live = begin_event();
        }
        if (live) dabs_todo = count_dabs_to(P, S, base_radius, tx, ty, (float)dtime_left);
    }
    if (cursor > cursor_max) err |= 2;

    // ---- write back ----
    bs[BS_X] = S.X; bs[BS_Y] = S.Y; bs[BS_PRESSURE] = S.P; bs[BS_PARTIAL_DABS] = S.partial;
    bs[BS_ACTUAL_RADIUS] = S.AR; bs[BS_ACTUAL_X] = S.AX; bs[BS_ACTUAL_Y] = S.AY;
    bs[BS_SPEED1_SLOW] = S.s1; bs[BS_SPEED2_SLOW] = S.s2;
    bs[BS_FLAGS] = __int_as_float(S.flags);
    bs[BS_STEP] = __int_as_float(step + 1);
    bs[BS_RNG_POS] = __int_as_float(cursor);
    es[ES_SX] = st.x;
    es[ES_SY] = st.y;
    es[ES_ENTRY_PRESSURE] = st.curved ? cur_pressure_var : pressure;
    es[ES_HAVE_START] = 1.0;
    dab_count[2 * b] = n_dabs;
    dab_count[2 * b + 1] = color_idx;
    if (err) atomicOr(status, err);
    if (STREAM) {
        __threadfence_block();
        prog[0] = n_dabs;                                // the final count (dabs since the last publication included)
        __threadfence_block();
        prog[1] = 1;                                     // done: prog[0] is final
    }
}

// ------------------------------------------------------------------------------------------
// K1a, FAST arithmetic mode (kMathFast).  The same state machine with a shorter serial chain; bit-identical to
// oracle/paint_oracle.c's stroke_to_fast (PO_MATH_FAST), whose header lists the differences from libmypaint's
// arithmetic (all at rounding level):
//   * every operation is binary32 (dtime_left too); exp / log / hypot are fastmath.cuh's polynomials
//   * a stroke_to event moves linearly in time, so its normalised speed is a constant of the event,
//     dist0 / (dtime * base_radius), taken once when the event starts instead of per dab from
//     hypot(step_dx, step_dy) / step_dtime / base_radius (where the dab's fraction cancels)
//   * dist / ACTUAL_RADIUS multiplies by exp(-radius_logarithmic), which shares its range reduction with ACTUAL_RADIUS
//   * exp_decay(T, t) = exp(t * (-1 / T)) with -1 / T hoisted when the slowness setting has no mapping
//   * mapping segments are fmaf(slope, x, intercept)
// The position update keeps libmypaint's own operations: the dab count is discontinuous in the position, and every
// rounding shared with the reference keeps the two trajectories on the same dabs (tools/raster_tolerance.py).
// ------------------------------------------------------------------------------------------
struct BrushF {
    float X, Y, P, partial, AR, iAR, s1, s2;
    int flags;
    float v_rlog;
    float in[kIn];
};

__device__ __forceinline__ float exp_decay_fast(float T_const, float t)
{
    if (T_const < 0.001f) return 0.0f;
    return fastmath::fast_expf(t * div_exact(-1.0f, T_const));
}

template <bool STREAM>
__device__ __forceinline__ void expand_one_fast(const EnvParams &P, const int b, const int32_t *__restrict__ actions,
                                                float *__restrict__ brush_state, double *__restrict__ env_state,
                                                float *__restrict__ dabs, int32_t *__restrict__ dab_count,
                                                int32_t *__restrict__ status, const float *__restrict__ gauss_tab,
                                                volatile int *prog)
{
    using fastmath::fast_exp_pm;
    using fastmath::fast_hypotf;
    using fastmath::fast_logf;
    float *bs = brush_state + (size_t)b * kBrushFloats;
    double *es = env_state + (size_t)b * kEnvDoubles;
    BrushF S;
    S.X = bs[BS_X]; S.Y = bs[BS_Y]; S.P = bs[BS_PRESSURE]; S.partial = bs[BS_PARTIAL_DABS];
    S.AR = bs[BS_ACTUAL_RADIUS]; S.iAR = bs[BS_INV_RADIUS];
    S.s1 = bs[BS_SPEED1_SLOW]; S.s2 = bs[BS_SPEED2_SLOW];
    S.flags = __float_as_int(bs[BS_FLAGS]);
    const int step = __float_as_int(bs[BS_STEP]);
    int cursor = __float_as_int(bs[BS_RNG_POS]);
    const int cursor_max = P.rng_table_len - 16;
    S.v_rlog = 0.0f;
    S.in[0] = S.in[1] = S.in[2] = 0.0f;

    StepSetup st;
    decode_step(P, actions + (size_t)b * P.K, es, step, st);
    const double pressure = st.pressure;
    const int color_idx = st.color_idx | (st.size_idx << 8);
    const float br_raw = P.base_radius[st.size_idx];          // fast_expf(base radius_logarithmic), unclamped
    const float br_c = clamp_radius(br_raw);
    const float rlog_base = P.base_rlog[st.size_idx];
    const float inv01 = P.f_inv_norm01[st.size_idx], inv1 = P.f_inv_norm1[st.size_idx];
    const float inv001 = P.f_inv_norm001[st.size_idx], k_basic = P.f_k_basic[st.size_idx];
    const float cull_pad = P.cull_pad[st.size_idx];
    double cur_pressure_var = pressure;

    if (STREAM) {
        prog[2] = color_idx;
        __threadfence_block();
        prog[3] = 1;
    }
    float *my_dabs = dabs + (size_t)b * P.max_dabs * kDabFloats;
    int n_dabs = 0;
    int err = 0;

    int ev = 0;
    int split = 0;
    float ex = 0, ey = 0, ep = 0;
    double edt = 0;
    float tx = 0, ty = 0, tp = 0;
    float dtl = 0;                    // dtime_left
    float dist = 0;                   // distance to the event's target (count_dabs_to)
    float ev_speed = 0;               // the event's normalised speed
    float dabs_moved = 0, dabs_todo = 0;

    // count_dabs_to (dist measured by the caller)
    auto count_dabs = [&]() {
        if (S.iAR == 0.0f) { S.AR = br_c; S.iAR = div_exact(1.0f, br_c); }     // ACTUAL_RADIUS == 0: first event after the reset
        const bool init = (S.flags & 2) != 0;
        const float k = init ? fmaf(S.iAR, P.dabs_actual, k_basic) : 0.0f;
        dabs_todo = fmaf(dist, k, dtl * (init ? P.dabs_second : 0.0f));
    };
    auto begin_event = [&]() -> bool {
        for (;;) {
            if (ev >= st.n_events) return false;
            if (split == 2) {
                edt = 0.0001;
                split = 0;
            } else {
                double px, py, pp;
                event_raw(st, ev, px, py, pp, edt, cur_pressure_var);
                ex = (float)px; ey = (float)py; ep = (float)pp;
                if (ep <= 0.0f) ep = 0.0f;
                if (edt > 0.100 && ep != 0.0f && S.P == 0.0f) split = 1;
            }
            float evp = ep;
            double dt = edt;
            if (split == 1) { evp = 0.0f; dt = edt - 0.0001; }
            float fac, inv_dtl;
            dtl = (float)dt;
            if (dt == 0.1) { fac = P.track_fac[0]; inv_dtl = inv01; }
            else if (dt == 1.0) { fac = P.track_fac[1]; inv_dtl = inv1; }
            else { fac = 1.0f - exp_decay_fast(P.slow_tracking, (float)(100.0 * dt)); inv_dtl = div_exact(1.0f, dtl * br_raw); }
            tx = S.X + (ex - S.X) * fac;
            ty = S.Y + (ey - S.Y) * fac;
            tp = evp;
            dabs_moved = S.partial;
            if (!(S.flags & 1)) {
                dist = fast_hypotf(tx - S.X, ty - S.Y);
                ev_speed = dist * inv_dtl;
                count_dabs();
                return true;
            }
            S.flags &= ~1;
            S.X = tx; S.Y = ty; S.P = tp;
            S.partial = 0; S.AR = 0; S.iAR = 0; S.s1 = 0; S.s2 = 0;
            if (split == 1) split = 2; else ev++;
        }
    };

    bool live = begin_event();

    const int mm = P.mapped_mask;
    while (live) {
        const int c0 = cursor;
        const bool is_dab = (dabs_moved + dabs_todo >= 1.0f);
        float step_dx, step_dy, step_dp, step_dtime, corner_speed;
        if (is_dab) {
            float step_ddab = 1.0f;
            if (dabs_moved > 0.0f) {
                step_ddab = 1.0f - dabs_moved;
                dabs_moved = 0.0f;
            }
            const float frac = div_exact(step_ddab, dabs_todo);
            step_dx = frac * (tx - S.X);
            step_dy = frac * (ty - S.Y);
            step_dp = frac * (tp - S.P);
            step_dtime = frac * dtl;
            corner_speed = frac * dist * inv001;
        } else {
            step_dx = tx - S.X;
            step_dy = ty - S.Y;
            step_dp = tp - S.P;
            step_dtime = dtl;
            corner_speed = dist * inv001;
        }
        float norm_speed = ev_speed, t_eff = step_dtime;
        if (!(step_dtime > 0.0f)) {                          // step_dtime clamped to 0.001
            norm_speed = corner_speed;
            t_eff = 0.001f;
        }

        // ---- update_states_and_setting_values ----
        S.X += step_dx;
        S.Y += step_dy;
        S.P += step_dp;
        if (S.P <= 0.0f) S.P = 0.0f;
        S.in[0] = S.P;
        S.in[1] = 0.0f;
        S.in[2] = 0.0f;
        if (P.uses_speed1) S.in[1] = fmaf(fast_logf(P.sp_gamma[0] + S.s1), P.sp_m[0], P.sp_q[0]);
        if (P.uses_speed2) S.in[2] = fmaf(fast_logf(P.sp_gamma[1] + S.s2), P.sp_m[1], P.sp_q[1]);
        if (P.dry_structure) S.v_rlog = rlog_base + map_seg_fast(P.dyn[2], 2, S.in[2]);
        else S.v_rlog = map_eval_fast(P.dyn[2], mm & 4, S.in, rlog_base);
        if (P.uses_speed1) {
            float decay;
            if (mm & 64) decay = exp_decay_fast(map_eval_fast(P.dyn[6], true, S.in, P.dyn[6].base), t_eff);
            else decay = P.dyn[6].base < 0.001f ? 0.0f : fastmath::fast_expf(t_eff * P.f_neg_inv_slow[0]);
            S.s1 += (norm_speed - S.s1) * (1.0f - decay);
        }
        if (P.uses_speed2) {
            float decay;
            if (mm & 128) decay = exp_decay_fast(map_eval_fast(P.dyn[7], true, S.in, P.dyn[7].base), t_eff);
            else decay = P.dyn[7].base < 0.001f ? 0.0f : fastmath::fast_expf(t_eff * P.f_neg_inv_slow[1]);
            S.s2 += (norm_speed - S.s2) * (1.0f - decay);
        }
        {
            float inv;
            float ar = fast_exp_pm(S.v_rlog, &inv);
            if (ar < 0.2f) { ar = 0.2f; inv = 5.0f; }
            if (ar > 1000.0f) { ar = 1000.0f; inv = 0.001f; }
            S.AR = ar;
            S.iAR = inv;
        }
        S.flags |= 2;
        cursor += 1;

        if (is_dab) {
            float v_obr, v_rbr;
            jitter_settings<true>(P, S.in, v_obr, v_rbr);
            if (v_obr != 0.0f) cursor += 8;
            if (v_rbr != 0.0f) cursor += 4;
            if (raw_dab_may_paint(P, S.X, S.Y, S.AR, S.in[0], cull_pad)) {
                if (n_dabs < P.max_dabs) {
                    SP_CHECK(n_dabs >= 0 && n_dabs < P.max_dabs && b >= 0 && b < P.B);
                    float4 *d = reinterpret_cast<float4 *>(my_dabs + (size_t)n_dabs * kDabFloats);
                    d[0] = make_float4(S.X, S.Y, S.in[0], S.v_rlog);             // ACTUAL_X/Y == X/Y (slow_tracking_per_dab 0)
                    d[1] = make_float4(S.AR, __int_as_float(c0), S.in[1], S.in[2]);
                    n_dabs++;
                    if (STREAM && (n_dabs & (kPublishEvery - 1)) == 0) {
                        __threadfence_block();
                        prog[0] = n_dabs;
                    }
                } else {
                    err |= 1;
                }
            }
            dtl -= step_dtime;
            dist = fast_hypotf(tx - S.X, ty - S.Y);
            count_dabs();
        } else {
            S.partial = dabs_moved + dabs_todo;
            if (split == 1) split = 2; else ev++;
            live = begin_event();
        }
    }
    if (cursor > cursor_max) err |= 2;

    bs[BS_X] = S.X; bs[BS_Y] = S.Y; bs[BS_PRESSURE] = S.P; bs[BS_PARTIAL_DABS] = S.partial;
    bs[BS_ACTUAL_RADIUS] = S.AR; bs[BS_INV_RADIUS] = S.iAR; bs[BS_ACTUAL_X] = S.X; bs[BS_ACTUAL_Y] = S.Y;
    bs[BS_SPEED1_SLOW] = S.s1; bs[BS_SPEED2_SLOW] = S.s2;
    bs[BS_FLAGS] = __int_as_float(S.flags);
    bs[BS_STEP] = __int_as_float(step + 1);
    bs[BS_RNG_POS] = __int_as_float(cursor);
    es[ES_SX] = st.x;
    es[ES_SY] = st.y;
    es[ES_ENTRY_PRESSURE] = st.curved ? cur_pressure_var : pressure;
    es[ES_HAVE_START] = 1.0;
    dab_count[2 * b] = n_dabs;
    dab_count[2 * b + 1] = color_idx;
    if (err) atomicOr(status, err);
    if (STREAM) {
        __threadfence_block();
        prog[0] = n_dabs;
        __threadfence_block();
        prog[1] = 1;
    }
}

template <int MATH>
__global__ void __launch_bounds__(kExpandThreads, EXPAND_MIN_BLOCKS)
expand_kernel(const __grid_constant__ EnvParams P, const int32_t *__restrict__ actions,
              float *__restrict__ brush_state, double *__restrict__ env_state,
              float *__restrict__ dabs, int32_t *__restrict__ dab_count, int32_t *__restrict__ status,
              const float *__restrict__ gauss_tab, int lanes, const int32_t *__restrict__ perm, int n_solo)
{
    // `lanes` canvases per warp (1..32, host-chosen): the brush state machine is a latency-bound
    // serial chain, so with few canvases it pays to give each its own warp (no intra-warp divergence
    // or load imbalance, more resident warps per SM to hide latency) and leave the other lanes idle.
    const int tid = threadIdx.x;
    // slots are in order of decreasing predicted chain length (perm): the first n_solo canvases -- the ones
    // that decide the kernel's duration -- get a warp each (no divergence against shorter neighbours, lowest
    // per-iteration latency), the rest are packed `lanes` per warp next to canvases of similar length
    int slot;
    if ((int)blockIdx.x < n_solo) {
        if (tid != 0) return;
        slot = blockIdx.x;
    } else {
        if (tid >= lanes) return;
        slot = n_solo + ((int)blockIdx.x - n_solo) * lanes + tid;
    }
    if (slot >= P.B) return;
    // canvases sorted by predicted chain length (plan_kernel / scatter_kernel): the lanes of a warp walk
    // chains of similar length instead of idling behind the longest one
    const int b = perm ? perm[slot] : slot;

    if (MATH == kMathFast)
        expand_one_fast<false>(P, b, actions, brush_state, env_state, dabs, dab_count, status, gauss_tab, nullptr);
    else
        expand_one<false>(P, b, actions, brush_state, env_state, dabs, dab_count, status, gauss_tab, nullptr);
}

// ------------------------------------------------------------------------------------------
// Work-sorted scheduling of expand_kernel.  The state machine's iteration count per canvas-step
// ranges from 2 (nothing to draw) to ~3000 (long curve with an off-canvas excursion); with canvases
// in batch order a warp's lanes idle behind its longest chain (ncu, round 1: 6 of 16 lanes active).
// plan_kernel predicts the count from the action alone (polyline length of the Bezier x dabs per
// pixel), a 256-bucket counting sort (histogram + scatter, order inside a bucket is irrelevant: every
// canvas is computed independently) yields the slot -> canvas permutation expand_kernel reads.
// ------------------------------------------------------------------------------------------
constexpr int kPlanBuckets = 256;

__global__ void __launch_bounds__(256)
plan_kernel(const __grid_constant__ EnvParams P, const int32_t *__restrict__ actions,
            const double *__restrict__ env_state, uint8_t *__restrict__ key_out, int32_t *__restrict__ hist)
{
    __shared__ int s_hist[kPlanBuckets];
    for (int i = threadIdx.x; i < kPlanBuckets; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < P.B) {
        const int32_t *ac = actions + (size_t)b * P.K;
        const double *es = env_state + (size_t)b * kEnvDoubles;
        const int L = P.L;
        float ex = (float)P.lin[0], ey = ex, cx = ex, cy = ex;
        if (P.idx_end >= 0) { const int a = ac[P.idx_end]; ex = (float)P.lin[a % L]; ey = (float)P.lin[a / L]; }
        if (P.idx_control >= 0) { const int a = ac[P.idx_control]; cx = (float)P.lin[a % L]; cy = (float)P.lin[a / L]; }
        float pressure = (float)P.default_pressure;
        if (P.idx_pressure >= 0) pressure = (float)P.pressures[ac[P.idx_pressure]];
        const int size_idx = P.idx_size >= 0 ? ac[P.idx_size] : 8;
        const bool have_start = es[ES_HAVE_START] != 0.0;
        float sx = 0.0f, sy = 0.0f;
        if (have_start) { sx = (float)es[ES_SX]; sy = (float)es[ES_SY]; }
        if (!have_start || (P.idx_jump >= 0 && ac[P.idx_jump])) pressure = 0.0f;
        const bool curved = !(P.idx_control < 0 || pressure == 0.0f);
        float len, events;
        if (curved) {
            const float mx = 0.5f * (sx + ex), my = 0.5f * (sy + ey);
            const float qx = mx + 2.0f * (cx - mx), qy = my + 2.0f * (cy - my);      // mnist.py:179-181
            len = 0.0f;
            float px = sx, py = sy;
#pragma unroll
            for (int i = 1; i <= 8; i++) {
                const float t = (float)i * 0.125f, u = 1.0f - t;
                const float x = u * u * sx + 2.0f * u * t * qx + t * t * ex;
                const float y = u * u * sy + 2.0f * u * t * qy + t * t * ey;
                len += sqrtf((x - px) * (x - px) + (y - py) * (y - py));
                px = x; py = y;
            }
            events = 102.0f;
        } else {
            len = sqrtf((ex - sx) * (ex - sx) + (ey - sy) * (ey - sy));
            events = 2.0f;
        }
        float r = P.base_radius[size_idx];
        r = r < 0.2f ? 0.2f : r;
        const float iters = len * (P.dabs_basic + P.dabs_actual) / r + events;
        int key = (int)(iters * (1.0f / 16.0f));
        key = key > kPlanBuckets - 1 ? kPlanBuckets - 1 : key;
        key = kPlanBuckets - 1 - key;                      // longest chains first
        key_out[b] = (uint8_t)key;
        SP_CHECK(key >= 0 && key < kPlanBuckets);
        atomicAdd(&s_hist[key], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kPlanBuckets; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&hist[i], s_hist[i]);
}

// hist[0..255] counts, hist[256..511] cursors (both zeroed before plan_kernel)
__global__ void __launch_bounds__(256)
scatter_kernel(int B, const uint8_t *__restrict__ key, int32_t *__restrict__ hist, int32_t *__restrict__ perm)
{
    __shared__ int s_start[kPlanBuckets];
    if (threadIdx.x < 32) {
        // exclusive prefix of the 256 counts by one warp (8 per lane)
        int v[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) { v[j] = hist[threadIdx.x * 8 + j]; sum += v[j]; }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)threadIdx.x >= o) incl += t;
        }
        int run = incl - sum;
#pragma unroll
        for (int j = 0; j < 8; j++) { s_start[threadIdx.x * 8 + j] = run; run += v[j]; }
    }
    __syncthreads();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) {
        const int k = key[b];
        const int pos = s_start[k] + atomicAdd(&hist[kPlanBuckets + k], 1);
        SP_CHECK(pos >= 0 && pos < B && k >= 0 && k < kPlanBuckets);
        perm[pos] = b;
    }
}

// ------------------------------------------------------------------------------------------
// K1b: composite
// ------------------------------------------------------------------------------------------
struct FinalDab {
    float x, y, radius, hardness;
    float one_over_r2, r_aa_start, seg1_slope, seg2_offset, seg2_slope;
    int opacity;          // 0 => rejected
    int x0, x1, y0, y1;   // bbox clipped to the tile
};

// prepare_and_draw_dab's second half + draw_dab_internal + render_dab_mask's per-dab constants
__device__ __forceinline__ FinalDab finalize_dab(const float4 a, const float4 b, const bool fast_math)
{
    FinalDab f;
    float x = a.x, y = a.y;
    float radius = a.w;          // ACTUAL_RADIUS
    float opq = b.x;
    if (b.w != 0.0f) {
        radius = clamp_radius(fast_math ? fastmath::fast_expf(a.z) : det_expf(a.z));
        float ac = a.w / radius;
        ac = ac * ac;
        if (ac <= 1.0f) opq *= ac;
    }
    float hardness = b.y;
    hardness = hardness > 1.0f ? 1.0f : (hardness < 0.0f ? 0.0f : hardness);
    const float fadeout = (float)((double)radius * (1.0 - (double)hardness));
    const float min_fadeout = b.z;
    if (fadeout < min_fadeout) {
        const float opt = (float)((double)radius - (1.0 - (double)hardness) * (double)radius / 2.0);
        const float hardness_new = (float)(((double)opt - ((double)min_fadeout / 2.0)) /
                                           ((double)opt + ((double)min_fadeout / 2.0)));
        const float radius_new = (float)((double)min_fadeout / (1.0 - (double)hardness_new));
        hardness = hardness_new;
        radius = radius_new;
    }
    // draw_dab_internal
    opq = opq > 1.0f ? 1.0f : (opq < 0.0f ? 0.0f : opq);
    hardness = hardness > 1.0f ? 1.0f : (hardness < 0.0f ? 0.0f : hardness);
    f.x = x; f.y = y; f.radius = radius; f.hardness = hardness;
    f.opacity = 0;
    f.x0 = 1; f.x1 = 0; f.y0 = 1; f.y1 = 0;
    f.one_over_r2 = 0; f.r_aa_start = 0; f.seg1_slope = 0; f.seg2_offset = 0; f.seg2_slope = 0;
    if (radius < 0.1f || hardness == 0.0f || opq == 0.0f) return f;
    const float r_fringe = radius + 1.0f;
    int x0 = (int)floorf(x - r_fringe), y0 = (int)floorf(y - r_fringe);
    int x1 = (int)floorf(x + r_fringe), y1 = (int)floorf(y + r_fringe);
    if (x0 > kTile - 1 || x1 < 0 || y0 > kTile - 1 || y1 < 0) return f;   // tile (0,0) untouched
    if (x0 < 0) x0 = 0;
    if (y0 < 0) y0 = 0;
    if (x1 > kTile - 1) x1 = kTile - 1;
    if (y1 > kTile - 1) y1 = kTile - 1;
    f.x0 = x0; f.x1 = x1; f.y0 = y0; f.y1 = y1;
    f.opacity = (int)(unsigned short)(1.0f * opq * 32768.0f);
    f.seg1_slope = -(1.0f / hardness - 1.0f);
    f.seg2_offset = hardness / (1.0f - hardness);
    f.seg2_slope = -hardness / (1.0f - hardness);
    f.one_over_r2 = 1.0f / (radius * radius);
    float ras = (radius > 1.0f) ? (radius - 1.0f) : 0.0f;
    ras *= ras / 1.0f;
    f.r_aa_start = ras;
    return f;
}

// prepare_and_draw_dab from the chain's raw record (see raw_dab_may_paint): the dab's settings from the brush inputs of its
// update_states call, opacity, the position jitter and the radius jitter from the window of the gaussian stream the
// dab owns (cursor + 1, + 5, + 9), then finalize_dab.  size_idx: the step's size action (base radius of the jitter).
template <bool FASTM>
__device__ __forceinline__ FinalDab finalize_raw(const EnvParams &P, const float4 r0, const float4 r1, const int size_idx,
                                                 const float *__restrict__ gauss_tab)
{
    const float in[kIn] = {r0.z, r1.z, r1.w};
    const DabSettings v = dab_settings<FASTM>(P, in);
    float opq = v.opaque;
    if (opq < 0.0f) opq = 0.0f;
    opq = opq * v.opaque_mul;
    opq = opq > 1.0f ? 1.0f : (opq < 0.0f ? 0.0f : opq);
    float x = r0.x, y = r0.y;
    const int c0 = min(__float_as_int(r1.y), P.rng_table_len - 16);
    SP_CHECK(size_idx >= 0 && size_idx < 9 && c0 >= 0 && c0 + 12 < P.rng_table_len);
    const float base_radius = P.base_radius[size_idx];
    int g = c0 + 1;
    if (v.obr != 0.0f) {
        float amp = v.obr;
        if (amp < 0.0f) amp = 0.0f;
        x += __ldg(gauss_tab + c0 + 1) * amp * base_radius;
        y += __ldg(gauss_tab + c0 + 5) * amp * base_radius;
        g = c0 + 9;
    }
    float rl = r0.w;
    const bool has_rand = v.rbr != 0.0f;
    if (has_rand) rl += __ldg(gauss_tab + g) * v.rbr;
    return finalize_dab(make_float4(x, y, rl, r1.x), make_float4(opq, v.hard, v.aa, has_rand ? 1.0f : 0.0f), FASTM);
}

// mypaint-tiled-surface.c calculate_rr_antialiased / calculate_rr with aspect_ratio == 1
__device__ __forceinline__ float dab_rr(const EnvParams &P, int xp, int yp, float x, float y,
                                        float one_over_radius2, float r_aa_start, bool aa)
{
    const float cs = P.cs, sn = P.sn;
    if (!aa) {
        const float yy = ((float)yp + 0.5f - y);
        const float xx = ((float)xp + 0.5f - x);
        const float yyr = (yy * cs - xx * sn) * 1.0f;
        const float xxr = yy * sn + xx * cs;
        return (yyr * yyr + xxr * xxr) * one_over_radius2;
    }
    const float pixel_right = x - (float)xp;
    const float pixel_bottom = y - (float)yp;
    const float pixel_center_x = pixel_right - 0.5f;
    const float pixel_center_y = pixel_bottom - 0.5f;
    const float pixel_left = pixel_right - 1.0f;
    const float pixel_top = pixel_bottom - 1.0f;
    float nearest_x, nearest_y, rr_near;
    if (pixel_left < 0 && pixel_right > 0 && pixel_top < 0 && pixel_bottom > 0) {
        nearest_x = 0;
        nearest_y = 0;
        rr_near = 0;
    } else {
        const float ltp_dot = pixel_center_x * cs + pixel_center_y * sn;
        const float t = (P.l2 == 1.0f) ? ltp_dot : ltp_dot / P.l2;
        nearest_x = cs * t;
        nearest_y = sn * t;
        nearest_x = nearest_x < pixel_left ? pixel_left : (nearest_x > pixel_right ? pixel_right : nearest_x);
        nearest_y = nearest_y < pixel_top ? pixel_top : (nearest_y > pixel_bottom ? pixel_bottom : nearest_y);
        const float yyr = (nearest_y * cs - nearest_x * sn) * 1.0f;
        const float xxr = nearest_y * sn + nearest_x * cs;
        const float r_near = (yyr * yyr + xxr * xxr);
        rr_near = r_near * one_over_radius2;
    }
    if (rr_near > 1.0f) return rr_near;
    // sign_point_in_line(pcx, pcy, cs, -sn) = (px - vx) * (-vy) - vx * (py - vy)
    const float center_sign = (pixel_center_x - cs) * (sn) - (cs) * (pixel_center_y - (-sn));
    float farthest_x, farthest_y;
    if (center_sign < 0) {
        farthest_x = nearest_x - sn * P.rad_area_1;
        farthest_y = nearest_y + cs * P.rad_area_1;
    } else {
        farthest_x = nearest_x + sn * P.rad_area_1;
        farthest_y = nearest_y - cs * P.rad_area_1;
    }
    const float yyr = (farthest_y * cs - farthest_x * sn) * 1.0f;
    const float xxr = farthest_y * sn + farthest_x * cs;
    const float r_far = (yyr * yyr + xxr * xxr);
    const float rr_far = r_far * one_over_radius2;
    if (r_far < r_aa_start) return (rr_far + rr_near) * 0.5f;
    float visibilityNear = 1.0f - rr_near;
    const float delta = rr_far - rr_near;
    const float delta2 = 1.0f + delta;
    visibilityNear /= delta2;
    return 1.0f - visibilityNear;
}

// calculate_rr_antialiased for the upright round dab (P.fast_spans: sn == 1.0f, l2 == cs*cs + 1 == 1.0f): the same
// expression tree as dab_rr with the multiplications by sn / l2 (exact identities in IEEE arithmetic: x * 1.0f == x,
// x / 1.0f == x, a - (-1.0f) == a + 1.0f) removed and the uniform product cs * rad_area_1 hoisted.  Bit-identical.
__device__ __forceinline__ float dab_rr_round(float cs, float rad_area_1, float cs_rad_area_1, int xp, int yp, float x, float y,
                                              float one_over_radius2, float r_aa_start)
{
    const float pixel_right = x - (float)xp;
    const float pixel_bottom = y - (float)yp;
    const float pixel_center_x = pixel_right - 0.5f;
    const float pixel_center_y = pixel_bottom - 0.5f;
    const float pixel_left = pixel_right - 1.0f;
    const float pixel_top = pixel_bottom - 1.0f;
    float nearest_x, nearest_y, rr_near;
    if (pixel_left < 0 && pixel_right > 0 && pixel_top < 0 && pixel_bottom > 0) {
        nearest_x = 0;
        nearest_y = 0;
        rr_near = 0;
    } else {
        const float t = pixel_center_x * cs + pixel_center_y;
        nearest_x = cs * t;
        nearest_y = t;
        nearest_x = nearest_x < pixel_left ? pixel_left : (nearest_x > pixel_right ? pixel_right : nearest_x);
        nearest_y = nearest_y < pixel_top ? pixel_top : (nearest_y > pixel_bottom ? pixel_bottom : nearest_y);
        const float yyr = nearest_y * cs - nearest_x;
        const float xxr = nearest_y + nearest_x * cs;
        const float r_near = (yyr * yyr + xxr * xxr);
        rr_near = r_near * one_over_radius2;
    }
    if (rr_near > 1.0f) return rr_near;
    const float center_sign = (pixel_center_x - cs) - cs * (pixel_center_y + 1.0f);
    float farthest_x, farthest_y;
    if (center_sign < 0) {
        farthest_x = nearest_x - rad_area_1;
        farthest_y = nearest_y + cs_rad_area_1;
    } else {
        farthest_x = nearest_x + rad_area_1;
        farthest_y = nearest_y - cs_rad_area_1;
    }
    const float yyr = farthest_y * cs - farthest_x;
    const float xxr = farthest_y + farthest_x * cs;
    const float r_far = (yyr * yyr + xxr * xxr);
    const float rr_far = r_far * one_over_radius2;
    if (r_far < r_aa_start) return (rr_far + rr_near) * 0.5f;
    float visibilityNear = 1.0f - rr_near;
    const float delta = rr_far - rr_near;
    const float delta2 = 1.0f + delta;
    visibilityNear /= delta2;
    return 1.0f - visibilityNear;
}

// Half-width of a candidate span.  The candidate list only has to be a SUPERSET of the pixels with a non-zero mask
// (the disc is inflated by 0.1 % + 0.01 px), so the hardware's approximate square root (error ~2^-22) is enough; the
// list is built twice (count, write) from the same instruction, so both passes agree.
__device__ __forceinline__ float span_sqrt(float h2)
{
    float h;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(h) : "f"(h2));
    return h;
}

// Sparse in-place compositing: ONE WARP per canvas.
//
// A stroke is a serial chain of heavily overlapping dabs (6 + 6 per radius), each touching fewer than
// 32 pixels, so the natural unit is a warp that walks its canvas's dab list in order; throughput
// comes from 32 independent canvases resident per SM, not from splitting a tile among warps (which
// leaves most of them waiting at the barrier: ncu, round 1).  The fix15 tile stays where it lives --
// in HBM/L2, pixels read-modify-written through L1 as one 64-bit RGBA word -- so a step only touches
// the sectors under the stroke and only the rows it dirtied are converted into the observation (the
// observation buffer is persistent: it already holds the previous step's image).  Canvases without
// dabs (jump steps, the first step) cost one 8-byte load.
//
// Candidate pixels: libmypaint walks the whole (r+1) bounding box, but for the round dabs this
// engine draws (ratio 1, angle 90) calculate_rr_antialiased returns rr_near > 1 -- opacity 0 --
// exactly when  dx_near^2 + dyc^2 > r^2  (dx_near: distance in x from the dab centre to the pixel's
// [xp, xp+1] extent, dyc: distance in y to the pixel-centre row).  Only pixels inside that disc
// (radius inflated by 0.1 % + 0.01 px, far above float rounding) are evaluated; skipped pixels are
// exactly pixels the reference blends with opa_ == 0.
//
// Per round of up to 32 dabs:
//   phase 0  lane-parallel: finalise one dab per lane, count its candidate pixels, warp-scan the
//            counts, write (owner dab, pixel) pairs into one DENSE list in shared memory
//   phase 1  the expensive, canvas-independent part -- the anti-aliased coverage rr, the opacity
//            ramp, opa_a = mask * opacity -- for ALL (dab, pixel) pairs of the round, 32 per warp
//            instruction regardless of dab boundaries (a dab alone fills ~17 of 32 lanes)
//   phase 2  the cheap, order-dependent part: dab by dab, blend the pixels whose opa_a != 0
// Dabs the disc test does not cover (radius >= 3, elliptical brushes) take the bounding-box walk.
#ifndef FUSED_STEP_MAX_DEFAULT
#define FUSED_STEP_MAX_DEFAULT (148 * 8 * 2)      // measured (B200): wins up to ~3k canvases, loses from 4k
#endif
#ifndef COMP_WARPS_PER_SM
#define COMP_WARPS_PER_SM 40
#endif
constexpr int kCompWarps = COMP_WARPS_PER_SM > 32 ? 2 : 1;   // canvases (warps) per CTA (at most 32 CTAs per SM)
constexpr int kListCap = 64;                  // candidate pixels per dab on the fast path
constexpr int kCompResident = COMP_WARPS_PER_SM;   // resident warps (= canvases in flight) per SM
constexpr int kDense = kCompResident > 40 ? 512 : (kCompResident > 32 ? 768 : 1024);   // (dab, pixel) pairs per round

struct __align__(16) SharedDab {
    float x, y, hardness, one_over_r2;
    float r_aa_start, seg1_slope, seg2_offset, seg2_slope;
    uint32_t opacity;       // fix15
    uint32_t box;           // x0 | y0 << 8 | x1 << 16 | y1 << 24, clipped to the tile
    uint32_t aa;            // radius < 3: anti-aliased rr
    uint32_t pad;
};

// One canvas's compositing step (the body of composite_kernel): walks the canvas's dab list in order.  STREAM: the
// list is still being written by this CTA's expansion thread (step_small_kernel); prog = {dabs published, done,
// colour index, started} in shared memory.
template <int C, bool SMT, bool STREAM>
__device__ __forceinline__ void composite_one(const EnvParams &P, const int b, uint16_t *__restrict__ canvas,
                                              const float *__restrict__ dabs, const int32_t *__restrict__ dab_count,
                                              const uint16_t *__restrict__ dither, float *__restrict__ obs,
                                              const float *__restrict__ gauss_tab,
                                              SharedDab *sd, uint16_t *sp, uint16_t *so, int *soff, uint2 *stile,
                                              const int lane, volatile int *prog)
{
    int n_dabs = STREAM ? 0 : dab_count[2 * b];
    if (!STREAM && n_dabs == 0) return;              // canvas and observation unchanged
    int color_idx;
    if (STREAM) {                                    // the producer publishes the colour before its first dab
        if (lane == 0) while (prog[3] == 0) __nanosleep(50);
        __syncwarp();
        color_idx = prog[2];
    } else {
        color_idx = dab_count[2 * b + 1];
    }
    bool done = !STREAM;
    const int size_idx = color_idx >> 8;             // packed by the chain: colour | size action << 8
    color_idx &= 0xff;
    const uint32_t col_r = P.colors[color_idx][0], col_g = P.colors[color_idx][1],
                   col_b = P.colors[color_idx][2];
    uint2 *gtile = reinterpret_cast<uint2 *>(canvas + (size_t)b * kTile * kTile * 4);   // RGBA u16x4 per pixel
    uint2 *tile = gtile;
    if (SMT) {
        tile = stile;
        const uint4 *g4 = reinterpret_cast<const uint4 *>(gtile);
        uint4 *t4 = reinterpret_cast<uint4 *>(tile);
#pragma unroll 8
        for (int i = lane; i < kTile * kTile / 2; i += 32) t4[i] = g4[i];
        __syncwarp();
    }
    const float4 *dl = reinterpret_cast<const float4 *>(dabs + (size_t)b * P.max_dabs * kDabFloats);
    int ymin = kTile, ymax = -1;

    // calculate_opa: mask value (fix15) of one pixel of a dab
    auto mask_of = [&](const SharedDab &d, int xp, int yp) -> uint32_t {
        const float rr = dab_rr(P, xp, yp, d.x, d.y, d.one_over_r2, d.r_aa_start, d.aa != 0);
        const float fac = rr <= d.hardness ? d.seg1_slope : d.seg2_slope;
        float opa = rr <= d.hardness ? 1.0f : d.seg2_offset;
        opa += rr * fac;
        if (rr > 1.0f) opa = 0.0f;
        return (uint32_t)(unsigned short)(opa * 32768.0f);
    };
    // draw_dab_pixels_BlendMode_Normal for one pixel
    auto blend = [&](int idx, uint32_t opa_a) {
        SP_CHECK(idx >= 0 && idx < kTile * kTile && opa_a <= (1u << 15));
        const uint2 px = tile[idx];
        const uint32_t opa_b = (1u << 15) - opa_a;
        uint32_t r = px.x & 0xffffu, g = px.x >> 16, bl = px.y & 0xffffu, al = px.y >> 16;
        r = (opa_a * col_r + opa_b * r) / (1u << 15);
        g = (opa_a * col_g + opa_b * g) / (1u << 15);
        bl = (opa_a * col_b + opa_b * bl) / (1u << 15);
        al = opa_a + opa_b * al / (1u << 15);
        tile[idx] = make_uint2(r | (g << 16), bl | (al << 16));
    };

    const float cs_rad_area_1 = P.cs * P.rad_area_1;
    const bool fast_math = P.math_mode == kMathFast;
    int base = 0;
    for (;;) {
        if (STREAM && !done && n_dabs - base < 32) {
            // wait until a full round of dabs is published, or the producer is done (read `done` first: its store
            // follows the last update of the count)
            int d = 0, a = 0;
            if (lane == 0) {
                for (;;) {
                    d = prog[1];
                    a = prog[0];
                    if (d || a - base >= 32) break;
                    __nanosleep(100);
                }
            }
            done = __shfl_sync(0xffffffffu, d, 0) != 0;
            n_dabs = __shfl_sync(0xffffffffu, a, 0);
            __threadfence_block();                   // the dab loads below come after the count they depend on
        }
        if (base >= n_dabs) break;
        // ---- phase 0: finalise, count candidates ----
        const int mine = base + lane;
        FinalDab f;
        f.opacity = 0; f.x0 = 1; f.x1 = 0; f.y0 = 1; f.y1 = 0;
        f.x = f.y = f.radius = f.hardness = f.one_over_r2 = f.r_aa_start = 0;
        f.seg1_slope = f.seg2_offset = f.seg2_slope = 0;
        if (mine < n_dabs) {
            // STREAM: the records were written by this CTA's producer (L2-coherent loads)
            SP_CHECK(mine >= 0 && mine < P.max_dabs);
            const float4 r0 = STREAM ? __ldcg(dl + 2 * mine) : __ldg(dl + 2 * mine);
            const float4 r1 = STREAM ? __ldcg(dl + 2 * mine + 1) : __ldg(dl + 2 * mine + 1);
            f = fast_math ? finalize_raw<true>(P, r0, r1, size_idx, gauss_tab) : finalize_raw<false>(P, r0, r1, size_idx, gauss_tab);
        }
        const bool hit = f.opacity != 0;             // 0 also when the dab misses tile (0,0)
        const float rlim = f.radius * 1.001f + 0.01f;
        const float rlim2 = rlim * rlim;
        int cnt = 0;
        bool generic = false;
        if (hit) {
            if (P.fast_spans && f.radius < 3.0f && f.x1 - f.x0 < 16 && f.y1 - f.y0 < 16) {
                for (int yp = f.y0; yp <= f.y1; yp++) {
                    const float dyc = f.y - ((float)yp + 0.5f);
                    const float h2 = rlim2 - dyc * dyc;
                    if (h2 < 0.0f) continue;
                    const float h = span_sqrt(h2);
                    const int xa = max(f.x0, (int)floorf(f.x - h - 1.0f));
                    const int xb = min(f.x1, (int)floorf(f.x + h));
                    cnt += max(0, xb - xa + 1);
                }
                if (cnt > kListCap) { generic = true; cnt = 0; }
            } else {
                generic = true;
            }
        }
        int incl = cnt;                              // inclusive warp scan of the counts
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        // this round takes dabs [0, m): up to the first bounding-box dab, and while the pairs fit
        const unsigned stop = __ballot_sync(0xffffffffu, generic || incl > kDense);
        const int m = stop ? __ffs(stop) - 1 : 32;

        SharedDab d;
        d.x = f.x; d.y = f.y; d.hardness = f.hardness; d.one_over_r2 = f.one_over_r2;
        d.r_aa_start = f.r_aa_start; d.seg1_slope = f.seg1_slope; d.seg2_offset = f.seg2_offset;
        d.seg2_slope = f.seg2_slope;
        d.opacity = (uint32_t)f.opacity;
        d.box = (uint32_t)f.x0 | ((uint32_t)f.y0 << 8) | ((uint32_t)f.x1 << 16) | ((uint32_t)f.y1 << 24);
        d.aa = f.radius < 3.0f ? 1u : 0u;
        d.pad = 0;

        if (m == 0) {
            // the first dab of the round walks its bounding box (libmypaint's own loop)
            if (lane == 0) {
                sd[0] = d;
                ymin = min(ymin, f.y0);
                ymax = max(ymax, f.y1);
            }
            __syncwarp();
            const SharedDab cur = sd[0];
            const int bx0 = cur.box & 0xff, by0 = (cur.box >> 8) & 0xff;
            const int bx1 = (cur.box >> 16) & 0xff, by1 = cur.box >> 24;
            const int w = bx1 - bx0 + 1;
            const int n = w * (by1 - by0 + 1);
            for (int p = lane; p < n; p += 32) {
                const int ry = p / w;
                const int xp = bx0 + (p - ry * w), yp = by0 + ry;
                const uint32_t opa_ = mask_of(cur, xp, yp);
                if (opa_) blend(yp * kTile + xp, opa_ * cur.opacity / (1u << 15));
            }
            __syncwarp();
            base += 1;
            continue;
        }

        const int off = incl - cnt;
        const int total = __shfl_sync(0xffffffffu, incl, m - 1);
        soff[lane] = lane < m ? off : total;
        if (lane == 0) soff[32] = total;
        if (lane < m && hit) {
            ymin = min(ymin, f.y0);
            ymax = max(ymax, f.y1);
            sd[lane] = d;
            int n = off;
            for (int yp = f.y0; yp <= f.y1; yp++) {
                const float dyc = f.y - ((float)yp + 0.5f);
                const float h2 = rlim2 - dyc * dyc;
                if (h2 < 0.0f) continue;
                const float h = span_sqrt(h2);
                const int xa = max(f.x0, (int)floorf(f.x - h - 1.0f));
                const int xb = min(f.x1, (int)floorf(f.x + h));
                for (int xp = xa; xp <= xb; xp++) {
                    SP_CHECK(n >= 0 && n < kDense && yp - f.y0 >= 0 && yp - f.y0 < 16 && xp - f.x0 >= 0 && xp - f.x0 < 16);
                    sp[n++] = (uint16_t)((lane << 8) | ((yp - f.y0) << 4) | (xp - f.x0));
                }
            }
        }
        __syncwarp();

        // ---- phase 1: coverage and opacity of every (dab, pixel) pair, densely packed ----
        for (int k = lane; k < total; k += 32) {
            SP_CHECK(k >= 0 && k < kDense && total <= kDense);
            const uint32_t e = sp[k];
            SP_CHECK((e >> 8) < 32u);
            const SharedDab &cur = sd[e >> 8];
            const uint32_t box = cur.box;
            const int xp = (int)(box & 0xff) + (int)(e & 15), yp = (int)((box >> 8) & 0xff) + (int)((e >> 4) & 15);
            // pairs only exist on the fast path: upright round dabs with radius < 3 (anti-aliased rr)
            const float rr = dab_rr_round(P.cs, P.rad_area_1, cs_rad_area_1, xp, yp, cur.x, cur.y, cur.one_over_r2, cur.r_aa_start);
            const float fac = rr <= cur.hardness ? cur.seg1_slope : cur.seg2_slope;
            float opa = rr <= cur.hardness ? 1.0f : cur.seg2_offset;
            opa += rr * fac;
            if (rr > 1.0f) opa = 0.0f;
            const uint32_t opa_ = (uint32_t)(unsigned short)(opa * 32768.0f);
            so[k] = (uint16_t)(opa_ * cur.opacity / (1u << 15));
            sp[k] = (uint16_t)(yp * kTile + xp);             // phase 2 only needs the pixel's index in the tile
        }
        __syncwarp();

        // ---- phase 2: ordered blend ----
        for (int j = 0; j < m; j++) {
            const int o0 = soff[j], o1 = soff[j + 1];
            SP_CHECK(o0 >= 0 && o0 <= o1 && o1 <= kDense && o1 - o0 <= kListCap);
            if (o0 == o1) continue;
            // a dab on this path has at most kListCap = 64 candidate pixels: two passes, no loop
            int k = o0 + lane;
            if (k < o1) {
                const uint32_t opa_a = so[k];
                if (opa_a) blend((int)sp[k], opa_a);
            }
            if (o1 - o0 > 32) {                          // warp-uniform, rare (radius > ~2.5)
                k += 32;
                if (k < o1) {
                    const uint32_t opa_a = so[k];
                    if (opa_a) blend((int)sp[k], opa_a);
                }
            }
            __syncwarp();                                // orders this dab's stores before the next dab's loads
        }
        base += m;
    }

    // ---- fused read-back of the dirtied rows (pixops.cpp tile_convert_rgbu16_to_rgbu8, utils.py:3-5) ----
#pragma unroll
    for (int off = 16; off; off >>= 1) {
        ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, off));
        ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, off));
    }
    if (ymax < ymin) return;
    SP_CHECK(ymin >= 0 && ymax < kTile && b >= 0 && b < P.B);
    const uint4 *src = reinterpret_cast<const uint4 *>(tile);        // uint4 = 2 RGBA pixels
    float *ob = obs + (size_t)b * kTile * kTile * C;
#pragma unroll 4
    for (int y = ymin; y <= ymax; y++) {
        const int i = y * (kTile / 2) + lane;                          // pixel-pair index
        const uint4 v = src[i];
        if (SMT) reinterpret_cast<uint4 *>(gtile)[i] = v;            // dirty rows back to the canvas
        const uint32_t R0 = v.x & 0xffffu, G0 = v.x >> 16, B0 = v.y & 0xffffu;
        const uint32_t R1 = v.z & 0xffffu, G1 = v.z >> 16, B1 = v.w & 0xffffu;
        uint32_t add0 = 16384u, add1 = 16384u;
        if (P.dither) {
            const uint32_t nz = __ldg(reinterpret_cast<const uint32_t *>(dither) + i);
            add0 = nz & 0xffffu;
            add1 = nz >> 16;
        }
        const uint32_t r0 = (R0 * 255u + add0) >> 15, r1 = (R1 * 255u + add1) >> 15;
        const uint32_t g0 = (G0 * 255u + add0) >> 15, g1 = (G1 * 255u + add1) >> 15;
        const uint32_t b0 = (B0 * 255u + add0) >> 15, b1 = (B1 * 255u + add1) >> 15;
        if (C == 1) {
            // np.dot(rgb, [0.299, 0.587, 0.114]) in float64, cast to float32
            double a0 = (double)r0 * 0.299;
            a0 = a0 + (double)g0 * 0.587;
            a0 = a0 + (double)b0 * 0.114;
            double a1 = (double)r1 * 0.299;
            a1 = a1 + (double)g1 * 0.587;
            a1 = a1 + (double)b1 * 0.114;
            reinterpret_cast<float2 *>(ob)[i] = make_float2((float)a0, (float)a1);
        } else {
            float2 *o2 = reinterpret_cast<float2 *>(ob + (size_t)i * 6);
            o2[0] = make_float2((float)r0, (float)g0);
            o2[1] = make_float2((float)b0, (float)r1);
            o2[2] = make_float2((float)g1, (float)b1);
        }
    }
}

// SMT (small batches): the warp's whole fix15 tile (32 KB) lives in shared memory for the step.  With few canvases per SM
// nothing hides the latency of the ordered blend's load -> blend -> store chain through L1/L2 (two thirds of the longest
// canvas's time at 1,024 canvases per GPU); from shared memory the chain is ~4x shorter, and the 64 KB in / dirty rows out
// per canvas are noise when the batch is small.  Same arithmetic, bit-identical canvases.
template <int C, bool SMT>
__global__ void __launch_bounds__(32 * kCompWarps, SMT ? 3 : kCompResident / kCompWarps)
composite_kernel(const __grid_constant__ EnvParams P, uint16_t *__restrict__ canvas,
                 const float *__restrict__ dabs, const int32_t *__restrict__ dab_count,
                 const uint16_t *__restrict__ dither, float *__restrict__ obs, const float *__restrict__ gauss_tab,
                 int32_t *__restrict__ work_counter)
{
    extern __shared__ __align__(16) uint8_t s_tiles[];   // SMT: [kCompWarps][64*64] RGBA u16x4
    __shared__ SharedDab s_dab[kCompWarps][32];
    __shared__ uint16_t s_pair[kCompWarps][kDense];     // owner dab << 8 | (yp - y0) << 4 | (xp - x0)
    __shared__ uint16_t s_opa[kCompWarps][kDense];      // opa_a of the pair
    __shared__ int s_off[kCompWarps][33];               // first pair of each dab of the round
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // persistent warps: canvases differ by 1000x in work (0 .. ~1750 dabs), so every warp pulls the next
    // canvas from a global counter until the batch is drained instead of owning a fixed one
    for (;;) {
    int b = 0;
    if (lane == 0) b = atomicAdd(work_counter, 1);
    b = __shfl_sync(0xffffffffu, b, 0);
    if (b >= P.B) return;
    composite_one<C, SMT, false>(P, b, canvas, dabs, dab_count, dither, obs, gauss_tab, s_dab[warp], s_pair[warp], s_opa[warp], s_off[warp],
                                 SMT ? reinterpret_cast<uint2 *>(s_tiles) + warp * (kTile * kTile) : nullptr, lane, nullptr);
    __syncwarp();
    }   // next canvas
}

// ------------------------------------------------------------------------------------------
// Fused env step for small batches: CTA = one canvas, warp 0 (one thread) runs the brush state machine, warp 1
// composites the dabs while they are being produced.  With a few canvases per SM the step is the longest canvas's
// serial chain -- expansion (~1.9 ms) and THEN its compositing (~0.5 ms) as two kernels; here the consumer keeps up
// with the producer (it is ~4x faster per dab), so the step ends a round of dabs after the chain does.  Same device
// functions as the two-kernel path: bit-identical canvases, dab lists and observations.
// ------------------------------------------------------------------------------------------
// SMT: the canvas's tile lives in shared memory for the step (32 KB, 5 CTAs per SM), as in composite_kernel<C, true>: alone
// on its canvas the compositing warp is bound by the latency of the ordered blend's load -> blend -> store chain, which is
// several times shorter from shared memory than through L1/L2, so it keeps up with the chain on the densest canvases.
template <int C, bool SMT, int MATH>
__global__ void __launch_bounds__(64, SMT ? 5 : 8)
step_small_kernel(const __grid_constant__ EnvParams P, const int32_t *__restrict__ actions,
                  float *__restrict__ brush_state, double *__restrict__ env_state, float *__restrict__ dabs,
                  int32_t *__restrict__ dab_count, int32_t *__restrict__ status, const float *__restrict__ gauss_tab,
                  uint16_t *__restrict__ canvas, const uint16_t *__restrict__ dither, float *__restrict__ obs,
                  const int32_t *__restrict__ perm)
{
    __shared__ SharedDab s_dab[32];
    __shared__ uint16_t s_pair[kDense];
    __shared__ uint16_t s_opa[kDense];
    __shared__ int s_off[33];
    __shared__ int s_prog[4];              // dabs published, done, colour index, started
    extern __shared__ __align__(16) uint8_t s_fused_tile[];   // SMT: [64*64] RGBA u16x4
    // more than one wave of CTAs: canvases in order of decreasing predicted chain length (plan_kernel / scatter_kernel),
    // so the longest chains start first and the short ones fill the slots they leave (the mean chain is ~1/5 of the longest)
    const int b = perm ? perm[blockIdx.x] : blockIdx.x;
    if (threadIdx.x < 4) s_prog[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        if (MATH == kMathFast)
            expand_one_fast<true>(P, b, actions, brush_state, env_state, dabs, dab_count, status, gauss_tab, s_prog);
        else
            expand_one<true>(P, b, actions, brush_state, env_state, dabs, dab_count, status, gauss_tab, s_prog);
    } else if (threadIdx.x >= 32) {
        composite_one<C, SMT, true>(P, b, canvas, dabs, dab_count, dither, obs, gauss_tab, s_dab, s_pair, s_opa, s_off,
                                    SMT ? reinterpret_cast<uint2 *>(s_fused_tile) : nullptr, threadIdx.x & 31, s_prog);
    }
}

// ------------------------------------------------------------------------------------------
// reset, debug, reward
// ------------------------------------------------------------------------------------------
__global__ void reset_kernel(const __grid_constant__ EnvParams P, uint16_t *__restrict__ canvas,
                             float *__restrict__ brush_state, double *__restrict__ env_state,
                             float *__restrict__ obs)
{
    const size_t n_words = (size_t)P.B * kTile * kTile * 4 / 8;     // uint4 = 8 u16
    const size_t n_obs = (size_t)P.B * kTile * kTile * P.C / 4;     // float4
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t w = 32768u | (32768u << 16);
    for (size_t i = t; i < n_words; i += stride) reinterpret_cast<uint4 *>(canvas)[i] = make_uint4(w, w, w, w);
    for (size_t i = t; i < n_obs; i += stride)
        reinterpret_cast<float4 *>(obs)[i] = make_float4(255.0f, 255.0f, 255.0f, 255.0f);
    for (size_t i = t; i < (size_t)P.B; i += stride) {
        float *bs = brush_state + i * kBrushFloats;
        for (int j = 0; j < kBrushFloats; j++) bs[j] = 0.0f;
        bs[BS_FLAGS] = __int_as_float(1);
        double *es = env_state + i * kEnvDoubles;
        for (int j = 0; j < kEnvDoubles; j++) es[j] = 0.0;
        for (int j = 0; j < 10; j++) es[ES_RNG0 + j] = P.rng_seed_block[j];
        es[ES_ENTRY_PRESSURE] = P.entry_pressure0;
    }
}

__global__ void debug_dabs_kernel(const __grid_constant__ EnvParams P, const float *__restrict__ dabs,
                                  const int32_t *__restrict__ dab_count, const float *__restrict__ gauss_tab,
                                  float *__restrict__ out, int32_t *__restrict__ out_count)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= P.B) return;
    const float4 *dl = reinterpret_cast<const float4 *>(dabs + (size_t)b * P.max_dabs * kDabFloats);
    float *o = out + (size_t)b * P.max_dabs * 5;
    const int n = dab_count[2 * b];
    int m = 0;
    for (int i = 0; i < n; i++) {
        const int size_idx = dab_count[2 * b + 1] >> 8;
        const FinalDab f = P.math_mode == kMathFast ? finalize_raw<true>(P, dl[2 * i], dl[2 * i + 1], size_idx, gauss_tab)
                                                    : finalize_raw<false>(P, dl[2 * i], dl[2 * i + 1], size_idx, gauss_tab);
        if (f.opacity == 0) continue;
        o[m * 5 + 0] = f.x; o[m * 5 + 1] = f.y; o[m * 5 + 2] = f.radius; o[m * 5 + 3] = f.hardness;
        o[m * 5 + 4] = (float)f.opacity;
        m++;
    }
    out_count[b] = m;
}

// envs/utils.py:7-8 l2 + mnist.py:237-239: one CTA per canvas, float64 accumulation like numpy
__global__ void __launch_bounds__(256)
l2_reward_kernel(const float *__restrict__ obs, const float *__restrict__ target,
                 float *__restrict__ reward, int n)
{
    __shared__ double s_part[8];
    const int b = blockIdx.x;
    const float4 *o = reinterpret_cast<const float4 *>(obs + (size_t)b * n);
    const float4 *t = reinterpret_cast<const float4 *>(target + (size_t)b * n);
    double acc = 0.0;
    for (int i = threadIdx.x; i < n / 4; i += blockDim.x) {
        const float4 a = o[i], c = t[i];
        const double d0 = (double)a.x - (double)c.x, d1 = (double)a.y - (double)c.y;
        const double d2 = (double)a.z - (double)c.z, d3 = (double)a.w - (double)c.w;
        acc += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) s += s_part[i];
        reward[b] = (float)(1.0 - sqrt(s) / (double)n);
    }
}

__global__ void detmath_kernel(const double *__restrict__ x, double *__restrict__ e,
                               double *__restrict__ l, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    e[i] = det_exp(x[i]);
    l[i] = x[i] > 0 ? det_log(x[i]) : 0.0;
}

__global__ void fastmath_kernel(const float *__restrict__ x, float *__restrict__ e, float *__restrict__ en,
                                float *__restrict__ l, float *__restrict__ h, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float neg;
    e[i] = fastmath::fast_exp_pm(x[i], &neg);
    en[i] = neg;
    l[i] = x[i] > 0 ? fastmath::fast_logf(x[i]) : 0.0f;
    h[i] = fastmath::fast_hypotf(x[i], x[n - 1 - i]);
}

// div_exact (fastmath.cuh) against the IEEE division over pseudo-random operand pairs: magnitudes 2^-24 .. 2^24 with random
// mantissas and signs (everything the dab loop divides lies well inside).  mismatches: pairs whose bits differ.
__global__ void divcheck_kernel(unsigned long long seed, long long n, unsigned long long *__restrict__ mismatches)
{
    unsigned long long bad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long z = seed + (unsigned long long)i * 0x9E3779B97F4A7C15ULL;      // splitmix64
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        z ^= z >> 31;
        const unsigned lo = (unsigned)z, hi = (unsigned)(z >> 32);
        const unsigned ea = 127u - 24u + (lo >> 23) % 49u, eb = 127u - 24u + (hi >> 23) % 49u;
        const float a = __uint_as_float((lo & 0x807fffffu) | (ea << 23));
        const float b = __uint_as_float((hi & 0x807fffffu) | (eb << 23));
        if (__float_as_uint(fastmath::div_exact(a, b)) != __float_as_uint(__fdiv_rn(a, b))) bad++;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// ------------------------------------------------------------------------------------------
// host launchers (called from api.cu)
// ------------------------------------------------------------------------------------------
cudaError_t launch_reset(const EnvParams &P, uint16_t *canvas, float *bs, double *es, float *obs,
                         cudaStream_t st)
{
    const int blocks = 148 * 8;
    reset_kernel<<<blocks, 256, 0, st>>>(P, canvas, bs, es, obs);
    return cudaGetLastError();
}
cudaError_t launch_expand(const EnvParams &P, const int32_t *actions, float *bs, double *es,
                          float *dabs, int32_t *dab_count, int32_t *status, const float *gauss_tab,
                          int32_t *plan_scratch, cudaStream_t st)
{
    // aim for >= 4 warps per SM before packing more than one canvas into a warp (tools/sweep_expand_lanes.sh)
    int lanes = P.B / (148 * 4);
    if (lanes < 1) lanes = 1;
    if (lanes > 32) lanes = 32;
    while (lanes & (lanes - 1)) lanes &= lanes - 1;      // round down to a power of two
    if (P.expand_lanes > 0) lanes = P.expand_lanes;
    const int32_t *perm = nullptr;
    int n_solo = 0;
    if (lanes > 1 && plan_scratch && P.expand_sort) {
        // plan_scratch: i32 [512] histogram + cursors | i32 [B] permutation | u8 [B] keys
        int32_t *hist = plan_scratch;
        int32_t *pm = plan_scratch + 2 * kPlanBuckets;
        uint8_t *key = reinterpret_cast<uint8_t *>(pm + P.B);
        cudaMemsetAsync(hist, 0, 2 * kPlanBuckets * sizeof(int32_t), st);
        const int pb = (P.B + 255) / 256;
        plan_kernel<<<pb, 256, 0, st>>>(P, actions, es, key, hist);
        scatter_kernel<<<pb, 256, 0, st>>>(P.B, key, hist, pm);
        perm = pm;
        // measured (B200): 128 solo warps at B = 16384, 256 at B = 65536; more of them exceed the 16 resident
        // 128-register warps per SM and push the packed warps into a second wave
        n_solo = P.B / P.expand_solo_div;
        if (n_solo > 256) n_solo = 256;
    }
    const int blocks = n_solo + (P.B - n_solo + lanes - 1) / lanes;
    if (P.math_mode == kMathFast)
        expand_kernel<kMathFast><<<blocks, kExpandThreads, 0, st>>>(P, actions, bs, es, dabs, dab_count, status, gauss_tab, lanes, perm, n_solo);
    else
        expand_kernel<kMathDet><<<blocks, kExpandThreads, 0, st>>>(P, actions, bs, es, dabs, dab_count, status, gauss_tab, lanes, perm, n_solo);
    return cudaGetLastError();
}
cudaError_t launch_composite(const EnvParams &P, uint16_t *canvas, const float *dabs,
                             const int32_t *dab_count, const uint16_t *dither, float *obs, const float *gauss_tab,
                             int32_t *work_counter, cudaStream_t st)
{
    cudaMemsetAsync(work_counter, 0, sizeof(int32_t), st);
    static bool carveout = false;
    constexpr int kSmtBytes = kCompWarps * kTile * kTile * 8;
    if (!carveout) {      // 32 resident warps need 32 x 5.8 KB of shared memory: ask for the large carve-out
        cudaFuncSetAttribute(composite_kernel<1, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(composite_kernel<3, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(composite_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmtBytes);
        cudaFuncSetAttribute(composite_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmtBytes);
        carveout = true;
    }
    // small batches: tiles in shared memory (SPIRAL_COMPOSITE_SMEM_MAX overrides the largest batch that uses it; 0 = never)
    int smt_max = 2048;
    if (const char *e = getenv("SPIRAL_COMPOSITE_SMEM_MAX")) smt_max = atoi(e);
    if (P.B <= smt_max) {
        int blocks = (P.B + kCompWarps - 1) / kCompWarps;
        if (blocks > 148 * 3) blocks = 148 * 3;
        if (P.C == 1)
            composite_kernel<1, true><<<blocks, 32 * kCompWarps, kSmtBytes, st>>>(P, canvas, dabs, dab_count, dither, obs, gauss_tab, work_counter);
        else
            composite_kernel<3, true><<<blocks, 32 * kCompWarps, kSmtBytes, st>>>(P, canvas, dabs, dab_count, dither, obs, gauss_tab, work_counter);
        return cudaGetLastError();
    }
    int blocks = (P.B + kCompWarps - 1) / kCompWarps;
    if (blocks > 148 * kCompResident / kCompWarps) blocks = 148 * kCompResident / kCompWarps;   // one resident warp per slot
    if (P.C == 1)
        composite_kernel<1, false><<<blocks, 32 * kCompWarps, 0, st>>>(P, canvas, dabs, dab_count, dither, obs, gauss_tab, work_counter);
    else
        composite_kernel<3, false><<<blocks, 32 * kCompWarps, 0, st>>>(P, canvas, dabs, dab_count, dither, obs, gauss_tab, work_counter);
    return cudaGetLastError();
}
// fused small-batch step (SPIRAL_FUSED_STEP_MAX: largest batch that uses it; 0 = never)
constexpr int kFusedWave = 148 * 8;        // CTAs of step_small_kernel resident at once (registers)
constexpr int kFusedWaveSmt = 148 * 5;     // with the 32 KB tile in shared memory (37.5 KB per CTA)
bool fused_step_applies(const EnvParams &P)
{
    // measured (B200, tools/fused_step_probe.py): the fused kernel wins up to ~3k canvases with the det chain and up to ~6k
    // with the fast one (22.1 vs 30.0 ms per 20-step episode at 4,096; 41.9 vs 34.2 at 8,192)
    int mx = P.math_mode == kMathFast ? 148 * 8 * 5 : FUSED_STEP_MAX_DEFAULT;
    if (const char *e = getenv("SPIRAL_FUSED_STEP_MAX")) mx = atoi(e);
    return P.B <= mx;
}
cudaError_t launch_step_small(const EnvParams &P, const int32_t *actions, float *bs, double *es, float *dabs,
                              int32_t *dab_count, int32_t *status, const float *gauss_tab, uint16_t *canvas,
                              const uint16_t *dither, float *obs, int32_t *plan_scratch, cudaStream_t st)
{
    // tile in shared memory while two waves of 5 CTAs per SM hold the batch (SPIRAL_FUSED_SMEM=0: always in place)
    bool smt = P.B <= 2 * kFusedWaveSmt;
    if (const char *e = getenv("SPIRAL_FUSED_SMEM")) smt = smt && atoi(e) != 0;
    const int32_t *perm = nullptr;
    if (P.B > (smt ? kFusedWaveSmt : kFusedWave) && plan_scratch) {
        // plan_scratch: i32 [512] histogram + cursors | i32 [B] permutation | u8 [B] keys  (as in launch_expand)
        int32_t *hist = plan_scratch;
        int32_t *pm = plan_scratch + 2 * kPlanBuckets;
        uint8_t *key = reinterpret_cast<uint8_t *>(pm + P.B);
        cudaMemsetAsync(hist, 0, 2 * kPlanBuckets * sizeof(int32_t), st);
        const int pb = (P.B + 255) / 256;
        plan_kernel<<<pb, 256, 0, st>>>(P, actions, es, key, hist);
        scatter_kernel<<<pb, 256, 0, st>>>(P.B, key, hist, pm);
        perm = pm;
    }
    constexpr int kTileBytes = kTile * kTile * 8;
#define SPIRAL_STEP_SMALL(C_, SMT_, M_, SMEM_) \
    step_small_kernel<C_, SMT_, M_><<<P.B, 64, SMEM_, st>>>(P, actions, bs, es, dabs, dab_count, status, gauss_tab, canvas, dither, obs, perm)
    const bool fast = P.math_mode == kMathFast;
    if (smt) {
        static bool attr = false;
        if (!attr) {
            cudaFuncSetAttribute(step_small_kernel<1, true, kMathDet>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(step_small_kernel<3, true, kMathDet>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(step_small_kernel<1, true, kMathFast>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(step_small_kernel<3, true, kMathFast>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            attr = true;
        }
        if (P.C == 1) { if (fast) SPIRAL_STEP_SMALL(1, true, kMathFast, kTileBytes); else SPIRAL_STEP_SMALL(1, true, kMathDet, kTileBytes); }
        else { if (fast) SPIRAL_STEP_SMALL(3, true, kMathFast, kTileBytes); else SPIRAL_STEP_SMALL(3, true, kMathDet, kTileBytes); }
    } else if (P.C == 1) {
        if (fast) SPIRAL_STEP_SMALL(1, false, kMathFast, 0); else SPIRAL_STEP_SMALL(1, false, kMathDet, 0);
    } else {
        if (fast) SPIRAL_STEP_SMALL(3, false, kMathFast, 0); else SPIRAL_STEP_SMALL(3, false, kMathDet, 0);
    }
#undef SPIRAL_STEP_SMALL
    return cudaGetLastError();
}
cudaError_t launch_debug_dabs(const EnvParams &P, const float *dabs, const int32_t *dab_count, const float *gauss_tab,
                              float *out, int32_t *out_count, cudaStream_t st)
{
    debug_dabs_kernel<<<(P.B + 63) / 64, 64, 0, st>>>(P, dabs, dab_count, gauss_tab, out, out_count);
    return cudaGetLastError();
}
cudaError_t launch_l2_reward(const float *obs, const float *target, float *reward, int batch, int n,
                             cudaStream_t st)
{
    l2_reward_kernel<<<batch, 256, 0, st>>>(obs, target, reward, n);
    return cudaGetLastError();
}
cudaError_t launch_detmath(const double *x, double *e, double *l, int n, cudaStream_t st)
{
    detmath_kernel<<<(n + 255) / 256, 256, 0, st>>>(x, e, l, n);
    return cudaGetLastError();
}

long long bounds_violations_raster()
{
#ifdef SPIRAL_BOUNDS_CHECK
    return spiral_tu_bounds_violations();
#else
    return -1;                                     // not a checked build
#endif
}
cudaError_t launch_divcheck(unsigned long long seed, long long n, unsigned long long *mismatches, cudaStream_t st)
{
    divcheck_kernel<<<148 * 8, 256, 0, st>>>(seed, n, mismatches);
    return cudaGetLastError();
}
cudaError_t launch_fastmath(const float *x, float *e, float *en, float *l, float *h, int n, cudaStream_t st)
{
    fastmath_kernel<<<(n + 255) / 256, 256, 0, st>>>(x, e, en, l, h, n);
    return cudaGetLastError();
}

}  // namespace spiral
