/*
 * oracle/paint_oracle.c -- TEST INFRASTRUCTURE (CPU oracle), NOT product code.
 *
 * Plain-C restatement of the renderer that carpedm20/SPIRAL-tensorflow drives
 * from envs/mnist.py.  The arithmetic lives in two un-vendored third-party
 * libraries pinned by the reference's install.sh:16-33:
 *     libmypaint 1.3.0  (mypaint-brush.c, mypaint-mapping.c, rng-double.c,
 *                        helpers.c, mypaint-tiled-surface.c, brushmodes.c)
 *     mypaint    1.2.1  (lib/tiledsurface.py, lib/surface.py, lib/pixops.cpp,
 *                        lib/fill.cpp, lib/brush.py)
 * Neither is present in /root/reference nor installable here, so this file
 * restates their published algorithms from recall and
 * anchors on the reference's own call sites:
 *     envs/mnist.py:81-105   reset  (new Surface, flood_fill white, new Brush)
 *     envs/mnist.py:107-148  draw   (action decode, pre-positioning event)
 *     envs/mnist.py:150-167  _draw  (straight/jump vs curve, atomic flush)
 *     envs/mnist.py:173-214  curve  (101 stroke_to events, pressure ramps)
 *     envs/mnist.py:216-222  _stroke_to (duration 0.1)
 *     envs/mnist.py:251-260  image/state (scanline strip -> rgbu8 -> gray)
 *     envs/mypaint_utils.py:5-11,99-106,129-160 (curve math)
 *
 * PARITY UNPINNED: the reference ships no tests, golden images or fixtures,
 * and the real libmypaint/mypaint binaries cannot be run here, so parity of
 * this oracle with the upstream binaries cannot be established.  What IS
 * pinned: the analytically derivable known answers of SURVEY.md 8(c), and the
 * glibc rand() stream used by the read-back dither (checked against libc).
 *
 * Math modes:  PO_MATH_LIBM = glibc expf/exp/log/hypotf exactly where the C
 * sources call them (upstream-faithful); PO_MATH_DET = det_math.h replacements
 * (bit-reproducible on the GPU).  Everything else is identical.
 * PO_MATH_FAST = the rasteriser's binary32 arithmetic mode: fast_math.h's
 * polynomial expf/logf/hypotf AND a restructured dab loop (stroke_to_fast below:
 * every operation in binary32, the loop's divisions hoisted or replaced by
 * exp(-x)); it is NOT a restatement of libmypaint's arithmetic but of the CUDA
 * kernel's, bit for bit, and tests/ measures its distance from the libm mode
 * (north-star tolerance: max abs <= 1/255 per channel, mean <= 1e-3).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "det_math.h"
#include "fast_math.h"
#include "paint_oracle.h"

/* ------------------------------------------------------------------ */
/* math dispatch                                                        */
/* ------------------------------------------------------------------ */
static inline float m_expf(int mode, float x)
{
    if (mode == PO_MATH_FAST) return fast_expf(x);
    return mode == PO_MATH_DET ? (float)det_exp((double)x) : expf(x);
}
static inline double m_exp(int mode, double x)
{
    if (mode == PO_MATH_FAST) return (double)fast_expf((float)x);
    return mode == PO_MATH_DET ? det_exp(x) : exp(x);
}
static inline double m_log(int mode, double x)
{
    if (mode == PO_MATH_FAST) return (double)fast_logf((float)x);
    return mode == PO_MATH_DET ? det_log(x) : log(x);
}
static inline float m_hypotf(int mode, float a, float b)
{
    if (mode == PO_MATH_FAST) return fast_hypotf(a, b);
    return mode == PO_MATH_DET ? det_hypotf(a, b) : hypotf(a, b);
}

/* ------------------------------------------------------------------ */
/* rng-double.c  (Knuth ranf_array, libmypaint's reduced parameters)   */
/* ------------------------------------------------------------------ */
#define KK 10
#define LL 7
#define TT 7
#define QUALITY 19
#define mod_sum(x, y) (((x) + (y)) - (int)((x) + (y)))
#define is_odd(s) ((s) & 1)

typedef struct {
    double ran_u[KK];
    double buf[QUALITY + 1];
    int ptr;            /* index into buf of next value; -1 = "started" sentinel */
} RngDouble;

static void rng_get_array(RngDouble *self, double aa[], int n)
{
    int i, j;
    for (j = 0; j < KK; j++) aa[j] = self->ran_u[j];
    for (; j < n; j++) aa[j] = mod_sum(aa[j - KK], aa[j - LL]);
    for (i = 0; i < LL; i++, j++) self->ran_u[i] = mod_sum(aa[j - KK], aa[j - LL]);
    for (; i < KK; i++, j++) self->ran_u[i] = mod_sum(aa[j - KK], self->ran_u[i - LL]);
}

static void rng_set_seed(RngDouble *self, long seed)
{
    int t, s, j;
    double u[KK + KK - 1];
    double ulp = (1.0 / (1L << 30)) / (1L << 22);       /* 2^-52 */
    double ss = 2.0 * ulp * ((seed & 0x3fffffff) + 2);

    for (j = 0; j < KK; j++) {
        u[j] = ss;
        ss += ss;
        if (ss >= 1.0) ss -= 1.0 - 2 * ulp;
    }
    u[1] += ulp;
    for (s = seed & 0x3fffffff, t = TT - 1; t;) {
        for (j = KK - 1; j > 0; j--) { u[j + j] = u[j]; u[j + j - 1] = 0.0; }
        for (j = KK + KK - 2; j >= KK; j--) {
            u[j - (KK - LL)] = mod_sum(u[j - (KK - LL)], u[j]);
            u[j - KK] = mod_sum(u[j - KK], u[j]);
        }
        if (is_odd(s)) {
            for (j = KK; j > 0; j--) u[j] = u[j - 1];
            u[0] = u[KK];
            u[LL] = mod_sum(u[LL], u[KK]);
        }
        if (s) s >>= 1; else t--;
    }
    for (j = 0; j < LL; j++) self->ran_u[j + KK - LL] = u[j];
    for (; j < KK; j++) self->ran_u[j - LL] = u[j];
    for (j = 0; j < 10; j++) rng_get_array(self, u, KK + KK - 1);
    self->ptr = -1;
}

static double rng_next(RngDouble *self)
{
    if (self->ptr >= 0 && self->buf[self->ptr] >= 0) return self->buf[self->ptr++];
    rng_get_array(self, self->buf, QUALITY);
    self->buf[KK] = -1;
    self->ptr = 1;
    return self->buf[0];
}

/* helpers.c: rand_gauss */
static float rand_gauss(RngDouble *rng)
{
    double sum = 0.0;
    sum += rng_next(rng);
    sum += rng_next(rng);
    sum += rng_next(rng);
    sum += rng_next(rng);
    return sum * 1.73205080757 - 3.46410161514;
}

/* ------------------------------------------------------------------ */
/* mypaint-mapping.c                                                    */
/* ------------------------------------------------------------------ */
typedef struct {
    int n;
    float xvalues[PO_MAX_POINTS];
    float yvalues[PO_MAX_POINTS];
} ControlPoints;

typedef struct {
    float base_value;
    int inputs_used;
    ControlPoints points[PO_INPUT_COUNT];
} Mapping;

static float mapping_calculate(const Mapping *self, const float *data)
{
    int j;
    float result = self->base_value;
    if (self->inputs_used == 0) return result;
    for (j = 0; j < PO_INPUT_COUNT; j++) {
        const ControlPoints *p = &self->points[j];
        if (p->n) {
            float x, y;
            x = data[j];
            float x0, y0, x1, y1;
            x0 = p->xvalues[0]; y0 = p->yvalues[0];
            x1 = p->xvalues[1]; y1 = p->yvalues[1];
            int i;
            for (i = 2; i < p->n && x > x1; i++) {
                x0 = x1; y0 = y1;
                x1 = p->xvalues[i]; y1 = p->yvalues[i];
            }
            if (x0 == x1 || y0 == y1) {
                y = y0;
            } else {
                y = (y1 * (x - x0) + y0 * (x1 - x)) / (x1 - x0);
            }
            result += y;
        }
    }
    return result;
}

/* ------------------------------------------------------------------ */
/* tiled surface: ONE 64x64 fix15 RGBA tile (only tile (0,0) is ever    */
/* read back by envs/mnist.py:251-256), op executed immediately -- the  */
/* per-tile operation queue preserves order, so this is equivalent.     */
/* ------------------------------------------------------------------ */
#define TILE 64

struct POSurface {
    uint16_t rgba[TILE * TILE * 4];
    /* dab log (debug / parity of the expansion stage) */
    PODab *dablog;
    int dablog_cap;
    int dablog_n;
    long dabs_total;     /* every prepare_and_draw_dab call                    */
    long dabs_drawn;     /* passed draw_dab's rejects (radius/hardness/opaque) */
    long dabs_on_tile;   /* ... and touch tile (0,0)                           */
};

POSurface *po_surface_new(void)
{
    POSurface *s = (POSurface *)calloc(1, sizeof(POSurface));
    return s;
}
void po_surface_free(POSurface *s)
{
    if (!s) return;
    free(s->dablog);
    free(s);
}
void po_surface_clear(POSurface *s)
{
    memset(s->rgba, 0, sizeof(s->rgba));
    s->dablog_n = 0;
    s->dabs_total = s->dabs_drawn = s->dabs_on_tile = 0;
}
/* tiledsurface.flood_fill(0,0,(255,255,255),(0,0,64,64),0,s) on an empty surface:
 * fill colour is clamped to 1.0 -> fix15_one on all four channels of tile (0,0)
 * (mypaint lib/fill.cpp; envs/mnist.py:90). */
void po_surface_fill_white(POSurface *s)
{
    for (int i = 0; i < TILE * TILE * 4; i++) s->rgba[i] = 1 << 15;
}
uint16_t *po_surface_data(POSurface *s) { return s->rgba; }
void po_surface_enable_dablog(POSurface *s, int cap)
{
    free(s->dablog);
    s->dablog = (PODab *)calloc((size_t)cap, sizeof(PODab));
    s->dablog_cap = cap;
    s->dablog_n = 0;
}
int po_surface_dablog_count(POSurface *s) { return s->dablog_n; }
const PODab *po_surface_dablog(POSurface *s) { return s->dablog; }
void po_surface_dablog_reset(POSurface *s) { s->dablog_n = 0; }
void po_surface_counters(POSurface *s, long *total, long *drawn, long *on_tile)
{
    *total = s->dabs_total; *drawn = s->dabs_drawn; *on_tile = s->dabs_on_tile;
}

#define CLAMPF(x, lo, hi) ((x) < (lo) ? (lo) : ((x) > (hi) ? (hi) : (x)))

static inline float calculate_r_sample(float x, float y, float aspect_ratio, float sn, float cs)
{
    const float yyr = (y * cs - x * sn) * aspect_ratio;
    const float xxr = y * sn + x * cs;
    const float r = (yyr * yyr + xxr * xxr);
    return r;
}

static inline float calculate_rr(int xp, int yp, float x, float y, float aspect_ratio,
                                 float sn, float cs, float one_over_radius2)
{
    const float yy = (yp + 0.5f - y);
    const float xx = (xp + 0.5f - x);
    const float yyr = (yy * cs - xx * sn) * aspect_ratio;
    const float xxr = yy * sn + xx * cs;
    const float rr = (yyr * yyr + xxr * xxr) * one_over_radius2;
    return rr;
}

static inline float sign_point_in_line(float px, float py, float vx, float vy)
{
    return (px - vx) * (-vy) - (vx) * (py - vy);
}

static inline void closest_point_to_line(float lx, float ly, float px, float py, float *ox, float *oy)
{
    const float l2 = lx * lx + ly * ly;
    const float ltp_dot = px * lx + py * ly;
    const float t = ltp_dot / l2;
    *ox = lx * t;
    *oy = ly * t;
}

static inline float calculate_rr_antialiased(int xp, int yp, float x, float y, float aspect_ratio,
                                             float sn, float cs, float one_over_radius2,
                                             float r_aa_start)
{
    float pixel_right = x - (float)xp;
    float pixel_bottom = y - (float)yp;
    float pixel_center_x = pixel_right - 0.5f;
    float pixel_center_y = pixel_bottom - 0.5f;
    float pixel_left = pixel_right - 1.0f;
    float pixel_top = pixel_bottom - 1.0f;

    float nearest_x, nearest_y;
    float farthest_x, farthest_y;
    float r_near, r_far, rr_near, rr_far;
    if (pixel_left < 0 && pixel_right > 0 && pixel_top < 0 && pixel_bottom > 0) {
        nearest_x = 0;
        nearest_y = 0;
        r_near = rr_near = 0;
    } else {
        closest_point_to_line(cs, sn, pixel_center_x, pixel_center_y, &nearest_x, &nearest_y);
        nearest_x = CLAMPF(nearest_x, pixel_left, pixel_right);
        nearest_y = CLAMPF(nearest_y, pixel_top, pixel_bottom);
        r_near = calculate_r_sample(nearest_x, nearest_y, aspect_ratio, sn, cs);
        rr_near = r_near * one_over_radius2;
    }
    if (rr_near > 1.0f) return rr_near;

    float center_sign = sign_point_in_line(pixel_center_x, pixel_center_y, cs, -sn);
    const float rad_area_1 = sqrtf(1.0f / M_PI);
    if (center_sign < 0) {
        farthest_x = nearest_x - sn * rad_area_1;
        farthest_y = nearest_y + cs * rad_area_1;
    } else {
        farthest_x = nearest_x + sn * rad_area_1;
        farthest_y = nearest_y - cs * rad_area_1;
    }
    r_far = calculate_r_sample(farthest_x, farthest_y, aspect_ratio, sn, cs);
    rr_far = r_far * one_over_radius2;
    if (r_far < r_aa_start) return (rr_far + rr_near) * 0.5f;

    float visibilityNear = 1.0f - rr_near;
    float delta = rr_far - rr_near;
    float delta2 = 1.0f + delta;
    visibilityNear /= delta2;
    return 1.0f - visibilityNear;
}

static inline float calculate_opa(float rr, float hardness,
                                  float segment1_offset, float segment1_slope,
                                  float segment2_offset, float segment2_slope)
{
    const float fac = rr <= hardness ? segment1_slope : segment2_slope;
    float opa = rr <= hardness ? segment1_offset : segment2_offset;
    opa += rr * fac;
    if (rr > 1.0f) opa = 0.0f;
    return opa;
}

/* render_dab_mask + draw_dab_pixels_BlendMode_Normal fused per pixel: the RLE
 * mask is only a transport between the two loops, so blending pixel by pixel in
 * the same raster order is equivalent. */
static void process_op_normal(uint16_t *rgba, float x, float y, float radius, float hardness,
                              float aspect_ratio, float angle,
                              uint16_t color_r, uint16_t color_g, uint16_t color_b,
                              uint16_t opacity)
{
    hardness = CLAMPF(hardness, 0.0, 1.0);
    if (aspect_ratio < 1.0) aspect_ratio = 1.0;

    float segment1_offset = 1.0f;
    float segment1_slope = -(1.0f / hardness - 1.0f);
    float segment2_offset = hardness / (1.0f - hardness);
    float segment2_slope = -hardness / (1.0f - hardness);

    float angle_rad = angle / 360 * 2 * M_PI;
    float cs = cos(angle_rad);
    float sn = sin(angle_rad);

    const float r_fringe = radius + 1.0f;
    int x0 = floor(x - r_fringe);
    int y0 = floor(y - r_fringe);
    int x1 = floor(x + r_fringe);
    int y1 = floor(y + r_fringe);
    if (x0 < 0) x0 = 0;
    if (y0 < 0) y0 = 0;
    if (x1 > TILE - 1) x1 = TILE - 1;
    if (y1 > TILE - 1) y1 = TILE - 1;
    const float one_over_radius2 = 1.0f / (radius * radius);

    float r_aa_start = 0.0f;
    const int aa = radius < 3.0f;
    if (aa) {
        const float aa_border = 1.0f;
        r_aa_start = ((radius > aa_border) ? (radius - aa_border) : 0);
        r_aa_start *= r_aa_start / aspect_ratio;
    }
    for (int yp = y0; yp <= y1; yp++) {
        for (int xp = x0; xp <= x1; xp++) {
            float rr;
            if (aa)
                rr = calculate_rr_antialiased(xp, yp, x, y, aspect_ratio, sn, cs,
                                              one_over_radius2, r_aa_start);
            else
                rr = calculate_rr(xp, yp, x, y, aspect_ratio, sn, cs, one_over_radius2);
            const float opa = calculate_opa(rr, hardness, segment1_offset, segment1_slope,
                                            segment2_offset, segment2_slope);
            const uint16_t opa_ = opa * (1 << 15);
            if (!opa_) continue;
            uint16_t *px = rgba + (yp * TILE + xp) * 4;
            uint32_t opa_a = opa_ * (uint32_t)opacity / (1 << 15);
            uint32_t opa_b = (1 << 15) - opa_a;
            px[3] = opa_a + opa_b * px[3] / (1 << 15);
            px[0] = (opa_a * color_r + opa_b * px[0]) / (1 << 15);
            px[1] = (opa_a * color_g + opa_b * px[1]) / (1 << 15);
            px[2] = (opa_a * color_b + opa_b * px[2]) / (1 << 15);
        }
    }
}

/* mypaint-tiled-surface.c draw_dab_internal, restricted to BlendMode_Normal
 * (color_a == 1, lock_alpha == 0, colorize == 0: the only mode the reference's
 * brushes reach). */
static int surface_draw_dab(POSurface *s, float x, float y, float radius,
                            float color_r, float color_g, float color_b,
                            float opaque, float hardness, float color_a,
                            float aspect_ratio, float angle,
                            float lock_alpha, float colorize)
{
    s->dabs_total++;
    opaque = CLAMPF(opaque, 0.0f, 1.0f);
    hardness = CLAMPF(hardness, 0.0f, 1.0f);
    lock_alpha = CLAMPF(lock_alpha, 0.0f, 1.0f);
    colorize = CLAMPF(colorize, 0.0f, 1.0f);
    if (radius < 0.1f) return 0;
    if (hardness == 0.0f) return 0;
    if (opaque == 0.0f) return 0;

    color_r = CLAMPF(color_r, 0.0f, 1.0f);
    color_g = CLAMPF(color_g, 0.0f, 1.0f);
    color_b = CLAMPF(color_b, 0.0f, 1.0f);
    color_a = CLAMPF(color_a, 0.0f, 1.0f);

    uint16_t cr = color_r * (1 << 15);
    uint16_t cg = color_g * (1 << 15);
    uint16_t cb = color_b * (1 << 15);

    float normal = 1.0f;
    normal *= 1.0f - lock_alpha;
    normal *= 1.0f - colorize;
    if (aspect_ratio < 1.0f) aspect_ratio = 1.0f;
    if (color_a != 1.0f || lock_alpha != 0.0f || colorize != 0.0f) {
        fprintf(stderr, "paint_oracle: unsupported blend mode (color_a=%f lock_alpha=%f colorize=%f)\n",
                color_a, lock_alpha, colorize);
        abort();
    }
    s->dabs_drawn++;

    float r_fringe = radius + 1.0f;
    int tx1 = floor(floor(x - r_fringe) / TILE);
    int tx2 = floor(floor(x + r_fringe) / TILE);
    int ty1 = floor(floor(y - r_fringe) / TILE);
    int ty2 = floor(floor(y + r_fringe) / TILE);
    /* only tile (0,0) is ever read back */
    if (tx1 > 0 || tx2 < 0 || ty1 > 0 || ty2 < 0) return 1;
    s->dabs_on_tile++;

    uint16_t opacity = normal * opaque * (1 << 15);
    if (s->dablog && s->dablog_n < s->dablog_cap) {
        PODab *d = &s->dablog[s->dablog_n];
        d->x = x; d->y = y; d->radius = radius; d->hardness = hardness;
        d->opacity = opacity; d->r = cr; d->g = cg; d->b = cb;
    }
    if (s->dablog) s->dablog_n++;   /* may exceed cap: caller detects overflow */
    process_op_normal(s->rgba, x, y, radius, hardness, aspect_ratio, angle, cr, cg, cb, opacity);
    return 1;
}

/* ------------------------------------------------------------------ */
/* read-back: mypaint lib/pixops.cpp tile_convert_rgbu16_to_rgbu8       */
/* ------------------------------------------------------------------ */

/* glibc rand(): TYPE_3 additive feedback generator (r[i] = r[i-3] + r[i-31]),
 * default seed 1.  pixops.cpp fills its dithering table from the UNSEEDED
 * libc rand(), so in a process where nothing else called rand() first the
 * table is this fixed stream.  tests/ checks this against the real libc. */
void po_glibc_rand_stream(unsigned int seed, int n, int32_t *out)
{
    int32_t r[34];
    int i;
    if (seed == 0) seed = 1;
    r[0] = (int32_t)seed;
    for (i = 1; i < 31; i++) {
        long hi = r[i - 1] / 127773;
        long lo = r[i - 1] % 127773;
        long word = 16807 * lo - 2836 * hi;
        if (word < 0) word += 2147483647;
        r[i] = (int32_t)word;
    }
    /* state as a ring of 31 words; discard the first 310 outputs */
    uint32_t ring[31];
    for (i = 0; i < 31; i++) ring[i] = (uint32_t)r[i];
    /* glibc: fptr = &state[3], rptr = &state[0]; *fptr += *rptr */
    int f = 3, b = 0;
    for (i = 0; i < 310 + n; i++) {
        ring[f] += ring[b];
        uint32_t res = ring[f] >> 1;
        if (i >= 310) out[i - 310] = (int32_t)res;
        f = (f + 1) % 31;
        b = (b + 1) % 31;
    }
}

/* dithering_noise[i] = (rand() % (1<<15)) * 240/256 + (1<<15) * 8/256  (int arithmetic) */
void po_dither_table(unsigned int seed, int n, uint16_t *out)
{
    int32_t *v = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    po_glibc_rand_stream(seed, n, v);
    for (int i = 0; i < n; i++) out[i] = (uint16_t)((v[i] % (1 << 15)) * 240 / 256 + (1 << 15) * 8 / 256);
    free(v);
}

/* noise == NULL -> round-half (add = 1<<14), the commented-out alternative in pixops.cpp */
void po_surface_read_rgbu8(const POSurface *s, const uint16_t *noise, uint8_t *out)
{
    int noise_idx = 0;
    const uint16_t *src = s->rgba;
    for (int y = 0; y < TILE; y++) {
        for (int x = 0; x < TILE; x++) {
            uint32_t r = *src++, g = *src++, b = *src++;
            src++;
            const uint32_t add = noise ? noise[noise_idx] : (uint32_t)((1 << 15) / 2);
            noise_idx++;
            *out++ = (r * 255 + add) / (1 << 15);
            *out++ = (g * 255 + add) / (1 << 15);
            *out++ = (b * 255 + add) / (1 << 15);
            *out++ = 255;
        }
    }
}

/* ------------------------------------------------------------------ */
/* mypaint-brush.c                                                      */
/* ------------------------------------------------------------------ */
enum {
    ST_X, ST_Y, ST_PRESSURE, ST_PARTIAL_DABS, ST_ACTUAL_RADIUS,
    ST_SMUDGE_RA, ST_SMUDGE_GA, ST_SMUDGE_BA, ST_SMUDGE_A,
    ST_LAST_GETCOLOR_R, ST_LAST_GETCOLOR_G, ST_LAST_GETCOLOR_B, ST_LAST_GETCOLOR_A,
    ST_LAST_GETCOLOR_RECENTNESS,
    ST_ACTUAL_X, ST_ACTUAL_Y,
    ST_NORM_DX_SLOW, ST_NORM_DY_SLOW, ST_NORM_SPEED1_SLOW, ST_NORM_SPEED2_SLOW,
    ST_STROKE, ST_STROKE_STARTED, ST_CUSTOM_INPUT, ST_RNG_SEED,
    ST_ACTUAL_ELLIPTICAL_DAB_RATIO, ST_ACTUAL_ELLIPTICAL_DAB_ANGLE,
    ST_DIRECTION_DX, ST_DIRECTION_DY, ST_DECLINATION, ST_ASCENSION,
    ST_COUNT
};

#define ACTUAL_RADIUS_MIN 0.2
#define ACTUAL_RADIUS_MAX 1000

struct POBrush {
    int math_mode;
    float states[ST_COUNT];
    RngDouble rng;
    Mapping settings[PO_SETTING_COUNT];
    float settings_value[PO_SETTING_COUNT];
    float speed_mapping_gamma[2];
    float speed_mapping_m[2];
    float speed_mapping_q[2];
    int reset_requested;
    long n_events;
    float inv_actual_radius;      /* PO_MATH_FAST: exp(-radius_logarithmic), kept beside states[ST_ACTUAL_RADIUS] */
};

static const char *const setting_names[PO_SETTING_COUNT] = {
    "opaque", "opaque_multiply", "opaque_linearize", "radius_logarithmic", "hardness",
    "anti_aliasing", "dabs_per_basic_radius", "dabs_per_actual_radius", "dabs_per_second",
    "radius_by_random", "speed1_slowness", "speed2_slowness", "speed1_gamma", "speed2_gamma",
    "offset_by_random", "offset_by_speed", "offset_by_speed_slowness", "slow_tracking",
    "slow_tracking_per_dab", "tracking_noise", "color_h", "color_s", "color_v", "restore_color",
    "change_color_h", "change_color_l", "change_color_hsl_s", "change_color_v",
    "change_color_hsv_s", "smudge", "smudge_length", "smudge_radius_log", "eraser",
    "stroke_threshold", "stroke_duration_logarithmic", "stroke_holdtime", "custom_input",
    "custom_input_slowness", "elliptical_dab_ratio", "elliptical_dab_angle", "direction_filter",
    "lock_alpha", "colorize", "snap_to_pixel", "pressure_gain_log",
};
static const char *const input_names[PO_INPUT_COUNT] = {
    "pressure", "speed1", "speed2", "random", "stroke", "direction",
    "tilt_declination", "tilt_ascension", "custom",
};
/* libmypaint brushsettings.json defaults (recalled) for settings a .myb may omit */
static const float setting_defaults[PO_SETTING_COUNT] = {
    1.0f, 0.0f, 0.9f, 2.0f, 0.8f,
    1.0f, 0.0f, 2.0f, 0.0f,
    0.0f, 0.04f, 0.8f, 4.0f, 4.0f,
    0.0f, 0.0f, 1.0f, 0.0f,
    0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f,
    0.0f, 0.0f, 0.0f, 0.0f,
    0.0f, 0.0f, 0.5f, 0.0f, 0.0f,
    0.0f, 4.0f, 0.0f, 0.0f,
    0.0f, 1.0f, 90.0f, 2.0f,
    0.0f, 0.0f, 0.0f, 0.0f,
};

int po_setting_count(void) { return PO_SETTING_COUNT; }
int po_input_count(void) { return PO_INPUT_COUNT; }
const char *po_setting_name(int i) { return (i >= 0 && i < PO_SETTING_COUNT) ? setting_names[i] : NULL; }
const char *po_input_name(int i) { return (i >= 0 && i < PO_INPUT_COUNT) ? input_names[i] : NULL; }
float po_setting_default(int i) { return setting_defaults[i]; }

enum {
    S_OPAQUE, S_OPAQUE_MULTIPLY, S_OPAQUE_LINEARIZE, S_RADIUS_LOGARITHMIC, S_HARDNESS,
    S_ANTI_ALIASING, S_DABS_PER_BASIC_RADIUS, S_DABS_PER_ACTUAL_RADIUS, S_DABS_PER_SECOND,
    S_RADIUS_BY_RANDOM, S_SPEED1_SLOWNESS, S_SPEED2_SLOWNESS, S_SPEED1_GAMMA, S_SPEED2_GAMMA,
    S_OFFSET_BY_RANDOM, S_OFFSET_BY_SPEED, S_OFFSET_BY_SPEED_SLOWNESS, S_SLOW_TRACKING,
    S_SLOW_TRACKING_PER_DAB, S_TRACKING_NOISE, S_COLOR_H, S_COLOR_S, S_COLOR_V, S_RESTORE_COLOR,
    S_CHANGE_COLOR_H, S_CHANGE_COLOR_L, S_CHANGE_COLOR_HSL_S, S_CHANGE_COLOR_V,
    S_CHANGE_COLOR_HSV_S, S_SMUDGE, S_SMUDGE_LENGTH, S_SMUDGE_RADIUS_LOG, S_ERASER,
    S_STROKE_THRESHOLD, S_STROKE_DURATION_LOGARITHMIC, S_STROKE_HOLDTIME, S_CUSTOM_INPUT,
    S_CUSTOM_INPUT_SLOWNESS, S_ELLIPTICAL_DAB_RATIO, S_ELLIPTICAL_DAB_ANGLE, S_DIRECTION_FILTER,
    S_LOCK_ALPHA, S_COLORIZE, S_SNAP_TO_PIXEL, S_PRESSURE_GAIN_LOG,
};
enum { I_PRESSURE, I_SPEED1, I_SPEED2, I_RANDOM, I_STROKE, I_DIRECTION, I_TILT_DECLINATION,
       I_TILT_ASCENSION, I_CUSTOM };

static float base_value(const POBrush *b, int s) { return b->settings[s].base_value; }

static void settings_base_values_have_changed(POBrush *self)
{
    for (int i = 0; i < 2; i++) {
        float gamma;
        gamma = base_value(self, (i == 0) ? S_SPEED1_GAMMA : S_SPEED2_GAMMA);
        gamma = m_expf(self->math_mode, gamma);
        float fix1_x, fix1_y, fix2_x, fix2_dy;
        fix1_x = 45.0; fix1_y = 0.5; fix2_x = 45.0; fix2_dy = 0.015;
        float m, q;
        float c1;
        c1 = m_log(self->math_mode, fix1_x + gamma);
        m = fix2_dy * (fix2_x + gamma);
        q = fix1_y - m * c1;
        self->speed_mapping_gamma[i] = gamma;
        self->speed_mapping_m[i] = m;
        self->speed_mapping_q[i] = q;
    }
}

POBrush *po_brush_new(int math_mode)
{
    POBrush *self = (POBrush *)calloc(1, sizeof(POBrush));
    self->math_mode = math_mode;
    for (int i = 0; i < PO_SETTING_COUNT; i++) self->settings[i].base_value = setting_defaults[i];
    rng_set_seed(&self->rng, 1000);
    /* settings_value[] is left zero (calloc): upstream mallocs the struct and never
     * initialises it; the first count_dabs_to() after the reset event reads
     * dabs_per_* from it before any update -- zero assumed (fresh heap page). */
    settings_base_values_have_changed(self);
    self->reset_requested = 1;
    return self;
}
void po_brush_free(POBrush *b) { free(b); }

void po_brush_set_base_value(POBrush *self, int id, float value)
{
    self->settings[id].base_value = value;
    settings_base_values_have_changed(self);
}
void po_brush_set_mapping_n(POBrush *self, int id, int input, int n)
{
    Mapping *m = &self->settings[id];
    ControlPoints *p = &m->points[input];
    if (n != 0 && p->n == 0) m->inputs_used++;
    if (n == 0 && p->n != 0) m->inputs_used--;
    p->n = n;
}
void po_brush_set_mapping_point(POBrush *self, int id, int input, int index, float x, float y)
{
    ControlPoints *p = &self->settings[id].points[input];
    p->xvalues[index] = x;
    p->yvalues[index] = y;
}
float po_brush_get_state(const POBrush *b, int i) { return b->states[i]; }
int po_brush_state_count(void) { return ST_COUNT; }
void po_brush_get_rng(const POBrush *b, double *ran_u, double *buf, int *ptr)
{
    memcpy(ran_u, b->rng.ran_u, sizeof(double) * KK);
    memcpy(buf, b->rng.buf, sizeof(double) * (QUALITY + 1));
    *ptr = b->rng.ptr;
}
double po_brush_rng_next(POBrush *b) { return rng_next(&b->rng); }

static float exp_decay(int mode, float T_const, float t)
{
    if (T_const <= 0.001) {
        return 0.0;
    } else {
        return m_exp(mode, -t / T_const);
    }
}

#define SQR(x) ((x) * (x))
#define MIN2(a, b) ((a) < (b) ? (a) : (b))

static void update_states_and_setting_values(POBrush *self, float step_ddab, float step_dx,
                                             float step_dy, float step_dpressure,
                                             float step_declination, float step_ascension,
                                             float step_dtime)
{
    const int mode = self->math_mode;
    float pressure;
    float inputs[PO_INPUT_COUNT];

    if (step_dtime < 0.0) {
        step_dtime = 0.001;
    } else if (step_dtime == 0.0) {
        step_dtime = 0.001;
    }

    self->states[ST_X] += step_dx;
    self->states[ST_Y] += step_dy;
    self->states[ST_PRESSURE] += step_dpressure;
    self->states[ST_DECLINATION] += step_declination;
    self->states[ST_ASCENSION] += step_ascension;

    float base_radius = m_expf(mode, base_value(self, S_RADIUS_LOGARITHMIC));

    if (self->states[ST_PRESSURE] <= 0.0) self->states[ST_PRESSURE] = 0.0;
    pressure = self->states[ST_PRESSURE];

    {
        if (!self->states[ST_STROKE_STARTED]) {
            if (pressure > base_value(self, S_STROKE_THRESHOLD) + 0.0001) {
                self->states[ST_STROKE_STARTED] = 1;
                self->states[ST_STROKE] = 0.0;
            }
        } else {
            if (pressure <= base_value(self, S_STROKE_THRESHOLD) * 0.9 + 0.0001) {
                self->states[ST_STROKE_STARTED] = 0;
            }
        }
    }

    float norm_dx, norm_dy, norm_dist, norm_speed;
    norm_dx = step_dx / step_dtime / base_radius;
    norm_dy = step_dy / step_dtime / base_radius;
    norm_speed = m_hypotf(mode, norm_dx, norm_dy);
    norm_dist = norm_speed * step_dtime;

    inputs[I_PRESSURE] = pressure * m_expf(mode, base_value(self, S_PRESSURE_GAIN_LOG));
    inputs[I_SPEED1] = m_log(mode, self->speed_mapping_gamma[0] + self->states[ST_NORM_SPEED1_SLOW])
                       * self->speed_mapping_m[0] + self->speed_mapping_q[0];
    inputs[I_SPEED2] = m_log(mode, self->speed_mapping_gamma[1] + self->states[ST_NORM_SPEED2_SLOW])
                       * self->speed_mapping_m[1] + self->speed_mapping_q[1];
    inputs[I_RANDOM] = rng_next(&self->rng);
    inputs[I_STROKE] = MIN2(self->states[ST_STROKE], 1.0);
    /* direction/tilt inputs: libm in both modes -- no mapping of the supported
     * brushes reads them, so they cannot influence any output. */
    inputs[I_DIRECTION] = fmodf(atan2f(self->states[ST_DIRECTION_DY], self->states[ST_DIRECTION_DX])
                                / (2 * M_PI) * 360 + 180.0, 180.0);
    inputs[I_TILT_DECLINATION] = self->states[ST_DECLINATION];
    inputs[I_TILT_ASCENSION] = fmodf(self->states[ST_ASCENSION] + 180.0, 360.0) - 180.0;
    inputs[I_CUSTOM] = self->states[ST_CUSTOM_INPUT];

    for (int i = 0; i < PO_SETTING_COUNT; i++) {
        self->settings_value[i] = mapping_calculate(&self->settings[i], inputs);
    }

    {
        float fac = 1.0 - exp_decay(mode, self->settings_value[S_SLOW_TRACKING_PER_DAB], step_ddab);
        self->states[ST_ACTUAL_X] += (self->states[ST_X] - self->states[ST_ACTUAL_X]) * fac;
        self->states[ST_ACTUAL_Y] += (self->states[ST_Y] - self->states[ST_ACTUAL_Y]) * fac;
    }
    {
        float fac;
        fac = 1.0 - exp_decay(mode, self->settings_value[S_SPEED1_SLOWNESS], step_dtime);
        self->states[ST_NORM_SPEED1_SLOW] += (norm_speed - self->states[ST_NORM_SPEED1_SLOW]) * fac;
        fac = 1.0 - exp_decay(mode, self->settings_value[S_SPEED2_SLOWNESS], step_dtime);
        self->states[ST_NORM_SPEED2_SLOW] += (norm_speed - self->states[ST_NORM_SPEED2_SLOW]) * fac;
    }
    {
        float time_constant = m_exp(mode, self->settings_value[S_OFFSET_BY_SPEED_SLOWNESS] * 0.01) - 1.0;
        if (time_constant < 0.002) time_constant = 0.002;
        float fac = 1.0 - exp_decay(mode, time_constant, step_dtime);
        self->states[ST_NORM_DX_SLOW] += (norm_dx - self->states[ST_NORM_DX_SLOW]) * fac;
        self->states[ST_NORM_DY_SLOW] += (norm_dy - self->states[ST_NORM_DY_SLOW]) * fac;
    }
    {
        float dx = step_dx / base_radius;
        float dy = step_dy / base_radius;
        float step_in_dabtime = m_hypotf(mode, dx, dy);
        float fac = 1.0 - exp_decay(mode, m_exp(mode, self->settings_value[S_DIRECTION_FILTER] * 0.5) - 1.0,
                                    step_in_dabtime);
        float dx_old = self->states[ST_DIRECTION_DX];
        float dy_old = self->states[ST_DIRECTION_DY];
        if (SQR(dx_old - dx) + SQR(dy_old - dy) > SQR(dx_old - (-dx)) + SQR(dy_old - (-dy))) {
            dx = -dx;
            dy = -dy;
        }
        self->states[ST_DIRECTION_DX] += (dx - self->states[ST_DIRECTION_DX]) * fac;
        self->states[ST_DIRECTION_DY] += (dy - self->states[ST_DIRECTION_DY]) * fac;
    }
    {
        float fac;
        fac = 1.0 - exp_decay(mode, self->settings_value[S_CUSTOM_INPUT_SLOWNESS], 0.1);
        self->states[ST_CUSTOM_INPUT] += (self->settings_value[S_CUSTOM_INPUT] - self->states[ST_CUSTOM_INPUT]) * fac;
    }
    {
        float frequency;
        float wrap;
        frequency = m_expf(mode, -self->settings_value[S_STROKE_DURATION_LOGARITHMIC]);
        self->states[ST_STROKE] += norm_dist * frequency;
        if (self->states[ST_STROKE] < 0) self->states[ST_STROKE] = 0;
        wrap = 1.0 + self->settings_value[S_STROKE_HOLDTIME];
        if (self->states[ST_STROKE] > wrap) {
            if (wrap > 9.9 + 1.0) {
                self->states[ST_STROKE] = 1.0;
            } else {
                self->states[ST_STROKE] = fmodf(self->states[ST_STROKE], wrap);
                if (self->states[ST_STROKE] < 0) self->states[ST_STROKE] = 0;
            }
        }
    }

    float radius_log;
    radius_log = self->settings_value[S_RADIUS_LOGARITHMIC];
    self->states[ST_ACTUAL_RADIUS] = m_expf(mode, radius_log);
    if (self->states[ST_ACTUAL_RADIUS] < ACTUAL_RADIUS_MIN) self->states[ST_ACTUAL_RADIUS] = ACTUAL_RADIUS_MIN;
    if (self->states[ST_ACTUAL_RADIUS] > ACTUAL_RADIUS_MAX) self->states[ST_ACTUAL_RADIUS] = ACTUAL_RADIUS_MAX;

    self->states[ST_ACTUAL_ELLIPTICAL_DAB_RATIO] = self->settings_value[S_ELLIPTICAL_DAB_RATIO];
    self->states[ST_ACTUAL_ELLIPTICAL_DAB_ANGLE] = self->settings_value[S_ELLIPTICAL_DAB_ANGLE];
}

/* helpers.c hsv_to_rgb_float */
void po_hsv_to_rgb_float(float *h_, float *s_, float *v_)
{
    int i;
    double f, w, q, t;
    float h, s, v;
    float r, g, b;
    r = g = b = 0.0;
    h = *h_; s = *s_; v = *v_;
    h = h - floor(h);
    s = CLAMPF(s, 0.0, 1.0);
    v = CLAMPF(v, 0.0, 1.0);
    double hue;
    if (s == 0.0) {
        r = v; g = v; b = v;
    } else {
        hue = h;
        if (hue == 1.0) hue = 0.0;
        hue *= 6.0;
        i = (int)hue;
        f = hue - i;
        w = v * (1.0 - s);
        q = v * (1.0 - (s * f));
        t = v * (1.0 - (s * (1.0 - f)));
        switch (i) {
        case 0: r = v; g = t; b = w; break;
        case 1: r = q; g = v; b = w; break;
        case 2: r = w; g = v; b = t; break;
        case 3: r = w; g = q; b = v; break;
        case 4: r = t; g = w; b = v; break;
        case 5: r = v; g = w; b = q; break;
        }
    }
    *h_ = r; *s_ = g; *v_ = b;
}

static void unsupported(const char *what)
{
    fprintf(stderr, "paint_oracle: brush feature not restated: %s\n", what);
    abort();
}

static int prepare_and_draw_dab(POBrush *self, POSurface *surface)
{
    const int mode = self->math_mode;
    float x, y, opaque;
    float radius;

    if (self->settings_value[S_OPAQUE] < 0) self->settings_value[S_OPAQUE] = 0;
    opaque = self->settings_value[S_OPAQUE] * self->settings_value[S_OPAQUE_MULTIPLY];
    opaque = CLAMPF(opaque, 0.0, 1.0);
    if (self->settings_value[S_OPAQUE_LINEARIZE]) {
        float alpha, beta, alpha_dab, beta_dab;
        float dabs_per_pixel;
        dabs_per_pixel = (base_value(self, S_DABS_PER_ACTUAL_RADIUS) +
                          base_value(self, S_DABS_PER_BASIC_RADIUS)) * 2.0;
        if (dabs_per_pixel < 1.0) dabs_per_pixel = 1.0;
        dabs_per_pixel = 1.0 + base_value(self, S_OPAQUE_LINEARIZE) * (dabs_per_pixel - 1.0);
        alpha = opaque;
        beta = 1.0 - alpha;
        beta_dab = powf(beta, 1.0 / dabs_per_pixel);
        alpha_dab = 1.0 - beta_dab;
        opaque = alpha_dab;
    }

    x = self->states[ST_ACTUAL_X];
    y = self->states[ST_ACTUAL_Y];

    float base_radius = m_expf(mode, base_value(self, S_RADIUS_LOGARITHMIC));

    if (self->settings_value[S_OFFSET_BY_SPEED]) {
        x += self->states[ST_NORM_DX_SLOW] * self->settings_value[S_OFFSET_BY_SPEED] * 0.1 * base_radius;
        y += self->states[ST_NORM_DY_SLOW] * self->settings_value[S_OFFSET_BY_SPEED] * 0.1 * base_radius;
    }

    if (self->settings_value[S_OFFSET_BY_RANDOM]) {
        float amp = self->settings_value[S_OFFSET_BY_RANDOM];
        if (amp < 0.0) amp = 0.0;
        x += rand_gauss(&self->rng) * amp * base_radius;
        y += rand_gauss(&self->rng) * amp * base_radius;
    }

    radius = self->states[ST_ACTUAL_RADIUS];
    if (self->settings_value[S_RADIUS_BY_RANDOM]) {
        float radius_log, alpha_correction;
        radius_log = self->settings_value[S_RADIUS_LOGARITHMIC];
        radius_log += rand_gauss(&self->rng) * self->settings_value[S_RADIUS_BY_RANDOM];
        radius = m_expf(mode, radius_log);
        radius = CLAMPF(radius, ACTUAL_RADIUS_MIN, ACTUAL_RADIUS_MAX);
        alpha_correction = self->states[ST_ACTUAL_RADIUS] / radius;
        alpha_correction = SQR(alpha_correction);
        if (alpha_correction <= 1.0) {
            opaque *= alpha_correction;
        }
    }

    if (self->settings_value[S_SMUDGE_LENGTH] < 1.0 &&
        (self->settings_value[S_SMUDGE] != 0.0 || self->settings[S_SMUDGE].inputs_used != 0)) {
        unsupported("smudge");
    }

    float color_h = base_value(self, S_COLOR_H);
    float color_s = base_value(self, S_COLOR_S);
    float color_v = base_value(self, S_COLOR_V);
    float eraser_target_alpha = 1.0;
    if (self->settings_value[S_SMUDGE] > 0.0) unsupported("smudge");
    if (self->settings_value[S_ERASER]) unsupported("eraser");

    color_h += self->settings_value[S_CHANGE_COLOR_H];
    color_s += color_s * color_v * self->settings_value[S_CHANGE_COLOR_HSV_S];
    color_v += self->settings_value[S_CHANGE_COLOR_V];
    if (self->settings_value[S_CHANGE_COLOR_L] || self->settings_value[S_CHANGE_COLOR_HSL_S])
        unsupported("change_color_l / change_color_hsl_s");

    float hardness = CLAMPF(self->settings_value[S_HARDNESS], 0.0, 1.0);

    float current_fadeout_in_pixels = radius * (1.0 - hardness);
    float min_fadeout_in_pixels = self->settings_value[S_ANTI_ALIASING];
    if (current_fadeout_in_pixels < min_fadeout_in_pixels) {
        float current_optical_radius = radius - (1.0 - hardness) * radius / 2.0;
        float hardness_new = ((current_optical_radius - (min_fadeout_in_pixels / 2.0)) /
                              (current_optical_radius + (min_fadeout_in_pixels / 2.0)));
        float radius_new = (min_fadeout_in_pixels / (1.0 - hardness_new));
        hardness = hardness_new;
        radius = radius_new;
    }

    float snapToPixel = self->settings_value[S_SNAP_TO_PIXEL];
    if (snapToPixel > 0.0) unsupported("snap_to_pixel");

    po_hsv_to_rgb_float(&color_h, &color_s, &color_v);
    return surface_draw_dab(surface, x, y, radius, color_h, color_s, color_v, opaque, hardness,
                            eraser_target_alpha,
                            self->states[ST_ACTUAL_ELLIPTICAL_DAB_RATIO],
                            self->states[ST_ACTUAL_ELLIPTICAL_DAB_ANGLE],
                            self->settings_value[S_LOCK_ALPHA], self->settings_value[S_COLORIZE]);
}

static float count_dabs_to(POBrush *self, float x, float y, float pressure, float dt)
{
    const int mode = self->math_mode;
    float xx, yy;
    float res1, res2, res3;
    float dist;
    (void)pressure;

    if (self->states[ST_ACTUAL_RADIUS] == 0.0)
        self->states[ST_ACTUAL_RADIUS] = m_expf(mode, base_value(self, S_RADIUS_LOGARITHMIC));
    if (self->states[ST_ACTUAL_RADIUS] < ACTUAL_RADIUS_MIN) self->states[ST_ACTUAL_RADIUS] = ACTUAL_RADIUS_MIN;
    if (self->states[ST_ACTUAL_RADIUS] > ACTUAL_RADIUS_MAX) self->states[ST_ACTUAL_RADIUS] = ACTUAL_RADIUS_MAX;

    float base_radius = m_expf(mode, base_value(self, S_RADIUS_LOGARITHMIC));
    if (base_radius < ACTUAL_RADIUS_MIN) base_radius = ACTUAL_RADIUS_MIN;
    if (base_radius > ACTUAL_RADIUS_MAX) base_radius = ACTUAL_RADIUS_MAX;

    xx = x - self->states[ST_X];
    yy = y - self->states[ST_Y];

    if (self->states[ST_ACTUAL_ELLIPTICAL_DAB_RATIO] > 1.0) {
        unsupported("elliptical dabs");
        dist = 0;
    } else {
        dist = m_hypotf(mode, xx, yy);
    }

    res1 = dist / self->states[ST_ACTUAL_RADIUS] * self->settings_value[S_DABS_PER_ACTUAL_RADIUS];
    res2 = dist / base_radius * self->settings_value[S_DABS_PER_BASIC_RADIUS];
    res3 = dt * self->settings_value[S_DABS_PER_SECOND];
    return res1 + res2 + res3;
}

/* ------------------------------------------------------------------ */
/* PO_MATH_FAST: the CUDA kernel's binary32 dab loop (csrc/raster.cu expand_one_fast), restated.
 * Same state machine as update_states_and_setting_values / count_dabs_to / the loop of
 * mypaint_brush_stroke_to above; every operation is binary32 and the algebra is rearranged so that the
 * serial chain per dab is short.  Differences from libmypaint's arithmetic (all at rounding level):
 *   - dtime_left is a float; X += frac * (x - X) is one fmaf; the event's last update lands exactly on the
 *     target (X = x instead of X += x - X)
 *   - a stroke_to event moves linearly in time, so its normalised speed is a constant of the event:
 *     norm_speed = dist0 * (1 / (dtime * base_radius)) with dist0 the distance measured when the event
 *     starts (libmypaint re-derives it per dab as hypot(step_dx, step_dy) / step_dtime / base_radius, where
 *     step_d{x,y} = frac * (x - X), step_dtime = frac * dtime_left and frac cancels); the remaining
 *     distance follows dist *= 1 - frac instead of a hypot per dab
 *   - dist / ACTUAL_RADIUS multiplies by exp(-radius_logarithmic)
 *   - exp_decay(T, t) = fast_expf(t * (-1 / T));  speed inputs = fmaf(fast_logf(gamma + s), m, q)
 *   - a mapping segment is evaluated as fmaf(slope, x, intercept), slope = (y1 - y0) / (x1 - x0),
 *     intercept = y0 - slope * x0 (libmypaint: (y1 (x - x0) + y0 (x1 - x)) / (x1 - x0))
 *   - dabs_todo = fmaf(dist, fmaf(1/AR, dabs_per_actual_radius, dabs_per_basic_radius / base_radius),
 *                      dtime_left * dabs_per_second)
 * Inputs no supported mapping reads (stroke, direction, tilt, custom) and the states behind them are
 * not evolved; the RANDOM input is still drawn on every call (it moves the RNG stream).
 * ------------------------------------------------------------------ */
static float exp_decay_fast(float T_const, float t)
{
    if (T_const < 0.001f) return 0.0f;
    return fast_expf(t * (-1.0f / T_const));
}

static float clamp_radius_f(float r)
{
    if (r < 0.2f) r = 0.2f;
    if (r > 1000.0f) r = 1000.0f;
    return r;
}

static float mapping_calculate_fast(const Mapping *self, const float *data)
{
    float result = self->base_value;
    if (self->inputs_used == 0) return result;
    for (int j = 0; j < PO_INPUT_COUNT; j++) {
        const ControlPoints *p = &self->points[j];
        if (p->n) {
            const float x = data[j];
            float x0 = p->xvalues[0], y0 = p->yvalues[0];
            float x1 = p->xvalues[1], y1 = p->yvalues[1];
            for (int i = 2; i < p->n && x > x1; i++) {
                x0 = x1; y0 = y1;
                x1 = p->xvalues[i]; y1 = p->yvalues[i];
            }
            float slope = 0.0f, icpt = y0;
            if (!(x0 == x1 || y0 == y1)) {
                slope = (y1 - y0) / (x1 - x0);
                icpt = y0 - slope * x0;
            }
            result += fmaf(slope, x, icpt);
        }
    }
    return result;
}

/* count_dabs_to with the distance handed in (measured at the start of the event, then dist *= 1 - frac) */
static float count_dabs_fast(POBrush *self, float br_raw, float dist, float dtl)
{
    if (self->inv_actual_radius == 0.0f) {                      /* ACTUAL_RADIUS == 0: first event after the reset */
        self->states[ST_ACTUAL_RADIUS] = clamp_radius_f(br_raw);
        self->inv_actual_radius = 1.0f / self->states[ST_ACTUAL_RADIUS];
    }
    const float br_c = clamp_radius_f(br_raw);
    if (self->states[ST_ACTUAL_ELLIPTICAL_DAB_RATIO] > 1.0f) unsupported("elliptical dabs");
    const float k = fmaf(self->inv_actual_radius, self->settings_value[S_DABS_PER_ACTUAL_RADIUS],
                         self->settings_value[S_DABS_PER_BASIC_RADIUS] / br_c);
    return fmaf(dist, k, dtl * self->settings_value[S_DABS_PER_SECOND]);
}

/* update_states_and_setting_values after the position update; norm_speed and step_dtime (> 0) from the caller */
static void update_states_fast(POBrush *self, float norm_speed, float step_dtime)
{
    float inputs[PO_INPUT_COUNT];
    if (self->states[ST_PRESSURE] <= 0.0f) self->states[ST_PRESSURE] = 0.0f;
    for (int i = 0; i < PO_INPUT_COUNT; i++) inputs[i] = 0.0f;
    inputs[I_PRESSURE] = self->states[ST_PRESSURE];                /* pressure_gain_log must be 0 */
    inputs[I_SPEED1] = fmaf(fast_logf(self->speed_mapping_gamma[0] + self->states[ST_NORM_SPEED1_SLOW]),
                            self->speed_mapping_m[0], self->speed_mapping_q[0]);
    inputs[I_SPEED2] = fmaf(fast_logf(self->speed_mapping_gamma[1] + self->states[ST_NORM_SPEED2_SLOW]),
                            self->speed_mapping_m[1], self->speed_mapping_q[1]);
    inputs[I_RANDOM] = (float)rng_next(&self->rng);
    for (int i = 0; i < PO_SETTING_COUNT; i++) {
        for (int j = I_STROKE; j < PO_INPUT_COUNT; j++)
            if (self->settings[i].points[j].n) unsupported("mapping on stroke/direction/tilt/custom input (fast mode)");
        self->settings_value[i] = mapping_calculate_fast(&self->settings[i], inputs);
    }
    if (base_value(self, S_PRESSURE_GAIN_LOG) != 0.0f) unsupported("pressure_gain_log (fast mode)");
    if (self->settings_value[S_SLOW_TRACKING_PER_DAB] != 0.0f) unsupported("slow_tracking_per_dab (fast mode)");
    self->states[ST_ACTUAL_X] = self->states[ST_X];                /* slow_tracking_per_dab == 0: fac = 1 */
    self->states[ST_ACTUAL_Y] = self->states[ST_Y];
    {
        float fac = 1.0f - exp_decay_fast(self->settings_value[S_SPEED1_SLOWNESS], step_dtime);
        self->states[ST_NORM_SPEED1_SLOW] += (norm_speed - self->states[ST_NORM_SPEED1_SLOW]) * fac;
        fac = 1.0f - exp_decay_fast(self->settings_value[S_SPEED2_SLOWNESS], step_dtime);
        self->states[ST_NORM_SPEED2_SLOW] += (norm_speed - self->states[ST_NORM_SPEED2_SLOW]) * fac;
    }
    {
        float inv;
        float ar = fast_exp_pm(self->settings_value[S_RADIUS_LOGARITHMIC], &inv);
        if (ar < 0.2f) { ar = 0.2f; inv = 5.0f; }
        if (ar > 1000.0f) { ar = 1000.0f; inv = 0.001f; }
        self->states[ST_ACTUAL_RADIUS] = ar;
        self->inv_actual_radius = inv;
    }
    self->states[ST_ACTUAL_ELLIPTICAL_DAB_RATIO] = self->settings_value[S_ELLIPTICAL_DAB_RATIO];
    self->states[ST_ACTUAL_ELLIPTICAL_DAB_ANGLE] = self->settings_value[S_ELLIPTICAL_DAB_ANGLE];
}

/* the body of mypaint_brush_stroke_to after its prologue (pressure clamp, dtime, the split recursion) */
static int stroke_to_fast(POBrush *self, POSurface *surface, float x, float y, float pressure, double dtime)
{
    const float br_raw = fast_expf(base_value(self, S_RADIUS_LOGARITHMIC));
    if (base_value(self, S_TRACKING_NOISE)) unsupported("tracking_noise (fast mode)");
    {
        const float fac = 1.0f - exp_decay_fast(base_value(self, S_SLOW_TRACKING), (float)(100.0 * dtime));
        x = self->states[ST_X] + (x - self->states[ST_X]) * fac;
        y = self->states[ST_Y] + (y - self->states[ST_Y]) * fac;
    }
    float dabs_moved = self->states[ST_PARTIAL_DABS];
    if (dtime > 5 || self->reset_requested) {
        self->reset_requested = 0;
        for (int i = 0; i < ST_COUNT; i++) self->states[i] = 0;
        self->inv_actual_radius = 0.0f;
        self->states[ST_X] = x;
        self->states[ST_Y] = y;
        self->states[ST_PRESSURE] = pressure;
        self->states[ST_ACTUAL_X] = self->states[ST_X];
        self->states[ST_ACTUAL_Y] = self->states[ST_Y];
        self->states[ST_STROKE] = 1.0;
        return 1;
    }
    float dtl = (float)dtime;
    float dist = fast_hypotf(x - self->states[ST_X], y - self->states[ST_Y]);
    const float norm_speed = dist * (1.0f / (dtl * br_raw));       /* constant over the event */
    const float inv001 = 1.0f / (0.001f * br_raw);
    float dabs_todo = count_dabs_fast(self, br_raw, dist, dtl);
    while (dabs_moved + dabs_todo >= 1.0f) {
        float step_ddab = 1.0f;
        if (dabs_moved > 0.0f) {
            step_ddab = 1.0f - dabs_moved;
            dabs_moved = 0.0f;
        }
        const float frac = step_ddab / dabs_todo;
        const float step_dtime = frac * dtl;
        /* the position update keeps libmypaint's own operations (multiply, then add): the dab count is discontinuous
         * in the position, so every rounding it shares with the reference keeps the two trajectories on the same dabs */
        self->states[ST_X] += frac * (x - self->states[ST_X]);
        self->states[ST_Y] += frac * (y - self->states[ST_Y]);
        self->states[ST_PRESSURE] += frac * (pressure - self->states[ST_PRESSURE]);
        if (step_dtime > 0.0f) update_states_fast(self, norm_speed, step_dtime);
        else update_states_fast(self, frac * dist * inv001, 0.001f);       /* step_dtime clamped to 0.001 */
        prepare_and_draw_dab(self, surface);
        dtl -= step_dtime;
        dist = fast_hypotf(x - self->states[ST_X], y - self->states[ST_Y]);
        dabs_todo = count_dabs_fast(self, br_raw, dist, dtl);
    }
    self->states[ST_X] += x - self->states[ST_X];
    self->states[ST_Y] += y - self->states[ST_Y];
    self->states[ST_PRESSURE] += pressure - self->states[ST_PRESSURE];
    if (dtl > 0.0f) update_states_fast(self, norm_speed, dtl);
    else update_states_fast(self, dist * inv001, 0.001f);
    self->states[ST_PARTIAL_DABS] = dabs_moved + dabs_todo;
    return 0;
}

int po_brush_stroke_to(POBrush *self, POSurface *surface, float x, float y, float pressure,
                       float xtilt, float ytilt, double dtime)
{
    const int mode = self->math_mode;
    float tilt_ascension = 0.0;
    float tilt_declination = 90.0;
    if (xtilt != 0 || ytilt != 0) unsupported("tilt");
    self->n_events++;

    if (pressure <= 0.0) pressure = 0.0;
    if (!isfinite(x) || !isfinite(y) || (x > 1e10 || y > 1e10 || x < -1e10 || y < -1e10)) {
        x = 0.0; y = 0.0; pressure = 0.0;
    }
    if (dtime <= 0) dtime = 0.0001;

    if (dtime > 0.100 && pressure && self->states[ST_PRESSURE] == 0) {
        po_brush_stroke_to(self, surface, x, y, 0.0, 0.0, 0.0, dtime - 0.0001);
        dtime = 0.0001;
    }
    if (mode == PO_MATH_FAST) return stroke_to_fast(self, surface, x, y, pressure, dtime);

    {
        if (base_value(self, S_TRACKING_NOISE)) {
            const float base_radius = m_expf(mode, base_value(self, S_RADIUS_LOGARITHMIC));
            x += rand_gauss(&self->rng) * base_value(self, S_TRACKING_NOISE) * base_radius;
            y += rand_gauss(&self->rng) * base_value(self, S_TRACKING_NOISE) * base_radius;
        }
        const float fac = 1.0 - exp_decay(mode, base_value(self, S_SLOW_TRACKING), 100.0 * dtime);
        x = self->states[ST_X] + (x - self->states[ST_X]) * fac;
        y = self->states[ST_Y] + (y - self->states[ST_Y]) * fac;
    }

    float dabs_moved = self->states[ST_PARTIAL_DABS];
    float dabs_todo = count_dabs_to(self, x, y, pressure, dtime);

    if (dtime > 5 || self->reset_requested) {
        self->reset_requested = 0;
        for (int i = 0; i < ST_COUNT; i++) self->states[i] = 0;
        self->states[ST_X] = x;
        self->states[ST_Y] = y;
        self->states[ST_PRESSURE] = pressure;
        self->states[ST_ACTUAL_X] = self->states[ST_X];
        self->states[ST_ACTUAL_Y] = self->states[ST_Y];
        self->states[ST_STROKE] = 1.0;
        return 1;
    }

    double dtime_left = dtime;
    float step_ddab, step_dx, step_dy, step_dpressure, step_dtime;
    float step_declination, step_ascension;
    while (dabs_moved + dabs_todo >= 1.0) {
        {
            float frac;
            if (dabs_moved > 0) {
                step_ddab = 1.0 - dabs_moved;
                dabs_moved = 0;
            } else {
                step_ddab = 1.0;
            }
            frac = step_ddab / dabs_todo;
            step_dx = frac * (x - self->states[ST_X]);
            step_dy = frac * (y - self->states[ST_Y]);
            step_dpressure = frac * (pressure - self->states[ST_PRESSURE]);
            step_dtime = frac * (dtime_left - 0.0);
            step_declination = frac * (tilt_declination - self->states[ST_DECLINATION]);
            /* smallest_angular_difference(ASCENSION, 0) with ASCENSION == 0 always */
            step_ascension = frac * (tilt_ascension - self->states[ST_ASCENSION]);
        }
        update_states_and_setting_values(self, step_ddab, step_dx, step_dy, step_dpressure,
                                         step_declination, step_ascension, step_dtime);
        prepare_and_draw_dab(self, surface);
        dtime_left -= step_dtime;
        dabs_todo = count_dabs_to(self, x, y, pressure, dtime_left);
    }
    {
        step_ddab = dabs_todo;
        step_dx = x - self->states[ST_X];
        step_dy = y - self->states[ST_Y];
        step_dpressure = pressure - self->states[ST_PRESSURE];
        step_declination = tilt_declination - self->states[ST_DECLINATION];
        step_ascension = tilt_ascension - self->states[ST_ASCENSION];
        step_dtime = dtime_left;
        update_states_and_setting_values(self, step_ddab, step_dx, step_dy, step_dpressure,
                                         step_declination, step_ascension, step_dtime);
    }
    self->states[ST_PARTIAL_DABS] = dabs_moved + dabs_todo;
    return 0;
}

/* ------------------------------------------------------------------ */
/* envs/mypaint_utils.py (live subset) -- Python floats are C doubles   */
/* ------------------------------------------------------------------ */
static void mu_multiply(double x, double y, double d, double *ox, double *oy) { *ox = x * d; *oy = y * d; }
static void mu_add(double x1, double y1, double x2, double y2, double *ox, double *oy) { *ox = x1 + x2; *oy = y1 + y2; }
static void mu_multiply_add(double x1, double y1, double x2, double y2, double d, double *ox, double *oy)
{
    double x3, y3;
    mu_multiply(x2, y2, d, &x3, &y3);
    mu_add(x1, y1, x3, y3, ox, oy);
}
static void mu_difference(double x1, double y1, double x2, double y2, double *ox, double *oy) { *ox = x2 - x1; *oy = y2 - y1; }
static void mu_midpoint(double x1, double y1, double x2, double y2, double *ox, double *oy)
{
    *ox = (x1 + x2) / 2.0; *oy = (y1 + y2) / 2.0;
}
static double mu_vector_length(double x, double y) { return sqrt(x * x + y * y); }
static void mu_length_and_normal(double x1, double y1, double x2, double y2, double *len, double *nx, double *ny)
{
    double x, y;
    mu_difference(x1, y1, x2, y2, &x, &y);
    double length = mu_vector_length(x, y);
    if (length == 0.0) { x = 0.0; y = 0.0; }
    else { x = x / length; y = y / length; }
    *len = length; *nx = x; *ny = y;
}
/* mypaint_utils.py:5-11 */
static void mu_point_on_curve_1(double t, double cx, double cy, double sx, double sy,
                                double x1, double y1, double x2, double y2, double *ox, double *oy)
{
    double ratio = t / 100.0;
    double x3, y3, x4, y4, x5, y5;
    mu_multiply_add(sx, sy, x1, y1, ratio, &x3, &y3);
    mu_multiply_add(cx, cy, x2, y2, ratio, &x4, &y4);
    mu_difference(x3, y3, x4, y4, &x5, &y5);
    mu_multiply_add(x3, y3, x5, y5, ratio, ox, oy);
}

/* ------------------------------------------------------------------ */
/* envs/mnist.py MNIST env, one canvas                                  */
/* ------------------------------------------------------------------ */
struct POEnv {
    POEnvCfg cfg;
    POSurface *s;
    POBrush *b;
    int step;
    int have_start;           /* s_x, s_y is not None */
    double s_x, s_y;
    double entry_pressure;
    /* event log */
    POEvent *evlog;
    int evlog_cap, evlog_n;
};

POEnv *po_env_new(const POEnvCfg *cfg)
{
    POEnv *e = (POEnv *)calloc(1, sizeof(POEnv));
    e->cfg = *cfg;
    e->s = po_surface_new();
    e->b = NULL;
    return e;
}
void po_env_free(POEnv *e)
{
    if (!e) return;
    po_surface_free(e->s);
    po_brush_free(e->b);
    free(e->evlog);
    free(e);
}
POSurface *po_env_surface(POEnv *e) { return e->s; }
POBrush *po_env_brush(POEnv *e) { return e->b; }
void po_env_enable_eventlog(POEnv *e, int cap)
{
    free(e->evlog);
    e->evlog = (POEvent *)calloc((size_t)cap, sizeof(POEvent));
    e->evlog_cap = cap;
    e->evlog_n = 0;
}
int po_env_eventlog_count(POEnv *e) { return e->evlog_n; }
const POEvent *po_env_eventlog(POEnv *e) { return e->evlog; }
void po_env_eventlog_reset(POEnv *e) { e->evlog_n = 0; }

static void env_apply_brush_cfg(POEnv *e)
{
    const POEnvCfg *c = &e->cfg;
    for (int i = 0; i < PO_SETTING_COUNT; i++) {
        e->b->settings[i].base_value = c->base[i];
        e->b->settings[i].inputs_used = 0;
        for (int j = 0; j < PO_INPUT_COUNT; j++) {
            int n = c->map_n[i][j];
            e->b->settings[i].points[j].n = n;
            if (n) e->b->settings[i].inputs_used++;
            for (int k = 0; k < n; k++) {
                e->b->settings[i].points[j].xvalues[k] = c->map_x[i][j][k];
                e->b->settings[i].points[j].yvalues[k] = c->map_y[i][j][k];
            }
        }
    }
    settings_base_values_have_changed(e->b);
}

/* mnist.py:81-105 */
void po_env_reset(POEnv *e)
{
    e->entry_pressure = e->cfg.entry_pressure0;     /* np.min(self.pressures) */
    po_surface_clear(e->s);
    po_surface_fill_white(e->s);                     /* Surface(); flood_fill(...) */
    po_brush_free(e->b);
    e->b = po_brush_new(e->cfg.math_mode);           /* BrushInfo(file) ; Brush(bi) */
    env_apply_brush_cfg(e);
    e->step = 0;
    e->have_start = 0;
    e->s_x = e->s_y = 0.0;
}

/* mnist.py:216-222 */
static void env_stroke_to(POEnv *e, double x, double y, double pressure, double duration)
{
    if (e->evlog) {
        if (e->evlog_n < e->evlog_cap) {
            POEvent *ev = &e->evlog[e->evlog_n];
            ev->x = (float)x; ev->y = (float)y; ev->pressure = (float)pressure; ev->dtime = duration;
        }
        e->evlog_n++;
    }
    po_brush_stroke_to(e->b, e->s, (float)x, (float)y, (float)pressure, 0.0f, 0.0f, duration);
}

/* mnist.py:173-214 */
static double env_curve(POEnv *e, double cx, double cy, double sx, double sy, double ex, double ey,
                        double pressure)
{
    /* _line_settings, mnist.py:270-278 */
    double p1 = e->entry_pressure;
    double p2 = (e->entry_pressure + pressure) / 2;
    double p3 = pressure;
    if (e->cfg.head == 0.0001) p1 = p2;
    double prange1 = p2 - p1;
    double prange2 = p3 - p2;
    double entry_p = p1, midpoint_p = p2, h = e->cfg.head, t = e->cfg.tail;

    int points_in_curve = 100;
    double mx, my, length, nx, ny, x1, y1, x2, y2;
    mu_midpoint(sx, sy, ex, ey, &mx, &my);
    mu_length_and_normal(mx, my, cx, cy, &length, &nx, &ny);
    mu_multiply_add(mx, my, nx, ny, length * 2, &cx, &cy);
    mu_difference(sx, sy, cx, cy, &x1, &y1);
    mu_difference(cx, cy, ex, ey, &x2, &y2);
    double head = points_in_curve * h;
    int head_range = (int)head + 1;
    double tail = points_in_curve * t;
    int tail_range = (int)tail + 1;
    double tail_length = points_in_curve - tail;

    double px, py, bx, by;
    mu_point_on_curve_1(1, cx, cy, sx, sy, x1, y1, x2, y2, &px, &py);
    mu_length_and_normal(sx, sy, px, py, &length, &nx, &ny);
    mu_multiply_add(sx, sy, nx, ny, 0.25, &bx, &by);
    env_stroke_to(e, bx, by, entry_p, 0.1);
    pressure = fabs(1 / head * prange1 + entry_p);
    env_stroke_to(e, px, py, pressure, 0.1);

    for (int i = 2; i < head_range; i++) {
        mu_point_on_curve_1(i, cx, cy, sx, sy, x1, y1, x2, y2, &px, &py);
        pressure = fabs(i / head * prange1 + entry_p);
        env_stroke_to(e, px, py, pressure, 0.1);
    }
    for (int i = head_range; i < tail_range; i++) {
        mu_point_on_curve_1(i, cx, cy, sx, sy, x1, y1, x2, y2, &px, &py);
        env_stroke_to(e, px, py, midpoint_p, 0.1);
    }
    for (int i = tail_range; i < points_in_curve + 1; i++) {
        mu_point_on_curve_1(i, cx, cy, sx, sy, x1, y1, x2, y2, &px, &py);
        pressure = fabs((i - tail) / tail_length * prange2 + midpoint_p);
        env_stroke_to(e, px, py, pressure, 0.1);
    }
    return pressure;
}

/* mnist.py:107-167.  Decoded values come from the caller (the per-name table
 * lookups of mnist.py:117-130 are done by po_env_step_actions / the Python side). */
void po_env_draw(POEnv *e, double x, double y, double c_x, double c_y, double pressure,
                 double size, int jump, int set_color, double col_h, double col_s, double col_v)
{
    const POEnvCfg *c = &e->cfg;
    if (set_color) {                       /* brushinfo.set_color_rgb -> color_h/s/v base values */
        e->b->settings[S_COLOR_H].base_value = (float)col_h;
        e->b->settings[S_COLOR_S].base_value = (float)col_s;
        e->b->settings[S_COLOR_V].base_value = (float)col_v;
    }
    if (c->has_size) po_brush_set_base_value(e->b, S_RADIUS_LOGARITHMIC, (float)size);

    if (!e->have_start) {
        pressure = 0;
        e->s_x = 0; e->s_y = 0;
        e->have_start = 1;
        env_stroke_to(e, e->s_x, e->s_y, pressure, 0.1);
    } else if (c->has_jump && jump) {
        pressure = 0;
        env_stroke_to(e, e->s_x, e->s_y, pressure, 0.1);
    } else {
        env_stroke_to(e, e->s_x, e->s_y, pressure, 0.1);
    }

    /* _draw */
    double end_pressure = pressure;
    if (!c->has_control || pressure == 0) {
        if (e->evlog) {
            if (e->evlog_n < e->evlog_cap) {
                POEvent *ev = &e->evlog[e->evlog_n];
                ev->x = (float)x; ev->y = (float)y; ev->pressure = (float)pressure; ev->dtime = 1.0;
            }
            e->evlog_n++;
        }
        po_brush_stroke_to(e->b, e->s, (float)x, (float)y, (float)pressure, 0, 0, 1.0 /* dtime */);
    } else {
        end_pressure = env_curve(e, c_x, c_y, e->s_x, e->s_y, x, y, pressure);
    }
    e->entry_pressure = end_pressure;
    e->s_x = x; e->s_y = y;
    /* s.end_atomic(); s.begin_atomic(): ops already applied in order */
    e->step++;
}

/* mnist.py:107-130 decode + draw.  ac[] is indexed by the env's acs order. */
void po_env_step_actions(POEnv *e, const int32_t *ac)
{
    const POEnvCfg *c = &e->cfg;
    int L = c->L;
    int jump = 0;
    double x = c->lin[0], y = c->lin[0];         /* self.ends[0] */
    double c_x = c->lin[0], c_y = c->lin[0];     /* self.controls[0] */
    double pressure = c->default_pressure, size = c->default_size;
    int set_color = 0;
    double ch = 0, cs = 0, cv = 0;
    if (c->idx_end >= 0) { int a = ac[c->idx_end]; x = c->lin[a % L]; y = c->lin[a / L]; }
    if (c->idx_control >= 0) { int a = ac[c->idx_control]; c_x = c->lin[a % L]; c_y = c->lin[a / L]; }
    if (c->idx_pressure >= 0) pressure = c->pressures[ac[c->idx_pressure]];
    if (c->idx_size >= 0) size = c->sizes[ac[c->idx_size]];
    if (c->idx_jump >= 0) jump = ac[c->idx_jump];
    if (c->idx_color >= 0) {                     /* EXTENSION: colour action head */
        int a = ac[c->idx_color];
        set_color = 1; ch = c->colors_hsv[a][0]; cs = c->colors_hsv[a][1]; cv = c->colors_hsv[a][2];
    } else if (c->colorize) {                    /* mnist.py:132-133 */
        int a = e->step % c->n_colors;
        set_color = 1; ch = c->colors_hsv[a][0]; cs = c->colors_hsv[a][1]; cv = c->colors_hsv[a][2];
    }
    po_env_draw(e, x, y, c_x, c_y, pressure, size, jump, set_color, ch, cs, cv);
}

int po_env_step_index(POEnv *e) { return e->step; }
double po_env_entry_pressure(POEnv *e) { return e->entry_pressure; }

/* mnist.py:251-260 + envs/utils.py:3-5.  channels==1: gray; channels==3 (EXTENSION): image[...,:3] */
void po_env_state(POEnv *e, const uint16_t *noise, int channels, double *out)
{
    uint8_t img[TILE * TILE * 4];
    po_surface_read_rgbu8(e->s, noise, img);
    for (int i = 0; i < TILE * TILE; i++) {
        if (channels == 1) {
            /* np.dot(rgb[...,:3], [0.299, 0.587, 0.114]) */
            double acc = (double)img[i * 4 + 0] * 0.299;
            acc += (double)img[i * 4 + 1] * 0.587;
            acc += (double)img[i * 4 + 2] * 0.114;
            out[i] = acc;
        } else {
            out[i * 3 + 0] = img[i * 4 + 0];
            out[i * 3 + 1] = img[i * 4 + 1];
            out[i * 3 + 2] = img[i * 4 + 2];
        }
    }
}

/* Whole-episode convenience used by the CPU baseline: reset, T steps, final state.
 * actions: [T,K] int32.  obs_out: [T+1, 64*64*channels] float32 or NULL; final canvas
 * fix15 copied to canvas_out if not NULL. */
void po_env_run_episode(POEnv *e, const int32_t *actions, int T, int K, const uint16_t *noise,
                        int channels, float *obs_out, uint16_t *canvas_out)
{
    double tmp[TILE * TILE * 3];
    int n = TILE * TILE * channels;
    po_env_reset(e);
    if (obs_out) {
        po_env_state(e, noise, channels, tmp);
        for (int i = 0; i < n; i++) obs_out[i] = (float)tmp[i];
    }
    for (int t = 0; t < T; t++) {
        po_env_step_actions(e, actions + (size_t)t * K);
        if (obs_out) {
            po_env_state(e, noise, channels, tmp);
            float *o = obs_out + (size_t)(t + 1) * n;
            for (int i = 0; i < n; i++) o[i] = (float)tmp[i];
        }
    }
    if (canvas_out) memcpy(canvas_out, e->s->rgba, sizeof(e->s->rgba));
}

/* exported det-math entry points (tests compare them with the CUDA copy and libm) */
double po_det_exp(double x) { return det_exp(x); }
double po_det_log(double x) { return det_log(x); }
float po_det_hypotf(float a, float b) { return det_hypotf(a, b); }
float po_fast_expf(float x) { return fast_expf(x); }
float po_fast_exp_neg(float x) { float n; fast_exp_pm(x, &n); return n; }
float po_fast_logf(float x) { return fast_logf(x); }
float po_fast_hypotf(float a, float b) { return fast_hypotf(a, b); }
float po_brush_get_inv_radius(const POBrush *b) { return b->inv_actual_radius; }
