// api_env.cu -- C-ABI glue for the painting environment (include/spiral_b200.h, K1).
//
// Host side only: validates the configuration, precomputes the per-brush constants the kernels
// read (the same float/double expressions libmypaint evaluates once per brush in
// settings_base_values_have_changed / render_dab_mask), seeds the RNG block, launches kernels.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "../../include/spiral_b200.h"
#include "common.h"
#include "detmath.cuh"
#include "fastmath.cuh"
#include "raster.cuh"

namespace spiral {
cudaError_t launch_reset(const EnvParams &P, uint16_t *canvas, float *bs, double *es, float *obs, cudaStream_t st);
cudaError_t launch_expand(const EnvParams &P, const int32_t *actions, float *bs, double *es, float *dabs,
                          int32_t *dab_count, int32_t *status, const float *gauss_tab, int32_t *plan_scratch,
                          cudaStream_t st);
cudaError_t launch_composite(const EnvParams &P, uint16_t *canvas, const float *dabs, const int32_t *dab_count,
                             const uint16_t *dither, float *obs, const float *gauss_tab, int32_t *work_counter, cudaStream_t st);
bool fused_step_applies(const EnvParams &P);
cudaError_t launch_step_small(const EnvParams &P, const int32_t *actions, float *bs, double *es, float *dabs,
                              int32_t *dab_count, int32_t *status, const float *gauss_tab, uint16_t *canvas,
                              const uint16_t *dither, float *obs, int32_t *plan_scratch, cudaStream_t st);
cudaError_t launch_debug_dabs(const EnvParams &P, const float *dabs, const int32_t *dab_count, const float *gauss_tab,
                              float *out, int32_t *out_count, cudaStream_t st);
cudaError_t launch_l2_reward(const float *obs, const float *target, float *reward, int batch, int n, cudaStream_t st);
cudaError_t launch_detmath(const double *x, double *e, double *l, int n, cudaStream_t st);
cudaError_t launch_fastmath(const float *x, float *e, float *en, float *l, float *h, int n, cudaStream_t st);
cudaError_t launch_divcheck(unsigned long long seed, long long n, unsigned long long *mismatches, cudaStream_t st);
long long bounds_violations_raster();
long long bounds_violations_learn();
}  // namespace spiral

using namespace spiral;

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

int spiral_set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
int spiral_cuda_error(cudaError_t e, const char *what)
{
    return spiral_set_error(SPIRAL_ECUDA, "%s: %s", what, cudaGetErrorString(e));
}

extern "C" int spiral_version(void) { return SPIRAL_ABI_VERSION; }
extern "C" const char *spiral_last_error(void) { return g_err; }

// ------------------------------------------------------------------------------------------
// rng-double.c set_seed (Knuth, ranf_start with KK=10, LL=7, TT=7): state after seeding = the
// first block of outputs
// ------------------------------------------------------------------------------------------
static inline double ms(double x, double y) { const double s = x + y; return s - (int)s; }

static void rng_array(double *ran_u, double *aa, int n)
{
    const int KK = 10, LL = 7;
    int i, j;
    for (j = 0; j < KK; j++) aa[j] = ran_u[j];
    for (; j < n; j++) aa[j] = ms(aa[j - KK], aa[j - LL]);
    for (i = 0; i < LL; i++, j++) ran_u[i] = ms(aa[j - KK], aa[j - LL]);
    for (; i < KK; i++, j++) ran_u[i] = ms(aa[j - KK], ran_u[i - LL]);
}

static void rng_seed_block(long seed, double *out10);

// rand_gauss (helpers.c) over every window u[i..i+3] of the rng-double output stream
static void rng_gauss_table(long seed, int n, float *g)
{
    double cur[10];
    rng_seed_block(seed, cur);
    const int need = n + 3;
    double *u = new double[(size_t)need + 10];
    int filled = 0;
    while (filled < need) {
        for (int j = 0; j < 10; j++) u[filled + j] = cur[j];     // one cycle() yields the state itself
        filled += 10;
        double a[29];                                            // ... and advances it by QUALITY = 19 steps
        for (int j = 0; j < 10; j++) a[j] = cur[j];
        for (int j = 10; j < 29; j++) a[j] = ms(a[j - 10], a[j - 7]);
        for (int j = 0; j < 10; j++) cur[j] = a[19 + j];
    }
    for (int i = 0; i < n; i++) {
        double sum = 0.0;
        sum += u[i];
        sum += u[i + 1];
        sum += u[i + 2];
        sum += u[i + 3];
        g[i] = (float)(sum * 1.73205080757 - 3.46410161514);
    }
    delete[] u;
}

static void rng_seed_block(long seed, double *out10)
{
    const int KK = 10, LL = 7, TT = 7;
    double u[KK + KK - 1], ran_u[KK];
    const double ulp = (1.0 / (1L << 30)) / (1L << 22);
    double ss = 2.0 * ulp * ((seed & 0x3fffffff) + 2);
    int j, s, t;
    for (j = 0; j < KK; j++) {
        u[j] = ss;
        ss += ss;
        if (ss >= 1.0) ss -= 1.0 - 2 * ulp;
    }
    u[1] += ulp;
    for (s = seed & 0x3fffffff, t = TT - 1; t;) {
        for (j = KK - 1; j > 0; j--) { u[j + j] = u[j]; u[j + j - 1] = 0.0; }
        for (j = KK + KK - 2; j >= KK; j--) {
            u[j - (KK - LL)] = ms(u[j - (KK - LL)], u[j]);
            u[j - KK] = ms(u[j - KK], u[j]);
        }
        if (s & 1) {
            for (j = KK; j > 0; j--) u[j] = u[j - 1];
            u[0] = u[KK];
            u[LL] = ms(u[LL], u[KK]);
        }
        if (s) s >>= 1; else t--;
    }
    for (j = 0; j < LL; j++) ran_u[j + KK - LL] = u[j];
    for (; j < KK; j++) ran_u[j - LL] = u[j];
    for (j = 0; j < 10; j++) rng_array(ran_u, u, KK + KK - 1);
    for (j = 0; j < KK; j++) out10[j] = ran_u[j];
}

// ------------------------------------------------------------------------------------------
// d_plan: i32 [64] work counters | i32 [512] plan histogram + cursors | i32 [B] permutation | u8 [B] keys
constexpr size_t kPlanHeader = 64 + 512;

struct SpiralEnv {
    EnvParams P;
    int device;
    uint16_t *d_dither;
    float *d_gauss;
    int32_t *d_plan;          // expand scheduling scratch: histogram, permutation, keys
};

static bool is_pow2f(float v)
{
    if (!(v > 0.0f) || !std::isfinite(v)) return false;
    int e;
    return std::frexp(v, &e) == 0.5f;
}

extern "C" int spiral_env_create(const SpiralEnvCfg *cfg, SpiralEnv **out)
{
    if (!cfg || !out) return spiral_set_error(SPIRAL_EINVAL, "spiral_env_create: null argument");
    *out = nullptr;
    if (cfg->abi_version != SPIRAL_ABI_VERSION)
        return spiral_set_error(SPIRAL_EINVAL, "ABI version mismatch: caller %d, library %d",
                                cfg->abi_version, SPIRAL_ABI_VERSION);
    if (cfg->batch <= 0) return spiral_set_error(SPIRAL_EINVAL, "batch must be > 0");
    if (cfg->channels != 1 && cfg->channels != 3)
        return spiral_set_error(SPIRAL_EINVAL, "channels must be 1 or 3 (got %d)", cfg->channels);
    if (cfg->location_size < 1 || cfg->location_size > SPIRAL_MAX_L)
        return spiral_set_error(SPIRAL_EINVAL, "location_size must be in 1..%d", SPIRAL_MAX_L);
    if (cfg->n_heads < 1 || cfg->n_heads > 8) return spiral_set_error(SPIRAL_EINVAL, "n_heads must be in 1..8");
    const int idx[6] = {cfg->idx_jump, cfg->idx_control, cfg->idx_end, cfg->idx_pressure, cfg->idx_size, cfg->idx_color};
    for (int i = 0; i < 6; i++)
        if (idx[i] < -1 || idx[i] >= cfg->n_heads)
            return spiral_set_error(SPIRAL_EINVAL, "action column index %d out of range", idx[i]);
    if (cfg->n_pressures > SPIRAL_MAX_TABLE || cfg->n_sizes > SPIRAL_MAX_TABLE || cfg->n_colors > SPIRAL_MAX_TABLE)
        return spiral_set_error(SPIRAL_EINVAL, "table sizes must be <= %d", SPIRAL_MAX_TABLE);
    if ((cfg->idx_color >= 0 || cfg->colorize) && cfg->n_colors < 1)
        return spiral_set_error(SPIRAL_EINVAL, "colour action / colorize needs a palette");
    if (cfg->max_dabs < 32) return spiral_set_error(SPIRAL_EINVAL, "max_dabs must be >= 32");
    if (cfg->math_mode != SPIRAL_MATH_DET && cfg->math_mode != SPIRAL_MATH_FAST)
        return spiral_set_error(SPIRAL_EINVAL, "math_mode must be SPIRAL_MATH_DET (1) or SPIRAL_MATH_FAST (2), got %d", cfg->math_mode);
    if (cfg->rng_table_len < 4096 || cfg->rng_table_len > (1 << 26))
        return spiral_set_error(SPIRAL_EINVAL, "rng_table_len must be in 4096..2^26 (got %d)", cfg->rng_table_len);
    {
        const double head = 100 * cfg->head, tail = 100 * cfg->tail;
        const int hr = (int)head + 1, tr = (int)tail + 1;
        if (hr < 2 || tr < hr || tr > 101)
            return spiral_set_error(SPIRAL_EUNSUPPORTED, "head/tail %.4f/%.4f outside the supported range", cfg->head, cfg->tail);
    }
    const SpiralBrush &br = cfg->brush;
    for (int s = 0; s < SPIRAL_DYN_COUNT; s++)
        for (int j = 0; j < SPIRAL_IN_COUNT; j++) {
            const int n = br.dyn[s].n[j];
            if (n != 0 && (n < 2 || n > SPIRAL_MAX_POINTS))
                return spiral_set_error(SPIRAL_EINVAL, "mapping %d/%d has %d points (need 0 or 2..%d)", s, j, n, SPIRAL_MAX_POINTS);
        }

    SpiralEnv *env = new (std::nothrow) SpiralEnv();
    if (!env) return spiral_set_error(SPIRAL_EINVAL, "out of host memory");
    memset(env, 0, sizeof(*env));
    EnvParams &P = env->P;
    env->device = cfg->device;
    P.B = cfg->batch; P.C = cfg->channels; P.L = cfg->location_size; P.K = cfg->n_heads;
    P.idx_jump = cfg->idx_jump; P.idx_control = cfg->idx_control; P.idx_end = cfg->idx_end;
    P.idx_pressure = cfg->idx_pressure; P.idx_size = cfg->idx_size; P.idx_color = cfg->idx_color;
    P.n_colors = cfg->n_colors; P.colorize = cfg->colorize; P.dither = cfg->dither; P.max_dabs = cfg->max_dabs;
    for (int i = 0; i < SPIRAL_MAX_L; i++) P.lin[i] = cfg->lin[i];
    for (int i = 0; i < SPIRAL_MAX_TABLE; i++) {
        P.pressures[i] = cfg->pressures[i];
        P.sizes[i] = cfg->sizes[i];
        for (int c = 0; c < 4; c++) P.colors[i][c] = cfg->colors_fix15[i][c];
    }
    P.default_pressure = cfg->default_pressure; P.default_size = cfg->default_size;
    P.head = cfg->head; P.tail = cfg->tail; P.entry_pressure0 = cfg->entry_pressure0;

    P.uses_speed1 = P.uses_speed2 = 0;
    P.mapped_mask = 0;
    P.rng_table_len = cfg->rng_table_len;
    {
        const char *e = getenv("SPIRAL_EXPAND_LANES");      // tuning knob for profiling
        P.expand_lanes = e ? atoi(e) : 0;
        if (P.expand_lanes < 0 || P.expand_lanes > 32) P.expand_lanes = 0;
        const char *so = getenv("SPIRAL_EXPAND_SORT");       // 0 disables the work-sorted schedule (profiling)
        P.expand_sort = so ? (atoi(so) != 0) : 1;
        const char *sd = getenv("SPIRAL_EXPAND_SOLO_DIV");
        P.expand_solo_div = sd ? atoi(sd) : 128;
        if (P.expand_solo_div < 1) P.expand_solo_div = 128;
    }
    for (int s = 0; s < SPIRAL_DYN_COUNT; s++) {
        MapDev &m = P.dyn[s];
        m.base = br.dyn[s].base;
        for (int j = 0; j < SPIRAL_IN_COUNT; j++) {
            m.n[j] = br.dyn[s].n[j];
            if (m.n[j]) P.mapped_mask |= 1 << s;
            if (m.n[j] && j == SPIRAL_IN_SPEED1) P.uses_speed1 = 1;
            if (m.n[j] && j == SPIRAL_IN_SPEED2) P.uses_speed2 = 1;
            for (int k = 0; k < SPIRAL_MAX_POINTS; k++) {
                m.x[j][k] = br.dyn[s].x[j][k];
                m.y[j][k] = br.dyn[s].y[j][k];
                m.inv[j][k] = 0.0f;
            }
            for (int k = 0; k + 1 < m.n[j]; k++) {
                const float den = m.x[j][k + 1] - m.x[j][k];
                if (is_pow2f(den)) m.inv[j][k] = 1.0f / den;
                // FAST mode: fmaf(slope, x, icpt) (oracle: mapping_calculate_fast)
                const float x0 = m.x[j][k], y0 = m.y[j][k], x1 = m.x[j][k + 1], y1 = m.y[j][k + 1];
                m.slope[j][k] = 0.0f;
                m.icpt[j][k] = y0;
                if (!(x0 == x1 || y0 == y1)) {
                    m.slope[j][k] = (y1 - y0) / (x1 - x0);
                    m.icpt[j][k] = y0 - m.slope[j][k] * x0;
                }
            }
        }
    }
    // dry_brush.myb structure: {opaque, opaque_multiply, offset_by_random} <- pressure, radius_logarithmic <- speed2,
    // 2 points each, nothing else mapped
    {
        bool ok = true;
        for (int s = 0; s < SPIRAL_DYN_COUNT; s++)
            for (int j = 0; j < SPIRAL_IN_COUNT; j++) {
                const bool want = (j == SPIRAL_IN_PRESSURE && (s == SPIRAL_DYN_OPAQUE || s == SPIRAL_DYN_OPAQUE_MULTIPLY || s == SPIRAL_DYN_OFFSET_BY_RANDOM)) ||
                                  (j == SPIRAL_IN_SPEED2 && s == SPIRAL_DYN_RADIUS_LOGARITHMIC);
                ok &= (br.dyn[s].n[j] == (want ? 2 : 0));
            }
        P.dry_structure = ok ? 1 : 0;
    }
    // arithmetic mode: the per-brush constants below are evaluated with the same functions the kernels (and the oracle's
    // settings_base_values_have_changed / exp_decay in that mode) use
    P.math_mode = cfg->math_mode;
    const bool fast = cfg->math_mode == SPIRAL_MATH_FAST;
    auto m_expf = [fast](float x) { return fast ? fastmath::fast_expf(x) : detmath::det_expf(x); };
    // base radius per size-table entry: set_base_value('radius_logarithmic', size) takes a float
    for (int i = 0; i < 8; i++) P.base_rlog[i] = (float)cfg->sizes[i];
    P.base_rlog[8] = br.dyn[SPIRAL_DYN_RADIUS_LOGARITHMIC].base;
    for (int i = 0; i < 9; i++) {
        P.base_radius[i] = m_expf(P.base_rlog[i]);
        // FAST mode tables (oracle: stroke_to_fast / count_dabs_fast / update_states_fast evaluate these expressions in place)
        const float r = P.base_radius[i];
        const float rc = r < 0.2f ? 0.2f : (r > 1000.0f ? 1000.0f : r);
        P.f_inv_norm01[i] = 1.0f / ((float)0.1 * r);
        P.f_inv_norm1[i] = 1.0f / (1.0f * r);
        P.f_inv_norm001[i] = 1.0f / (0.001f * r);
        P.f_k_basic[i] = br.dabs_per_basic_radius / rc;
    }
    for (int i = 0; i < 2; i++) {
        const float T = br.dyn[i == 0 ? SPIRAL_DYN_SPEED1_SLOWNESS : SPIRAL_DYN_SPEED2_SLOWNESS].base;
        P.f_neg_inv_slow[i] = T < 0.001f ? 0.0f : -1.0f / T;
    }

    P.slow_tracking = br.slow_tracking;
    // stroke_to: fac = 1.0 - exp_decay(slow_tracking, 100.0*dtime) for the two event durations
    {
        const double dts[2] = {0.1, 1.0};
        for (int i = 0; i < 2; i++) {
            const float t = (float)(100.0 * dts[i]);
            float ed = 0.0f;
            if (fast) {
                if (!(br.slow_tracking < 0.001f)) ed = fastmath::fast_expf(t * (-1.0f / br.slow_tracking));
                P.track_fac[i] = 1.0f - ed;
                continue;
            }
            if (!((double)br.slow_tracking <= 0.001)) ed = (float)detmath::det_exp((double)(-t / br.slow_tracking));
            P.track_fac[i] = (float)(1.0 - (double)ed);
        }
    }
    // settings_base_values_have_changed: speed input mapping  y = log(gamma + x) * m + q
    const float gam[2] = {br.speed1_gamma, br.speed2_gamma};
    for (int i = 0; i < 2; i++) {
        float gamma = m_expf(gam[i]);
        const float fix1_x = 45.0f, fix1_y = 0.5f, fix2_x = 45.0f, fix2_dy = 0.015f;
        const float c1 = fast ? fastmath::fast_logf(fix1_x + gamma) : (float)detmath::det_log((double)(fix1_x + gamma));
        const float m = fix2_dy * (fix2_x + gamma);
        const float q = fix1_y - m * c1;
        P.sp_gamma[i] = gamma; P.sp_m[i] = m; P.sp_q[i] = q;
    }
    P.dabs_basic = br.dabs_per_basic_radius;
    P.dabs_actual = br.dabs_per_actual_radius;
    P.dabs_second = br.dabs_per_second;
    // render_dab_mask: float angle_rad = angle/360*2*M_PI; float cs = cos(angle_rad); ...
    {
        const float angle = br.elliptical_dab_angle;
        const float angle_rad = angle / 360 * 2 * M_PI;
        P.cs = (float)::cos((double)angle_rad);     // C: float cs = cos(angle_rad)  (double cos)
        P.sn = (float)::sin((double)angle_rad);
        P.l2 = P.cs * P.cs + P.sn * P.sn;
        P.rad_area_1 = sqrtf(1.0f / M_PI);
        P.fast_spans = (P.sn == 1.0f && fabsf(P.cs) < 1e-6f) ? 1 : 0;   // aspect ratio is always 1 here
    }
    // |rand_gauss| <= 2*sqrt(3); radius <= ACTUAL_RADIUS * exp(|g| * |radius_by_random|)
    {
        const SpiralMapping &rb = br.dyn[SPIRAL_DYN_RADIUS_BY_RANDOM];
        float mx = fabsf(rb.base);
        bool mapped = false;
        for (int j = 0; j < SPIRAL_IN_COUNT; j++) mapped |= rb.n[j] != 0;
        P.cull_factor = mapped ? 1e30f : (float)(::exp(3.4642 * (double)mx) * 1.001);
    }
    // The chain culls on the un-jittered position (raw_dab_may_paint): the margin adds the largest position jitter,
    // 2 sqrt(3) |offset_by_random| base_radius, and the anti-aliasing fringe.  Bounds exist when those settings read the
    // pressure input only (0 <= pressure <= 1 here: the env's pressures are <= 0.8) through mappings that end at x >= 1.
    {
        auto bound = [&](int sidx, bool &ok) {
            const SpiralMapping &mp = br.dyn[sidx];
            float mxv = fabsf(mp.base);
            for (int j = 0; j < SPIRAL_IN_COUNT; j++) {
                if (!mp.n[j]) continue;
                if (j != SPIRAL_IN_PRESSURE || mp.x[j][mp.n[j] - 1] < 1.0f || mp.x[j][0] > 0.0f) { ok = false; continue; }
                float my = 0.0f;
                for (int k = 0; k < mp.n[j]; k++) my = fmaxf(my, fabsf(mp.y[j][k]));
                mxv += my;
            }
            return mxv;
        };
        bool ok = true;
        const float obr_max = bound(SPIRAL_DYN_OFFSET_BY_RANDOM, ok);
        const float aa_max = bound(SPIRAL_DYN_ANTI_ALIASING, ok);
        for (int i = 0; i < 9; i++)
            P.cull_pad[i] = ok ? 3.4642f * obr_max * P.base_radius[i] * 1.001f + 0.5f * aa_max + 3.0f : 1e30f;
        // dabs of zero pressure are blank when opaque_multiply is exactly 0 + y(pressure) with y(0) == 0 (dry_brush.myb)
        const SpiralMapping &om = br.dyn[SPIRAL_DYN_OPAQUE_MULTIPLY];
        P.zero_pressure_blank = (P.dry_structure && om.base == 0.0f && om.x[SPIRAL_IN_PRESSURE][0] == 0.0f &&
                                 om.y[SPIRAL_IN_PRESSURE][0] == 0.0f) ? 1 : 0;
    }
    rng_seed_block(1000, P.rng_seed_block);

    int prev = 0;
    cudaError_t e = cudaGetDevice(&prev);
    if (e == cudaSuccess) e = cudaSetDevice(cfg->device);
    if (e == cudaSuccess) e = cudaMalloc(&env->d_dither, sizeof(cfg->dither_table));
    if (e == cudaSuccess)
        e = cudaMemcpy(env->d_dither, cfg->dither_table, sizeof(cfg->dither_table), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&env->d_gauss, sizeof(float) * (size_t)cfg->rng_table_len);
    if (e == cudaSuccess) e = cudaMalloc(&env->d_plan, sizeof(int32_t) * (kPlanHeader + (size_t)cfg->batch) + (size_t)cfg->batch + 16);
    if (e == cudaSuccess) {
        float *g = new float[(size_t)cfg->rng_table_len];
        rng_gauss_table(1000, cfg->rng_table_len, g);
        e = cudaMemcpy(env->d_gauss, g, sizeof(float) * (size_t)cfg->rng_table_len, cudaMemcpyHostToDevice);
        delete[] g;
    }
    cudaSetDevice(prev);
    if (e != cudaSuccess) {
        if (env->d_dither) cudaFree(env->d_dither);
        if (env->d_gauss) cudaFree(env->d_gauss);
        if (env->d_plan) cudaFree(env->d_plan);
        delete env;
        return spiral_cuda_error(e, "spiral_env_create");
    }
    *out = env;
    return SPIRAL_OK;
}

extern "C" int spiral_env_destroy(SpiralEnv *env)
{
    if (!env) return SPIRAL_OK;
    if (env->d_dither) cudaFree(env->d_dither);
    if (env->d_gauss) cudaFree(env->d_gauss);
    if (env->d_plan) cudaFree(env->d_plan);
    delete env;
    return SPIRAL_OK;
}

#define REQUIRE(p) do { if (!(p)) return spiral_set_error(SPIRAL_EINVAL, "%s: null or invalid argument `%s`", __func__, #p); } while (0)

extern "C" int spiral_env_reset(SpiralEnv *env, uint16_t *canvas, float *bs, double *es, float *obs, void *stream)
{
    REQUIRE(env); REQUIRE(canvas); REQUIRE(bs); REQUIRE(es); REQUIRE(obs);
    const cudaError_t e = launch_reset(env->P, canvas, bs, es, obs, (cudaStream_t)stream);
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_env_reset");
}

extern "C" int spiral_env_expand(SpiralEnv *env, const int32_t *actions, float *bs, double *es, float *dabs,
                                 int32_t *dab_count, int32_t *status, void *stream)
{
    REQUIRE(env); REQUIRE(actions); REQUIRE(bs); REQUIRE(es); REQUIRE(dabs); REQUIRE(dab_count); REQUIRE(status);
    const cudaError_t e = launch_expand(env->P, actions, bs, es, dabs, dab_count, status, env->d_gauss, env->d_plan + 64, (cudaStream_t)stream);
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_env_expand");
}

extern "C" int spiral_env_composite(SpiralEnv *env, uint16_t *canvas, const float *dabs, const int32_t *dab_count,
                                    float *obs, void *stream)
{
    REQUIRE(env); REQUIRE(canvas); REQUIRE(dabs); REQUIRE(dab_count); REQUIRE(obs);
    const cudaError_t e = launch_composite(env->P, canvas, dabs, dab_count, env->d_dither, obs, env->d_gauss, env->d_plan, (cudaStream_t)stream);
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_env_composite");
}

extern "C" int spiral_env_step(SpiralEnv *env, const int32_t *actions, uint16_t *canvas, float *bs, double *es,
                               float *dabs, int32_t *dab_count, float *obs, int32_t *status, void *stream)
{
    if (env && fused_step_applies(env->P)) {
        // small batches: one kernel, a canvas's dabs are composited while its brush chain is still producing them
        REQUIRE(actions); REQUIRE(canvas); REQUIRE(bs); REQUIRE(es); REQUIRE(dabs); REQUIRE(dab_count); REQUIRE(obs); REQUIRE(status);
        const cudaError_t e = launch_step_small(env->P, actions, bs, es, dabs, dab_count, status, env->d_gauss, canvas,
                                                env->d_dither, obs, env->d_plan + 64, (cudaStream_t)stream);
        return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_env_step");
    }
    int rc = spiral_env_expand(env, actions, bs, es, dabs, dab_count, status, stream);
    if (rc != SPIRAL_OK) return rc;
    return spiral_env_composite(env, canvas, dabs, dab_count, obs, stream);
}

extern "C" int spiral_env_debug_dabs(SpiralEnv *env, const float *dabs, const int32_t *dab_count, float *out,
                                     int32_t *out_count, void *stream)
{
    REQUIRE(env); REQUIRE(dabs); REQUIRE(dab_count); REQUIRE(out); REQUIRE(out_count);
    const cudaError_t e = launch_debug_dabs(env->P, dabs, dab_count, env->d_gauss, out, out_count, (cudaStream_t)stream);
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_env_debug_dabs");
}

extern "C" int spiral_l2_reward(const float *obs, const float *target, float *reward, int32_t batch, int32_t n,
                                void *stream)
{
    REQUIRE(obs); REQUIRE(target); REQUIRE(reward);
    if (batch <= 0 || n <= 0 || (n & 3)) return spiral_set_error(SPIRAL_EINVAL, "spiral_l2_reward: n must be a positive multiple of 4");
    const cudaError_t e = launch_l2_reward(obs, target, reward, batch, n, (cudaStream_t)stream);
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_l2_reward");
}

extern "C" int spiral_debug_fastmath(const float *x, float *e_out, float *en_out, float *l_out, float *h_out, int32_t n,
                                     void *stream)
{
    REQUIRE(x); REQUIRE(e_out); REQUIRE(en_out); REQUIRE(l_out); REQUIRE(h_out);
    const cudaError_t e = launch_fastmath(x, e_out, en_out, l_out, h_out, n, (cudaStream_t)stream);
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_debug_fastmath");
}

extern "C" int64_t spiral_debug_bounds_violations(void)
{
    const long long a = bounds_violations_raster(), b = bounds_violations_learn();
    if (a < 0 || b < 0) return a < b ? a : b;      // -1: not a checked build
    return (int64_t)(a + b);
}

extern "C" int spiral_debug_divcheck(uint64_t seed, int64_t n_pairs, uint64_t *mismatches_dev, void *stream)
{
    REQUIRE(mismatches_dev);
    if (n_pairs <= 0) return spiral_set_error(SPIRAL_EINVAL, "spiral_debug_divcheck: n_pairs must be > 0");
    const cudaError_t e = launch_divcheck(seed, n_pairs, (unsigned long long *)mismatches_dev, (cudaStream_t)stream);
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_debug_divcheck");
}

extern "C" int spiral_debug_detmath(const double *x, double *e_out, double *l_out, int32_t n, void *stream)
{
    REQUIRE(x); REQUIRE(e_out); REQUIRE(l_out);
    const cudaError_t e = launch_detmath(x, e_out, l_out, n, (cudaStream_t)stream);
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_debug_detmath");
}
