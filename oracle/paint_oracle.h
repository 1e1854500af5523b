/*
 * oracle/paint_oracle.h -- TEST INFRASTRUCTURE (CPU oracle), NOT product code.
 * See paint_oracle.c for what this restates and the "parity unpinned" note.
 */
#ifndef SPIRAL_PAINT_ORACLE_H
#define SPIRAL_PAINT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PO_MATH_LIBM 0
#define PO_MATH_DET 1
#define PO_MATH_FAST 2   /* the CUDA kernel's binary32 mode (fast_math.h + stroke_to_fast) */

#define PO_SETTING_COUNT 45
#define PO_INPUT_COUNT 9
#define PO_MAX_POINTS 8
#define PO_MAX_TABLE 8
#define PO_MAX_L 64

typedef struct POSurface POSurface;
typedef struct POBrush POBrush;
typedef struct POEnv POEnv;

/* one dab as handed to the tile (after draw_dab's clamps/rejects, tile (0,0) only) */
typedef struct {
    float x, y, radius, hardness;
    uint16_t opacity;      /* (uint16)(normal*opaque*2^15) */
    uint16_t r, g, b;      /* fix15 colour */
} PODab;

/* one stroke_to event as it crosses the SWIG boundary (float32 x,y,pressure; double dtime) */
typedef struct {
    float x, y, pressure;
    double dtime;
} POEvent;

typedef struct {
    int math_mode;
    /* brush (parsed .myb, v2): base values + mapping points */
    float base[PO_SETTING_COUNT];
    int map_n[PO_SETTING_COUNT][PO_INPUT_COUNT];
    float map_x[PO_SETTING_COUNT][PO_INPUT_COUNT][PO_MAX_POINTS];
    float map_y[PO_SETTING_COUNT][PO_INPUT_COUNT][PO_MAX_POINTS];
    /* env tables (envs/mnist.py:37-79) */
    int L;
    double lin[PO_MAX_L];              /* np.linspace(0, screen_size, L) */
    int has_jump, has_control, has_size;
    int idx_jump, idx_control, idx_end, idx_pressure, idx_size, idx_color;   /* -1 = absent */
    double pressures[PO_MAX_TABLE];
    double sizes[PO_MAX_TABLE];
    int n_colors;
    double colors_hsv[PO_MAX_TABLE][3];
    int colorize;
    double default_pressure, default_size;   /* MNIST.pressure=0.3, MNIST.size=0.2 */
    double head, tail;                        /* 0.25, 0.75 */
    double entry_pressure0;                   /* np.min(self.pressures) */
} POEnvCfg;

/* surface */
POSurface *po_surface_new(void);
void po_surface_free(POSurface *s);
void po_surface_clear(POSurface *s);
void po_surface_fill_white(POSurface *s);
uint16_t *po_surface_data(POSurface *s);
void po_surface_enable_dablog(POSurface *s, int cap);
int po_surface_dablog_count(POSurface *s);
const PODab *po_surface_dablog(POSurface *s);
void po_surface_dablog_reset(POSurface *s);
void po_surface_counters(POSurface *s, long *total, long *drawn, long *on_tile);
void po_surface_read_rgbu8(const POSurface *s, const uint16_t *noise, uint8_t *out);

/* glibc rand() stream / pixops dithering table */
void po_glibc_rand_stream(unsigned int seed, int n, int32_t *out);
void po_dither_table(unsigned int seed, int n, uint16_t *out);

/* brush */
int po_setting_count(void);
int po_input_count(void);
const char *po_setting_name(int i);
const char *po_input_name(int i);
float po_setting_default(int i);
POBrush *po_brush_new(int math_mode);
void po_brush_free(POBrush *b);
void po_brush_set_base_value(POBrush *b, int id, float value);
void po_brush_set_mapping_n(POBrush *b, int id, int input, int n);
void po_brush_set_mapping_point(POBrush *b, int id, int input, int index, float x, float y);
float po_brush_get_state(const POBrush *b, int i);
int po_brush_state_count(void);
void po_brush_get_rng(const POBrush *b, double *ran_u, double *buf, int *ptr);
double po_brush_rng_next(POBrush *b);
int po_brush_stroke_to(POBrush *b, POSurface *s, float x, float y, float pressure,
                       float xtilt, float ytilt, double dtime);
void po_hsv_to_rgb_float(float *h, float *s, float *v);

/* env (one canvas) */
POEnv *po_env_new(const POEnvCfg *cfg);
void po_env_free(POEnv *e);
POSurface *po_env_surface(POEnv *e);
POBrush *po_env_brush(POEnv *e);
void po_env_enable_eventlog(POEnv *e, int cap);
int po_env_eventlog_count(POEnv *e);
const POEvent *po_env_eventlog(POEnv *e);
void po_env_eventlog_reset(POEnv *e);
void po_env_reset(POEnv *e);
void po_env_draw(POEnv *e, double x, double y, double c_x, double c_y, double pressure,
                 double size, int jump, int set_color, double col_h, double col_s, double col_v);
void po_env_step_actions(POEnv *e, const int32_t *ac);
int po_env_step_index(POEnv *e);
double po_env_entry_pressure(POEnv *e);
void po_env_state(POEnv *e, const uint16_t *noise, int channels, double *out);
void po_env_run_episode(POEnv *e, const int32_t *actions, int T, int K, const uint16_t *noise,
                        int channels, float *obs_out, uint16_t *canvas_out);

double po_det_exp(double x);
double po_det_log(double x);
float po_det_hypotf(float a, float b);
float po_fast_expf(float x);
float po_fast_exp_neg(float x);
float po_fast_logf(float x);
float po_fast_hypotf(float a, float b);
float po_brush_get_inv_radius(const POBrush *b);

#ifdef __cplusplus
}
#endif
#endif
