/*
 * include/spiral_b200.h -- C ABI of libspiral_b200.so (sm_100a).
 *
 * Drop-in boundary for the SPIRAL inner loop (render -> D reward -> A2C loss).
 * The reference (carpedm20/SPIRAL-tensorflow) has no FFI of its own for this
 * path: it reaches native code through the mypaint SWIG API and tf.Session.run.
 * Each entry point below names the reference interface it replaces.
 *
 * Conventions
 *   - every call returns 0 on success, a negative SPIRAL_E* code otherwise;
 *     the message is available from spiral_last_error() (thread local)
 *   - all `dev` pointers are DEVICE pointers owned by the caller (PyTorch
 *     tensors: tensor.data_ptr()); the library allocates nothing on the step
 *     path and never synchronises the device
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream)
 *   - one handle per (process, GPU); calls on one handle are not re-entrant
 */
#ifndef SPIRAL_B200_H
#define SPIRAL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPIRAL_ABI_VERSION 3

#define SPIRAL_OK 0
#define SPIRAL_EINVAL (-1)      /* bad argument / unsupported configuration */
#define SPIRAL_ECUDA (-2)       /* CUDA runtime error (message has the cudaError string) */
#define SPIRAL_EUNSUPPORTED (-3)

int spiral_version(void);
const char *spiral_last_error(void);

/* ------------------------------------------------------------------ */
/* Painting environment (K1)                                            */
/* replaces: envs/mnist.py:81-105 reset, :107-222 draw/_draw/curve,     */
/*           :231-260 step/state, and beneath them libmypaint 1.3.0     */
/*           mypaint_brush_stroke_to + tiled-surface draw_dab and       */
/*           mypaint 1.2.1 tile_convert_rgbu16_to_rgbu8                 */
/* ------------------------------------------------------------------ */
#define SPIRAL_MAX_POINTS 4
#define SPIRAL_MAX_TABLE 8
#define SPIRAL_MAX_L 64
#define SPIRAL_TILE 64
#define SPIRAL_BRUSH_STATE_FLOATS 16     /* per canvas, caller-owned float[B,16]  */
#define SPIRAL_ENV_STATE_DOUBLES 16      /* per canvas, caller-owned double[B,16] */
#define SPIRAL_DAB_FLOATS 8              /* per pending dab, caller-owned float[B,max_dabs,8] */

/* settings whose value may depend on brush inputs (libmypaint "mappings") */
enum {
    SPIRAL_DYN_OPAQUE = 0,
    SPIRAL_DYN_OPAQUE_MULTIPLY,
    SPIRAL_DYN_RADIUS_LOGARITHMIC,
    SPIRAL_DYN_HARDNESS,
    SPIRAL_DYN_RADIUS_BY_RANDOM,
    SPIRAL_DYN_OFFSET_BY_RANDOM,
    SPIRAL_DYN_SPEED1_SLOWNESS,
    SPIRAL_DYN_SPEED2_SLOWNESS,
    SPIRAL_DYN_ANTI_ALIASING,
    SPIRAL_DYN_COUNT
};
/* brush inputs a mapping may read */
enum { SPIRAL_IN_PRESSURE = 0, SPIRAL_IN_SPEED1, SPIRAL_IN_SPEED2, SPIRAL_IN_COUNT };

/* libmypaint mapping: value = base + sum_i piecewise_linear_i(input_i) */
typedef struct {
    float base;
    int32_t n[SPIRAL_IN_COUNT];                       /* points per input, 0 = unused, else 2..4 */
    float x[SPIRAL_IN_COUNT][SPIRAL_MAX_POINTS];
    float y[SPIRAL_IN_COUNT][SPIRAL_MAX_POINTS];
} SpiralMapping;

/* the subset of a .myb brush the rasteriser implements (assets/brushes/dry_brush.myb) */
typedef struct {
    SpiralMapping dyn[SPIRAL_DYN_COUNT];
    float slow_tracking;
    float speed1_gamma, speed2_gamma;
    float dabs_per_basic_radius, dabs_per_actual_radius, dabs_per_second;
    float elliptical_dab_angle;                       /* degrees; ratio is fixed at 1 */
} SpiralBrush;

typedef struct {
    int32_t abi_version;          /* SPIRAL_ABI_VERSION */
    int32_t device;               /* CUDA device ordinal */
    int32_t batch;                /* canvases on this GPU */
    int32_t channels;             /* observation channels: 1 = gray (utils.py:3-5), 3 = RGB (extension) */
    int32_t location_size;        /* L: control/end heads index an L x L grid (utils.py:10-18) */
    int32_t n_heads;              /* K: columns of the action matrix, env.acs order (base.py:32) */
    int32_t idx_jump, idx_control, idx_end, idx_pressure, idx_size, idx_color;   /* column or -1 */
    int32_t n_pressures, n_sizes, n_colors;
    int32_t colorize;             /* mnist.py:40,132-133: colour = palette[step] */
    int32_t dither;               /* 0 = round-half, 1 = per-pixel table (pixops.cpp dithering_noise) */
    int32_t max_dabs;             /* capacity of the per-canvas pending-dab list */
    int32_t rng_table_len;        /* rng-double stream positions tabulated per handle (>= draws of one episode) */
    double lin[SPIRAL_MAX_L];     /* np.linspace(0, 64, L) */
    double pressures[SPIRAL_MAX_TABLE];
    double sizes[SPIRAL_MAX_TABLE];
    uint16_t colors_fix15[SPIRAL_MAX_TABLE][4];   /* dab colour after set_color_rgb -> hsv -> rgb */
    double default_pressure, default_size;        /* MNIST.pressure, MNIST.size (mnist.py:34-35) */
    double head, tail;                            /* mnist.py:23-24 */
    double entry_pressure0;                       /* np.min(self.pressures), mnist.py:82 */
    SpiralBrush brush;
    uint16_t dither_table[SPIRAL_TILE * SPIRAL_TILE];
    /* Arithmetic of the brush state machine (libmypaint's expf/logf/hypotf calls and the dab loop around them):
     *   SPIRAL_MATH_DET  (1)  double-precision, correctly rounded replacements: canvases bit-identical to the CPU
     *                         oracle's "det" mode, which agrees with glibc on >= 99 % of canvases
     *   SPIRAL_MATH_FAST (2)  binary32 polynomials and a division-light dab loop (~4x shorter serial chain):
     *                         bit-identical to the oracle's "fast" mode; against the libm-faithful mode every canvas
     *                         stays within mean |err| <= 1e-3, and ~80 % (T = 20) within max |err| <= 1/255 -- the
     *                         rest differ in the random jitter of single dabs (profiles/r2_raster_tolerance.json) */
    int32_t math_mode;
} SpiralEnvCfg;

#define SPIRAL_MATH_DET 1
#define SPIRAL_MATH_FAST 2

typedef struct SpiralEnv SpiralEnv;

/* Validates cfg, precomputes brush constants, uploads tables to `device`. */
int spiral_env_create(const SpiralEnvCfg *cfg, SpiralEnv **out);
int spiral_env_destroy(SpiralEnv *env);

/* envs/mnist.py:81-105.  canvas u16[B,64,64,4] <- opaque white (fix15 1.0), brush/env state
 * re-initialised (new Brush => RNG re-seeded with 1000), obs f32[B,64,64,C] <- 255. */
int spiral_env_reset(SpiralEnv *env, uint16_t *canvas_dev, float *brush_state_dev,
                     double *env_state_dev, float *obs_dev, void *stream);

/* envs/mnist.py:231-260 for every canvas.  actions i32[B,K].  dab_scratch f32[B,max_dabs,8] and
 * dab_count i32[B,2] are scratch owned by the caller (contents undefined between calls).
 * status_dev i32[1] is OR-ed with 1 if any canvas overflowed max_dabs and with 2 if a canvas ran past
 * rng_table_len (result then invalid).
 * Small batches (det arithmetic: at most 2 x 148 SMs x 8 = 2368 canvases; fast: 5920; SPIRAL_FUSED_STEP_MAX overrides, 0 = never)
 * run as ONE kernel in which a second warp composites a canvas's dabs while its brush chain is still producing them;
 * larger batches run spiral_env_expand then spiral_env_composite.  Results are bit-identical either way.
 * (SPIRAL_FUSED_SMEM=0: the fused kernel blends in place instead of keeping the tile in shared memory.) */
int spiral_env_step(SpiralEnv *env, const int32_t *actions_dev, uint16_t *canvas_dev,
                    float *brush_state_dev, double *env_state_dev, float *dab_scratch_dev,
                    int32_t *dab_count_dev, float *obs_dev, int32_t *status_dev, void *stream);

/* the two halves of spiral_env_step, exposed for profiling and tests:
 *   expand    : stroke expansion (action -> events -> brush state machine -> pending dabs)
 *   composite : dab mask + fix15 blend into the canvas tile (in place) + read-back of the rows it
 *               dirtied.  obs_dev is PERSISTENT: it must be the buffer the previous step / reset wrote
 *               (canvases a step does not paint keep their observation without being touched). */
int spiral_env_expand(SpiralEnv *env, const int32_t *actions_dev, float *brush_state_dev,
                      double *env_state_dev, float *dab_scratch_dev, int32_t *dab_count_dev,
                      int32_t *status_dev, void *stream);
int spiral_env_composite(SpiralEnv *env, uint16_t *canvas_dev, const float *dab_scratch_dev,
                         const int32_t *dab_count_dev, float *obs_dev, void *stream);

/* test hook: finalised dabs (x, y, radius, hardness, opacity) of the pending list as they are
 * handed to the tile, f32[B,max_dabs,5] (opacity as float), count per canvas in out_count i32[B] */
int spiral_env_debug_dabs(SpiralEnv *env, const float *dab_scratch_dev, const int32_t *dab_count_dev,
                          float *out_dev, int32_t *out_count_dev, void *stream);

/* envs/mnist.py:235-243 + envs/utils.py:7-8: reward[b] = 1 - ||obs_b - target_b||_2 / n,
 * n = 64*64*C elements per canvas, both in 0..255 units. */
int spiral_l2_reward(const float *obs_dev, const float *target_dev, float *reward_dev,
                     int32_t batch, int32_t n, void *stream);

/* deterministic math used by the expansion kernel (tests compare with the oracle's copy) */
int spiral_debug_detmath(const double *x_dev, double *exp_out_dev, double *log_out_dev, int32_t n,
                         void *stream);
/* the binary32 functions of SPIRAL_MATH_FAST (csrc/fastmath.cuh): exp(x), exp(-x), log(x) (0 where x <= 0) and
 * hypot(x, x_rev) with x_rev[i] = x[n-1-i]; tests compare the bits with the oracle's copy */
/* Index checks of the checked build (libspiral_b200_checked.so, -DSPIRAL_BOUNDS_CHECK: every data-dependent shared-memory,
 * tile, list and table index of the rasteriser and of the learner's convolution kernels is tested against its bound):
 * number of violations counted on the device since the library was loaded; -1 in the product build (no checks compiled in). */
int64_t spiral_debug_bounds_violations(void);
/* the dab loop's branch-free division (csrc/fastmath.cuh div_exact) against the IEEE division over n_pairs pseudo-random
 * operand pairs; *mismatches_dev (u64, zeroed by the caller) receives the number of pairs whose quotients differ */
int spiral_debug_divcheck(uint64_t seed, int64_t n_pairs, uint64_t *mismatches_dev, void *stream);
int spiral_debug_fastmath(const float *x_dev, float *exp_out_dev, float *expneg_out_dev, float *log_out_dev,
                          float *hypot_out_dev, int32_t n, void *stream);

/* ------------------------------------------------------------------ */
/* A2C loss (K3)                                                        */
/* replaces: rl_utils.py:18-57 (discount, multiple_process_rollout) and */
/*           agent.py:161-196 (per-head log-softmax PG / value / entropy */
/*           loss) together with its gradient w.r.t. logits and vf      */
/* ------------------------------------------------------------------ */
#define SPIRAL_MAX_HEADS 8

/* logits_dev[k]: f32[B,T,N_k] (device pointers in a HOST array of n_heads entries);
 * actions i32[B,T,K]; rewards, values, vf: f32[B,T].
 * adv_out, ret_out: f32[B,T] (rl_utils.py:32-40 with bootstrap r = 0).
 * out_scalars_dev f32[4] = {pi_loss, vf_loss, entropy, total} exactly as agent.py:172-194 sums them.
 * dlogits_dev[k] f32[B,T,N_k] and dvf_dev f32[B,T] = d total / d input (may be NULL to skip).
 * partial_dev: scratch f32[B*T*4 + 2048], 16-byte aligned. */
int spiral_a2c_loss(const float *const *logits_dev, const int32_t *head_sizes, int32_t n_heads,
                    const int32_t *actions_dev, const float *rewards_dev, const float *values_dev,
                    const float *vf_dev, int32_t batch, int32_t T, double gamma, double lambda_,
                    float entropy_coeff, float *adv_out_dev, float *ret_out_dev,
                    float *out_scalars_dev, float *const *dlogits_dev, float *dvf_dev,
                    float *partial_dev, void *stream);

/* ------------------------------------------------------------------ */
/* Discriminator forward / reward (K2)                                  */
/* replaces: models/discriminator.py:51-95 build_model and :157-163     */
/*           predict (sigmoid(D(norm(x))), BN in training mode)          */
/* ------------------------------------------------------------------ */
#define SPIRAL_DISC_MAX_LAYERS 8

typedef struct {
    int32_t abi_version;
    int32_t device;
    int32_t max_batch;
    int32_t in_channels;          /* C, or 2C when conditional (discriminator.py:36-38) */
    int32_t screen_size;          /* 64 -> 6 conv layers */
    int32_t disc_dim;
    int32_t batch_norm;           /* args.disc_batch_norm */
    int32_t normalize;            /* 1: norm_fn = (x-127.5)/127.5 fused into layer 0; 0: norm_fn=None */
    int32_t precision;            /* 0 = fp32 SIMT, 1 = tcgen05 bf16x3 (fp32-equivalent), 2 = tcgen05 bf16 */
} SpiralDiscCfg;

typedef struct SpiralDisc SpiralDisc;

/* parameter pointers (device, fp32), TF layouts: conv kernel HWIO [5,5,Cin,Cout], bias [Cout],
 * BN gamma/beta [Cout] (layers >= 1), dense kernel [disc_dim*32, 1], dense bias [1] */
typedef struct {
    const float *conv_w[SPIRAL_DISC_MAX_LAYERS];
    const float *conv_b[SPIRAL_DISC_MAX_LAYERS];
    const float *bn_gamma[SPIRAL_DISC_MAX_LAYERS];
    const float *bn_beta[SPIRAL_DISC_MAX_LAYERS];
    const float *dense_w;
    const float *dense_b;
} SpiralDiscWeights;

int spiral_disc_create(const SpiralDiscCfg *cfg, SpiralDisc **out);
int spiral_disc_destroy(SpiralDisc *d);
/* bytes of device workspace spiral_disc_forward needs for `batch` images */
int64_t spiral_disc_workspace_bytes(const SpiralDisc *d, int32_t batch);
/* (re)packs the weights into the layout the kernels read; call after every optimiser step */
int spiral_disc_load_weights(SpiralDisc *d, const SpiralDiscWeights *w, void *workspace_dev,
                             void *stream);
/* x f32[B,64,64,Cin] NHWC in 0..255 (norm_fn = (x-127.5)/127.5 applied inside, base.py:51-52);
 * logits/probs f32[B]; bn_stats_dev f32[layers,2,max Cout] receives per-layer batch mean / biased
 * variance (may be NULL). */
int spiral_disc_forward(SpiralDisc *d, const float *x_dev, int32_t batch, float *logits_dev,
                        float *probs_dev, float *bn_stats_dev, void *workspace_dev, void *stream);

/* The same forward in stages, for GLOBAL-BATCH batch norm over ranks: models/discriminator.py:76-79 normalises with the
 * statistics of the batch it is fed, so a batch sharded over GPUs scores like the gathered batch when every layer's
 * per-channel (sum x, sum x^2) is summed over the ranks before it is used.  Call stage = 0, 1, ..., 5 in order on the same
 * workspace: stage s >= 1 first finishes layer s's batch norm from sums_dev (f64 [Cout_s][2], as left by stage s-1 and
 * all-reduced by the caller -- NCCL sum) with global_count = the summed number of elements per channel (sum over ranks of
 * batch * (32 >> s)^2), then runs layer s+1 and leaves ITS sums in sums_dev; stage 0 runs layers 0 and 1; stage 5
 * finishes layer 5 and writes logits / probs.  One rank without all-reduce == spiral_disc_forward, bit for bit. */
int spiral_disc_forward_staged(SpiralDisc *d, const float *x_dev, int32_t batch, int32_t stage, double global_count,
                               double *sums_dev, float *logits_dev, float *probs_dev, float *bn_stats_dev,
                               void *workspace_dev, void *stream);

/* ------------------------------------------------------------------ */
/* Convolution layer and its gradients (K2b)                            */
/* replaces: what tf.gradients(d_loss) executes for                     */
/*           models/discriminator.py:97-136 (WGAN-GP step: three        */
/*           shared-weight passes :44-47,:113-114 and the double        */
/*           backward of the penalty on d sigmoid(D(x^))/dx^ :116-120)   */
/*           through tf.layers.conv2d(5x5, stride 2, 'same') :65-72      */
/* ------------------------------------------------------------------ */
/* x f32[B,H,H,Cin] NHWC, w f32 HWIO [5,5,Cin,Cout], y f32[B,H/2,H/2,Cout]; no bias (the host adds it).
 * 'SAME' padding = 1 before / 2 after.  The three maps are bilinear and closed under differentiation,
 * so first- and second-order gradients of the whole stack are compositions of them.
 * precision: 0 = fp32 SIMT; 1 = tcgen05 bf16x3 (16 operand bits, what the reward forward uses) where the
 * channel counts allow (forward: Cin % 64 == 0 and Cout % 128 == 0; dgrad / wgrad: Cout % 64 == 0), else
 * fp32 SIMT; 2 = tcgen05 plain bf16; 3 = tcgen05 bf16x6 (24 operand bits: fp32-exact products, for gradients).  workspace: spiral_conv_workspace_bytes() bytes, caller-owned. */
int64_t spiral_conv_workspace_bytes(int32_t B, int32_t H, int32_t Cin, int32_t Cout, int32_t precision);
int spiral_conv_forward(const float *x_dev, const float *w_dev, float *y_dev, int32_t B, int32_t H,
                        int32_t Cin, int32_t Cout, int32_t precision, void *workspace_dev, void *stream);
/* dx f32[B,H,H,Cin] = gradient of sum(y * dy) w.r.t. x */
int spiral_conv_dgrad(const float *dy_dev, const float *w_dev, float *dx_dev, int32_t B, int32_t H,
                      int32_t Cin, int32_t Cout, int32_t precision, void *workspace_dev, void *stream);
/* dw f32 HWIO [5,5,Cin,Cout] = gradient of sum(y * dy) w.r.t. w */
int spiral_conv_wgrad(const float *x_dev, const float *dy_dev, float *dw_dev, int32_t B, int32_t H,
                      int32_t Cin, int32_t Cout, int32_t precision, void *workspace_dev, void *stream);

/* The same kernel serves several calls of one optimiser step (build_optim evaluates the shared-weight discriminator on real,
 * fake and the interpolates, models/discriminator.py:99-117, and tf.gradients walks each pass): spiral_conv_pack splits it into
 * the tensor-core operand planes ONCE (op 0: for the forward, op 1: for the input gradient), the *_packed calls take the planes.
 * spiral_conv_packed_bytes: size of the planes, 0 when this shape / precision runs on a path without packed kernels (then
 * use the plain entry points), -1 on a bad shape. */
int64_t spiral_conv_packed_bytes(int32_t op, int32_t B, int32_t H, int32_t Cin, int32_t Cout, int32_t precision);
int spiral_conv_pack(const float *w_dev, void *packed_dev, int32_t op, int32_t B, int32_t H, int32_t Cin, int32_t Cout,
                     int32_t precision, void *stream);
int spiral_conv_forward_packed(const float *x_dev, const void *packed_dev, float *y_dev, int32_t B, int32_t H, int32_t Cin,
                               int32_t Cout, int32_t precision, void *workspace_dev, void *stream);
int spiral_conv_dgrad_packed(const float *dy_dev, const void *packed_dev, float *dx_dev, int32_t B, int32_t H, int32_t Cin,
                             int32_t Cout, int32_t precision, void *workspace_dev, void *stream);

/* ------------------------------------------------------------------ */
/* Categorical sampling for the action heads (K4)                       */
/* replaces: models/policy.py:364-368 categorical_sample                */
/*           (tf.multinomial(logits - max, 1)), one call per head per   */
/*           env step (policy.py:302-303)                               */
/* ------------------------------------------------------------------ */
/* logits f32[rows, n] -> out_idx i32[rows] ~ softmax(logits).  Gumbel-max over Philox4x32-10
 * uniforms: key = seed, counter = (class / 4, row, offset + *offset_dev); a (seed, offset) pair must not be
 * reused.  offset_dev (may be NULL) is a device-resident call counter the caller increments, so that a captured
 * CUDA graph draws fresh numbers on every replay. */
int spiral_categorical_sample(const float *logits_dev, int32_t rows, int32_t n, uint64_t seed,
                              uint64_t offset, const uint64_t *offset_dev, int32_t *out_idx_dev,
                              void *stream);
/* test hook: one Philox4x32-10 block; ctr_key6 is a HOST array {c0,c1,c2,c3,k0,k1} */
int spiral_debug_philox(const uint32_t *ctr_key6_host, uint32_t *out4_dev, void *stream);

/* ------------------------------------------------------------------ */
/* Learner update (K5)                                                  */
/* replaces: agent.py:231-239 (tf.clip_by_global_norm(grads, 40) +      */
/*           tf.train.AdamOptimizer(policy_lr)) and                     */
/*           models/discriminator.py:124-126 (Adam, beta1 .5, beta2 .9) */
/* ------------------------------------------------------------------ */
/* Pointer tables are HOST arrays of device pointers, sizes in elements.
 * spiral_grad_global_norm: norm_out_dev f32[1] = sqrt(sum of squares of every tensor); scratch_dev f64[32 * pieces] with
 * pieces = sum over tensors of ceil(size / 2^20) (tensors are processed in pieces of 2^20 elements; = n_tensors when none is larger).
 * spiral_adam_step: g' = g * grad_scale * clip / max(norm * grad_scale, clip)  (skipped when norm_dev is NULL or
 * clip_norm <= 0);  m = b1 m + (1-b1) g';  v = b2 v + (1-b2) g'^2;  p -= lr_t * m / (sqrt(v) + eps), with
 * lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t) computed by the caller (TensorFlow's formulation).  No host sync. */
int spiral_grad_global_norm(const float *const *grads_dev, const int64_t *sizes, int32_t n_tensors,
                            double *scratch_dev, float *norm_out_dev, void *stream);
int spiral_adam_step(float *const *params_dev, const float *const *grads_dev, float *const *m_dev,
                     float *const *v_dev, const int64_t *sizes, int32_t n_tensors, float lr_t, float beta1,
                     float beta2, float eps, float grad_scale, const float *norm_dev, float clip_norm,
                     void *stream);
/* Same update with lr_t read from device memory (f32[1]): a learner step captured into a CUDA graph then sees the
 * step size of every replay (the bias-correction factor changes with t) instead of the value baked in at capture. */
int spiral_adam_step_lr_dev(float *const *params_dev, const float *const *grads_dev, float *const *m_dev,
                            float *const *v_dev, const int64_t *sizes, int32_t n_tensors, const float *lr_t_dev,
                            float beta1, float beta2, float eps, float grad_scale, const float *norm_dev,
                            float clip_norm, void *stream);

/* ------------------------------------------------------------------ */
/* Policy convolutions for the acting path                              */
/* replaces: what TensorFlow runs inside policy.act                     */
/*           (models/policy.py:196-215) for the trunk's convolutions:   */
/*           :94-99 conv 5x5, :103-109 conv 4x4/2 valid, :349-362       */
/*           res-block conv 3x3, :242-247/:275-281 conv2d_transpose     */
/*           4x4/2 (as four parity-class 2x2 convolutions), :290-295     */
/* ------------------------------------------------------------------ */
/* x bf16 NHWC [B,Hin,Win,Cin] (Cin in 1,2,3,6,16,32); w_packed bf16 [round8(Cout)][KH*KW*Cin rounded up to 16, + 8]
 * with k = (ky*KW + kx)*Cin + c (rows channel-permuted when Cout == 32: row nt*8 + j holds channel (j>>1)*8 + nt*2 + (j&1));
 * out = relu?(conv + bias + residual) + post_add, written bf16 (or fp32) at pixel (oy*o_sy + o_oy, ox*o_sx + o_ox)
 * of an [oH,oW] map per frame.  residual (bf16, output layout) and post_add (f32 [B,Cout]) may be NULL.
 * bf16 operands, fp32 accumulation (mma.sync). */
int spiral_pconv_forward(const void *x_bf16_dev, const void *w_packed_bf16_dev, const float *bias_dev,
                         const void *residual_bf16_dev, const float *post_add_dev, void *out_dev, int32_t B,
                         int32_t Hin, int32_t Win, int32_t Cin, int32_t Hout, int32_t Wout, int32_t Cout,
                         int32_t KH, int32_t KW, int32_t stride, int32_t pad_t, int32_t pad_l, int32_t relu,
                         int32_t oH, int32_t oW, int32_t o_sy, int32_t o_sx, int32_t o_oy, int32_t o_ox,
                         int32_t out_f32, void *stream);
/* Packs a convolution kernel for spiral_pconv_forward: one bf16 plane [round8(rows)][round16(n_taps * Ck) + 8] with
 * packed[row][tap * Ck + c] = w[row * s_row + c * s_c + tap_ky[tap] * s_ky + tap_kx[tap] * s_kx] (element strides of the fp32
 * tensor: any layout / view / subset of taps, e.g. one parity class of a transposed convolution); permute_rows (rows == 32):
 * row nt*8 + j holds channel (j>>1)*8 + nt*2 + (j&1), the order spiral_pconv_forward expects for 32 output channels. */
int spiral_pconv_pack(const float *w_dev, void *packed_bf16_dev, int32_t rows, int32_t Ck, int32_t n_taps, int64_t s_row,
                      int64_t s_c, int64_t s_ky, int64_t s_kx, const int32_t *tap_ky, const int32_t *tap_kx,
                      int32_t permute_rows, void *stream);

/* Fused res-block stack (policy.py:111-114, :255-258, :349-362): n_convs / 2 blocks of [conv3x3 -> ReLU -> conv3x3, + x]
 * on bf16 NHWC [B,S,S,32] maps with S = 6 or 8; w_packed = n_convs kernels packed as for spiral_pconv_forward
 * (Cout = 32, channel-permuted rows), bias f32 [n_convs][32].  Whole frames stay in shared memory across all layers. */
int spiral_pres_stack(const void *x_bf16_dev, const void *w_packed_bf16_dev, const float *bias_dev,
                      void *out_bf16_dev, int32_t B, int32_t S, int32_t n_convs, void *stream);

/* Location head, last stage fused (policy.py:275-281 last conv2d_transpose 4x4/2 + ReLU, :290-295 conv 3x3 -> 1 logits):
 * u bf16 NHWC [B,Hh,Hh,32] -> logits f32 [B,2Hh,2Hh].  w_up: the four parity-class kernels [4][32][2*2*32 + 8] packed as
 * for spiral_pconv_forward (class py*2+px, channel-permuted rows); w_out: [8][3*3*32 + 8] (row 0 = the kernel). */
int spiral_phead_last(const void *u_bf16_dev, const void *w_up_packed_dev, const float *bias_up_dev,
                      const void *w_out_packed_dev, const float *bias_out_dev, float *logits_dev, int32_t B, int32_t Hh,
                      void *stream);

/* ------------------------------------------------------------------ */
/* Policy convolutions for the learner (forward, input gradient, weight */
/* gradient), fp32 parity on bf16 tensor cores (bf16x3)                 */
/* replaces: what tf.gradients(loss) executes through the conv layers   */
/*           of models/policy.py:94-114 (trunk), :242-295 (location     */
/*           heads) when agent.py:161-239 evaluates the A2C loss        */
/* ------------------------------------------------------------------ */
/* Operand precision: n_planes = 2 ("bf16x3": operands split into two bf16 planes, three MMAs per k-step, 16 operand bits)
 * or 3 ("bf16x6": three planes, six MMAs, 24 bits = fp32-exact operands -- the default of the learner, whose trunk is
 * ~20 convolutions deep).
 * spiral_lconv_pack: an fp32 kernel -> n_planes contiguous bf16 planes [rows_pad][KS], rows_pad = rows rounded up to 8,
 * K = n_taps * Ck, KS = (K rounded up to 16) + 8:
 *   packed[row][tap * Ck + c] = w[row * s_row + c * s_c + tap_ky[tap] * s_ky + tap_kx[tap] * s_kx]
 * (strides in elements).  The strides and the tap table select the forward kernel of a convolution, its adjoint (input
 * gradient: flipped taps, rows <-> channels) or one parity class of a stride-2 transposed convolution.
 * spiral_lconv_packed_elems = rows_pad * KS (elements of ONE plane). */
int64_t spiral_lconv_packed_elems(int32_t Cin, int32_t Cout, int32_t n_taps);
int spiral_lconv_pack(const float *w_dev, void *packed_planes_dev, int32_t n_planes, int32_t rows, int32_t Ck, int32_t n_taps,
                      int64_t s_row, int64_t s_c, int64_t s_ky, int64_t s_kx, const int32_t *tap_ky, const int32_t *tap_kx,
                      void *stream);
/* out[b, oy*o_sy + o_oy, ox*o_sx + o_ox, co] = bias[co] + sum_{ky,kx,c} x[b, oy*stride - pad_t + ky, ox*stride - pad_l + kx, c]
 *                                                                        * packed[co][(ky*KW + kx)*Cin + c]
 * for oy < Hout, ox < Wout (positions outside [oH, oW] are skipped, input positions outside the map read 0), then max(., 0)
 * when relu != 0.  x f32 [B,Hin,Wi,Cin] NHWC, out f32 [B,oH,oW,Cout]; Cin in {1,2,3,6,8,16,32} (counts below 8 are padded to 8
 * inside), Cout <= 32; bias may be NULL.  n_planes selects the operand format: 4 = f16x3 (two fp16 planes, three
 * products), 3 = bf16x6, 2 = bf16x3; w_planes must come from spiral_lconv_pack with the same value. */
int spiral_lconv_forward(const float *x_dev, const void *w_planes_dev, int32_t n_planes, const float *bias_dev, float *out_dev,
                         int32_t B, int32_t Hin, int32_t Win, int32_t Cin, int32_t Hout, int32_t Wout, int32_t Cout,
                         int32_t KH, int32_t KW, int32_t stride, int32_t pad_t, int32_t pad_l, int32_t oH, int32_t oW,
                         int32_t o_sy, int32_t o_sx, int32_t o_oy, int32_t o_ox, int32_t relu, void *stream);
/* The four parity classes (c = py * 2 + px) of a transposed convolution's forward (tf.layers.conv2d_transpose 4x4 / stride 2,
 * policy.py:242-247, 275-281) or of a 4x4 / stride-2 convolution's input gradient in ONE launch: 2x2-tap / stride-1
 * convolutions over the same input with per-class padding (0 or 1), kernel and output phase --
 *   out[b, 2 oy + py, 2 ox + px, co] = bias[co] + sum x[b, oy - cls_pad_t[c] + ky, ox - cls_pad_l[c] + kx, ci] * packed_c[co][(ky*2 + kx)*Cin + ci]
 * for oy < Hout, ox < Wout (pixels outside [oH, oW] skipped).  w_planes: the classes' spiral_lconv_pack outputs back to back. */
int spiral_lconv_forward_classes(const float *x_dev, const void *w_planes_dev, int32_t n_planes, const float *bias_dev,
                                 float *out_dev, int32_t B, int32_t Hin, int32_t Win, int32_t Cin, int32_t Hout, int32_t Wout,
                                 int32_t Cout, const int32_t *cls_pad_t, const int32_t *cls_pad_l, int32_t oH, int32_t oW,
                                 int32_t relu, void *stream);
/* dw[(ky*KW + kx)*Cin + c][co] = sum_{b,oy,ox} x[b, oy*stride - pad_t + ky, ox*stride - pad_l + kx, c] * dy[b, oy, ox, co];
 * dy f32 [B,Hout,Wout,Cout], dw f32 [KH*KW*Cin, Cout]; workspace: spiral_lconv_wgrad_workspace_floats() floats. */
int64_t spiral_lconv_wgrad_workspace_floats(int32_t Cin, int32_t Cout, int32_t n_taps);
int spiral_lconv_wgrad(const float *x_dev, const float *dy_dev, float *dw_dev, float *workspace_dev, int32_t n_planes, int32_t B,
                       int32_t Hin, int32_t Win, int32_t Cin, int32_t Hout, int32_t Wout, int32_t Cout, int32_t KH, int32_t KW,
                       int32_t stride, int32_t pad_t, int32_t pad_l, void *stream);

#ifdef __cplusplus
}
#endif
#endif
