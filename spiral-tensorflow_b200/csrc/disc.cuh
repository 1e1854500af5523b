// disc.cuh -- host object of the discriminator, shared by disc.cu (fp32 SIMT path, common
// kernels) and disc_tc.cu (tcgen05 implicit-GEMM path).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/spiral_b200.h"
#include "common.h"

struct SpiralDisc {
    SpiralDiscCfg cfg;
    int n_layers;
    int have_weights;
    SpiralDiscWeights w;
    void *tc;                 // opaque state of the tensor-core path (disc_tc.cu)
};

namespace spiral {

struct DiscPlan {
    int H[SPIRAL_DISC_MAX_LAYERS], Cin[SPIRAL_DISC_MAX_LAYERS], Cout[SPIRAL_DISC_MAX_LAYERS];
    int mtiles[SPIRAL_DISC_MAX_LAYERS];
    size_t act_off[SPIRAL_DISC_MAX_LAYERS], part_off[SPIRAL_DISC_MAX_LAYERS];
    size_t scale_off[SPIRAL_DISC_MAX_LAYERS], shift_off[SPIRAL_DISC_MAX_LAYERS];
    size_t tc_off[SPIRAL_DISC_MAX_LAYERS];        // wide tail layers of a narrow discriminator on K2b's tcgen05 forward: activated input + its workspace
    size_t pack_off[SPIRAL_DISC_MAX_LAYERS];      // packed kernel planes of the narrow layers (lconv_disc_layer)
    size_t simt_bytes;
};

size_t disc_align(size_t v);
void disc_plan(const SpiralDisc *d, int batch, DiscPlan *pl);

// kernels shared by both paths
size_t bn_partial_bytes(int n_part, int Cout);
void launch_bn_finalize(float *part, int n_part, int Cout, double count, const float *gamma, const float *beta, float eps,
                        float *scale, float *shift, float *stats, int stats_stride, cudaStream_t st);
// staged forward (global-batch batch norm): per-channel (sum x, sum x^2) out, affine from (possibly all-reduced) sums in
void launch_bn_sums(float *part, int n_part, int Cout, double *sums, cudaStream_t st);
void launch_bn_affine(const double *sums, int Cout, double count, const float *gamma, const float *beta, float eps,
                      float *scale, float *shift, float *stats, int stats_stride, cudaStream_t st);
__global__ void fill_affine_kernel(float *scale, float *shift, int n, float sc, float sh);
__global__ void dense_sigmoid_kernel(const float *in, const float *in_scale, const float *in_shift, const float *w,
                                     const float *bias, float *logits, float *probs, int B, int C);

int disc_conv0(SpiralDisc *d, const float *x, int batch, float *out, void *out_hi, void *out_lo, cudaStream_t st);

// narrow layers (disc_dim 8 / 16) on the mma.sync implicit GEMM of policy_learn.cu
bool lconv_disc_layer_supported(int Cin, int Cout);
size_t lconv_disc_pack_bytes(int Cin, int Cout);
int lconv_disc_layer(const float *in, int in_mode, const float *in_scale, const float *in_shift, float in_a, float in_b,
                     const float *w_hwio, const float *bias, float *out, float *bn_partial, void *pack_scratch, int B, int H,
                     int Cin, int Cout, int *n_part, cudaStream_t st);

// layer 0 on tensor cores (disc_conv0_mma.cu)
size_t conv0_mma_pack_bytes(int Cin);
bool conv0_mma_supported(const SpiralDisc *d);
void conv0_mma_pack(const SpiralDisc *d, void *w_hi, void *w_lo, cudaStream_t st);
int conv0_mma_forward(const SpiralDisc *d, const float *x, int batch, const void *w_hi, const void *w_lo, void *out_hi,
                      void *out_lo, int x3, cudaStream_t st);

// tensor-core path (disc_tc.cu)
int disc_tc_init(SpiralDisc *d);
size_t disc_tc_workspace_bytes(const SpiralDisc *d, int batch);
int disc_tc_pack_weights(SpiralDisc *d, char *ws, cudaStream_t st);
int disc_tc_forward(SpiralDisc *d, const float *x, int batch, float *logits, float *probs, float *bn_stats,
                    char *ws, cudaStream_t st, int stage, double count, double *sums);

}  // namespace spiral
