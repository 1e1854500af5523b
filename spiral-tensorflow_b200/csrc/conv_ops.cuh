// conv_ops.cuh -- internal interface between conv_ops.cu (C ABI, fp32 SIMT kernels) and
// conv_tc.cu (tcgen05 implicit GEMMs) for the standalone convolution layer and its gradients.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace spiral {

// op: 0 forward, 1 dgrad, 2 wgrad
bool conv_tc_supported(int op, int B, int H, int Cin, int Cout);
size_t conv_tc_workspace_bytes(int B, int H, int Cin, int Cout);
// packed: the kernel's operand planes from conv_tc_pack (nullptr: split `w` into the workspace)
size_t conv_tc_packed_bytes(int Cin, int Cout);
void conv_tc_pack(const float *w, char *packed, int op, int Cin, int Cout, int ns, cudaStream_t st);
int conv_tc_forward(const float *x, const float *w, float *y, int B, int H, int Cin, int Cout, int ns, char *ws,
                    cudaStream_t st, const char *packed = nullptr);
int conv_tc_dgrad(const float *dy, const float *w, float *dx, int B, int H, int Cin, int Cout, int ns, char *ws,
                  cudaStream_t st, const char *packed = nullptr);
int conv_tc_wgrad(const float *x, const float *dy, float *dw, int B, int H, int Cin, int Cout, int ns, char *ws,
                  cudaStream_t st);

// policy_learn.cu: kernel gradient of the image layer (Cin 1 / 2 / 3 / 6) on the mma.sync wgrad kernel; n_planes 2 = bf16x3, 3 = bf16x6
bool lwgrad2_image_layer_supported(int Cin, int Cout);
size_t lwgrad2_image_layer_workspace_bytes();
int lwgrad2_image_layer(const float *x, const float *dy, float *dw, float *workspace, int n_planes, int B, int H, int Cin, int Cout,
                        cudaStream_t st);

}  // namespace spiral
