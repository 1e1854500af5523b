// conv_tc.cu -- tcgen05 implicit GEMMs for the convolution layer's gradients (K2b).  (stub: filled in next)
#include "conv_ops.cuh"

namespace spiral {
bool conv_tc_supported(int, int, int, int, int) { return false; }
size_t conv_tc_workspace_bytes(int, int, int, int) { return 0; }
int conv_tc_forward(const float *, const float *, float *, int, int, int, int, bool, char *, cudaStream_t) { return -3; }
int conv_tc_dgrad(const float *, const float *, float *, int, int, int, int, bool, char *, cudaStream_t) { return -3; }
int conv_tc_wgrad(const float *, const float *, float *, int, int, int, int, bool, char *, cudaStream_t) { return -3; }
}  // namespace spiral
