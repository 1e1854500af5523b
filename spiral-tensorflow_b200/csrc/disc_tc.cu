// disc_tc.cu -- tcgen05 / TMEM / TMA implicit-GEMM path of the discriminator (placeholder until
// the kernels land: creating a discriminator with precision != 0 fails loudly).
#include "disc.cuh"

namespace spiral {
int disc_tc_init(SpiralDisc *) { return spiral_set_error(SPIRAL_EUNSUPPORTED, "tensor-core discriminator path not built yet"); }
size_t disc_tc_workspace_bytes(const SpiralDisc *, int) { return 0; }
int disc_tc_pack_weights(SpiralDisc *, char *, cudaStream_t) { return SPIRAL_EUNSUPPORTED; }
int disc_tc_forward(SpiralDisc *, const float *, int, float *, float *, float *, char *, cudaStream_t) { return SPIRAL_EUNSUPPORTED; }
}  // namespace spiral
