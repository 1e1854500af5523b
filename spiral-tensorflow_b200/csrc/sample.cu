// sample.cu -- K4: fused categorical sampling for the policy's action heads.
//
// Replaces models/policy.py:364-368 (categorical_sample: tf.multinomial(logits - max, 1) per head,
// called once per head per env step at policy.py:302-303).  TensorFlow's multinomial kernel draws
// with the Gumbel-max trick over Philox uniforms; this kernel does the same in one pass over the
// logits: idx = argmax_i (logit_i - log(-log(u_i))), u from Philox4x32-10 keyed by `seed`, counter
// = (i / 4, row, offset).  One warp per row, lanes stride over the N classes (N up to 4096 for the
// location heads), warp-shuffle arg-max, ties to the lowest index.  Sampling streams cannot match
// TensorFlow's op-seeded Philox schedule, so parity is distributional (tests) and bit-exact only
// against the in-repo oracle's restatement of this kernel.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/spiral_b200.h"
#include "common.h"

namespace spiral {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4])
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__global__ void __launch_bounds__(128)
categorical_sample_kernel(const float *__restrict__ logits, int rows, int n, uint32_t k0, uint32_t k1, uint64_t offset,
                          const uint64_t *__restrict__ offset_dev, int32_t *__restrict__ out)
{
    if (offset_dev) offset += *offset_dev;                 // device-resident call counter (CUDA-graph replays)
    const uint32_t off_lo = (uint32_t)offset, off_hi = (uint32_t)(offset >> 32);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 4 + warp;
    if (row >= rows) return;
    const float *lr = logits + (size_t)row * n;
    float best = -INFINITY;
    int best_i = 0x7fffffff;
    for (int q = lane; q * 4 < n; q += 32) {          // one Philox call = 4 classes
        uint32_t r[4];
        philox4x32_10((uint32_t)q, (uint32_t)row, off_lo, off_hi, k0, k1, r);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int i = q * 4 + j;
            if (i < n) {
                const float u = ((float)(r[j] >> 8) + 0.5f) * (1.0f / 16777216.0f);      // (0, 1), 24 bits
                const float v = lr[i] - logf(-logf(u));
                if (v > best) { best = v; best_i = i; }
            }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
    }
    if (lane == 0) out[row] = best_i;
}

__global__ void philox_debug_kernel(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t *out)
{
    uint32_t r[4];
    philox4x32_10(c0, c1, c2, c3, k0, k1, r);
    for (int j = 0; j < 4; j++) out[j] = r[j];
}

}  // namespace spiral

extern "C" int spiral_categorical_sample(const float *logits, int32_t rows, int32_t n, uint64_t seed, uint64_t offset,
                                         const uint64_t *offset_dev, int32_t *out_idx, void *stream)
{
    SPIRAL_REQUIRE(logits); SPIRAL_REQUIRE(out_idx);
    if (rows <= 0 || n <= 0) return spiral_set_error(SPIRAL_EINVAL, "categorical_sample: rows and n must be positive");
    spiral::categorical_sample_kernel<<<(rows + 3) / 4, 128, 0, (cudaStream_t)stream>>>(
        logits, rows, n, (uint32_t)seed, (uint32_t)(seed >> 32), offset, offset_dev, out_idx);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_categorical_sample");
}

extern "C" int spiral_debug_philox(const uint32_t *ctr_key6, uint32_t *out4_dev, void *stream)
{
    SPIRAL_REQUIRE(ctr_key6); SPIRAL_REQUIRE(out4_dev);
    spiral::philox_debug_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(ctr_key6[0], ctr_key6[1], ctr_key6[2], ctr_key6[3], ctr_key6[4],
                                                                ctr_key6[5], out4_dev);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_debug_philox");
}
