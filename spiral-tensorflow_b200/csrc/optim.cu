// optim.cu -- K5: the learner's update step as multi-tensor kernels.
//
// Replaces what TensorFlow runs for agent.py:231-239 (tf.clip_by_global_norm(grads, 40) followed by
// tf.train.AdamOptimizer(policy_lr).apply_gradients) and models/discriminator.py:124-126
// (AdamOptimizer(disc_lr, beta1=0.5, beta2=0.9)):
//
//   grad_sumsq   sum of squares of every gradient tensor -> global norm, two deterministic float64 stages
//   adam_step    g' = g * grad_scale * clip / max(norm, clip)        (grad_scale = 1 / world after the all-reduce)
//                m = b1 m + (1 - b1) g';  v = b2 v + (1 - b2) g'^2;  p -= lr_t * m / (sqrt(v) + eps)
//                with TensorFlow's epsilon placement and lr_t = lr sqrt(1 - b2^t) / (1 - b1^t) (host side)
//
// One launch covers up to kMaxTensors parameter tensors (pointer table passed by value); the norm stays on
// the device, so a whole optimiser step issues no host synchronisation.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/spiral_b200.h"
#include "common.h"

namespace spiral {

constexpr int kMaxTensors = 48;
constexpr int kOptBlocksX = 32;
// A table entry is a PIECE of at most 2^20 elements of a tensor (32 CTAs each): the discriminator's 279 MB are two big
// kernels and a handful of small tensors, and with one entry per tensor its update ran on 32 SMs (7.7 ms for a 2 GB sweep).
constexpr long long kPiece = 1ll << 20;

struct TensorTable {
    float *p[kMaxTensors];
    const float *g[kMaxTensors];
    float *m[kMaxTensors];
    float *v[kMaxTensors];
    long long n[kMaxTensors];
    int count;
};

__global__ void __launch_bounds__(256)
grad_sumsq_kernel(const __grid_constant__ TensorTable T, double *__restrict__ partial)
{
    __shared__ double s[256];
    const int t = blockIdx.y;
    const float *g = T.g[t];
    const long long n = T.n[t];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const double x = (double)g[i];
        acc += x * x;
    }
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[(size_t)t * gridDim.x + blockIdx.x] = s[0];
}

__global__ void __launch_bounds__(256)
norm_final_kernel(const double *__restrict__ partial, int n_partial, float *__restrict__ norm_out)
{
    __shared__ double s[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n_partial; i += 256) acc += partial[i];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) norm_out[0] = (float)sqrt(s[0]);
}

__global__ void __launch_bounds__(256)
adam_step_kernel(const __grid_constant__ TensorTable T, float lr_val, const float *__restrict__ lr_dev, float b1, float b2,
                 float eps, float grad_scale, const float *__restrict__ norm, float clip_norm)
{
    const int t = blockIdx.y;
    const float lr_t = lr_dev ? lr_dev[0] : lr_val;        // device-resident step size: a captured CUDA graph sees each step's value
    float scale = grad_scale;
    if (norm && clip_norm > 0.0f) scale *= clip_norm / fmaxf(norm[0] * grad_scale, clip_norm);   // the norm of the SCALED gradient
    float *p = T.p[t], *m = T.m[t], *v = T.v[t];
    const float *g = T.g[t];
    const long long n = T.n[t];
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float gi = g[i] * scale;
        const float mi = b1 * m[i] + (1.0f - b1) * gi;
        const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        p[i] -= lr_t * mi / (sqrtf(vi) + eps);
    }
}

}  // namespace spiral

using namespace spiral;

extern "C" int spiral_grad_global_norm(const float *const *grads, const int64_t *sizes, int32_t n_tensors,
                                       double *scratch, float *norm_out, void *stream)
{
    SPIRAL_REQUIRE(grads); SPIRAL_REQUIRE(sizes); SPIRAL_REQUIRE(scratch); SPIRAL_REQUIRE(norm_out);
    if (n_tensors <= 0 || n_tensors > 1024)
        return spiral_set_error(SPIRAL_EINVAL, "grad_global_norm: 1..1024 tensors");
    cudaStream_t st = (cudaStream_t)stream;
    // pieces in tensor order: the partial sums, and so the norm, do not depend on how the launches are grouped
    long long n_pieces = 0;
    TensorTable T;
    T.count = 0;
    auto flush = [&]() {
        if (T.count == 0) return;
        dim3 grid(kOptBlocksX, T.count);
        grad_sumsq_kernel<<<grid, 256, 0, st>>>(T, scratch + (size_t)(n_pieces - T.count) * kOptBlocksX);
        T.count = 0;
    };
    for (int t = 0; t < n_tensors; t++) {
        if (sizes[t] < 0) return spiral_set_error(SPIRAL_EINVAL, "grad_global_norm: negative size");
        for (long long off = 0; off < sizes[t]; off += kPiece) {
            const int i = T.count++;
            T.g[i] = grads[t] + off; T.n[i] = sizes[t] - off < kPiece ? sizes[t] - off : kPiece;
            T.p[i] = nullptr; T.m[i] = nullptr; T.v[i] = nullptr;
            n_pieces++;
            if (T.count == kMaxTensors) flush();
        }
    }
    flush();
    if (n_pieces == 0) return spiral_set_error(SPIRAL_EINVAL, "grad_global_norm: every tensor is empty");
    norm_final_kernel<<<1, 256, 0, st>>>(scratch, (int)(n_pieces * kOptBlocksX), norm_out);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_grad_global_norm");
}

static int adam_step_impl(float *const *params, const float *const *grads, float *const *m, float *const *v,
                          const int64_t *sizes, int32_t n_tensors, float lr_t, const float *lr_t_dev, float beta1, float beta2,
                          float eps, float grad_scale, const float *norm_dev, float clip_norm, void *stream)
{
    SPIRAL_REQUIRE(params); SPIRAL_REQUIRE(grads); SPIRAL_REQUIRE(m); SPIRAL_REQUIRE(v); SPIRAL_REQUIRE(sizes);
    if (n_tensors <= 0) return spiral_set_error(SPIRAL_EINVAL, "adam_step: no tensors");
    cudaStream_t st = (cudaStream_t)stream;
    TensorTable T;
    T.count = 0;
    auto flush = [&]() {
        if (T.count == 0) return;
        dim3 grid(kOptBlocksX, T.count);
        adam_step_kernel<<<grid, 256, 0, st>>>(T, lr_t, lr_t_dev, beta1, beta2, eps, grad_scale, norm_dev, clip_norm);
        T.count = 0;
    };
    for (int t = 0; t < n_tensors; t++) {
        for (long long off = 0; off < sizes[t]; off += kPiece) {
            const int i = T.count++;
            T.p[i] = params[t] + off; T.g[i] = grads[t] + off; T.m[i] = m[t] + off; T.v[i] = v[t] + off;
            T.n[i] = sizes[t] - off < kPiece ? sizes[t] - off : kPiece;
            if (T.count == kMaxTensors) flush();
        }
    }
    flush();
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_adam_step");
}

extern "C" int spiral_adam_step(float *const *params, const float *const *grads, float *const *m, float *const *v,
                                const int64_t *sizes, int32_t n_tensors, float lr_t, float beta1, float beta2, float eps,
                                float grad_scale, const float *norm_dev, float clip_norm, void *stream)
{
    return adam_step_impl(params, grads, m, v, sizes, n_tensors, lr_t, nullptr, beta1, beta2, eps, grad_scale, norm_dev, clip_norm, stream);
}

extern "C" int spiral_adam_step_lr_dev(float *const *params, const float *const *grads, float *const *m, float *const *v,
                                       const int64_t *sizes, int32_t n_tensors, const float *lr_t_dev, float beta1, float beta2,
                                       float eps, float grad_scale, const float *norm_dev, float clip_norm, void *stream)
{
    SPIRAL_REQUIRE(lr_t_dev);
    return adam_step_impl(params, grads, m, v, sizes, n_tensors, 0.0f, lr_t_dev, beta1, beta2, eps, grad_scale, norm_dev, clip_norm, stream);
}
