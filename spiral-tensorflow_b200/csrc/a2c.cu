// a2c.cu -- K3: fused A2C loss for sm_100a.
//
// Replaces rl_utils.py:18-57 (discount / multiple_process_rollout: returns and GAE with the
// reference's constant bootstrap r = 0) and agent.py:161-196 (per-head log-softmax policy
// gradient, value loss, entropy -- SUMS over batch and time, the value term added once per head)
// plus the gradient of the total w.r.t. every head's logits and the value output, which the
// reference gets from tf.gradients.
//
//   gae_kernel       one thread per canvas: reverse scans over T (float64 like numpy/lfilter)
//   a2c_row_kernel   one CTA per (b,t) row, looping over the K heads: 3 sweeps over the row's
//                    logits (max / sum-exp and entropy dot / gradient write); the row (<= 16 KB)
//                    stays in L1 between sweeps, so HBM sees each logit once in, once out.
//                    Block reductions are warp-shuffle trees + one smem hop.
//   a2c_final_kernel fixed-order float64 reduction of the per-row partials -> 4 scalars
//                    (deterministic: no atomics).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/spiral_b200.h"
#include "common.h"

namespace spiral {

struct A2CHeads {
    const float *logits[SPIRAL_MAX_HEADS];
    float *dlogits[SPIRAL_MAX_HEADS];
    int size[SPIRAL_MAX_HEADS];
    int K;
};

__global__ void gae_kernel(const float *__restrict__ rewards, const float *__restrict__ values,
                           float *__restrict__ adv, float *__restrict__ ret, int B, int T,
                           double gamma, double gl)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float *r = rewards + (size_t)b * T;
    const float *v = values + (size_t)b * T;
    // discount(x, g)[t] = x[t] + g * discount(x, g)[t+1]   (scipy.signal.lfilter on the reversed axis)
    double R = 0.0;          // rollout['r'] = 0.0: the bootstrap element appended to rewards
    double A = 0.0;
    double vnext = 0.0;      // vpred_t[:, T] = r = 0
    for (int t = T - 1; t >= 0; t--) {
        const double rt = (double)r[t], vt = (double)v[t];
        R = rt + gamma * R;
        const double delta = rt + gamma * vnext - vt;
        A = delta + gl * A;
        vnext = vt;
        ret[(size_t)b * T + t] = (float)R;
        adv[(size_t)b * T + t] = (float)A;
    }
}

constexpr int kRowThreads = 128;
#ifndef A2C_MIN_BLOCKS
#define A2C_MIN_BLOCKS 12      // 40 registers, 48 warps per SM, no spills (16 -> 32 registers with spills)
#endif
#ifndef A2C_UNROLL
#define A2C_UNROLL 4
#endif
constexpr int kGradUnroll = A2C_UNROLL;
// e^x for x <= 0 via ex2.approx: relative error ~1e-6 at |x| = 20, far inside the 2e-5 the parity tests hold.
// Three sweeps (the re-reads come from L1 / L2) with this exponential reach 96 % (B = 16384) / 88 % (B = 65536) of the
// measured HBM peak after the instruction-budget rewrite described at a2c_row_kernel; holding the row in registers
// (80 registers, 24 warps per SM) and a persistent bulk-async (cp.async.bulk + mbarrier) double-buffered variant were
// both slower on B200 (25 and 44 ms vs 19 ms at B = 65536 at the time): occupancy wins.
__device__ __forceinline__ float fast_exp(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));      // 2 instructions (__expf adds denormal handling)
    return y;
}

__device__ __forceinline__ float warp_max(float v)
{
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v)
{
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// all threads get the result
__device__ __forceinline__ float block_max(float v, float *s)
{
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = s[0];
#pragma unroll
    for (int i = 1; i < kRowThreads / 32; i++) r = fmaxf(r, s[i]);
    return r;
}
__device__ __forceinline__ float block_sum(float v, float *s)
{
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = s[0];
#pragma unroll
    for (int i = 1; i < kRowThreads / 32; i++) r += s[i];
    return r;
}

// both sums of sweep 2 in one pass over the barrier pair
__device__ __forceinline__ float2 block_sum2(float a, float b, float2 *s)
{
    a = warp_sum(a);
    b = warp_sum(b);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = make_float2(a, b);
    __syncthreads();
    float2 r = s[0];
#pragma unroll
    for (int i = 1; i < kRowThreads / 32; i++) { r.x += s[i].x; r.y += s[i].y; }
    return r;
}

// Instruction budget (ncu, round 1: 44 thread-instructions per logit, issue slots 91 % busy at 67 % of DRAM
// throughput -- the kernel was issue-bound).  Two things took most of it and are gone: (1) the scalar heads
// (jump / pressure / size / colour: 2..8 logits) ran the block path -- three block reductions with two barriers each
// for 128 threads, ~45 % of a row's instructions for 0.2 % of its logits; a head of <= 32 logits is now one WARP's
// work (warp k % 4, lane = logit, shuffles only), concurrent with the other warps' heads; (2) the gradient sweep
// evaluates  g = p (a + c (ent - logS) + c z)  with the row constants hoisted (5 instructions per logit instead of
// 15) and the taken action's  -a  applied after the loop by the thread that wrote that element.
__global__ void __launch_bounds__(kRowThreads, A2C_MIN_BLOCKS)
a2c_row_kernel(const __grid_constant__ A2CHeads H, const int32_t *__restrict__ actions,
               const float *__restrict__ adv, const float *__restrict__ ret,
               const float *__restrict__ vf, float entropy_coeff, float *__restrict__ dvf,
               float *__restrict__ partial)
{
    __shared__ float s_red[kRowThreads / 32];
    __shared__ float2 s_red2[kRowThreads / 32];
    __shared__ float2 s_warp[kRowThreads / 32];
    const int row = blockIdx.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const float a = adv[row];
    float pi_acc = 0.0f, ent_acc = 0.0f;        // block-path heads: the same value in every thread
    float pi_w = 0.0f, ent_w = 0.0f;            // warp-path heads: the same value in every lane of the owning warp
    constexpr float kLog2e = 1.4426950408889634f;

    for (int k = 0; k < H.K; k++) {
        const int n = H.size[k];
        const float *x = H.logits[k] + (size_t)row * n;
        float *dx = H.dlogits[k] ? H.dlogits[k] + (size_t)row * n : nullptr;
        const int act = actions[(size_t)row * H.K + k];

        if (n <= 32) {
            // ---- scalar heads: one warp, one logit per lane ----
            if (warp != (k & (kRowThreads / 32 - 1))) continue;
            const bool on = lane < n;
            const float xv = on ? x[lane] : -INFINITY;
            const float m = warp_max(xv);
            const float z = on ? xv - m : 0.0f;
            const float e = on ? fast_exp(z) : 0.0f;
            const float S = warp_sum(e);
            const float D = warp_sum(e * z);
            const float logS = logf(S);
            const float invS = 1.0f / S;
            const float ent = logS - D * invS;
            const float logp_a = __shfl_sync(0xffffffffu, z, act) - logS;
            pi_w += -(logp_a * a);
            ent_w += ent;
            if (dx && on) {
                const float p = e * invS;
                float g = a * p + entropy_coeff * p * ((z - logS) + ent);
                if (lane == act) g -= a;
                dx[lane] = g;
            }
            continue;
        }
        const bool vec = (n & 3) == 0;

        // sweep 1: max
        float m = -INFINITY;
        if (vec) {
            const float4 *x4 = reinterpret_cast<const float4 *>(x);
            for (int i = tid; i < n / 4; i += kRowThreads) {
                const float4 v = x4[i];
                m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
            }
        } else {
            for (int i = tid; i < n; i += kRowThreads) m = fmaxf(m, x[i]);
        }
        m = block_max(m, s_red);

        // sweep 2: S = sum e^(x-m),  D = sum e^(x-m) (x-m)
        float S = 0.0f, D = 0.0f;
        if (vec) {
            const float4 *x4 = reinterpret_cast<const float4 *>(x);
            for (int i = tid; i < n / 4; i += kRowThreads) {
                const float4 v = x4[i];
                const float z0 = v.x - m, z1 = v.y - m, z2 = v.z - m, z3 = v.w - m;
                const float e0 = fast_exp(z0), e1 = fast_exp(z1), e2 = fast_exp(z2), e3 = fast_exp(z3);
                S += (e0 + e1) + (e2 + e3);
                D += (e0 * z0 + e1 * z1) + (e2 * z2 + e3 * z3);
            }
        } else {
            for (int i = tid; i < n; i += kRowThreads) {
                const float z = x[i] - m;
                const float e = fast_exp(z);
                S += e;
                D += e * z;
            }
        }
        const float2 SD = block_sum2(S, D, s_red2);
        S = SD.x;
        D = SD.y;
        const float log2S = log2f(S);
        const float logS = log2S * 0.6931471805599453f;
        const float invS = 1.0f / S;
        // entropy = -sum p logp = logS - D/S ;  logp[act] = x[act] - m - logS
        const float ent = logS - D * invS;
        const float logp_a = x[act] - m - logS;
        pi_acc += -(logp_a * a);
        ent_acc += ent;

        // sweep 3: d total / d logit_j = adv (p_j - 1[j==act]) + coeff p_j (logp_j + ent)
        //        = p_j (c1 + coeff z_j) - adv 1[j==act],   p_j = 2^(z_j log2e - log2 S),  c1 = adv + coeff (ent - logS)
        if (dx) {
            const float c1 = a + entropy_coeff * (ent - logS);
            const float nl2S = -log2S;
#define A2C_GRAD(xv) ({ const float z_ = (xv) - m; float p_; \
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p_) : "f"(fmaf(z_, kLog2e, nl2S))); \
                        p_ * fmaf(entropy_coeff, z_, c1); })
            if (vec) {
                const float4 *x4 = reinterpret_cast<const float4 *>(x);
                float4 *d4 = reinterpret_cast<float4 *>(dx);
#pragma unroll kGradUnroll
                for (int i = tid; i < n / 4; i += kRowThreads) {
                    const float4 v = x4[i];
                    float4 g;
                    g.x = A2C_GRAD(v.x);
                    g.y = A2C_GRAD(v.y);
                    g.z = A2C_GRAD(v.z);
                    g.w = A2C_GRAD(v.w);
                    d4[i] = g;
                }
                if (((act >> 2) & (kRowThreads - 1)) == tid) dx[act] -= a;     // this thread wrote element `act`
            } else {
                for (int i = tid; i < n; i += kRowThreads) dx[i] = A2C_GRAD(x[i]);
                if ((act & (kRowThreads - 1)) == tid) dx[act] -= a;
            }
#undef A2C_GRAD
        }
    }
    // scalar heads were summed by their warps: combine in a fixed order
    if (lane == 0) s_warp[warp] = make_float2(pi_w, ent_w);
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int w = 0; w < kRowThreads / 32; w++) { pi_acc += s_warp[w].x; ent_acc += s_warp[w].y; }
        const float d = vf[row] - ret[row];
        // vf_loss = 0.5 * sum (vf - r)^2 is added (x 0.5) once per head: agent.py:186-190
        if (dvf) dvf[row] = 0.5f * (float)H.K * d;
        float4 out = make_float4(pi_acc, ent_acc, 0.5f * d * d, 0.0f);
        reinterpret_cast<float4 *>(partial)[row] = out;
    }
}

// stage 1: kFinalBlocks CTAs reduce contiguous chunks of the per-row partials (float64, fixed order)
constexpr int kFinalBlocks = 256;

__global__ void __launch_bounds__(256)
a2c_partial_kernel(const float *__restrict__ partial, int rows, double *__restrict__ scratch)
{
    __shared__ double s[3][256];
    const int per = (rows + kFinalBlocks - 1) / kFinalBlocks;
    const int r0 = blockIdx.x * per, r1 = min(rows, r0 + per);
    double pi = 0.0, en = 0.0, vf = 0.0;
    for (int i = r0 + threadIdx.x; i < r1; i += 256) {
        const float4 p = reinterpret_cast<const float4 *>(partial)[i];
        pi += (double)p.x;
        en += (double)p.y;
        vf += (double)p.z;
    }
    s[0][threadIdx.x] = pi; s[1][threadIdx.x] = en; s[2][threadIdx.x] = vf;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) {
            s[0][threadIdx.x] += s[0][threadIdx.x + o];
            s[1][threadIdx.x] += s[1][threadIdx.x + o];
            s[2][threadIdx.x] += s[2][threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x < 3) scratch[blockIdx.x * 4 + threadIdx.x] = s[threadIdx.x][0];
}

// stage 2: one CTA, fixed-order float64 reduction of the kFinalBlocks chunk sums -> 4 scalars
__global__ void __launch_bounds__(256)
a2c_final_kernel(const double *__restrict__ scratch, int K, float entropy_coeff, float *__restrict__ out)
{
    __shared__ double s[3][256];
    s[0][threadIdx.x] = scratch[threadIdx.x * 4 + 0];
    s[1][threadIdx.x] = scratch[threadIdx.x * 4 + 1];
    s[2][threadIdx.x] = scratch[threadIdx.x * 4 + 2];
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) {
            s[0][threadIdx.x] += s[0][threadIdx.x + o];
            s[1][threadIdx.x] += s[1][threadIdx.x + o];
            s[2][threadIdx.x] += s[2][threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double pi_loss = s[0][0], entropy = s[1][0];
        const double vf_loss = (double)K * s[2][0];      // self.vf_loss += vf_loss, once per head
        out[0] = (float)pi_loss;
        out[1] = (float)vf_loss;
        out[2] = (float)entropy;
        out[3] = (float)(pi_loss + 0.5 * vf_loss - entropy * (double)entropy_coeff);
    }
}

}  // namespace spiral

using namespace spiral;

extern "C" int spiral_a2c_loss(const float *const *logits, const int32_t *head_sizes, int32_t n_heads,
                               const int32_t *actions, const float *rewards, const float *values,
                               const float *vf, int32_t batch, int32_t T, double gamma, double lambda_,
                               float entropy_coeff, float *adv_out, float *ret_out, float *out_scalars,
                               float *const *dlogits, float *dvf, float *partial, void *stream)
{
    SPIRAL_REQUIRE(logits); SPIRAL_REQUIRE(head_sizes); SPIRAL_REQUIRE(actions); SPIRAL_REQUIRE(rewards);
    SPIRAL_REQUIRE(values); SPIRAL_REQUIRE(vf); SPIRAL_REQUIRE(adv_out); SPIRAL_REQUIRE(ret_out);
    SPIRAL_REQUIRE(out_scalars); SPIRAL_REQUIRE(partial);
    if (n_heads < 1 || n_heads > SPIRAL_MAX_HEADS)
        return spiral_set_error(SPIRAL_EINVAL, "spiral_a2c_loss: n_heads must be in 1..%d", SPIRAL_MAX_HEADS);
    if (batch <= 0 || T <= 0) return spiral_set_error(SPIRAL_EINVAL, "spiral_a2c_loss: batch and T must be > 0");
    A2CHeads H;
    H.K = n_heads;
    for (int k = 0; k < SPIRAL_MAX_HEADS; k++) { H.logits[k] = nullptr; H.dlogits[k] = nullptr; H.size[k] = 0; }
    for (int k = 0; k < n_heads; k++) {
        if (!logits[k] || head_sizes[k] < 1)
            return spiral_set_error(SPIRAL_EINVAL, "spiral_a2c_loss: head %d has no logits / bad size", k);
        H.logits[k] = logits[k];
        H.dlogits[k] = dlogits ? dlogits[k] : nullptr;
        H.size[k] = head_sizes[k];
    }
    cudaStream_t st = (cudaStream_t)stream;
    const double g = gamma, gl = gamma * lambda_;     // agent.py:340-341: gamma=0.99, lambda_=1.0
    gae_kernel<<<(batch + 127) / 128, 128, 0, st>>>(rewards, values, adv_out, ret_out, batch, T, g, gl);
    a2c_row_kernel<<<batch * T, kRowThreads, 0, st>>>(H, actions, adv_out, ret_out, vf, entropy_coeff, dvf, partial);
    // scratch for the two-stage reduction: the 2048 floats that follow the per-row partials
    double *scratch = reinterpret_cast<double *>(partial + (size_t)batch * T * 4);
    a2c_partial_kernel<<<kFinalBlocks, 256, 0, st>>>(partial, batch * T, scratch);
    a2c_final_kernel<<<1, 256, 0, st>>>(scratch, n_heads, entropy_coeff, out_scalars);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_a2c_loss");
}
