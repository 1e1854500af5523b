// disc.cu -- K2: discriminator forward D(x) / reward for sm_100a.
//
// Replaces models/discriminator.py:51-95 (build_model: log2(S) x [conv 5x5 stride 2 'SAME' ->
// batch-norm (training mode, layers >= 1) -> leaky-ReLU(0.2)] -> dense -> sigmoid) and
// :157-163 (predict), with norm_fn = (x - 127.5) / 127.5 (envs/base.py:51-52) fused into the
// first layer's load.
//
// TF semantics kept: NHWC activations, HWIO kernels, 'SAME' padding for
// k=5, s=2 on even sizes = 1 before / 2 after, BN epsilon 1e-3 with the BIASED batch variance,
// statistics of the batch being scored (training=True even in predict()).
//
// This file holds the host object, the fp32 SIMT path (precision 0, also the in-library
// reference for the tensor-core path) and the kernels shared by both paths (BN finalise,
// dense + sigmoid).  The tcgen05 implicit-GEMM path lives in disc_tc.cu.
#include "disc.cuh"
#include "conv_ops.cuh"

#include <cuda_bf16.h>

#include <cmath>
#include <cstring>
#include <new>

namespace spiral {

// ------------------------------------------------------------------------------------------
// layer 0: direct convolution, (x - 127.5) / 127.5 on load, bias, no BN; raw output + (optional)
// nothing else -- the consumer applies leaky-ReLU on load.   One CTA = 8x8 output pixels.
// ------------------------------------------------------------------------------------------
// Thread = one output pixel x G (8 or 16) output channels (register tile), CTA = 8x8 pixels x Cout.
// SPLIT = true writes leaky-ReLU(out) directly as the bf16 hi/lo pair the tensor-core layer 1
// reads (no BN at layer 0), SPLIT = false writes the raw fp32 output.
template <bool SPLIT, int G>
__global__ void __launch_bounds__(256)
conv0_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias,
             float *__restrict__ out, __nv_bfloat16 *__restrict__ out_hi, __nv_bfloat16 *__restrict__ out_lo,
             int B, int H, int Cin, int Cout, int normalize)
{
    extern __shared__ float sm[];
    const int Ho = H / 2;
    const int tiles = Ho / 8;
    const int b = blockIdx.x / (tiles * tiles);
    const int trem = blockIdx.x % (tiles * tiles);
    const int oy0 = (trem / tiles) * 8, ox0 = (trem % tiles) * 8;
    float *s_in = sm;                         // [19][19][Cin] normalised, zero padded
    float *s_w = sm + ((19 * 19 * Cin + 3) & ~3);   // [25*Cin][Cout], 16-byte aligned
    const int tid = threadIdx.x;
    for (int i = tid; i < 19 * 19 * Cin; i += 256) {
        const int c = i % Cin;
        const int p = i / Cin;
        const int ly = p / 19, lx = p % 19;
        const int iy = 2 * oy0 - 1 + ly, ix = 2 * ox0 - 1 + lx;
        float v = 0.0f;
        if (iy >= 0 && iy < H && ix >= 0 && ix < H) {
            v = x[(((size_t)b * H + iy) * H + ix) * Cin + c];
            if (normalize) v = (v - 127.5f) / 127.5f;
        }
        s_in[i] = v;
    }
    for (int i = tid; i < 25 * Cin * Cout / 4; i += 256)
        reinterpret_cast<float4 *>(s_w)[i] = reinterpret_cast<const float4 *>(w)[i];
    __syncthreads();
    const int ngroups = Cout / G;                        // channel groups of G
    // warp = 32 pixels of one channel group (weights broadcast across the warp)
    for (int item = tid; item < 64 * ngroups; item += 256) {
        const int p = item & 63, cg = item >> 6;
        const int py = p >> 3, px = p & 7;
        float acc[G];
#pragma unroll
        for (int j = 0; j < G; j++) acc[j] = bias[cg * G + j];
        for (int ky = 0; ky < 5; ky++)
            for (int kx = 0; kx < 5; kx++) {
                const float *ip = s_in + ((2 * py + ky) * 19 + (2 * px + kx)) * Cin;
                const float *wp = s_w + ((ky * 5 + kx) * Cin) * Cout + cg * G;
                for (int c = 0; c < Cin; c++) {
                    const float a = ip[c];
#pragma unroll
                    for (int j = 0; j < G; j += 4) {
                        const float4 w4 = *reinterpret_cast<const float4 *>(wp + c * Cout + j);
                        acc[j] = fmaf(a, w4.x, acc[j]); acc[j + 1] = fmaf(a, w4.y, acc[j + 1]);
                        acc[j + 2] = fmaf(a, w4.z, acc[j + 2]); acc[j + 3] = fmaf(a, w4.w, acc[j + 3]);
                    }
                }
            }
        const size_t o = (((size_t)b * Ho + oy0 + py) * Ho + ox0 + px) * Cout + cg * G;
        if (SPLIT) {
            uint32_t h[G / 2], l[G / 2];
#pragma unroll
            for (int j = 0; j < G / 2; j++) {
                float a0 = acc[2 * j], a1 = acc[2 * j + 1];
                a0 = a0 > 0.f ? a0 : 0.2f * a0;
                a1 = a1 > 0.f ? a1 : 0.2f * a1;
                const __nv_bfloat16 h0 = __float2bfloat16_rn(a0), h1 = __float2bfloat16_rn(a1);
                const __nv_bfloat16 l0 = __float2bfloat16_rn(a0 - __bfloat162float(h0));
                const __nv_bfloat16 l1 = __float2bfloat16_rn(a1 - __bfloat162float(h1));
                h[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                l[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
            }
            uint4 *ph = reinterpret_cast<uint4 *>(out_hi + o), *pl = reinterpret_cast<uint4 *>(out_lo + o);
#pragma unroll
            for (int j = 0; j < G / 8; j++) {
                ph[j] = make_uint4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
                pl[j] = make_uint4(l[4 * j], l[4 * j + 1], l[4 * j + 2], l[4 * j + 3]);
            }
        } else {
            float4 *po = reinterpret_cast<float4 *>(out + o);
#pragma unroll
            for (int j = 0; j < G / 4; j++) po[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// generic layer (fp32 SIMT implicit GEMM): M = B*Ho*Wo pixels, N = Cout, K = 25*Cin.
// 64x64 tile per CTA, 4x4 outputs per thread, K chunks of 16 channels of one tap.
// The producer layer's BN + leaky-ReLU is applied while loading A:  a = lrelu(x*scale[c]+shift[c]).
// Epilogue: + bias, store raw, per-CTA per-channel (sum, sum of squares) partials for this
// layer's BN.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
conv_simt_kernel(const float *__restrict__ in, const float *__restrict__ in_scale,
                 const float *__restrict__ in_shift, const float *__restrict__ w,
                 const float *__restrict__ bias, float *__restrict__ out,
                 float *__restrict__ bn_partial, int B, int H, int Cin, int Cout)
{
    __shared__ __align__(16) float As[16][64 + 4];
    __shared__ __align__(16) float Bs[16][64 + 4];
    __shared__ float s_sum[16][64], s_sq[16][64];
    const int Ho = H / 2;
    const int M = B * Ho * Ho;
    const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    // A-load role: row ar (0..63), channel quad aq (0..3)
    const int ar = tid >> 2, aq = tid & 3;
    const int am = m0 + ar;
    const bool a_ok = am < M;
    const int ab = a_ok ? am / (Ho * Ho) : 0;
    const int arem = a_ok ? am % (Ho * Ho) : 0;
    const int aoy = arem / Ho, aox = arem % Ho;
    // B-load role: k row bk (0..15), cout quad bq (0..15)
    const int bk = tid >> 4, bq = tid & 15;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.0f;

    for (int tap = 0; tap < 25; tap++) {
        const int ky = tap / 5, kx = tap % 5;
        const int iy = 2 * aoy + ky - 1, ix = 2 * aox + kx - 1;
        const bool pix_ok = a_ok && iy >= 0 && iy < H && ix >= 0 && ix < H;
        const float *ip = in + (((size_t)ab * H + iy) * H + ix) * Cin;
        for (int c0 = 0; c0 < Cin; c0 += 16) {
            float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
            const int c = c0 + aq * 4;
            if (pix_ok && c < Cin) {
                av = *reinterpret_cast<const float4 *>(ip + c);
                const float4 sc = *reinterpret_cast<const float4 *>(in_scale + c);
                const float4 sh = *reinterpret_cast<const float4 *>(in_shift + c);
                av.x = fmaf(av.x, sc.x, sh.x); av.y = fmaf(av.y, sc.y, sh.y);
                av.z = fmaf(av.z, sc.z, sh.z); av.w = fmaf(av.w, sc.w, sh.w);
                av.x = av.x > 0.f ? av.x : 0.2f * av.x; av.y = av.y > 0.f ? av.y : 0.2f * av.y;
                av.z = av.z > 0.f ? av.z : 0.2f * av.z; av.w = av.w > 0.f ? av.w : 0.2f * av.w;
            }
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c0 + bk < Cin && n0 + bq * 4 < Cout)
                bv = *reinterpret_cast<const float4 *>(w + ((size_t)(tap * Cin + c0 + bk)) * Cout + n0 + bq * 4);
            __syncthreads();
            As[aq * 4 + 0][ar] = av.x; As[aq * 4 + 1][ar] = av.y;
            As[aq * 4 + 2][ar] = av.z; As[aq * 4 + 3][ar] = av.w;
            *reinterpret_cast<float4 *>(&Bs[bk][bq * 4]) = bv;
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < 16; kk++) {
                const float4 a4 = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
                const float4 b4 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
                const float a[4] = {a4.x, a4.y, a4.z, a4.w};
                const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
            }
        }
    }
    // epilogue
    float csum[4] = {0, 0, 0, 0}, csq[4] = {0, 0, 0, 0};
    const bool n_ok = n0 + tx * 4 < Cout;
    float4 bz = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n_ok) bz = *reinterpret_cast<const float4 *>(bias + n0 + tx * 4);
    const float bzz[4] = {bz.x, bz.y, bz.z, bz.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int m = m0 + ty * 4 + i;
        if (m < M && n_ok) {
            float4 o;
            o.x = acc[i][0] + bzz[0]; o.y = acc[i][1] + bzz[1];
            o.z = acc[i][2] + bzz[2]; o.w = acc[i][3] + bzz[3];
            *reinterpret_cast<float4 *>(out + (size_t)m * Cout + n0 + tx * 4) = o;
            csum[0] += o.x; csum[1] += o.y; csum[2] += o.z; csum[3] += o.w;
            csq[0] += o.x * o.x; csq[1] += o.y * o.y; csq[2] += o.z * o.z; csq[3] += o.w * o.w;
        }
    }
    if (bn_partial) {
#pragma unroll
        for (int j = 0; j < 4; j++) { s_sum[ty][tx * 4 + j] = csum[j]; s_sq[ty][tx * 4 + j] = csq[j]; }
        __syncthreads();
        if (tid < 64 && n0 + tid < Cout) {
            float s = 0.f, q = 0.f;
#pragma unroll
            for (int r = 0; r < 16; r++) { s += s_sum[r][tid]; q += s_sq[r][tid]; }
            float *p = bn_partial + ((size_t)blockIdx.x * Cout + n0 + tid) * 2;
            p[0] = s; p[1] = q;
        }
    }
}

// Batch-norm statistics in two deterministic float64 stages (no atomics):
//   bn_reduce_kernel    kBnSlices x (Cout / 32) CTAs each sum a contiguous slice of the per-tile partials
//   bn_finalize_kernel  (Cout / 32) CTAs sum the slices in a fixed order, then emit the affine the consumer
//                       applies on load:  scale = gamma * rsqrt(var + eps), shift = beta - mean * scale
// Loads are coalesced float2 (sum, sum of squares); CTA = 32 channels x 8 partial-groups.
constexpr int kBnSlices = 64;

__global__ void __launch_bounds__(256)
bn_reduce_kernel(const float *__restrict__ partial, int n_part, int Cout, double *__restrict__ sliced)
{
    __shared__ double s_s[8][32], s_q[8][32];
    const int cl = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    const int per = (n_part + kBnSlices - 1) / kBnSlices;
    const int i0 = blockIdx.y * per, i1 = min(n_part, i0 + per);
    double s = 0.0, q = 0.0;
    if (c < Cout) {
        for (int i = i0 + g; i < i1; i += 8) {
            const float2 v = *reinterpret_cast<const float2 *>(partial + ((size_t)i * Cout + c) * 2);
            s += (double)v.x;
            q += (double)v.y;
        }
    }
    s_s[g][cl] = s;
    s_q[g][cl] = q;
    __syncthreads();
    if (g != 0 || c >= Cout) return;
    for (int k = 1; k < 8; k++) { s += s_s[k][cl]; q += s_q[k][cl]; }
    sliced[((size_t)blockIdx.y * Cout + c) * 2] = s;
    sliced[((size_t)blockIdx.y * Cout + c) * 2 + 1] = q;
}

__global__ void __launch_bounds__(256)
bn_finalize_kernel(const double *__restrict__ sliced, int Cout, double count,
                   const float *__restrict__ gamma, const float *__restrict__ beta,
                   float eps, float *__restrict__ scale, float *__restrict__ shift,
                   float *__restrict__ stats, int stats_stride, double *__restrict__ sums_out)
{
    __shared__ double s_s[8][32], s_q[8][32];
    const int cl = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    double s = 0.0, q = 0.0;
    if (c < Cout) {
        for (int i = g; i < kBnSlices; i += 8) {
            s += sliced[((size_t)i * Cout + c) * 2];
            q += sliced[((size_t)i * Cout + c) * 2 + 1];
        }
    }
    s_s[g][cl] = s;
    s_q[g][cl] = q;
    __syncthreads();
    if (g != 0 || c >= Cout) return;
    for (int k = 1; k < 8; k++) { s += s_s[k][cl]; q += s_q[k][cl]; }
    if (sums_out) {          // staged forward: hand (sum x, sum x^2) to the caller (all-reduce across ranks), no affine yet
        sums_out[2 * c] = s;
        sums_out[2 * c + 1] = q;
        return;
    }
    const double mean = s / count;
    double var = q / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float inv = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * inv;
    scale[c] = sc;
    shift[c] = beta[c] - (float)mean * sc;
    if (stats) { stats[c] = (float)mean; stats[stats_stride + c] = (float)var; }
}

// the second half of bn_finalize_kernel from per-channel sums [Cout][2] (the rank's own, or all-reduced over ranks with
// `count` the global element count): same expressions, so one rank staged == one rank unstaged, bit for bit
__global__ void __launch_bounds__(128)
bn_affine_kernel(const double *__restrict__ sums, int Cout, double count, const float *__restrict__ gamma,
                 const float *__restrict__ beta, float eps, float *__restrict__ scale, float *__restrict__ shift,
                 float *__restrict__ stats, int stats_stride)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= Cout) return;
    const double s = sums[2 * c], q = sums[2 * c + 1];
    const double mean = s / count;
    double var = q / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float inv = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * inv;
    scale[c] = sc;
    shift[c] = beta[c] - (float)mean * sc;
    if (stats) { stats[c] = (float)mean; stats[stats_stride + c] = (float)var; }
}

size_t bn_partial_bytes(int n_part, int Cout)
{
    // per-tile float partials, then the kBnSlices x Cout x 2 float64 slice sums
    return disc_align((size_t)n_part * Cout * 2 * sizeof(float)) + disc_align((size_t)kBnSlices * Cout * 2 * sizeof(double));
}

void launch_bn_finalize(float *part, int n_part, int Cout, double count, const float *gamma, const float *beta, float eps,
                        float *scale, float *shift, float *stats, int stats_stride, cudaStream_t st)
{
    double *sliced = reinterpret_cast<double *>(reinterpret_cast<char *>(part) + disc_align((size_t)n_part * Cout * 2 * sizeof(float)));
    dim3 g1((Cout + 31) / 32, kBnSlices);
    bn_reduce_kernel<<<g1, 256, 0, st>>>(part, n_part, Cout, sliced);
    bn_finalize_kernel<<<(Cout + 31) / 32, 256, 0, st>>>(sliced, Cout, count, gamma, beta, eps, scale, shift, stats, stats_stride, nullptr);
}

void launch_bn_sums(float *part, int n_part, int Cout, double *sums, cudaStream_t st)
{
    double *sliced = reinterpret_cast<double *>(reinterpret_cast<char *>(part) + disc_align((size_t)n_part * Cout * 2 * sizeof(float)));
    dim3 g1((Cout + 31) / 32, kBnSlices);
    bn_reduce_kernel<<<g1, 256, 0, st>>>(part, n_part, Cout, sliced);
    bn_finalize_kernel<<<(Cout + 31) / 32, 256, 0, st>>>(sliced, Cout, 1.0, nullptr, nullptr, 0.0f, nullptr, nullptr, nullptr, 0, sums);
}

void launch_bn_affine(const double *sums, int Cout, double count, const float *gamma, const float *beta, float eps,
                      float *scale, float *shift, float *stats, int stats_stride, cudaStream_t st)
{
    bn_affine_kernel<<<(Cout + 127) / 128, 128, 0, st>>>(sums, Cout, count, gamma, beta, eps, scale, shift, stats, stats_stride);
}

__global__ void fill_affine_kernel(float *scale, float *shift, int n, float sc, float sh)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { scale[i] = sc; shift[i] = sh; }
}

// flatten (1x1xC) -> dense(1) -> logits, probs = sigmoid(logits); one warp per image
__global__ void dense_sigmoid_kernel(const float *__restrict__ in, const float *__restrict__ in_scale,
                                     const float *__restrict__ in_shift, const float *__restrict__ w,
                                     const float *__restrict__ bias, float *__restrict__ logits,
                                     float *__restrict__ probs, int B, int C)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= B) return;
    const float *x = in + (size_t)warp * C;
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) {
        float v = fmaf(x[c], in_scale[c], in_shift[c]);
        v = v > 0.f ? v : 0.2f * v;
        acc = fmaf(v, w[c], acc);
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        const float l = acc + bias[0];
        logits[warp] = l;
        if (probs) probs[warp] = 1.0f / (1.0f + expf(-l));
    }
}

// ------------------------------------------------------------------------------------------
// Wide layers of a NARROW discriminator (disc_dim 8 / 16: its last layers have Cin >= 64, Cout >= 128 on 4x4 .. 1x1 maps): plain
// GEMM shapes that the fp32 SIMT tile runs at ~33 TFLOP/s.  They take K2b's tcgen05 forward with three bf16 planes (bf16x6:
// fp32-exact operands, conv_tc.cu) between two small elementwise passes -- the maps are a few hundred MB:
//   bn_act_kernel        y = leaky_relu(x * scale[c] + shift[c])   (this path's layers store RAW outputs, the consumer normalises)
//   bias_partial_kernel  out += bias[c]; partial[tile of 64 rows][c] = (sum, sum of squares) for the layer's batch norm
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bn_act_kernel(const float *__restrict__ x, const float *__restrict__ scale, const float *__restrict__ shift, float *__restrict__ y,
              size_t n4, int C)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const int c = (int)((i * 4) % (size_t)C);
        float4 v = __ldg(reinterpret_cast<const float4 *>(x) + i);
        const float4 sc = __ldg(reinterpret_cast<const float4 *>(scale + c)), sh = __ldg(reinterpret_cast<const float4 *>(shift + c));
        v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
        v.x = v.x > 0.f ? v.x : 0.2f * v.x; v.y = v.y > 0.f ? v.y : 0.2f * v.y;
        v.z = v.z > 0.f ? v.z : 0.2f * v.z; v.w = v.w > 0.f ? v.w : 0.2f * v.w;
        reinterpret_cast<float4 *>(y)[i] = v;
    }
}

// grid (tiles of 64 rows, Cout / 64); thread = (row group tid / 64, channel tid % 64): fixed order, deterministic
__global__ void __launch_bounds__(256)
bias_partial_kernel(float *__restrict__ out, const float *__restrict__ bias, float *__restrict__ partial, long long rows, int C)
{
    __shared__ float s_s[4][64], s_q[4][64];
    const int c = blockIdx.y * 64 + (threadIdx.x & 63), rg = threadIdx.x >> 6;
    const long long r0 = (long long)blockIdx.x * 64;
    const float b = __ldg(bias + c);
    float sum = 0.0f, sq = 0.0f;
    for (int j = rg; j < 64; j += 4) {
        const long long r = r0 + j;
        if (r < rows) {
            const float v = out[r * C + c] + b;
            out[r * C + c] = v;
            sum += v;
            sq += v * v;
        }
    }
    s_s[rg][threadIdx.x & 63] = sum;
    s_q[rg][threadIdx.x & 63] = sq;
    __syncthreads();
    if (rg == 0 && partial) {
        const int k = threadIdx.x;
        float *dst = partial + ((size_t)blockIdx.x * C + c) * 2;
        dst[0] = (s_s[0][k] + s_s[1][k]) + (s_s[2][k] + s_s[3][k]);
        dst[1] = (s_q[0][k] + s_q[1][k]) + (s_q[2][k] + s_q[3][k]);
    }
}

// ------------------------------------------------------------------------------------------
// host object
// ------------------------------------------------------------------------------------------
size_t disc_align(size_t v) { return (v + 255) & ~(size_t)255; }

void disc_plan(const SpiralDisc *d, int batch, DiscPlan *pl)
{
    size_t off = 0;
    int H = d->cfg.screen_size;
    int Cin = d->cfg.in_channels;
    for (int l = 0; l < d->n_layers; l++) {
        const int Cout = d->cfg.disc_dim << l;
        const int Ho = H / 2;
        pl->H[l] = H; pl->Cin[l] = Cin; pl->Cout[l] = Cout;
        pl->act_off[l] = off; off += disc_align((size_t)batch * Ho * Ho * Cout * sizeof(float));
        const int mtiles = (batch * Ho * Ho + 63) / 64;
        pl->mtiles[l] = mtiles;
        pl->part_off[l] = off; off += bn_partial_bytes(mtiles, Cout);
        pl->scale_off[l] = off; off += disc_align((size_t)Cout * sizeof(float));
        pl->shift_off[l] = off; off += disc_align((size_t)Cout * sizeof(float));
        pl->pack_off[l] = off;
        if (lconv_disc_layer_supported(Cin, Cout)) off += disc_align(lconv_disc_pack_bytes(Cin, Cout));
        pl->tc_off[l] = off;
        if (d->cfg.disc_dim < 64 && l > 0 && !lconv_disc_layer_supported(Cin, Cout) && conv_tc_supported(0, batch, H, Cin, Cout))
            off += disc_align((size_t)batch * H * H * Cin * sizeof(float)) + disc_align(conv_tc_workspace_bytes(batch, H, Cin, Cout));
        H = Ho; Cin = Cout;
    }
    pl->simt_bytes = off;
}

int disc_conv0(SpiralDisc *d, const float *x, int batch, float *out, void *out_hi, void *out_lo, cudaStream_t st)
{
    const int H = d->cfg.screen_size, Cin = d->cfg.in_channels, Cout = d->cfg.disc_dim, Ho = H / 2;
    const size_t smem = (size_t)(((19 * 19 * Cin + 3) & ~3) + 25 * Cin * Cout) * sizeof(float);
    if (smem > 200 * 1024) return spiral_set_error(SPIRAL_EUNSUPPORTED, "layer-0 weights do not fit shared memory");
    const int grid = batch * (Ho / 8) * (Ho / 8);
    const float *w0 = (const float *)d->w.conv_w[0], *b0 = (const float *)d->w.conv_b[0];
    if (out_hi) {        // tensor-core path: disc_dim % 64 == 0
        cudaFuncSetAttribute(conv0_kernel<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        conv0_kernel<true, 16><<<grid, 256, smem, st>>>(x, w0, b0, nullptr, (__nv_bfloat16 *)out_hi, (__nv_bfloat16 *)out_lo,
                                                        batch, H, Cin, Cout, d->cfg.normalize);
    } else if (Cout % 16 == 0) {
        cudaFuncSetAttribute(conv0_kernel<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        conv0_kernel<false, 16><<<grid, 256, smem, st>>>(x, w0, b0, out, nullptr, nullptr, batch, H, Cin, Cout, d->cfg.normalize);
    } else {
        cudaFuncSetAttribute(conv0_kernel<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        conv0_kernel<false, 8><<<grid, 256, smem, st>>>(x, w0, b0, out, nullptr, nullptr, batch, H, Cin, Cout, d->cfg.normalize);
    }
    return SPIRAL_OK;
}

int disc_forward_simt(SpiralDisc *d, const float *x, int batch, float *logits, float *probs,
                      float *bn_stats, char *ws, cudaStream_t st, int stage, double count, double *sums)
{
    DiscPlan pl;
    disc_plan(d, batch, &pl);
    const SpiralDiscWeights &W = d->w;
    const int maxC = d->cfg.disc_dim << (d->n_layers - 1);
    // staged (global-batch batch norm, spiral_disc_forward_staged): stage s finishes layer s's batch norm from `sums`
    // (s >= 1), runs layer s + 1 (stage 0: layers 0 and 1) and leaves that layer's per-channel sums in `sums`
    const int l_first = stage <= 0 ? 0 : stage + 1;                                   // stage n-1: no layer left (empty range)
    const int l_last = stage < 0 ? d->n_layers - 1 : (stage + 1 < d->n_layers ? stage + 1 : d->n_layers - 1);
    if (stage >= 1)
        launch_bn_affine(sums, pl.Cout[stage], count, (const float *)W.bn_gamma[stage], (const float *)W.bn_beta[stage], 1e-3f,
                         (float *)(ws + pl.scale_off[stage]), (float *)(ws + pl.shift_off[stage]),
                         bn_stats ? bn_stats + (size_t)stage * 2 * maxC : nullptr, maxC, st);
    for (int l = l_first; l <= l_last; l++) {
        const int H = pl.H[l], Cin = pl.Cin[l], Cout = pl.Cout[l], Ho = H / 2;
        float *out = (float *)(ws + pl.act_off[l]);
        float *scale = (float *)(ws + pl.scale_off[l]);
        float *shift = (float *)(ws + pl.shift_off[l]);
        float *part = (float *)(ws + pl.part_off[l]);
        const bool bn = l > 0 && d->cfg.batch_norm;
        int n_part = pl.mtiles[l];                   // rows of the batch-norm partials this layer's kernel writes
        // (disc_dim 64 with precision 0 stays on the fp32 SIMT kernels throughout: it is the independent reference of the
        // tensor-core reward path in the tests)
        const bool narrow = d->cfg.disc_dim < 64;
        // measured at 65,536 images (profiles/r2_mnist_uncond_disc_layers.md): the 1 -> 8 image layer is faster on the
        // scalar kernel (4.9 vs 7.6 ms: 25 FMAs per output against a staging-bound MMA tile that is 7/8 padding); wider
        // image layers and the 32 -> 64 layer on 4x4 maps (half of an 8-wide tile is padding) likewise stay where they are faster
        if (l == 0 && narrow && Cin * Cout > 8 && lconv_disc_layer_supported(Cin, Cout)) {
            // image layer: normalisation (x - 127.5) / 127.5 applied while the tile is staged ('SAME' pads the normalised image)
            const bool nz = d->cfg.normalize != 0;
            const int rc = lconv_disc_layer(x, nz ? 2 : 0, nullptr, nullptr, 1.0f / 127.5f, -1.0f, (const float *)W.conv_w[0],
                                            (const float *)W.conv_b[0], out, nullptr, ws + pl.pack_off[0], batch, H, Cin, Cout,
                                            &n_part, st);
            if (rc != SPIRAL_OK) return rc;
        } else if (l == 0) {
            const int rc = disc_conv0(d, x, batch, out, nullptr, nullptr, st);
            if (rc != SPIRAL_OK) return rc;
        } else {
            const float *in = (const float *)(ws + pl.act_off[l - 1]);
            const float *isc = (const float *)(ws + pl.scale_off[l - 1]);
            const float *ish = (const float *)(ws + pl.shift_off[l - 1]);
            if (narrow && Cout <= 32 && lconv_disc_layer_supported(Cin, Cout)) {
                // narrow layer (Cin <= 32, Cout <= 128): the mma.sync implicit GEMM of policy_learn.cu (lconv2_kernel, f16x3)
                const int rc = lconv_disc_layer(in, 1, isc, ish, 0.0f, 0.0f, (const float *)W.conv_w[l], (const float *)W.conv_b[l],
                                                out, bn ? part : nullptr, ws + pl.pack_off[l], batch, H, Cin, Cout, &n_part, st);
                if (rc != SPIRAL_OK) return rc;
            } else if (narrow && conv_tc_supported(0, batch, H, Cin, Cout)) {
                // wide tail of a narrow discriminator: activation pass, tcgen05 forward (bf16x6), bias + batch-norm partials
                float *act = (float *)(ws + pl.tc_off[l]);
                char *tcws = ws + pl.tc_off[l] + disc_align((size_t)batch * H * H * Cin * sizeof(float));
                const size_t n4 = (size_t)batch * H * H * Cin / 4;
                const size_t want = (n4 + 255) / 256;
                bn_act_kernel<<<(unsigned)(want < 148 * 8 ? want : 148 * 8), 256, 0, st>>>(in, isc, ish, act, n4, Cin);
                const int rc = conv_tc_forward(act, (const float *)W.conv_w[l], out, batch, H, Cin, Cout, 3, tcws, st);
                if (rc != SPIRAL_OK) return rc;
                dim3 grid(pl.mtiles[l], Cout / 64);
                bias_partial_kernel<<<grid, 256, 0, st>>>(out, (const float *)W.conv_b[l], bn ? part : nullptr,
                                                          (long long)batch * Ho * Ho, Cout);
            } else {
                dim3 grid(pl.mtiles[l], (Cout + 63) / 64);
                conv_simt_kernel<<<grid, 256, 0, st>>>(in, isc, ish, (const float *)W.conv_w[l], (const float *)W.conv_b[l],
                                                       out, bn ? part : nullptr, batch, H, Cin, Cout);
            }
        }
        if (bn && stage >= 0) {
            launch_bn_sums(part, n_part, Cout, sums, st);                // the caller all-reduces, the next stage finishes
        } else if (bn) {
            launch_bn_finalize(part, n_part, Cout, (double)batch * Ho * Ho, (const float *)W.bn_gamma[l],
                               (const float *)W.bn_beta[l], 1e-3f, scale, shift,
                               bn_stats ? bn_stats + (size_t)l * 2 * maxC : nullptr, maxC, st);
        } else {
            fill_affine_kernel<<<(Cout + 127) / 128, 128, 0, st>>>(scale, shift, Cout, 1.0f, 0.0f);
        }
    }
    const int last = d->n_layers - 1;
    if (stage < 0 || stage == last)
        dense_sigmoid_kernel<<<(batch * 32 + 127) / 128, 128, 0, st>>>(
            (const float *)(ws + pl.act_off[last]), (const float *)(ws + pl.scale_off[last]),
            (const float *)(ws + pl.shift_off[last]), (const float *)W.dense_w, (const float *)W.dense_b, logits, probs,
            batch, pl.Cout[last]);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_disc_forward(fp32)");
}

}  // namespace spiral

using namespace spiral;

extern "C" int spiral_disc_create(const SpiralDiscCfg *cfg, SpiralDisc **out)
{
    if (!cfg || !out) return spiral_set_error(SPIRAL_EINVAL, "spiral_disc_create: null argument");
    *out = nullptr;
    if (cfg->abi_version != SPIRAL_ABI_VERSION) return spiral_set_error(SPIRAL_EINVAL, "ABI version mismatch");
    if (cfg->screen_size != 64) return spiral_set_error(SPIRAL_EUNSUPPORTED, "screen_size must be 64");
    if (cfg->disc_dim < 8 || (cfg->disc_dim & (cfg->disc_dim - 1)))
        return spiral_set_error(SPIRAL_EUNSUPPORTED, "disc_dim must be a power of two >= 8 (got %d)", cfg->disc_dim);
    if (cfg->in_channels < 1 || cfg->in_channels > 8) return spiral_set_error(SPIRAL_EINVAL, "in_channels must be in 1..8");
    if (cfg->max_batch < 1) return spiral_set_error(SPIRAL_EINVAL, "max_batch must be >= 1");
    if (cfg->precision < 0 || cfg->precision > 2) return spiral_set_error(SPIRAL_EINVAL, "precision must be 0, 1 or 2");
    SpiralDisc *d = new (std::nothrow) SpiralDisc();
    if (!d) return spiral_set_error(SPIRAL_EINVAL, "out of host memory");
    memset(d, 0, sizeof(*d));
    d->cfg = *cfg;
    d->n_layers = 6;       // int(log2(64))
    if (cfg->precision != 0) {
        const int rc = disc_tc_init(d);
        if (rc != SPIRAL_OK) { delete d; return rc; }
    }
    *out = d;
    return SPIRAL_OK;
}

extern "C" int spiral_disc_destroy(SpiralDisc *d)
{
    if (!d) return SPIRAL_OK;
    delete d;
    return SPIRAL_OK;
}

extern "C" int64_t spiral_disc_workspace_bytes(const SpiralDisc *d, int32_t batch)
{
    if (!d || batch < 1) return -1;
    DiscPlan pl;
    disc_plan(d, batch, &pl);
    size_t total = pl.simt_bytes;
    if (d->cfg.precision != 0) total += disc_tc_workspace_bytes(d, batch);
    return (int64_t)total;
}

extern "C" int spiral_disc_load_weights(SpiralDisc *d, const SpiralDiscWeights *w, void *workspace, void *stream)
{
    SPIRAL_REQUIRE(d); SPIRAL_REQUIRE(w); SPIRAL_REQUIRE(workspace);
    for (int l = 0; l < d->n_layers; l++) {
        if (!w->conv_w[l] || !w->conv_b[l]) return spiral_set_error(SPIRAL_EINVAL, "missing conv weights for layer %d", l);
        if (l > 0 && d->cfg.batch_norm && (!w->bn_gamma[l] || !w->bn_beta[l]))
            return spiral_set_error(SPIRAL_EINVAL, "missing batch-norm parameters for layer %d", l);
    }
    if (!w->dense_w || !w->dense_b) return spiral_set_error(SPIRAL_EINVAL, "missing dense weights");
    d->w = *w;
    d->have_weights = 1;
    if (d->cfg.precision != 0) return disc_tc_pack_weights(d, (char *)workspace, (cudaStream_t)stream);
    return SPIRAL_OK;
}

extern "C" int spiral_disc_forward(SpiralDisc *d, const float *x, int32_t batch, float *logits, float *probs,
                                   float *bn_stats, void *workspace, void *stream)
{
    SPIRAL_REQUIRE(d); SPIRAL_REQUIRE(x); SPIRAL_REQUIRE(logits); SPIRAL_REQUIRE(workspace);
    if (!d->have_weights) return spiral_set_error(SPIRAL_EINVAL, "spiral_disc_forward: call spiral_disc_load_weights first");
    if (batch < 1 || batch > d->cfg.max_batch) return spiral_set_error(SPIRAL_EINVAL, "batch %d outside 1..max_batch=%d", batch, d->cfg.max_batch);
    if (d->cfg.precision == 0)
        return disc_forward_simt(d, x, batch, logits, probs, bn_stats, (char *)workspace, (cudaStream_t)stream, -1, 0.0, nullptr);
    return disc_tc_forward(d, x, batch, logits, probs, bn_stats, (char *)workspace, (cudaStream_t)stream, -1, 0.0, nullptr);
}

extern "C" int spiral_disc_forward_staged(SpiralDisc *d, const float *x, int32_t batch, int32_t stage, double global_count,
                                          double *sums, float *logits, float *probs, float *bn_stats, void *workspace,
                                          void *stream)
{
    SPIRAL_REQUIRE(d); SPIRAL_REQUIRE(x); SPIRAL_REQUIRE(logits); SPIRAL_REQUIRE(workspace); SPIRAL_REQUIRE(sums);
    if (!d->have_weights) return spiral_set_error(SPIRAL_EINVAL, "spiral_disc_forward_staged: call spiral_disc_load_weights first");
    if (!d->cfg.batch_norm) return spiral_set_error(SPIRAL_EINVAL, "spiral_disc_forward_staged needs batch_norm (nothing to reduce without it)");
    if (batch < 1 || batch > d->cfg.max_batch) return spiral_set_error(SPIRAL_EINVAL, "batch %d outside 1..max_batch=%d", batch, d->cfg.max_batch);
    if (stage < 0 || stage >= d->n_layers) return spiral_set_error(SPIRAL_EINVAL, "stage %d outside 0..%d", stage, d->n_layers - 1);
    if (stage >= 1 && !(global_count > 0.0)) return spiral_set_error(SPIRAL_EINVAL, "global_count must be > 0");
    if (d->cfg.precision == 0)
        return disc_forward_simt(d, x, batch, logits, probs, bn_stats, (char *)workspace, (cudaStream_t)stream, stage, global_count, sums);
    return disc_tc_forward(d, x, batch, logits, probs, bn_stats, (char *)workspace, (cudaStream_t)stream, stage, global_count, sums);
}
