// conv_ops.cu -- K2b: the discriminator's 5x5 / stride-2 / 'SAME' convolution as a standalone layer
// with its two gradients, for the WGAN-GP training step.
//
// Replaces what TensorFlow 1.6 runs for models/discriminator.py:97-136 (build_optim): tf.gradients of
// d_loss = critic + lambda * GP through three shared-weight passes of the conv stack (:44-47,
// :113-114), including the double backward of the penalty on d sigmoid(D(x^)) / d x^ (:116-120).
// The three bilinear maps
//
//     forward  y  = conv(x, w)            y [B,H/2,H/2,Cout]   x [B,H,H,Cin]   w HWIO [5,5,Cin,Cout]
//     dgrad    dx = conv^T(dy, w)         (gradient w.r.t. the layer input)
//     wgrad    dw = x (*) dy              (gradient w.r.t. the kernel)
//
// are closed under differentiation (the derivative of each is a combination of the other two), so
// first- and second-order gradients of the whole discriminator need nothing else that is dense;
// the host side (models/conv_ops.py) wires them as mutually recursive autograd functions.
//
// precision 0: fp32 SIMT kernels in this file (any channel count; layer 0 with Cin = 1/3/6 and
//              the disc_dim < 64 configurations always take this path).
// precision 1: tcgen05 bf16x3 implicit GEMMs (conv_tc.cu) when the channel counts allow it.
//
// TF semantics: NHWC, HWIO, 'SAME' padding for k=5, s=2 on even sizes = 1 before / 2 after.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/spiral_b200.h"
#include "common.h"
#include "conv_ops.cuh"

namespace spiral {

constexpr int kPx = 8;        // destination pixels per CTA
constexpr int kCk = 32;       // source channels staged per step
constexpr int kNt = 128;      // destination channels per CTA (one per thread)

// ------------------------------------------------------------------------------------------
// out[m, n] = sum_tap sum_c src[pix(m, tap), c] * wt[tap][c][n]
//   MODE 0 (forward): m = (img, oy, ox) on the Hd = Hs/2 grid, pix = (2oy-1+ky, 2ox-1+kx)
//   MODE 1 (dgrad)  : m = (img, iy, ix) on the Hd = 2*Hs grid, pix = ((iy+1-ky)/2, (ix+1-kx)/2)
//                     when both are integers; wt is the kernel with its channel axes swapped
//                     ([25][Cout][Cin], made by transpose_w_kernel)
// ------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kNt)
conv_taps_simt_kernel(const float *__restrict__ src, const float *__restrict__ wt, float *__restrict__ out,
                      int B, int Hs, int Cs, int Hd, int Cd)
{
    __shared__ float s_src[kPx][kCk];
    __shared__ int s_pix[kPx];
    const int tid = threadIdx.x;
    const int n = blockIdx.y * kNt + tid;
    const long long M = (long long)B * Hd * Hd;
    const long long m0 = (long long)blockIdx.x * kPx;
    float acc[kPx];
#pragma unroll
    for (int j = 0; j < kPx; j++) acc[j] = 0.0f;

    for (int tap = 0; tap < 25; tap++) {
        const int ky = tap / 5, kx = tap % 5;
        __syncthreads();
        if (tid < kPx) {
            const long long m = m0 + tid;
            int pix = -1;
            if (m < M) {
                const int img = (int)(m / (Hd * Hd));
                const int r = (int)(m % (Hd * Hd));
                const int dy = r / Hd, dx = r % Hd;
                int sy, sx;
                bool ok;
                if (MODE == 0) {
                    sy = 2 * dy - 1 + ky;
                    sx = 2 * dx - 1 + kx;
                    ok = true;
                } else {
                    const int ty = dy + 1 - ky, tx = dx + 1 - kx;
                    ok = ty >= 0 && tx >= 0 && !(ty & 1) && !(tx & 1);
                    sy = ty >> 1;
                    sx = tx >> 1;
                }
                if (ok && sy >= 0 && sy < Hs && sx >= 0 && sx < Hs) pix = (img * Hs + sy) * Hs + sx;
            }
            s_pix[tid] = pix;
        }
        __syncthreads();
        bool any = false;
#pragma unroll
        for (int j = 0; j < kPx; j++) any |= s_pix[j] >= 0;
        if (!any) continue;                                    // block-uniform
        for (int c0 = 0; c0 < Cs; c0 += kCk) {
            __syncthreads();
            for (int i = tid; i < kPx * kCk; i += kNt) {
                const int j = i / kCk, c = c0 + (i % kCk);
                const int pix = s_pix[j];
                s_src[j][i % kCk] = (pix >= 0 && c < Cs) ? src[(size_t)pix * Cs + c] : 0.0f;
            }
            __syncthreads();
            if (n < Cd) {
                const int cmax = min(kCk, Cs - c0);
                const float *wp = wt + ((size_t)tap * Cs + c0) * Cd + n;
                for (int c = 0; c < cmax; c++) {
                    const float w = __ldg(wp + (size_t)c * Cd);
#pragma unroll
                    for (int j = 0; j < kPx; j++) acc[j] = fmaf(s_src[j][c], w, acc[j]);
                }
            }
        }
    }
    if (n < Cd) {
#pragma unroll
        for (int j = 0; j < kPx; j++)
            if (m0 + j < M) out[(size_t)(m0 + j) * Cd + n] = acc[j];
    }
}

// HWIO [25][Cin][Cout] -> [25][Cout][Cin]
__global__ void transpose_w_kernel(const float *__restrict__ w, float *__restrict__ wt, int Cin, int Cout)
{
    const size_t total = (size_t)25 * Cin * Cout;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int ci = (int)(i % Cin);
        const int co = (int)((i / Cin) % Cout);
        const int tap = (int)(i / ((size_t)Cin * Cout));
        wt[i] = w[((size_t)tap * Cin + ci) * Cout + co];
    }
}

// ------------------------------------------------------------------------------------------
// wgrad: dw[tap][ci][co] = sum_p x[pix(p, tap), ci] * dy[p, co],  p = (img, oy, ox) in order
// CTA = one tap x kPx input channels x kNt output channels; the pixel range is split over
// gridDim.z / 25 slices whose partial sums are reduced in a fixed order (deterministic).
// ------------------------------------------------------------------------------------------
constexpr int kWgPix = 32;

__global__ void __launch_bounds__(kNt)
conv_wgrad_simt_kernel(const float *__restrict__ x, const float *__restrict__ dy, float *__restrict__ part,
                       int B, int H, int Cin, int Cout, int n_slices)
{
    __shared__ float s_x[kWgPix][kPx];
    __shared__ int s_pix[kWgPix];
    const int tid = threadIdx.x;
    const int co = blockIdx.x * kNt + tid;
    const int ci0 = blockIdx.y * kPx;
    const int tap = blockIdx.z % 25, slice = blockIdx.z / 25;
    const int ky = tap / 5, kx = tap % 5;
    const int Ho = H / 2;
    const long long P = (long long)B * Ho * Ho;
    const long long per = (P + n_slices - 1) / n_slices;
    const long long p_begin = slice * per, p_end = min(P, p_begin + per);
    float acc[kPx];
#pragma unroll
    for (int j = 0; j < kPx; j++) acc[j] = 0.0f;
    for (long long p0 = p_begin; p0 < p_end; p0 += kWgPix) {
        __syncthreads();
        if (tid < kWgPix) {
            const long long p = p0 + tid;
            int pix = -1;
            if (p < p_end) {
                const int img = (int)(p / (Ho * Ho));
                const int r = (int)(p % (Ho * Ho));
                const int sy = 2 * (r / Ho) - 1 + ky, sx = 2 * (r % Ho) - 1 + kx;
                if (sy >= 0 && sy < H && sx >= 0 && sx < H) pix = (img * H + sy) * H + sx;
            }
            s_pix[tid] = pix;
        }
        __syncthreads();
        for (int i = tid; i < kWgPix * kPx; i += kNt) {
            const int q = i / kPx, j = i % kPx;
            const int pix = s_pix[q];
            s_x[q][j] = (pix >= 0 && ci0 + j < Cin) ? x[(size_t)pix * Cin + ci0 + j] : 0.0f;
        }
        __syncthreads();
        if (co < Cout) {
            const int qmax = (int)min((long long)kWgPix, p_end - p0);
            for (int q = 0; q < qmax; q++) {
                if (s_pix[q] < 0) continue;                    // block-uniform
                const float d = __ldg(dy + (size_t)(p0 + q) * Cout + co);
#pragma unroll
                for (int j = 0; j < kPx; j++) acc[j] = fmaf(s_x[q][j], d, acc[j]);
            }
        }
    }
    if (co < Cout) {
#pragma unroll
        for (int j = 0; j < kPx; j++)
            if (ci0 + j < Cin) part[((size_t)slice * 25 * Cin + (size_t)tap * Cin + ci0 + j) * Cout + co] = acc[j];
    }
}

__global__ void reduce_slices_kernel(const float *__restrict__ part, float *__restrict__ out, size_t n, int n_slices)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float s = 0.0f;
        for (int k = 0; k < n_slices; k++) s += part[(size_t)k * n + i];
        out[i] = s;
    }
}

static int wgrad_slices(int B, int H, int Cin, int Cout)
{
    const long long P = (long long)B * (H / 2) * (H / 2);
    const long long ctas = 25LL * ((Cin + kPx - 1) / kPx) * ((Cout + kNt - 1) / kNt);
    long long s = (148 * 4 + ctas - 1) / ctas;                // fill the machine
    const long long smax = (P + 4 * kWgPix - 1) / (4 * kWgPix);
    if (s > smax) s = smax;
    if (s < 1) s = 1;
    if (s > 64) s = 64;
    return (int)s;
}

size_t conv_simt_workspace_bytes(int B, int H, int Cin, int Cout)
{
    const size_t wt = (size_t)25 * Cin * Cout * sizeof(float);
    const size_t part = (size_t)wgrad_slices(B, H, Cin, Cout) * 25 * Cin * Cout * sizeof(float);
    return ((wt > part ? wt : part) + 255) & ~(size_t)255;
}

int conv_forward_simt(const float *x, const float *w, float *y, int B, int H, int Cin, int Cout, cudaStream_t st)
{
    const int Ho = H / 2;
    const long long M = (long long)B * Ho * Ho;
    dim3 grid((unsigned)((M + kPx - 1) / kPx), (Cout + kNt - 1) / kNt);
    conv_taps_simt_kernel<0><<<grid, kNt, 0, st>>>(x, w, y, B, H, Cin, Ho, Cout);
    return 0;
}

int conv_dgrad_simt(const float *dy, const float *w, float *dx, int B, int H, int Cin, int Cout, char *ws, cudaStream_t st)
{
    float *wt = reinterpret_cast<float *>(ws);
    transpose_w_kernel<<<148 * 2, 256, 0, st>>>(w, wt, Cin, Cout);
    const long long M = (long long)B * H * H;
    dim3 grid((unsigned)((M + kPx - 1) / kPx), (Cin + kNt - 1) / kNt);
    conv_taps_simt_kernel<1><<<grid, kNt, 0, st>>>(dy, wt, dx, B, H / 2, Cout, H, Cin);
    return 0;
}

int conv_wgrad_simt(const float *x, const float *dy, float *dw, int B, int H, int Cin, int Cout, char *ws, cudaStream_t st)
{
    const int slices = wgrad_slices(B, H, Cin, Cout);
    float *part = slices > 1 ? reinterpret_cast<float *>(ws) : dw;
    dim3 grid((Cout + kNt - 1) / kNt, (Cin + kPx - 1) / kPx, 25 * slices);
    conv_wgrad_simt_kernel<<<grid, kNt, 0, st>>>(x, dy, part, B, H, Cin, Cout, slices);
    if (slices > 1) reduce_slices_kernel<<<148 * 2, 256, 0, st>>>(part, dw, (size_t)25 * Cin * Cout, slices);
    return 0;
}

}  // namespace spiral

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
using namespace spiral;

// precision -> bf16 planes per operand: 1 = bf16x3 (2 planes), 2 = bf16 (1 plane), 3 = bf16x6 (3 planes)
static int precision_planes(int precision) { return precision == 3 ? 3 : (precision == 2 ? 1 : 2); }

static int check_shape(int B, int H, int Cin, int Cout, int precision)
{
    if (B <= 0 || H < 2 || (H & 1) || Cin <= 0 || Cout <= 0)
        return spiral_set_error(SPIRAL_EINVAL, "conv: bad shape B=%d H=%d Cin=%d Cout=%d (H must be even)", B, H, Cin, Cout);
    if (precision < 0 || precision > 3) return spiral_set_error(SPIRAL_EINVAL, "conv: precision must be 0, 1, 2 or 3");
    if ((long long)B * H * H * (long long)(Cin > Cout ? Cin : Cout) >= (1LL << 31) * 8)
        return spiral_set_error(SPIRAL_EINVAL, "conv: tensor too large");
    return SPIRAL_OK;
}

extern "C" int64_t spiral_conv_workspace_bytes(int32_t B, int32_t H, int32_t Cin, int32_t Cout, int32_t precision)
{
    if (check_shape(B, H, Cin, Cout, precision) != SPIRAL_OK) return -1;
    size_t n = conv_simt_workspace_bytes(B, H, Cin, Cout);
    if (precision != 0) {
        const size_t t = conv_tc_workspace_bytes(B, H, Cin, Cout);
        if (t > n) n = t;
        if (lwgrad2_image_layer_supported(Cin, Cout) && lwgrad2_image_layer_workspace_bytes() > n) n = lwgrad2_image_layer_workspace_bytes();
    }
    return (int64_t)n;
}

extern "C" int64_t spiral_conv_packed_bytes(int32_t op, int32_t B, int32_t H, int32_t Cin, int32_t Cout, int32_t precision)
{
    if (check_shape(B, H, Cin, Cout, precision) != SPIRAL_OK || (op != 0 && op != 1)) return -1;
    return (precision != 0 && conv_tc_supported(op, B, H, Cin, Cout)) ? (int64_t)conv_tc_packed_bytes(Cin, Cout) : 0;
}

extern "C" int spiral_conv_pack(const float *w, void *packed, int32_t op, int32_t B, int32_t H, int32_t Cin, int32_t Cout,
                                int32_t precision, void *stream)
{
    SPIRAL_REQUIRE(w); SPIRAL_REQUIRE(packed);
    if (spiral_conv_packed_bytes(op, B, H, Cin, Cout, precision) <= 0)
        return spiral_set_error(SPIRAL_EINVAL, "conv_pack: this shape / precision has no packed kernel (op %d Cin=%d Cout=%d precision %d)", op, Cin, Cout, precision);
    conv_tc_pack(w, (char *)packed, op, Cin, Cout, precision_planes(precision), (cudaStream_t)stream);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_conv_pack");
}

extern "C" int spiral_conv_forward_packed(const float *x, const void *packed, float *y, int32_t B, int32_t H, int32_t Cin,
                                          int32_t Cout, int32_t precision, void *ws, void *stream)
{
    SPIRAL_REQUIRE(x); SPIRAL_REQUIRE(packed); SPIRAL_REQUIRE(y); SPIRAL_REQUIRE(ws);
    if (spiral_conv_packed_bytes(0, B, H, Cin, Cout, precision) <= 0)
        return spiral_set_error(SPIRAL_EINVAL, "conv_forward_packed: this shape / precision has no packed kernel");
    const int rc = conv_tc_forward(x, nullptr, y, B, H, Cin, Cout, precision_planes(precision), (char *)ws, (cudaStream_t)stream,
                                   (const char *)packed);
    if (rc != SPIRAL_OK) return rc;
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_conv_forward_packed");
}

extern "C" int spiral_conv_dgrad_packed(const float *dy, const void *packed, float *dx, int32_t B, int32_t H, int32_t Cin,
                                        int32_t Cout, int32_t precision, void *ws, void *stream)
{
    SPIRAL_REQUIRE(dy); SPIRAL_REQUIRE(packed); SPIRAL_REQUIRE(dx); SPIRAL_REQUIRE(ws);
    if (spiral_conv_packed_bytes(1, B, H, Cin, Cout, precision) <= 0)
        return spiral_set_error(SPIRAL_EINVAL, "conv_dgrad_packed: this shape / precision has no packed kernel");
    const int rc = conv_tc_dgrad(dy, nullptr, dx, B, H, Cin, Cout, precision_planes(precision), (char *)ws, (cudaStream_t)stream,
                                 (const char *)packed);
    if (rc != SPIRAL_OK) return rc;
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_conv_dgrad_packed");
}

extern "C" int spiral_conv_forward(const float *x, const float *w, float *y, int32_t B, int32_t H, int32_t Cin,
                                   int32_t Cout, int32_t precision, void *ws, void *stream)
{
    SPIRAL_REQUIRE(x); SPIRAL_REQUIRE(w); SPIRAL_REQUIRE(y);
    int rc = check_shape(B, H, Cin, Cout, precision);
    if (rc != SPIRAL_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (precision != 0 && conv_tc_supported(0, B, H, Cin, Cout)) {
        SPIRAL_REQUIRE(ws);
        rc = conv_tc_forward(x, w, y, B, H, Cin, Cout, precision_planes(precision), (char *)ws, st);
        if (rc != SPIRAL_OK) return rc;
    } else {
        conv_forward_simt(x, w, y, B, H, Cin, Cout, st);
    }
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_conv_forward");
}

extern "C" int spiral_conv_dgrad(const float *dy, const float *w, float *dx, int32_t B, int32_t H, int32_t Cin,
                                 int32_t Cout, int32_t precision, void *ws, void *stream)
{
    SPIRAL_REQUIRE(dy); SPIRAL_REQUIRE(w); SPIRAL_REQUIRE(dx); SPIRAL_REQUIRE(ws);
    int rc = check_shape(B, H, Cin, Cout, precision);
    if (rc != SPIRAL_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (precision != 0 && conv_tc_supported(1, B, H, Cin, Cout)) {
        rc = conv_tc_dgrad(dy, w, dx, B, H, Cin, Cout, precision_planes(precision), (char *)ws, st);
        if (rc != SPIRAL_OK) return rc;
    } else {
        conv_dgrad_simt(dy, w, dx, B, H, Cin, Cout, (char *)ws, st);
    }
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_conv_dgrad");
}

extern "C" int spiral_conv_wgrad(const float *x, const float *dy, float *dw, int32_t B, int32_t H, int32_t Cin,
                                 int32_t Cout, int32_t precision, void *ws, void *stream)
{
    SPIRAL_REQUIRE(x); SPIRAL_REQUIRE(dy); SPIRAL_REQUIRE(dw); SPIRAL_REQUIRE(ws);
    int rc = check_shape(B, H, Cin, Cout, precision);
    if (rc != SPIRAL_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (precision != 0 && lwgrad2_image_layer_supported(Cin, Cout)) {
        // the image layer: policy_learn.cu's mma.sync wgrad over 32-channel slices of dy (bf16x6 unless the caller asked for less)
        rc = lwgrad2_image_layer(x, dy, dw, (float *)ws, precision_planes(precision) == 3 ? 3 : 2, B, H, Cin, Cout, st);
        if (rc != SPIRAL_OK) return rc;
    } else if (precision != 0 && conv_tc_supported(2, B, H, Cin, Cout)) {
        rc = conv_tc_wgrad(x, dy, dw, B, H, Cin, Cout, precision_planes(precision), (char *)ws, st);
        if (rc != SPIRAL_OK) return rc;
    } else {
        conv_wgrad_simt(x, dy, dw, B, H, Cin, Cout, (char *)ws, st);
    }
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_conv_wgrad");
}
