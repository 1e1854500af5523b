// policy_conv.cu -- the policy network's convolutions for the ACTING path (rollouts), bf16 tensor cores.
//
// Replaces, for policy.act (models/policy.py:196-215, one call per env step on every actor), what TensorFlow runs for
// the trunk's convolutions: conv 5x5 'same' (policy.py:94-99), three conv 4x4 / stride 2 'valid' (:103-109), the
// res-blocks' conv 3x3 'same' (:349-362), and in the location heads the conv2d_transpose 4x4 / stride 2 'same'
// (:242-247, :275-281) and the final conv 3x3 -> 1 channel (:290-295).  Every one of them is a <= 32-channel
// convolution over a <= 64x64 map: K = taps * Cin <= 512, N = Cout <= 32 -- far too narrow for a TMA / tcgen05
// pipeline, and what cuDNN runs at ~35 TFLOP/s (tools/policy_act_probe.py).  One kernel covers them all:
//
//   * NHWC bf16 activations; a CTA computes an 8 x 16 tile of output pixels x all output channels of one frame
//   * the input tile (with halo, zero padded) is staged once in shared memory; the im2col is never built: each
//     thread gathers its mma.sync.m16n8k16 A fragment from the tile through a k -> offset table (as in
//     disc_conv0_mma.cu); B fragments come from the packed kernel [Cout][K] staged in shared memory
//   * epilogue: + bias, + residual (res-blocks), ReLU, + per-frame vector (the action embedding of policy.py:101),
//     bf16 (or fp32 for the logits) stores; output pixels may be scattered with a stride / offset, which turns the
//     four parity classes of a stride-2 transposed convolution into four 2x2 stride-1 convolutions
//
// Plain bf16 operands with fp32 accumulation: this is the actors' behaviour policy, whose samples need no fp32
// parity (the learner's logits / loss go through the fp32 PyTorch module and K3).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/spiral_b200.h"
#include "common.h"

namespace spiral {

constexpr int kPcTw = 16, kPcTh = 8;             // output tile: 8 warps x 16 pixels

struct PConvParams {
    int B, Hin, Win, Hout, Wout;                 // input map and the grid of outputs this launch computes, per frame
    int Cout, CoutPad;                           // CoutPad: multiple of 8, <= 32
    int KH, KW, stride, pad_t, pad_l;
    int K, KP, KS;                               // K = KH*KW*Cin; KP = K rounded up to 16; KS = KP + 8 (row stride, bf16)
    int IH, IW;                                  // staged input tile
    int tiles_x, tiles_y;
    int relu;
    int oH, oW, o_sy, o_sx, o_oy, o_ox;          // output pixel (oy, ox) -> (oy*o_sy + o_oy, ox*o_sx + o_ox) of an [oH, oW] map
    int out_f32;                                 // 1: fp32 output [.., Cout]; 0: bf16
    int permuted;                                // packed rows are channel-permuted (CoutPad == 32): 128-bit stores
};

__device__ __forceinline__ void pc_mma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int CIN>
__global__ void __launch_bounds__(256)
pconv_kernel(const __grid_constant__ PConvParams p, const __nv_bfloat16 *__restrict__ x, const __nv_bfloat16 *__restrict__ wpk,
             const float *__restrict__ bias, const __nv_bfloat16 *__restrict__ residual, const float *__restrict__ post_add,
             void *__restrict__ out)
{
    extern __shared__ __align__(16) uint8_t sm_raw[];
    uint16_t *s_w = reinterpret_cast<uint16_t *>(sm_raw);                     // [CoutPad][KS]
    const int n_w = p.CoutPad * p.KS;
    uint16_t *s_in = s_w + ((n_w + 7) & ~7);                                   // [IH][IW][CIN]
    const int n_in = p.IH * p.IW * CIN;
    int *s_koff = reinterpret_cast<int *>(s_in + ((n_in + 7) & ~7));           // [KP]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles = p.tiles_x * p.tiles_y;
    const int img = blockIdx.x / tiles;
    const int trem = blockIdx.x % tiles;
    const int oy0 = (trem / p.tiles_x) * kPcTh, ox0 = (trem % p.tiles_x) * kPcTw;

    // ---- stage: packed kernel (16-byte copies), k -> offset table, input tile ----
    {
        const uint4 *gw = reinterpret_cast<const uint4 *>(wpk);
        uint4 *sw = reinterpret_cast<uint4 *>(s_w);
        for (int i = tid; i < n_w / 8; i += 256) sw[i] = __ldg(gw + i);
        for (int k = tid; k < p.KP; k += 256) {
            int off = 0;
            if (k < p.K) {
                const int tap = k / CIN, c = k - tap * CIN;
                const int ky = tap / p.KW, kx = tap - ky * p.KW;
                off = (ky * p.IW + kx) * CIN + c;
            }
            s_koff[k] = off;
        }
        const uint16_t *xi = reinterpret_cast<const uint16_t *>(x) + (size_t)img * p.Hin * p.Win * CIN;
        const int iy0 = oy0 * p.stride - p.pad_t, ix0 = ox0 * p.stride - p.pad_l;
        if (CIN % 8 == 0) {
            constexpr int V = CIN / 8 > 0 ? CIN / 8 : 1;                       // uint4 per pixel
            for (int i = tid; i < p.IH * p.IW * V; i += 256) {
                const int pix = i / V, v = i - pix * V;
                const int ly = pix / p.IW, lx = pix - ly * p.IW;
                const int iy = iy0 + ly, ix = ix0 + lx;
                uint4 val = make_uint4(0, 0, 0, 0);
                if (iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win)
                    val = __ldg(reinterpret_cast<const uint4 *>(xi + ((size_t)iy * p.Win + ix) * CIN) + v);
                reinterpret_cast<uint4 *>(s_in)[i] = val;
            }
        } else {
            for (int i = tid; i < n_in; i += 256) {
                const int pix = i / CIN, c = i - pix * CIN;
                const int ly = pix / p.IW, lx = pix - ly * p.IW;
                const int iy = iy0 + ly, ix = ix0 + lx;
                uint16_t val = 0;
                if (iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win) val = __ldg(xi + ((size_t)iy * p.Win + ix) * CIN + c);
                s_in[i] = val;
            }
        }
    }
    __syncthreads();

    // ---- main loop: warp = output row `warp` of the tile, 16 pixels x CoutPad channels ----
    const int g = lane >> 2, t = lane & 3;
    const int base0 = ((warp * p.stride) * p.IW + g * p.stride) * CIN;         // pixel g
    const int base1 = base0 + 8 * p.stride * CIN;                               // pixel g + 8
    const int NT = p.CoutPad >> 3;
    float acc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; nt++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[nt][j] = 0.0f;

#pragma unroll 1
    for (int ks = 0; ks < p.KP / 16; ks++) {
        const int k0 = ks * 16 + 2 * t;
        uint32_t a[4];
        if (CIN % 2 == 0) {
            // k0 is even and CIN is even: k0 and k0+1 are adjacent channels of one tap -> one 32-bit load
            const int o0 = s_koff[k0], o8 = s_koff[k0 + 8];
            a[0] = *reinterpret_cast<const uint32_t *>(s_in + base0 + o0);
            a[1] = *reinterpret_cast<const uint32_t *>(s_in + base1 + o0);
            a[2] = *reinterpret_cast<const uint32_t *>(s_in + base0 + o8);
            a[3] = *reinterpret_cast<const uint32_t *>(s_in + base1 + o8);
            if (k0 >= p.K) { a[0] = 0; a[1] = 0; }                              // padded k (weights are zero there, keep NaN-free)
            if (k0 + 8 >= p.K) { a[2] = 0; a[3] = 0; }
        } else {
            const int o0 = s_koff[k0], o1 = s_koff[k0 + 1], o8 = s_koff[k0 + 8], o9 = s_koff[k0 + 9];
            a[0] = (uint32_t)s_in[base0 + o0] | ((uint32_t)s_in[base0 + o1] << 16);
            a[1] = (uint32_t)s_in[base1 + o0] | ((uint32_t)s_in[base1 + o1] << 16);
            a[2] = (uint32_t)s_in[base0 + o8] | ((uint32_t)s_in[base0 + o9] << 16);
            a[3] = (uint32_t)s_in[base1 + o8] | ((uint32_t)s_in[base1 + o9] << 16);
        }
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
            if (nt < NT) {
                const int wrow = (nt * 8 + g) * p.KS + k0;
                uint32_t b[2];
                b[0] = *reinterpret_cast<const uint32_t *>(s_w + wrow);
                b[1] = *reinterpret_cast<const uint32_t *>(s_w + wrow + 8);
                pc_mma(acc[nt], a, b);
            }
        }
    }

    // ---- epilogue ----
    const int oy = oy0 + warp;
    if (oy >= p.Hout) return;
    const size_t frame_out = (size_t)img * p.oH * p.oW;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const int ox = ox0 + g + 8 * half;
        if (ox >= p.Wout) continue;
        const size_t pix = frame_out + (size_t)(oy * p.o_sy + p.o_oy) * p.oW + (ox * p.o_sx + p.o_ox);
        if (p.permuted) {
            // thread t owns channels t*8 .. t*8+7 (column (nt, 2t+e) of the MMA <-> channel t*8 + nt*2 + e)
            float v[8];
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
                v[2 * nt] = acc[nt][2 * half] + __ldg(bias + t * 8 + 2 * nt);
                v[2 * nt + 1] = acc[nt][2 * half + 1] + __ldg(bias + t * 8 + 2 * nt + 1);
            }
            if (residual) {
                const uint4 r = __ldg(reinterpret_cast<const uint4 *>(residual + pix * p.Cout + t * 8));
                const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    v[2 * j] += __bfloat162float(__ushort_as_bfloat16((unsigned short)(rr[j] & 0xffffu)));
                    v[2 * j + 1] += __bfloat162float(__ushort_as_bfloat16((unsigned short)(rr[j] >> 16)));
                }
            }
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (p.relu) v[j] = fmaxf(v[j], 0.0f);
                if (post_add) v[j] += __ldg(post_add + (size_t)img * p.Cout + t * 8 + j);
            }
            if (p.out_f32) {
                float4 *o = reinterpret_cast<float4 *>(reinterpret_cast<float *>(out) + pix * p.Cout + t * 8);
                o[0] = make_float4(v[0], v[1], v[2], v[3]);
                o[1] = make_float4(v[4], v[5], v[6], v[7]);
            } else {
                uint32_t w[4];
#pragma unroll
                for (int j = 0; j < 4; j++)
                    w[j] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v[2 * j])) |
                           ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v[2 * j + 1])) << 16);
                *reinterpret_cast<uint4 *>(reinterpret_cast<__nv_bfloat16 *>(out) + pix * p.Cout + t * 8) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        } else {
            // generic (narrow) output: column (nt, 2t+e) <-> channel nt*8 + 2t + e
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
                if (nt >= NT) continue;
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int ch = nt * 8 + 2 * t + e;
                    if (ch >= p.Cout) continue;
                    float v = acc[nt][2 * half + e] + __ldg(bias + ch);
                    if (residual) v += __bfloat162float(residual[pix * p.Cout + ch]);
                    if (p.relu) v = fmaxf(v, 0.0f);
                    if (post_add) v += __ldg(post_add + (size_t)img * p.Cout + ch);
                    if (p.out_f32) reinterpret_cast<float *>(out)[pix * p.Cout + ch] = v;
                    else reinterpret_cast<__nv_bfloat16 *>(out)[pix * p.Cout + ch] = __float2bfloat16_rn(v);
                }
            }
        }
    }
}

}  // namespace spiral

using namespace spiral;

extern "C" int spiral_pconv_forward(const void *x_bf16, const void *w_packed_bf16, const float *bias, const void *residual_bf16,
                                    const float *post_add, void *out, int32_t B, int32_t Hin, int32_t Win, int32_t Cin,
                                    int32_t Hout, int32_t Wout, int32_t Cout, int32_t KH, int32_t KW, int32_t stride,
                                    int32_t pad_t, int32_t pad_l, int32_t relu, int32_t oH, int32_t oW, int32_t o_sy,
                                    int32_t o_sx, int32_t o_oy, int32_t o_ox, int32_t out_f32, void *stream)
{
    SPIRAL_REQUIRE(x_bf16); SPIRAL_REQUIRE(w_packed_bf16); SPIRAL_REQUIRE(bias); SPIRAL_REQUIRE(out);
    if (B <= 0 || Cout <= 0 || Cout > 32 || !(Cin == 1 || Cin == 2 || Cin == 3 || Cin == 6 || Cin == 16 || Cin == 32))
        return spiral_set_error(SPIRAL_EINVAL, "pconv: unsupported channels Cin=%d Cout=%d", Cin, Cout);
    PConvParams p;
    p.B = B; p.Hin = Hin; p.Win = Win; p.Hout = Hout; p.Wout = Wout;
    p.Cout = Cout; p.CoutPad = (Cout + 7) & ~7;
    p.KH = KH; p.KW = KW; p.stride = stride; p.pad_t = pad_t; p.pad_l = pad_l;
    p.K = KH * KW * Cin; p.KP = (p.K + 15) & ~15; p.KS = p.KP + 8;
    p.IH = (kPcTh - 1) * stride + KH; p.IW = (kPcTw - 1) * stride + KW;
    p.tiles_x = (Wout + kPcTw - 1) / kPcTw; p.tiles_y = (Hout + kPcTh - 1) / kPcTh;
    p.relu = relu;
    p.oH = oH; p.oW = oW; p.o_sy = o_sy; p.o_sx = o_sx; p.o_oy = o_oy; p.o_ox = o_ox;
    p.out_f32 = out_f32;
    p.permuted = (Cout == 32) ? 1 : 0;
    const size_t n_w = (size_t)p.CoutPad * p.KS, n_in = (size_t)p.IH * p.IW * Cin;
    const size_t smem = (((n_w + 7) & ~(size_t)7) + ((n_in + 7) & ~(size_t)7)) * 2 + (size_t)p.KP * 4;
    if (smem > 200 * 1024) return spiral_set_error(SPIRAL_EUNSUPPORTED, "pconv: tile does not fit shared memory (%zu bytes)", smem);
    const int grid = B * p.tiles_x * p.tiles_y;
    cudaStream_t st = (cudaStream_t)stream;
#define PCONV_LAUNCH(C)                                                                                                     \
    do {                                                                                                                    \
        static size_t attr = 0;                                                                                             \
        if (smem > attr) {                                                                                                  \
            const cudaError_t e = cudaFuncSetAttribute(pconv_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
            if (e != cudaSuccess) return spiral_cuda_error(e, "cudaFuncSetAttribute(pconv_kernel)");                        \
            attr = 200 * 1024;                                                                                              \
        }                                                                                                                   \
        pconv_kernel<C><<<grid, 256, smem, st>>>(p, (const __nv_bfloat16 *)x_bf16, (const __nv_bfloat16 *)w_packed_bf16, bias, \
                                                 (const __nv_bfloat16 *)residual_bf16, post_add, out);                     \
    } while (0)
    switch (Cin) {
    case 1: PCONV_LAUNCH(1); break;
    case 2: PCONV_LAUNCH(2); break;
    case 3: PCONV_LAUNCH(3); break;
    case 6: PCONV_LAUNCH(6); break;
    case 16: PCONV_LAUNCH(16); break;
    default: PCONV_LAUNCH(32); break;
    }
#undef PCONV_LAUNCH
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_pconv_forward");
}

// ------------------------------------------------------------------------------------------
// Fused res-block stack (policy.py:111-114 encoder, :255-258 location heads): n_blocks x [conv3x3 -> ReLU -> conv3x3,
// + x] on a tiny map (6x6 after the encoder, 8x8 in the heads), 32 channels.  As separate launches these are 16
// convolutions per stack whose CTAs are mostly empty (36 or 64 of 128 pixels) and whose activations bounce through
// L2 between every pair; here a CTA keeps several whole frames (7 at 6x6, 4 at 8x8: ~256 pixels = two m16 tiles
// per warp) in shared memory across ALL layers, streams each layer's packed kernel in with cp.async while the
// previous layer computes, and touches global memory once in and once out.
// ------------------------------------------------------------------------------------------
namespace spiral {

constexpr int kRsC = 32;
constexpr int kRsK = 9 * kRsC;                   // 288
constexpr int kRsKS = kRsK + 8;                  // 296: conflict-free row stride (bf16)
constexpr int kRsWElems = kRsC * kRsKS;          // one packed kernel

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int S>
__global__ void __launch_bounds__(256)
pres_stack_kernel(const __nv_bfloat16 *__restrict__ x, const __nv_bfloat16 *__restrict__ wpk, const float *__restrict__ bias,
                  __nv_bfloat16 *__restrict__ out, int B, int n_convs)
{
    constexpr int P = S + 2;                      // padded frame
    constexpr int F = 256 / (S * S);              // frames per CTA
    constexpr int FE = P * P * kRsC;              // elements per padded frame
    extern __shared__ __align__(16) uint8_t sm_raw[];
    uint16_t *s_x = reinterpret_cast<uint16_t *>(sm_raw);          // [F][P][P][32] block input / output (residual)
    uint16_t *s_t = s_x + F * FE;                                  // [F][P][P][32] after the first conv
    uint16_t *s_w = s_t + F * FE;                                  // [2][32][KS] double-buffered packed kernel
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int f0 = blockIdx.x * F;
    const int nf = min(F, B - f0);

    // zero both activation buffers (halo stays zero), start the first kernel's copy, load the frames
    for (int i = tid; i < 2 * F * FE / 8; i += 256) reinterpret_cast<uint4 *>(s_x)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < kRsWElems / 8; i += 256) cp_async16(s_w + i * 8, wpk + i * 8);
    __syncthreads();
    for (int i = tid; i < nf * S * S * (kRsC / 8); i += 256) {
        const int v = i % (kRsC / 8), pix = i / (kRsC / 8);
        const int f = pix / (S * S), r = pix % (S * S);
        const int y = r / S, xx = r % S;
        const uint4 val = __ldg(reinterpret_cast<const uint4 *>(x + ((size_t)(f0 + f) * S * S + r) * kRsC) + v);
        *reinterpret_cast<uint4 *>(s_x + f * FE + ((y + 1) * P + xx + 1) * kRsC + v * 8) = val;
    }

    // this warp's 32 pixels: two m16 tiles; lane (g, t) handles rows g and g+8 of each
    const int g = lane >> 2, t = lane & 3;
    int pbase[4];                                 // top-left of the 3x3 window in the padded frame, per fragment row
    int pint[4];                                  // interior offset of the pixel itself (for residual / stores)
    bool pok[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int p = warp * 32 + (r >> 1) * 16 + g + 8 * (r & 1);
        const int f = p / (S * S), q = p % (S * S);
        const int y = q / S, xx = q % S;
        pok[r] = p < nf * S * S;
        const int fo = (pok[r] ? f : 0) * FE;
        pbase[r] = fo + (y * P + xx) * kRsC;
        pint[r] = fo + ((y + 1) * P + xx + 1) * kRsC;
    }

    for (int l = 0; l < n_convs; l++) {
        cp_async_wait_all();
        __syncthreads();                          // kernel l is in s_w[l & 1]; the previous layer's stores are visible
        if (l + 1 < n_convs) {
            uint16_t *dst = s_w + ((l + 1) & 1) * kRsWElems;
            const __nv_bfloat16 *srcw = wpk + (size_t)(l + 1) * kRsWElems;
            for (int i = tid; i < kRsWElems / 8; i += 256) cp_async16(dst + i * 8, srcw + i * 8);
        }
        const uint16_t *src = (l & 1) ? s_t : s_x;       // conv1 reads x, conv2 reads t
        uint16_t *dstb = (l & 1) ? s_x : s_t;            // conv1 writes t, conv2 writes x (in place: + its own residual)
        const uint16_t *w = s_w + (l & 1) * kRsWElems;
        float acc[2][4][4];
#pragma unroll
        for (int m = 0; m < 2; m++)
#pragma unroll
            for (int nt = 0; nt < 4; nt++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[m][nt][j] = 0.0f;
#pragma unroll 1
        for (int ks = 0; ks < kRsK / 16; ks++) {
            const int k0 = ks * 16 + 2 * t;
            // k -> (tap, c): c = k % 32 (even), tap = k / 32; offsets inside the padded frame
            const int tap0 = k0 >> 5, c0 = k0 & 31, tap8 = (k0 + 8) >> 5, c8 = (k0 + 8) & 31;
            const int o0 = ((tap0 / 3) * P + tap0 % 3) * kRsC + c0, o8 = ((tap8 / 3) * P + tap8 % 3) * kRsC + c8;
            uint32_t a[2][4];
#pragma unroll
            for (int m = 0; m < 2; m++) {
                a[m][0] = *reinterpret_cast<const uint32_t *>(src + pbase[2 * m] + o0);
                a[m][1] = *reinterpret_cast<const uint32_t *>(src + pbase[2 * m + 1] + o0);
                a[m][2] = *reinterpret_cast<const uint32_t *>(src + pbase[2 * m] + o8);
                a[m][3] = *reinterpret_cast<const uint32_t *>(src + pbase[2 * m + 1] + o8);
            }
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
                const int wrow = (nt * 8 + g) * kRsKS + k0;
                uint32_t b[2];
                b[0] = *reinterpret_cast<const uint32_t *>(w + wrow);
                b[1] = *reinterpret_cast<const uint32_t *>(w + wrow + 8);
                pc_mma(acc[0][nt], a[0], b);
                pc_mma(acc[1][nt], a[1], b);
            }
        }
        __syncthreads();                          // everyone has read `src` (conv2 overwrites x in place only at own pixels,
                                                  // but conv1 of the NEXT block reads x's neighbours: keep the barrier simple)
        const float *bl = bias + (size_t)l * kRsC;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            if (!pok[r]) continue;
            const int m = r >> 1, half = r & 1;
            float v[8];
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
                v[2 * nt] = acc[m][nt][2 * half] + __ldg(bl + t * 8 + 2 * nt);
                v[2 * nt + 1] = acc[m][nt][2 * half + 1] + __ldg(bl + t * 8 + 2 * nt + 1);
            }
            uint16_t *o = dstb + pint[r] + t * 8;
            if (l & 1) {                          // second conv of the block: + x, no ReLU
                const uint4 rv = *reinterpret_cast<const uint4 *>(o);
                const uint32_t rr[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    v[2 * j] += __bfloat162float(__ushort_as_bfloat16((unsigned short)(rr[j] & 0xffffu)));
                    v[2 * j + 1] += __bfloat162float(__ushort_as_bfloat16((unsigned short)(rr[j] >> 16)));
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] = fmaxf(v[j], 0.0f);
            }
            uint32_t wv[4];
#pragma unroll
            for (int j = 0; j < 4; j++)
                wv[j] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v[2 * j])) |
                        ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v[2 * j + 1])) << 16);
            *reinterpret_cast<uint4 *>(o) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
        }
    }
    __syncthreads();
    for (int i = tid; i < nf * S * S * (kRsC / 8); i += 256) {
        const int v = i % (kRsC / 8), pix = i / (kRsC / 8);
        const int f = pix / (S * S), r = pix % (S * S);
        const int y = r / S, xx = r % S;
        const uint4 val = *reinterpret_cast<const uint4 *>(s_x + f * FE + ((y + 1) * P + xx + 1) * kRsC + v * 8);
        reinterpret_cast<uint4 *>(out + ((size_t)(f0 + f) * S * S + r) * kRsC)[v] = val;
    }
}

}  // namespace spiral

extern "C" int spiral_pres_stack(const void *x_bf16, const void *w_packed_bf16, const float *bias, void *out_bf16, int32_t B,
                                 int32_t S, int32_t n_convs, void *stream)
{
    SPIRAL_REQUIRE(x_bf16); SPIRAL_REQUIRE(w_packed_bf16); SPIRAL_REQUIRE(bias); SPIRAL_REQUIRE(out_bf16);
    if (B <= 0 || n_convs <= 0 || (n_convs & 1) || !(S == 6 || S == 8))
        return spiral_set_error(SPIRAL_EINVAL, "pres_stack: S must be 6 or 8 and n_convs even (got S=%d n_convs=%d)", S, n_convs);
    cudaStream_t st = (cudaStream_t)stream;
    const int F = 256 / (S * S), P = S + 2;
    const size_t smem = ((size_t)2 * F * P * P * kRsC + (size_t)2 * kRsWElems) * 2;
    const int grid = (B + F - 1) / F;
    static bool attr6 = false, attr8 = false;
    if (S == 6) {
        if (!attr6) { cudaFuncSetAttribute(pres_stack_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr6 = true; }
        pres_stack_kernel<6><<<grid, 256, smem, st>>>((const __nv_bfloat16 *)x_bf16, (const __nv_bfloat16 *)w_packed_bf16, bias,
                                                      (__nv_bfloat16 *)out_bf16, B, n_convs);
    } else {
        if (!attr8) { cudaFuncSetAttribute(pres_stack_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr8 = true; }
        pres_stack_kernel<8><<<grid, 256, smem, st>>>((const __nv_bfloat16 *)x_bf16, (const __nv_bfloat16 *)w_packed_bf16, bias,
                                                      (__nv_bfloat16 *)out_bf16, B, n_convs);
    }
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_pres_stack");
}

// ------------------------------------------------------------------------------------------
// Location head, last stage (policy.py:275-281 last conv2d_transpose 4x4/2 + ReLU, then :290-295 the 3x3 -> 1 channel
// "conv_1x1" logits): fused.  As two layers the L x L x 32 activation between them is written and read back -- 4 GB
// per head per step at 8,192 canvases and L = 64 -- to produce 33 MB of logits.  Here a CTA computes a 16 x 16 tile
// of logits: the transposed convolution's outputs for the 18 x 18 region the 3x3 needs (81 pixels of each of the four
// parity classes = 24 m16 tiles, three per warp, all of one class) go to shared memory as bf16, then the logits
// convolution runs from there (N = 8 MMA, one useful column).
// ------------------------------------------------------------------------------------------
namespace spiral {

constexpr int kHlT = 16;                         // logits tile
constexpr int kHlR = kHlT + 2;                   // region of up-conv outputs
constexpr int kHlI = kHlT / 2 + 2;               // input tile (10 x 10)
constexpr int kHlKu = 4 * 32, kHlKSu = kHlKu + 8;    // up-conv class kernel: K = 2*2*32
constexpr int kHlKo = 9 * 32, kHlKSo = kHlKo + 8;    // logits kernel: K = 3*3*32

__global__ void __launch_bounds__(256)
phead_last_kernel(const __nv_bfloat16 *__restrict__ u, const __nv_bfloat16 *__restrict__ w_up /*[4][32][KSu] permuted*/,
                  const float *__restrict__ bias_up, const __nv_bfloat16 *__restrict__ w_out /*[8][KSo]*/, const float *__restrict__ bias_out_p,
                  float *__restrict__ logits, int B, int Hh /* input map = L/2 */)
{
    extern __shared__ __align__(16) uint8_t sm_raw[];
    uint16_t *s_in = reinterpret_cast<uint16_t *>(sm_raw);                  // [10][10][32]
    uint16_t *s_wu = s_in + kHlI * kHlI * 32;                                // [4][32][KSu]
    uint16_t *s_up = s_wu + 4 * 32 * kHlKSu;                                 // [18][18][32]
    uint16_t *s_wo = s_up + kHlR * kHlR * 32;                                // [8][KSo]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = 2 * Hh, tiles = L / kHlT;
    const int img = blockIdx.x / (tiles * tiles);
    const int trem = blockIdx.x % (tiles * tiles);
    const int Y0 = (trem / tiles) * kHlT, X0 = (trem % tiles) * kHlT;
    const int a0 = Y0 / 2 - 1, b0 = X0 / 2 - 1;                              // input tile origin

    for (int i = tid; i < 4 * 32 * kHlKSu / 8; i += 256) reinterpret_cast<uint4 *>(s_wu)[i] = __ldg(reinterpret_cast<const uint4 *>(w_up) + i);
    for (int i = tid; i < 8 * kHlKSo / 8; i += 256) reinterpret_cast<uint4 *>(s_wo)[i] = __ldg(reinterpret_cast<const uint4 *>(w_out) + i);
    const uint16_t *ui = reinterpret_cast<const uint16_t *>(u) + (size_t)img * Hh * Hh * 32;
    for (int i = tid; i < kHlI * kHlI * 4; i += 256) {
        const int pix = i >> 2, v = i & 3;
        const int ly = pix / kHlI, lx = pix % kHlI;
        const int iy = a0 + ly, ix = b0 + lx;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (iy >= 0 && iy < Hh && ix >= 0 && ix < Hh) val = __ldg(reinterpret_cast<const uint4 *>(ui + ((size_t)iy * Hh + ix) * 32) + v);
        reinterpret_cast<uint4 *>(s_in)[i] = val;
    }
    __syncthreads();

    const int g = lane >> 2, t = lane & 3;
    // ---- transposed convolution: warp -> class (warp / 2), m-tiles 3 * (warp & 1) .. +2 of that class's 81 pixels ----
    {
        const int cls = warp >> 1, py = cls >> 1, px = cls & 1;
        const uint16_t *w = s_wu + cls * 32 * kHlKSu;
        int base[6];
        int q[6];
#pragma unroll
        for (int r = 0; r < 6; r++) {
            q[r] = ((warp & 1) * 3 + (r >> 1)) * 16 + g + 8 * (r & 1);
            const int qq = q[r] < 81 ? q[r] : 0;
            base[r] = ((qq / 9) * kHlI + (qq % 9)) * 32;
        }
        float acc[3][4][4];
#pragma unroll
        for (int m = 0; m < 3; m++)
#pragma unroll
            for (int nt = 0; nt < 4; nt++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[m][nt][j] = 0.0f;
#pragma unroll 1
        for (int ks = 0; ks < kHlKu / 16; ks++) {
            const int k0 = ks * 16 + 2 * t;
            const int tap0 = k0 >> 5, tap8 = (k0 + 8) >> 5;
            const int o0 = ((tap0 >> 1) * kHlI + (tap0 & 1)) * 32 + (k0 & 31), o8 = ((tap8 >> 1) * kHlI + (tap8 & 1)) * 32 + ((k0 + 8) & 31);
            uint32_t a[3][4];
#pragma unroll
            for (int m = 0; m < 3; m++) {
                a[m][0] = *reinterpret_cast<const uint32_t *>(s_in + base[2 * m] + o0);
                a[m][1] = *reinterpret_cast<const uint32_t *>(s_in + base[2 * m + 1] + o0);
                a[m][2] = *reinterpret_cast<const uint32_t *>(s_in + base[2 * m] + o8);
                a[m][3] = *reinterpret_cast<const uint32_t *>(s_in + base[2 * m + 1] + o8);
            }
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
                const int wrow = (nt * 8 + g) * kHlKSu + k0;
                uint32_t b[2];
                b[0] = *reinterpret_cast<const uint32_t *>(w + wrow);
                b[1] = *reinterpret_cast<const uint32_t *>(w + wrow + 8);
#pragma unroll
                for (int m = 0; m < 3; m++) pc_mma(acc[m][nt], a[m], b);
            }
        }
#pragma unroll
        for (int r = 0; r < 6; r++) {
            if (q[r] >= 81) continue;
            const int m = r >> 1, half = r & 1;
            const int i = q[r] / 9, j = q[r] % 9;
            const int ry = 2 * i + (py == 0 ? 1 : 0), rx = 2 * j + (px == 0 ? 1 : 0);      // position in the 18 x 18 region
            const int Y = Y0 - 1 + ry, X = X0 - 1 + rx;
            const bool inside = Y >= 0 && Y < L && X >= 0 && X < L;                          // outside the map: the 3x3's zero padding
            uint32_t wv[4];
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
                float v0 = acc[m][nt][2 * half] + __ldg(bias_up + t * 8 + 2 * nt);
                float v1 = acc[m][nt][2 * half + 1] + __ldg(bias_up + t * 8 + 2 * nt + 1);
                v0 = inside ? fmaxf(v0, 0.0f) : 0.0f;
                v1 = inside ? fmaxf(v1, 0.0f) : 0.0f;
                wv[nt] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v0)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v1)) << 16);
            }
            *reinterpret_cast<uint4 *>(s_up + (ry * kHlR + rx) * 32 + t * 8) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
        }
    }
    __syncthreads();

    // ---- logits: 3x3 over the region, 256 pixels = 16 m-tiles, two per warp; column 0 of an N = 8 MMA ----
    {
        int base[4];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int p = (warp * 2 + (r >> 1)) * 16 + g + 8 * (r & 1);       // pixel of the 16 x 16 tile
            base[r] = ((p >> 4) * kHlR + (p & 15)) * 32;                      // top-left of its 3x3 window in the region
        }
        float acc[2][4];
#pragma unroll
        for (int m = 0; m < 2; m++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[m][j] = 0.0f;
#pragma unroll 1
        for (int ks = 0; ks < kHlKo / 16; ks++) {
            const int k0 = ks * 16 + 2 * t;
            const int tap0 = k0 >> 5, tap8 = (k0 + 8) >> 5;
            const int o0 = ((tap0 / 3) * kHlR + tap0 % 3) * 32 + (k0 & 31), o8 = ((tap8 / 3) * kHlR + tap8 % 3) * 32 + ((k0 + 8) & 31);
            const int wrow = g * kHlKSo + k0;
            uint32_t b[2];
            b[0] = *reinterpret_cast<const uint32_t *>(s_wo + wrow);
            b[1] = *reinterpret_cast<const uint32_t *>(s_wo + wrow + 8);
#pragma unroll
            for (int m = 0; m < 2; m++) {
                uint32_t a[4];
                a[0] = *reinterpret_cast<const uint32_t *>(s_up + base[2 * m] + o0);
                a[1] = *reinterpret_cast<const uint32_t *>(s_up + base[2 * m + 1] + o0);
                a[2] = *reinterpret_cast<const uint32_t *>(s_up + base[2 * m] + o8);
                a[3] = *reinterpret_cast<const uint32_t *>(s_up + base[2 * m + 1] + o8);
                pc_mma(acc[m], a, b);
            }
        }
        if (t == 0) {                                                          // column 0 of the accumulator tile
            const float bias_out = __ldg(bias_out_p);
            float *lo = logits + (size_t)img * L * L;
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int p = (warp * 2 + (r >> 1)) * 16 + g + 8 * (r & 1);
                lo[(size_t)(Y0 + (p >> 4)) * L + X0 + (p & 15)] = acc[r >> 1][2 * (r & 1)] + bias_out;
            }
        }
    }
}

}  // namespace spiral

extern "C" int spiral_phead_last(const void *u_bf16, const void *w_up_packed, const float *bias_up, const void *w_out_packed,
                                 const float *bias_out, float *logits, int32_t B, int32_t Hh, void *stream)
{
    SPIRAL_REQUIRE(u_bf16); SPIRAL_REQUIRE(w_up_packed); SPIRAL_REQUIRE(bias_up); SPIRAL_REQUIRE(w_out_packed); SPIRAL_REQUIRE(bias_out); SPIRAL_REQUIRE(logits);
    if (B <= 0 || Hh < 8 || (Hh % 8) != 0) return spiral_set_error(SPIRAL_EINVAL, "phead_last: input map must be a multiple of 8 (got %d)", Hh);
    const size_t smem = ((size_t)kHlI * kHlI * 32 + 4 * 32 * kHlKSu + kHlR * kHlR * 32 + 8 * kHlKSo) * 2;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(phead_last_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
    const int tiles = (2 * Hh) / kHlT;
    phead_last_kernel<<<B * tiles * tiles, 256, smem, (cudaStream_t)stream>>>((const __nv_bfloat16 *)u_bf16, (const __nv_bfloat16 *)w_up_packed,
                                                                            bias_up, (const __nv_bfloat16 *)w_out_packed, bias_out, logits, B, Hh);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_phead_last");
}
