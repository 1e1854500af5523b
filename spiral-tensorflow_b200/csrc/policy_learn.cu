// policy_learn.cu -- the policy network's convolutions for the LEARNER (forward, input gradient, weight gradient), fp32
// parity on bf16 tensor cores.
//
// Replaces what tf.gradients runs through the trunk and the location heads of models/policy.py when the learner
// evaluates the A2C loss over [B*T] frames (agent.py:161-239): conv 5x5 'same' (policy.py:94-99), three conv 4x4 /
// stride 2 'valid' (:103-109), the res-blocks' conv 3x3 'same' (:349-362), conv2d_transpose 4x4 / stride 2 'same'
// (:242-247, :275-281) and the final conv 3x3 -> 1 channel (:290-295).  All of them are <= 32-channel convolutions on
// <= 64x64 maps: K = taps * Cin <= 512, N = Cout <= 32 -- too narrow for a TMA / tcgen05 pipeline (a 128-wide N tile
// would be 3/4 padding), so they are mma.sync.m16n8k16 implicit GEMMs like the acting path's (policy_conv.cu), with the
// learner's two extra requirements:
//
//   * fp32 parity: activations and kernels are fp32 tensors; operands are split x = hi + lo into two bf16 planes while
//     they are staged in shared memory and every k-step issues three MMAs (hi*hi + hi*lo + lo*hi): 16 operand bits, fp32
//     accumulation -- the same split the discriminator's reward forward uses (disc_tc.cu)
//   * gradients: the input gradient of every one of these layers is again one of these convolutions with a re-indexed
//     kernel (flipped taps for stride 1; the four parity classes of a transposed convolution for stride 2; a strided
//     convolution for the transposed ones), so `lconv_kernel` + `lpack_kernel` (a strided gather of the kernel into the
//     [Cout][taps * Cin] layout the MMA reads) cover forward AND dgrad; `lwgrad2_kernel` is the one new contraction:
//     dW[tap, ci][co] = sum over frames and output pixels of x[pixel * stride + tap - pad][ci] * dy[pixel][co], an implicit
//     GEMM whose reduction runs over pixels, accumulated in registers by persistent CTAs and reduced over CTAs in a fixed
//     order (deterministic, no atomics).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/spiral_b200.h"
#include "common.h"

namespace spiral {

constexpr int kLcTw = 16, kLcTh = 8;             // output tile: 8 warps x 16 pixels
constexpr int kLcMaxTaps = 25;

struct LConvParams {
    int B, Hin, Win, Hout, Wout;                 // input map and the grid of outputs this launch computes, per frame
    int Cout, CoutPad;                           // CoutPad: multiple of 8, <= 32
    int KH, KW, stride, pad_t, pad_l;
    int K, KP, KS;                               // K = KH*KW*Cin; KP = K rounded up to 16; KS = KP + 8 (row stride, bf16)
    int IH, IW;                                  // staged input tile
    int tiles_x, tiles_y;
    int oH, oW, o_sy, o_sx, o_oy, o_ox;          // output pixel (oy, ox) -> (oy*o_sy + o_oy, ox*o_sx + o_ox) of an [oH, oW] map
    // v2 kernels (lconv2_kernel / lwgrad2_kernel): a CTA covers fr frames x th rows x tw columns of outputs
    int tw, th, fr, mt_rows;                     // tw = 16 or 8; mt_rows = 16 / tw rows per 16-pixel m-tile; th a power of two
    int th_log;
    int n_in;                                    // halfwords per plane of the staged input tile (multiple of 64)
    // forward: n_cls > 1 = the four parity classes of a transposed / stride-2-adjoint convolution from ONE staged tile: class c
    // reads taps shifted by (cls_dy, cls_dx) inside the tile (staged with the largest padding) and writes output pixel
    // (oy * o_sy + cls_oy[c], ox * o_sx + cls_ox[c]); its kernel planes follow class c - 1's
    int n_cls, cls_dy[4], cls_dx[4], cls_oy[4], cls_ox[4];
    int lw_ks;                                   // wgrad: k-steps (16 output pixels each) per item: 8 or 16
    uint32_t mIH, mIW, mTH, mTW;                 // lc_magic of IH, IW (input tile) and th, tw (wgrad's dy tile)
    int relu;                                    // forward epilogue: max(v, 0) after the bias (tf.nn.relu of the layer's activation)
    int dy_pitch, dy_c0;                         // wgrad: dy is [.., dy_pitch] channels per pixel, this launch takes [dy_c0, dy_c0 + Cout)
};

// What the discriminator's narrow layers need on top of a plain convolution (lconv2_kernel<..., XF = true>): the producer
// layer's batch norm + leaky-ReLU applied to the input while it is staged (in_mode 1: v <- lrelu(v * scale[c] + shift[c],
// 0.2), disc.cu's convention: a layer stores its RAW output and the consumer normalises), or the image normalisation of layer
// 0 (in_mode 2: v <- v * in_a + in_b); padding stays zero in the transformed space, as 'SAME' padding pads the activation.
// Output columns [out_c0, out_c0 + Cout) of an [.., out_ld]-channel map, and per-CTA batch-norm partials
// bn_partial[cta][bn_ld][2] = (sum, sum of squares) of this launch's raw outputs per channel.
struct LConvExtra {
    const float *in_scale, *in_shift;
    float in_a, in_b;
    int in_mode;
    float *bn_partial;
    int out_ld, out_c0, bn_ld;
};

__device__ __forceinline__ void lc_mma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ void lc_mma_h(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ uint32_t lc_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void lc_ldsm4(uint32_t (&r)[4], uint32_t addr)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

__device__ __forceinline__ void lc_ldsm4t(uint32_t (&r)[4], uint32_t addr)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

__device__ __forceinline__ void lc_cp_async16(uint32_t dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

__device__ __forceinline__ void lc_cp_async_wait()
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// The third operand format, "f16x3": v = h + l * 2^-11 with h = fp16(v) and l = fp16((v - h) * 2^11) -- 22 operand bits in TWO
// planes and three products (h*h into `hi`; h*l + l*h into `lo`, which the epilogue scales by 2^-11; the dropped l*l term is
// 2^-22 of the result).  The scaling keeps the residual plane in fp16's normal range whenever h is: |v| >= 6.1e-5 carries 22
// bits, smaller values an absolute error <= 3e-11; |v| is clamped to fp16's 65504 (activations, kernels and gradients of the
// policy network are orders of magnitude inside that).  Half the MMAs and two thirds of the shared memory of bf16x6 at
// the same parity (tests/test_gpu_policy_native.py compares both with float64).
constexpr float kF16LoScale = 2048.0f, kF16LoInv = 1.0f / 2048.0f;

__device__ __forceinline__ uint32_t lc_h2_bits(__half2 h) { return *reinterpret_cast<uint32_t *>(&h); }
__device__ __forceinline__ uint32_t lc_b2_bits(__nv_bfloat162 h) { return *reinterpret_cast<uint32_t *>(&h); }

// two values -> NPL packed pairs (low half = v0)
template <int NPL, bool F16>
__device__ __forceinline__ void lc_split_pair(float v0, float v1, uint32_t (&pl)[NPL])
{
    if (F16) {
        v0 = fminf(fmaxf(v0, -65504.0f), 65504.0f);
        v1 = fminf(fmaxf(v1, -65504.0f), 65504.0f);
        const __half2 h = __floats2half2_rn(v0, v1);
        const float2 f = __half22float2(h);
        pl[0] = lc_h2_bits(h);
        pl[1] = lc_h2_bits(__floats2half2_rn((v0 - f.x) * kF16LoScale, (v1 - f.y) * kF16LoScale));
    } else {
#pragma unroll
        for (int q = 0; q < NPL; q++) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
            const float2 f = __bfloat1622float2(h);
            pl[q] = lc_b2_bits(h);
            v0 -= f.x;
            v1 -= f.y;
        }
    }
}

// v = p[0] + p[1] (+ p[2]) + O(2^-8NPL |v|): NPL bf16 planes of 8 bits each (16 bits: "bf16x3", the three products
// p0*q0 + p0*q1 + p1*q0; 24 bits: "bf16x6", fp32-exact operands, six products)
template <int NPL>
__device__ __forceinline__ void lc_split(float v, uint16_t (&pl)[NPL])
{
#pragma unroll
    for (int i = 0; i < NPL; i++) {
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        pl[i] = __bfloat16_as_ushort(h);
        v -= __bfloat162float(h);
    }
}

// The tensor core adds every MMA into its fp32 accumulator with truncation, so the error of an accumulator grows with the
// number of MMAs chained into it.  Two accumulators per tile: `hi` takes only the dominant p0*q0 product of each k-step
// (K/16 MMAs), `lo` every cross product (2^-8 .. 2^-16 of the result: their truncation is invisible); the epilogue adds the
// two in registers with round-to-nearest.
template <int NPL>
__device__ __forceinline__ void lc_mma_planes(float (&hi)[4], float (&lo)[4], const uint32_t (&a)[NPL][4], const uint32_t (&b)[NPL][2])
{
    if (NPL == 3) {
        lc_mma(lo, a[2], b[0]);
        lc_mma(lo, a[0], b[2]);
        lc_mma(lo, a[1], b[1]);
    }
    lc_mma(lo, a[1], b[0]);
    lc_mma(lo, a[0], b[1]);
    lc_mma(hi, a[0], b[0]);
}

template <int NPL, bool F16>
__device__ __forceinline__ void lc_mma_planes2(float (&hi)[4], float (&lo)[4], const uint32_t (&a)[NPL][4], const uint32_t (&b)[NPL][2])
{
    if (F16) {
        lc_mma_h(lo, a[1], b[0]);
        lc_mma_h(lo, a[0], b[1]);
        lc_mma_h(hi, a[0], b[0]);
    } else {
        lc_mma_planes<NPL>(hi, lo, a, b);
    }
}

// ------------------------------------------------------------------------------------------
// kernel packing: packed_{hi,lo}[row][tap * Ck + c] = w[row * s_row + c * s_c + ky(tap) * s_ky + kx(tap) * s_kx]
// (zero for row >= rows and for the k padding).  With the strides and the tap table chosen by the caller this is the
// forward kernel of a convolution, its flipped / transposed adjoint, or one parity class of a transposed convolution.
// ------------------------------------------------------------------------------------------
struct LPackParams {
    int rows, rows_pad, Ck, n_taps, K, KS;       // Ck: channels per tap of the packed layout
    int Ck_src;                                  // channels of the kernel tensor (< Ck: the rest of each tap is zero)
    int perm;                                    // acting path (pconv_kernel with 32 output channels): row nt*8 + j holds
                                                 // channel (j>>1)*8 + nt*2 + (j&1), so a thread's accumulators are 8 adjacent channels
    long long s_row, s_c, s_ky, s_kx;
    int tap_ky[kLcMaxTaps], tap_kx[kLcMaxTaps];
};

template <int NPL, bool F16 = false>
__global__ void __launch_bounds__(256)
lpack_kernel(const __grid_constant__ LPackParams p, const float *__restrict__ w, uint16_t *__restrict__ planes)
{
    const int n = p.rows_pad * p.KS;                                          // elements per plane; planes are contiguous
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int prow = i / p.KS, k = i - prow * p.KS;
        const int row = p.perm ? (((prow & 7) >> 1) * 8 + (prow >> 3) * 2 + (prow & 1)) : prow;
        float v = 0.0f;
        if (row < p.rows && k < p.K) {
            const int tap = k / p.Ck, c = k - tap * p.Ck;
            if (c < p.Ck_src) v = w[row * p.s_row + c * p.s_c + p.tap_ky[tap] * p.s_ky + p.tap_kx[tap] * p.s_kx];
        }
        if (F16) {
            uint32_t pp[NPL];
            lc_split_pair<NPL, F16>(v, 0.0f, pp);
#pragma unroll
            for (int q = 0; q < NPL; q++) planes[(size_t)q * n + i] = (uint16_t)(pp[q] & 0xffffu);
        } else {
            uint16_t pl[NPL];
            lc_split<NPL>(v, pl);
#pragma unroll
            for (int q = 0; q < NPL; q++) planes[(size_t)q * n + i] = pl[q];
        }
    }
}

// ------------------------------------------------------------------------------------------
// wgrad: dW[k = tap * CIN + c][co] = sum_{frame, oy, ox} x[frame, oy*stride - pad + ky, ox*stride - pad + kx, c] * dy[frame, oy, ox, co]
// (lwgrad2_kernel below).  Persistent CTAs walk work items of 128 output pixels; warp w owns the m-tiles (16 rows of k)
// w, w + 8, ... and keeps their accumulators in registers across all items; partial[cta][k][co] at the end, reduced over
// CTAs in a fixed order.
// ------------------------------------------------------------------------------------------
constexpr int kLwMaxMt = 4;                      // m-tiles per warp: K <= 8 * 4 * 16 = 512
constexpr int kLwPix = kLcTw * kLcTh;            // 128 reduction elements per item

// dw[k][co] (k < K, co < Cout) = sum over CTAs, fixed order
// (cin < cin_pad: the partials' k index is tap * cin_pad + c, dw's is tap * cin + c)
// Eight lanes per output: lane j adds the CTAs j, j + 8, ... in order, the eight sums are combined as a fixed tree.
__global__ void __launch_bounds__(256)
lwgrad_reduce_kernel(const float *__restrict__ partial, float *__restrict__ dw, int n_cta, int K, int KP, int Cout, int CoutPad,
                     int cin = 1, int cin_pad = 1, int ld = 0, int co0 = 0)    // dw[k][ld] columns [co0, co0 + Cout); ld 0: Cout
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = t >> 3, j = t & 7;
    const bool ok = i < K * Cout;                                             // whole groups of 8 lanes agree
    const int k = ok ? i / Cout : 0, co = ok ? i - k * Cout : 0;
    const int kp = (k / cin) * cin_pad + k % cin;
    float s = 0.0f;
    if (ok)
        for (int c = j; c < n_cta; c += 8) s += partial[((size_t)c * KP + kp) * CoutPad + co];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (ok && j == 0) dw[(size_t)k * (ld ? ld : Cout) + co0 + co] = s;
}


// ==========================================================================================
// v2 kernels (Cin a multiple of 8): the learner's hot shapes.
//
// What the first kernels above paid for (ncu / per-call timings, tools/learner_probe.py): 8-way bank conflicts on the A
// fragments (a pixel's 32 channels are 64 B, so the 8 pixels of a fragment column start in the same banks), scalar 32-bit /
// 16-bit fragment loads, the 100 KB kernel re-read per 128 output pixels, and half-empty 8 x 16 tiles on the 8 x 8 / 6 x 6
// maps of the res-blocks.  Here:
//   * the staged tile is [pixel][channel] with the 16-byte chunks of every 128-byte line XOR-swizzled by the line index, so
//     the eight rows of an ldmatrix land in eight different bank groups for stride 1 and stride 2; fragments come from
//     ldmatrix.x4 (forward) / ldmatrix.x4.trans (wgrad: the reduction runs over pixels, the transposed load IS the
//     pixel-major -> channel-major transpose the first kernel did with 16-bit loads)
//   * a CTA covers 256 outputs (two m-tiles per warp, B fragments used twice) as fr frames x th rows x tw columns with
//     tw = 8 for maps up to 8 wide (four 8 x 8 frames per CTA instead of half of one 8 x 16 tile)
//   * the kernel planes arrive by cp.async while the input tile is converted
//   * operand format f16x3 (above) next to bf16x3 / bf16x6
// ==========================================================================================
#ifndef LC2_STAGE_U
#define LC2_STAGE_U 8
#endif
__device__ __forceinline__ int lc_swz(int L) { return (L & ~7) | ((L ^ (L >> 3)) & 7); }

// n / d for n < 2^16 with the host-computed magic m = ceil(2^32 / d) (d >= 2): exact, two instructions
__device__ __forceinline__ int lc_fastdiv(int n, uint32_t m) { return (int)__umulhi((uint32_t)n, m); }
static uint32_t lc_magic(int d) { return (uint32_t)((0x100000000ull + (uint64_t)d - 1) / (uint64_t)d); }

// stage the box [fr frames][RH rows][RW columns] of an NHWC fp32 tensor [., H, W, CSRC] starting at (y0, x0) of frame 0 of
// `src` (zero outside the map and for frames >= frames_left) as NPL swizzled planes [pixel][C] of n_plane halfwords.
// CSRC == C: float4 loads.  CSRC < C = 8 (1, 2, 3 or 6 image / scalar-map channels): the pixel's channels are padded with
// zeros to one 16-byte chunk, so these layers run on the same kernels (the padded k rows meet zero kernel columns).
// The box is walked as ONE flat list by all 256 threads, up to eight loads in flight per thread: a tile is a few thousand
// float4, and staging it is latency-bound -- row-per-warp loops serialised three to nine round trips where this takes one
// to five.  (mRW, mRH: lc_magic of RW, RH; the box has < 2^16 pixels.)
template <int CSRC, int C, int NPL, bool F16, bool XF = false>
__device__ __forceinline__ void lc2_stage(const float *__restrict__ src, size_t frame_elems, int frames_left, int H, int W,
                                          int RH, int RW, uint32_t mRH, uint32_t mRW, int y0, int x0, int fr, uint16_t *s,
                                          int n_plane, int tid,
                                          int pitch = CSRC,                    // floats per pixel of the source (>= CSRC: a channel slice)
                                          const LConvExtra *ex = nullptr)      // XF: input transform (LConvExtra)
{
    static_assert(CSRC == C || C == 8, "channel padding goes to one chunk");
    const int npix = fr * RH * RW;
    if (CSRC == C) {
        constexpr int V = C / 4;                                               // float4 per pixel
        constexpr int U = LC2_STAGE_U;                                         // loads in flight per thread
        const int n_vec = npix * V;
        for (int i0 = tid; i0 < n_vec; i0 += 256 * U) {
            float4 val[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int i = i0 + 256 * u;
                val[u] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (i < n_vec) {
                    const int pixel = i / V, iv = i - pixel * V;
                    const int grow = lc_fastdiv(pixel, mRW), lx = pixel - grow * RW;
                    const int f = lc_fastdiv(grow, mRH), ly = grow - f * RH;
                    const int iy = y0 + ly, ix = x0 + lx;
                    if (f < frames_left && iy >= 0 && iy < H && ix >= 0 && ix < W) {
                        val[u] = __ldg(reinterpret_cast<const float4 *>(src + (ptrdiff_t)f * (ptrdiff_t)frame_elems +
                                                                        ((ptrdiff_t)iy * W + ix) * pitch) + iv);
                        if (XF) {
                            float4 &q = val[u];
                            if (ex->in_mode == 1) {
                                const float4 sc = __ldg(reinterpret_cast<const float4 *>(ex->in_scale) + iv);
                                const float4 sh = __ldg(reinterpret_cast<const float4 *>(ex->in_shift) + iv);
                                q.x = fmaf(q.x, sc.x, sh.x); q.y = fmaf(q.y, sc.y, sh.y); q.z = fmaf(q.z, sc.z, sh.z); q.w = fmaf(q.w, sc.w, sh.w);
                                q.x = q.x > 0.f ? q.x : 0.2f * q.x; q.y = q.y > 0.f ? q.y : 0.2f * q.y;
                                q.z = q.z > 0.f ? q.z : 0.2f * q.z; q.w = q.w > 0.f ? q.w : 0.2f * q.w;
                            } else if (ex->in_mode == 2) {
                                q.x = fmaf(q.x, ex->in_a, ex->in_b); q.y = fmaf(q.y, ex->in_a, ex->in_b);
                                q.z = fmaf(q.z, ex->in_a, ex->in_b); q.w = fmaf(q.w, ex->in_a, ex->in_b);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int i = i0 + 256 * u;
                if (i < n_vec) {
                    const int e = i * 4;                                       // element index in the unswizzled tile
                    const int L = e >> 3;
                    uint32_t p0[NPL], p1[NPL];
                    lc_split_pair<NPL, F16>(val[u].x, val[u].y, p0);
                    lc_split_pair<NPL, F16>(val[u].z, val[u].w, p1);
                    SP_CHECK(lc_swz(L) * 8 + 8 <= n_plane);
                    uint16_t *dst = s + lc_swz(L) * 8 + (e & 4);
#pragma unroll
                    for (int q = 0; q < NPL; q++) *reinterpret_cast<uint2 *>(dst + (size_t)q * n_plane) = make_uint2(p0[q], p1[q]);
                }
            }
        }
    } else {
        constexpr int U = 4;                                                   // pixels in flight per thread
        for (int i0 = tid; i0 < npix; i0 += 256 * U) {
            float val[U][8];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int i = i0 + 256 * u;
                const int grow = lc_fastdiv(i, mRW), lx = i - grow * RW;
                const int f = lc_fastdiv(grow, mRH), ly = grow - f * RH;
                const int iy = y0 + ly, ix = x0 + lx;
                const bool ok = i < npix && f < frames_left && iy >= 0 && iy < H && ix >= 0 && ix < W;
                const float *pp = src + (ptrdiff_t)f * (ptrdiff_t)frame_elems + ((ptrdiff_t)iy * W + ix) * pitch;
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    float v = (c < CSRC && ok) ? __ldg(pp + c) : 0.0f;
                    if (XF && c < CSRC && ok) {
                        if (ex->in_mode == 1) {
                            v = fmaf(v, __ldg(ex->in_scale + c), __ldg(ex->in_shift + c));
                            v = v > 0.f ? v : 0.2f * v;
                        } else if (ex->in_mode == 2) {
                            v = fmaf(v, ex->in_a, ex->in_b);
                        }
                    }
                    val[u][c] = v;
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int i = i0 + 256 * u;
                if (i < npix) {
                    uint32_t pp[4][NPL];                                       // one chunk per pixel
#pragma unroll
                    for (int h = 0; h < 4; h++) lc_split_pair<NPL, F16>(val[u][2 * h], val[u][2 * h + 1], pp[h]);
                    SP_CHECK(lc_swz(i) * 8 + 8 <= n_plane);
                    uint16_t *dst = s + lc_swz(i) * 8;
#pragma unroll
                    for (int q = 0; q < NPL; q++)
                        *reinterpret_cast<uint4 *>(dst + (size_t)q * n_plane) = make_uint4(pp[0][q], pp[1][q], pp[2][q], pp[3][q]);
                }
            }
        }
    }
}

// chunk q (8 consecutive k) of the packed reduction index -> chunk offset in the staged tile
template <int CIN>
__device__ __forceinline__ void lc2_qoff_table(const LConvParams &p, int *s_qoff, int tid)
{
    constexpr int CH = CIN / 8;
    const int nq = p.KP / 8;
    for (int i = tid; i < nq * p.n_cls; i += 256) {
        const int cls = i / nq, q = i - cls * nq;
        int off = 0;
        if (q * 8 < p.K) {
            const int tap = q / CH, c8 = q - tap * CH;
            const int ky = tap / p.KW, kx = tap - ky * p.KW;
            off = ((ky + p.cls_dy[cls]) * p.IW + kx + p.cls_dx[cls]) * CH + c8;
        }
        s_qoff[i] = off;
    }
}

template <int CSRC, int NPL, bool F16, int MT, bool XF = false>
__global__ void __launch_bounds__(256, (NPL == 2) ? 2 : 1)
lconv2_kernel(const __grid_constant__ LConvParams p, const float *__restrict__ x, const uint16_t *__restrict__ w_planes,
              const float *__restrict__ bias, float *__restrict__ out, const __grid_constant__ LConvExtra ex)
{
    extern __shared__ __align__(16) uint8_t sm_raw[];
    constexpr int CIN = CSRC < 8 ? 8 : CSRC;                                    // channels of the staged tile (lc2_stage pads)
    constexpr int CH = CIN / 8;
    const int n_w = p.CoutPad * p.KS;
    uint16_t *s_i = reinterpret_cast<uint16_t *>(sm_raw);                     // [NPL][n_in], swizzled
    uint16_t *s_w = s_i + (size_t)NPL * p.n_in;                                // [NPL][CoutPad][KS]
    int *s_qoff = reinterpret_cast<int *>(s_w + (size_t)p.n_cls * NPL * n_w);  // [n_cls][KP / 8]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles = p.tiles_x * p.tiles_y;
    const int fg = blockIdx.x / tiles, trem = blockIdx.x - fg * tiles;
    const int oy0 = (trem / p.tiles_x) * p.th, ox0 = (trem % p.tiles_x) * p.tw;
    const int img0 = fg * p.fr;
    {
        const uint32_t sw = lc_smem_u32(s_w);
        const int n16 = p.n_cls * NPL * n_w / 8;
        const uint4 *gw = reinterpret_cast<const uint4 *>(w_planes);
        for (int i = tid; i < n16; i += 256) lc_cp_async16(sw + (uint32_t)i * 16u, gw + i);
    }
    lc2_qoff_table<CIN>(p, s_qoff, tid);
    lc2_stage<CSRC, CIN, NPL, F16, XF>(x + (size_t)img0 * p.Hin * p.Win * CSRC, (size_t)p.Hin * p.Win * CSRC, p.B - img0, p.Hin,
                                       p.Win, p.IH, p.IW, p.mIH, p.mIW, oy0 * p.stride - p.pad_t, ox0 * p.stride - p.pad_l, p.fr, s_i,
                                       p.n_in, tid, CSRC, &ex);
    lc_cp_async_wait();
    __syncthreads();

    // ldmatrix roles of this lane: matrix j = lane / 8, row r = lane % 8
    const int j = lane >> 3, r = lane & 7;
    const int rows_w = MT * p.mt_rows;                                          // output rows of this warp
    int pixA[MT];
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
        const int i_m = r + ((j & 1) << 3);                                     // A: matrices 1, 3 are pixels 8..15 of the m-tile
        const int rowg = warp * rows_w + mt * p.mt_rows + (p.tw == 8 ? (i_m >> 3) : 0);
        const int col = p.tw == 8 ? (i_m & 7) : i_m;
        const int f = rowg >> p.th_log, ry = rowg & (p.th - 1);
        pixA[mt] = ((f * p.IH + ry * p.stride) * p.IW + col * p.stride) * CH;
    }
    const int qsel = j >> 1;                                                    // A: matrices 2, 3 are the k-step's second chunk
    int wrow[2];
#pragma unroll
    for (int ntp = 0; ntp < 2; ntp++) {                                         // B: matrices 0, 1 = n-tile 2 ntp (k chunks 0, 1), 2, 3 = n-tile 2 ntp + 1
        int n = ntp * 16 + ((j >> 1) << 3) + r;
        n = n < p.CoutPad ? n : p.CoutPad - 1;
        wrow[ntp] = n * p.KS + ((j & 1) << 3);
    }
    const uint32_t si_base = lc_smem_u32(s_i), sw_base = lc_smem_u32(s_w);
    const uint32_t in_plane = (uint32_t)p.n_in * 2u, w_plane = (uint32_t)n_w * 2u;
    const int NT = p.CoutPad >> 3;
    const int g = lane >> 2, t = lane & 3;
    const float lo_scale = F16 ? kF16LoInv : 1.0f;
    const int out_ld = XF ? ex.out_ld : p.Cout, out_c0 = XF ? ex.out_c0 : 0;
    float bs[4][2], bq[4][2];                                                  // XF: batch-norm partials of this thread's channels
#pragma unroll
    for (int nt = 0; nt < 4; nt++) { bs[nt][0] = bs[nt][1] = bq[nt][0] = bq[nt][1] = 0.0f; }

#pragma unroll 1
  for (int cls = 0; cls < p.n_cls; cls++) {
    const uint32_t sw_cls = sw_base + (uint32_t)cls * NPL * w_plane;
    const int *qoff_cls = s_qoff + cls * (p.KP / 8);
    float acc[MT][4][4], acl[MT][4][4];
#pragma unroll
    for (int mt = 0; mt < MT; mt++)
#pragma unroll
        for (int nt = 0; nt < 4; nt++)
#pragma unroll
            for (int e = 0; e < 4; e++) { acc[mt][nt][e] = 0.0f; acl[mt][nt][e] = 0.0f; }

#pragma unroll 1
    for (int ks = 0; ks < p.KP / 16; ks++) {
        const int qo = qoff_cls[2 * ks + qsel];
        uint32_t a[MT][NPL][4];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            const int L = pixA[mt] + qo;
            SP_CHECK(L >= 0 && lc_swz(L) * 8 + 8 <= p.n_in);
            const uint32_t ad = si_base + ((uint32_t)lc_swz(L) << 4);
#pragma unroll
            for (int q = 0; q < NPL; q++) lc_ldsm4(a[mt][q], ad + (uint32_t)q * in_plane);
        }
        uint32_t b[4][NPL][2];
#pragma unroll
        for (int ntp = 0; ntp < 2; ntp++) {
            if (ntp * 2 < NT) {
                const uint32_t ad = sw_cls + ((uint32_t)(wrow[ntp] + ks * 16) << 1);
#pragma unroll
                for (int q = 0; q < NPL; q++) {
                    uint32_t t4[4];
                    lc_ldsm4(t4, ad + (uint32_t)q * w_plane);
                    b[2 * ntp][q][0] = t4[0]; b[2 * ntp][q][1] = t4[1];
                    b[2 * ntp + 1][q][0] = t4[2]; b[2 * ntp + 1][q][1] = t4[3];
                }
            }
        }
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int nt = 0; nt < 4; nt++)
                if (nt < NT) lc_mma_planes2<NPL, F16>(acc[mt][nt], acl[mt][nt], a[mt], b[nt]);
    }

#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const int i_m = g + 8 * half;
            const int rowg = warp * rows_w + mt * p.mt_rows + (p.tw == 8 ? (i_m >> 3) : 0);
            const int col = p.tw == 8 ? (i_m & 7) : i_m;
            const int f = rowg >> p.th_log, ry = rowg & (p.th - 1);
            const int img = img0 + f, oy = oy0 + ry, ox = ox0 + col;
            const int py = oy * p.o_sy + p.cls_oy[cls], px = ox * p.o_sx + p.cls_ox[cls];
            if (img >= p.B || oy >= p.Hout || ox >= p.Wout || py < 0 || py >= p.oH || px < 0 || px >= p.oW) continue;
            const size_t pix = ((size_t)img * p.oH + py) * p.oW + px;
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
                if (nt >= NT) continue;
                const int ch = nt * 8 + 2 * t;
                float v0 = fmaf(acl[mt][nt][2 * half], lo_scale, acc[mt][nt][2 * half]);
                float v1 = fmaf(acl[mt][nt][2 * half + 1], lo_scale, acc[mt][nt][2 * half + 1]);
                if (bias) {
                    if (ch < p.Cout) v0 += __ldg(bias + ch);
                    if (ch + 1 < p.Cout) v1 += __ldg(bias + ch + 1);
                }
                if (p.relu) { v0 = fmaxf(v0, 0.0f); v1 = fmaxf(v1, 0.0f); }
                if (XF) {
                    bs[nt][0] += v0; bq[nt][0] += v0 * v0;
                    bs[nt][1] += v1; bq[nt][1] += v1 * v1;
                }
                float *o = out + pix * out_ld + out_c0 + ch;
                if ((p.Cout & 1) == 0 && ch + 1 < p.Cout && ((out_ld | out_c0) & 1) == 0) {
                    *reinterpret_cast<float2 *>(o) = make_float2(v0, v1);
                } else {
                    if (ch < p.Cout) o[0] = v0;
                    if (ch + 1 < p.Cout) o[1] = v1;
                }
            }
        }
    }
  }   // next class
    if constexpr (XF) {
        if (ex.bn_partial) {
            // per-CTA per-channel (sum, sum of squares) in a fixed order: lanes of a warp by shuffles, warps through shared memory
            __shared__ float s_bn[8][32][2];
#pragma unroll
            for (int nt = 0; nt < 4; nt++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
#pragma unroll
                    for (int off = 4; off < 32; off <<= 1) {
                        bs[nt][e] += __shfl_xor_sync(0xffffffffu, bs[nt][e], off);
                        bq[nt][e] += __shfl_xor_sync(0xffffffffu, bq[nt][e], off);
                    }
                    if (g == 0 && nt < NT) {
                        s_bn[warp][nt * 8 + 2 * t + e][0] = bs[nt][e];
                        s_bn[warp][nt * 8 + 2 * t + e][1] = bq[nt][e];
                    }
                }
            __syncthreads();
            if (tid < p.Cout) {
                float sum = 0.0f, sq = 0.0f;
#pragma unroll
                for (int w = 0; w < 8; w++) { sum += s_bn[w][tid][0]; sq += s_bn[w][tid][1]; }
                float *dst = ex.bn_partial + ((size_t)blockIdx.x * ex.bn_ld + ex.out_c0 + tid) * 2;
                dst[0] = sum;
                dst[1] = sq;
            }
        }
    }
}

// wgrad v2: an item is 128 output pixels (8 k-steps of one 16-pixel group each: a row of a 16-wide tile, or two rows of an
// 8-wide one) of fr frames; x tile and dy tile are staged by the same routine ([pixel][channel], swizzled) and both
// fragments are transposed ldmatrix loads.  `hi` (dominant products) is flushed into the fp32 sum `tot` after every item, so
// that an accumulator never chains more than 8 MMAs; `lo` (cross products) runs through all items.
template <int CSRC, int NPL, bool F16>
__global__ void __launch_bounds__(256, 1)
lwgrad2_kernel(const __grid_constant__ LConvParams p, const float *__restrict__ x, const float *__restrict__ dy,
               float *__restrict__ partial, int n_items)
{
    extern __shared__ __align__(16) uint8_t sm_raw[];
    constexpr int CIN = CSRC < 8 ? 8 : CSRC;
    constexpr int CH = CIN / 8;
    const int CC = p.CoutPad, CCH = CC >> 3;                                    // staged dy channels (8, 16 or 32) and chunks per pixel
    const int n_d = p.lw_ks * 16 * CC;                                          // multiple of 64
    uint16_t *s_i = reinterpret_cast<uint16_t *>(sm_raw);                     // [NPL][n_in]
    uint16_t *s_d = s_i + (size_t)NPL * p.n_in;                                // [NPL][128][CC]
    int *s_qoff = reinterpret_cast<int *>(s_d + (size_t)NPL * n_d);            // [KP / 8]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int j = lane >> 3, r = lane & 7;
    const int NT = CC >> 3;
    const int n_mt = p.KP / 16;
    const int tiles = p.tiles_x * p.tiles_y;
    lc2_qoff_table<CIN>(p, s_qoff, tid);

    float acc[kLwMaxMt][4][4], acl[kLwMaxMt][4][4], tot[kLwMaxMt][4][4];
#pragma unroll
    for (int m = 0; m < kLwMaxMt; m++)
#pragma unroll
        for (int nt = 0; nt < 4; nt++)
#pragma unroll
            for (int e = 0; e < 4; e++) { acc[m][nt][e] = 0.0f; acl[m][nt][e] = 0.0f; tot[m][nt][e] = 0.0f; }

    // A (transposed load): matrix j = (chunk j & 1 of the m-tile, pixels 8 (j >> 1) .. + 7 of the k-step)
    const int i_a = r + ((j >> 1) << 3);
    // B (transposed load): matrix j = (pixels 8 (j & 1) .. + 7, n-tile 2 ntp + (j >> 1))
    const int i_b = r + ((j & 1) << 3);
    const uint32_t si_base = lc_smem_u32(s_i), sd_base = lc_smem_u32(s_d);
    const uint32_t in_plane = (uint32_t)p.n_in * 2u, d_plane = (uint32_t)n_d * 2u;

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int fg = item / tiles, trem = item - fg * tiles;
        const int oy0 = (trem / p.tiles_x) * p.th, ox0 = (trem % p.tiles_x) * p.tw;
        const int img0 = fg * p.fr;
        __syncthreads();                                                       // the previous item's MMAs are done with the tiles
        lc2_stage<CSRC, CIN, NPL, F16>(x + (size_t)img0 * p.Hin * p.Win * CSRC, (size_t)p.Hin * p.Win * CSRC, p.B - img0, p.Hin,
                                       p.Win, p.IH, p.IW, p.mIH, p.mIW, oy0 * p.stride - p.pad_t, ox0 * p.stride - p.pad_l, p.fr, s_i,
                                       p.n_in, tid);
        {
            const float *dsrc = dy + (size_t)img0 * p.Hout * p.Wout * p.dy_pitch + p.dy_c0;
            const size_t fe = (size_t)p.Hout * p.Wout * p.dy_pitch;
            const int rem = p.B - img0, pt = p.dy_pitch;
            if (p.Cout == 32) lc2_stage<32, 32, NPL, F16>(dsrc, fe, rem, p.Hout, p.Wout, p.th, p.tw, p.mTH, p.mTW, oy0, ox0, p.fr, s_d, n_d, tid, pt);
            else if (p.Cout == 16) lc2_stage<16, 16, NPL, F16>(dsrc, fe, rem, p.Hout, p.Wout, p.th, p.tw, p.mTH, p.mTW, oy0, ox0, p.fr, s_d, n_d, tid, pt);
            else if (p.Cout == 8) lc2_stage<8, 8, NPL, F16>(dsrc, fe, rem, p.Hout, p.Wout, p.th, p.tw, p.mTH, p.mTW, oy0, ox0, p.fr, s_d, n_d, tid, pt);
            else lc2_stage<1, 8, NPL, F16>(dsrc, fe, rem, p.Hout, p.Wout, p.th, p.tw, p.mTH, p.mTW, oy0, ox0, p.fr, s_d, n_d, tid, pt);   // scalar map
        }
        __syncthreads();

#pragma unroll 1
        for (int ks = 0; ks < p.lw_ks; ks++) {
            uint32_t b[4][NPL][2];
#pragma unroll
            for (int ntp = 0; ntp < 2; ntp++) {
                if (ntp * 2 < NT) {
                    int chunk = ntp * 2 + (j >> 1);
                    chunk = chunk < CCH ? chunk : CCH - 1;
                    const int L = (ks * 16 + i_b) * CCH + chunk;
                    const uint32_t ad = sd_base + ((uint32_t)lc_swz(L) << 4);
#pragma unroll
                    for (int q = 0; q < NPL; q++) {
                        uint32_t t4[4];
                        lc_ldsm4t(t4, ad + (uint32_t)q * d_plane);
                        b[2 * ntp][q][0] = t4[0]; b[2 * ntp][q][1] = t4[1];
                        b[2 * ntp + 1][q][0] = t4[2]; b[2 * ntp + 1][q][1] = t4[3];
                    }
                }
            }
            // staged pixel of this lane's row of the A matrices
            const int rowg = ks * p.mt_rows + (p.tw == 8 ? (i_a >> 3) : 0);
            const int col = p.tw == 8 ? (i_a & 7) : i_a;
            const int f = rowg >> p.th_log, ry = rowg & (p.th - 1);
            const int pix = ((f * p.IH + ry * p.stride) * p.IW + col * p.stride) * CH;
#pragma unroll
            for (int m = 0; m < kLwMaxMt; m++) {
                const int mt = warp + 8 * m;
                if (mt < n_mt) {
                    const int L = pix + s_qoff[2 * mt + (j & 1)];
                    SP_CHECK(L >= 0 && lc_swz(L) * 8 + 8 <= p.n_in);
                    const uint32_t ad = si_base + ((uint32_t)lc_swz(L) << 4);
                    uint32_t a[NPL][4];
#pragma unroll
                    for (int q = 0; q < NPL; q++) lc_ldsm4t(a[q], ad + (uint32_t)q * in_plane);
#pragma unroll
                    for (int nt = 0; nt < 4; nt++)
                        if (nt < NT) lc_mma_planes2<NPL, F16>(acc[m][nt], acl[m][nt], a, b[nt]);
                }
            }
        }
#pragma unroll
        for (int m = 0; m < kLwMaxMt; m++)
#pragma unroll
            for (int nt = 0; nt < 4; nt++)
#pragma unroll
                for (int e = 0; e < 4; e++) { tot[m][nt][e] += acc[m][nt][e]; acc[m][nt][e] = 0.0f; }
    }
    const float lo_scale = F16 ? kF16LoInv : 1.0f;
    const int g = lane >> 2, t = lane & 3;
    float *mine = partial + (size_t)blockIdx.x * p.KP * p.CoutPad;
#pragma unroll
    for (int m = 0; m < kLwMaxMt; m++) {
        const int mt = warp + 8 * m;
        if (mt >= n_mt) continue;
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
            if (nt >= NT) continue;
            const int col = nt * 8 + 2 * t;
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; e++) v[e] = fmaf(acl[m][nt][e], lo_scale, tot[m][nt][e]);
            *reinterpret_cast<float2 *>(mine + (size_t)(mt * 16 + g) * p.CoutPad + col) = make_float2(v[0], v[1]);
            *reinterpret_cast<float2 *>(mine + (size_t)(mt * 16 + g + 8) * p.CoutPad + col) = make_float2(v[2], v[3]);
        }
    }
}


// wgrad of a convolution to ONE output channel (the location heads' final 3x3 -> 1, policy.py:290-295), stride 1, 32 input
// channels: dW[tap][c] = sum_q x[q][c] * dy[q + pad - tap] is a correlation of each input channel with a scalar map -- 288
// outputs from a 1.3 GB read, memory-bound, and on the MMA path 7 of 8 columns would be padding.  Plain fp32 FMAs (exact
// products): a warp walks input rows, lane = channel, one coalesced 128-byte load per pixel; the KH x KW window of dy values
// that pixel meets slides along the row in registers (KH uniform loads per pixel).  Fixed row -> warp assignment and a
// fixed-order reduction: deterministic.
template <int KH, int KW>
__global__ void __launch_bounds__(256)
lwgrad_scalar_kernel(const float *__restrict__ x, const float *__restrict__ dy, float *__restrict__ partial, int B, int Hin, int Win,
                     int Hout, int Wout, int pad_t, int pad_l)
{
    __shared__ float s_red[8][KH * KW][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * 8 + warp, nw = gridDim.x * 8;
    float acc[KH * KW];
#pragma unroll
    for (int i = 0; i < KH * KW; i++) acc[i] = 0.0f;
    const int rows = B * Hin;
    for (int rw = gw; rw < rows; rw += nw) {
        const int img = rw / Hin, iy = rw - img * Hin;
        const float *xr = x + (size_t)rw * Win * 32 + lane;
        const float *dr[KH];
        bool ok[KH];
#pragma unroll
        for (int ky = 0; ky < KH; ky++) {
            const int oy = iy + pad_t - ky;
            ok[ky] = oy >= 0 && oy < Hout;
            dr[ky] = dy + ((size_t)img * Hout + (ok[ky] ? oy : 0)) * Wout;
        }
        float w[KH][KW];                                                       // w[ky][kx] = dy[oy(ky)][ix + pad_l - kx]
#pragma unroll
        for (int ky = 0; ky < KH; ky++)
#pragma unroll
            for (int kx = 0; kx < KW; kx++) {
                const int o = pad_l - kx - 1;                                  // the window of ix = -1
                w[ky][kx] = (ok[ky] && o >= 0 && o < Wout) ? __ldg(dr[ky] + o) : 0.0f;
            }
        constexpr int U = 8;
        for (int ix0 = 0; ix0 < Win; ix0 += U) {
            float xv[U], dn[U][KH];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int ix = ix0 + u;
                xv[u] = ix < Win ? __ldg(xr + (size_t)ix * 32) : 0.0f;
                const int o = ix + pad_l;
#pragma unroll
                for (int ky = 0; ky < KH; ky++) dn[u][ky] = (ok[ky] && o < Wout) ? __ldg(dr[ky] + o) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
#pragma unroll
                for (int ky = 0; ky < KH; ky++) {
#pragma unroll
                    for (int kx = KW - 1; kx > 0; kx--) w[ky][kx] = w[ky][kx - 1];
                    w[ky][0] = dn[u][ky];
#pragma unroll
                    for (int kx = 0; kx < KW; kx++) acc[ky * KW + kx] = fmaf(xv[u], w[ky][kx], acc[ky * KW + kx]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < KH * KW; i++) s_red[warp][i][lane] = acc[i];
    __syncthreads();
    for (int k = threadIdx.x; k < KH * KW * 32; k += 256) {
        float sum = 0.0f;
#pragma unroll
        for (int wv = 0; wv < 8; wv++) sum += s_red[wv][k >> 5][k & 31];
        partial[(size_t)blockIdx.x * (KH * KW * 32) + k] = sum;
    }
}

__global__ void __launch_bounds__(256)
lwgrad_scalar_reduce_kernel(const float *__restrict__ partial, float *__restrict__ dw, int n_cta, int K)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    float s = 0.0f;
    for (int c = 0; c < n_cta; c++) s += partial[(size_t)c * K + k];
    dw[k] = s;
}

static void lc_fill(LConvParams &p, int B, int Hin, int Win, int Cin, int Hout, int Wout, int Cout, int KH, int KW, int stride,
                    int pad_t, int pad_l, int oH, int oW, int o_sy, int o_sx, int o_oy, int o_ox)
{
    p.B = B; p.Hin = Hin; p.Win = Win; p.Hout = Hout; p.Wout = Wout;
    p.Cout = Cout; p.CoutPad = (Cout + 7) & ~7;
    p.KH = KH; p.KW = KW; p.stride = stride; p.pad_t = pad_t; p.pad_l = pad_l;
    p.K = KH * KW * Cin; p.KP = (p.K + 15) & ~15; p.KS = p.KP + 8;
    p.IH = (kLcTh - 1) * stride + KH; p.IW = (kLcTw - 1) * stride + KW;
    p.tiles_x = (Wout + kLcTw - 1) / kLcTw; p.tiles_y = (Hout + kLcTh - 1) / kLcTh;
    p.oH = oH; p.oW = oW; p.o_sy = o_sy; p.o_sx = o_sx; p.o_oy = o_oy; p.o_ox = o_ox;
    p.tw = kLcTw; p.th = kLcTh; p.fr = 1; p.mt_rows = 1; p.th_log = 3; p.n_in = 0;
    p.dy_pitch = Cout; p.dy_c0 = 0; p.relu = 0; p.lw_ks = 8;
    p.n_cls = 1;
    for (int c = 0; c < 4; c++) { p.cls_dy[c] = 0; p.cls_dx[c] = 0; p.cls_oy[c] = o_oy; p.cls_ox[c] = o_ox; }
}

// tile of the v2 kernels: `rows` = output rows a CTA covers over all its frames (8 warps x m-tiles per warp x mt_rows)
static void lc2_tile(LConvParams &p, int Cin, int m_tiles_per_warp)
{
    p.tw = p.Wout <= 8 ? 8 : 16;
    p.mt_rows = 16 / p.tw;
    const int rows_w = m_tiles_per_warp * p.mt_rows, rows = 8 * rows_w;
    int th = rows_w;
    while (th < p.Hout && th < rows) th *= 2;
    p.th = th;
    p.th_log = 0;
    while ((1 << p.th_log) < th) p.th_log++;
    p.fr = rows / th;
    const int ext = p.n_cls > 1 ? 1 : 0;                                      // classes shift their taps by 0 or 1
    p.IH = (p.th - 1) * p.stride + p.KH + ext; p.IW = (p.tw - 1) * p.stride + p.KW + ext;
    p.tiles_x = (p.Wout + p.tw - 1) / p.tw; p.tiles_y = (p.Hout + p.th - 1) / p.th;
    p.n_in = (p.fr * p.IH * p.IW * Cin + 63) & ~63;
    p.mIH = lc_magic(p.IH); p.mIW = lc_magic(p.IW); p.mTH = lc_magic(p.th); p.mTW = lc_magic(p.tw);
}

constexpr size_t kLcSmemMax = 227 * 1024;        // per-CTA shared memory on sm_100

long long bounds_violations_learn()
{
#ifdef SPIRAL_BOUNDS_CHECK
    return spiral_tu_bounds_violations();
#else
    return -1;
#endif
}

}  // namespace spiral

using namespace spiral;

static bool lc_cin_ok(int Cin) { return Cin == 1 || Cin == 2 || Cin == 3 || Cin == 6 || Cin == 8 || Cin == 16 || Cin == 32; }
static int lc_cin_pad(int Cin) { return Cin < 8 ? 8 : Cin; }                  // channels per tap of the staged tile / packed kernel
// operand formats (the `n_planes` argument of the C ABI): 2 = bf16x3, 3 = bf16x6, 4 = f16x3 (two fp16 planes)
static bool lc_prec_ok(int code) { return code == 2 || code == 3 || code == 4; }
static int lc_prec_planes(int code) { return code == 3 ? 3 : 2; }

// elements of ONE plane of a packed kernel
extern "C" int64_t spiral_lconv_packed_elems(int32_t Cin, int32_t Cout, int32_t n_taps)
{
    const int K = n_taps * lc_cin_pad(Cin), KP = (K + 15) & ~15, KS = KP + 8, CoutPad = (Cout + 7) & ~7;
    return (int64_t)CoutPad * KS;
}

extern "C" int spiral_lconv_pack(const float *w, void *packed_planes, int32_t n_planes, int32_t rows, int32_t Ck, int32_t n_taps,
                                 int64_t s_row, int64_t s_c, int64_t s_ky, int64_t s_kx, const int32_t *tap_ky,
                                 const int32_t *tap_kx, void *stream)
{
    SPIRAL_REQUIRE(w); SPIRAL_REQUIRE(packed_planes); SPIRAL_REQUIRE(tap_ky); SPIRAL_REQUIRE(tap_kx);
    if (rows < 1 || rows > 32 || !lc_cin_ok(Ck) || n_taps < 1 || n_taps > kLcMaxTaps || !lc_prec_ok(n_planes))
        return spiral_set_error(SPIRAL_EINVAL, "lconv_pack: unsupported shape rows=%d Ck=%d taps=%d planes=%d", rows, Ck, n_taps, n_planes);
    LPackParams p;
    p.rows = rows; p.rows_pad = (rows + 7) & ~7; p.Ck = lc_cin_pad(Ck); p.Ck_src = Ck; p.n_taps = n_taps; p.perm = 0;
    p.K = n_taps * p.Ck; p.KS = ((p.K + 15) & ~15) + 8;
    p.s_row = s_row; p.s_c = s_c; p.s_ky = s_ky; p.s_kx = s_kx;
    for (int i = 0; i < kLcMaxTaps; i++) { p.tap_ky[i] = i < n_taps ? tap_ky[i] : 0; p.tap_kx[i] = i < n_taps ? tap_kx[i] : 0; }
    const int n = p.rows_pad * p.KS;
    if (n_planes == 2) lpack_kernel<2><<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p, w, (uint16_t *)packed_planes);
    else if (n_planes == 4) lpack_kernel<2, true><<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p, w, (uint16_t *)packed_planes);
    else lpack_kernel<3><<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p, w, (uint16_t *)packed_planes);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_lconv_pack");
}

// The ACTING path's kernels (policy_conv.cu: one bf16 plane [round8(rows)][round16(taps * Ck) + 8], channels NOT padded,
// rows channel-permuted when there are 32 of them) from the same strided gather -- one launch per layer when the actors pull
// new weights (NativeActor.sync inside the learner step's CUDA graph) instead of a dozen indexing kernels.
extern "C" int spiral_pconv_pack(const float *w, void *packed_bf16, int32_t rows, int32_t Ck, int32_t n_taps, int64_t s_row,
                                 int64_t s_c, int64_t s_ky, int64_t s_kx, const int32_t *tap_ky, const int32_t *tap_kx,
                                 int32_t permute_rows, void *stream)
{
    SPIRAL_REQUIRE(w); SPIRAL_REQUIRE(packed_bf16); SPIRAL_REQUIRE(tap_ky); SPIRAL_REQUIRE(tap_kx);
    if (rows < 1 || rows > 32 || Ck < 1 || Ck > 32 || n_taps < 1 || n_taps > kLcMaxTaps || (permute_rows && rows != 32))
        return spiral_set_error(SPIRAL_EINVAL, "pconv_pack: unsupported shape rows=%d Ck=%d taps=%d", rows, Ck, n_taps);
    LPackParams p;
    p.rows = rows; p.rows_pad = (rows + 7) & ~7; p.Ck = Ck; p.Ck_src = Ck; p.n_taps = n_taps; p.perm = permute_rows ? 1 : 0;
    p.K = n_taps * Ck; p.KS = ((p.K + 15) & ~15) + 8;
    p.s_row = s_row; p.s_c = s_c; p.s_ky = s_ky; p.s_kx = s_kx;
    for (int i = 0; i < kLcMaxTaps; i++) { p.tap_ky[i] = i < n_taps ? tap_ky[i] : 0; p.tap_kx[i] = i < n_taps ? tap_kx[i] : 0; }
    const int n = p.rows_pad * p.KS;
    lpack_kernel<1><<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p, w, (uint16_t *)packed_bf16);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_pconv_pack");
}

template <int CSRC, int NPL, bool F16, int MT>
static int lconv2_launch(const LConvParams &p, const float *x, const void *w, const float *bias, float *out, int grid, size_t smem,
                         cudaStream_t st)
{
    static bool attr = false;
    if (!attr) {
        const cudaError_t e = cudaFuncSetAttribute(lconv2_kernel<CSRC, NPL, F16, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLcSmemMax);
        if (e != cudaSuccess) return spiral_cuda_error(e, "cudaFuncSetAttribute(lconv2_kernel)");
        attr = true;
    }
    LConvExtra none;
    memset(&none, 0, sizeof(none));
    lconv2_kernel<CSRC, NPL, F16, MT><<<grid, 256, smem, st>>>(p, x, (const uint16_t *)w, bias, out, none);
    return SPIRAL_OK;
}

// the discriminator's variant (input transform, output slice, batch-norm partials), operand format f16x3
template <int CSRC, int MT>
static int lconv2x_launch(const LConvParams &p, const float *x, const void *w, const float *bias, float *out, const LConvExtra &ex,
                          int grid, size_t smem, cudaStream_t st)
{
    static bool attr = false;
    if (!attr) {
        // 2 KB of static shared memory next to the dynamic tile
        const cudaError_t e = cudaFuncSetAttribute(lconv2_kernel<CSRC, 2, true, MT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (int)(kLcSmemMax - 2048));
        if (e != cudaSuccess) return spiral_cuda_error(e, "cudaFuncSetAttribute(lconv2_kernel, discriminator variant)");
        attr = true;
    }
    lconv2_kernel<CSRC, 2, true, MT, true><<<grid, 256, smem, st>>>(p, x, (const uint16_t *)w, bias, out, ex);
    return SPIRAL_OK;
}

template <int CSRC, int NPL, bool F16>
static int lwgrad2_launch(const LConvParams &p, const float *x, const float *dy, float *ws, int n_items, int grid, size_t smem,
                          cudaStream_t st)
{
    static bool attr = false;
    if (!attr) {
        const cudaError_t e = cudaFuncSetAttribute(lwgrad2_kernel<CSRC, NPL, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLcSmemMax);
        if (e != cudaSuccess) return spiral_cuda_error(e, "cudaFuncSetAttribute(lwgrad2_kernel)");
        attr = true;
    }
    lwgrad2_kernel<CSRC, NPL, F16><<<grid, 256, smem, st>>>(p, x, dy, ws, n_items);
    return SPIRAL_OK;
}

template <int CSRC, int NPL, bool F16>
static int lconv2_launch_mt(int mt, const LConvParams &p, const float *x, const void *w, const float *bias, float *out, int grid,
                            size_t smem, cudaStream_t st)
{
    return mt == 2 ? lconv2_launch<CSRC, NPL, F16, 2>(p, x, w, bias, out, grid, smem, st)
                   : lconv2_launch<CSRC, NPL, F16, 1>(p, x, w, bias, out, grid, smem, st);
}

// dispatch over the tensor's channel count x the operand format; FN<CSRC, NPL, F16>
#define LC2_FORMAT(FN, CS, ...)                                                                                       \
    (n_planes == 4 ? FN<CS, 2, true>(__VA_ARGS__) : n_planes == 2 ? FN<CS, 2, false>(__VA_ARGS__) : FN<CS, 3, false>(__VA_ARGS__))
#define LC2_DISPATCH(FN, ...)                                                                                         \
    do {                                                                                                              \
        int rc_;                                                                                                      \
        switch (Cin) {                                                                                                \
        case 1: rc_ = LC2_FORMAT(FN, 1, __VA_ARGS__); break;                                                          \
        case 2: rc_ = LC2_FORMAT(FN, 2, __VA_ARGS__); break;                                                          \
        case 3: rc_ = LC2_FORMAT(FN, 3, __VA_ARGS__); break;                                                          \
        case 6: rc_ = LC2_FORMAT(FN, 6, __VA_ARGS__); break;                                                          \
        case 8: rc_ = LC2_FORMAT(FN, 8, __VA_ARGS__); break;                                                          \
        case 16: rc_ = LC2_FORMAT(FN, 16, __VA_ARGS__); break;                                                        \
        default: rc_ = LC2_FORMAT(FN, 32, __VA_ARGS__); break;                                                        \
        }                                                                                                             \
        if (rc_ != SPIRAL_OK) return rc_;                                                                             \
    } while (0)

extern "C" int spiral_lconv_forward(const float *x, const void *w_planes, int32_t n_planes, const float *bias, float *out,
                                    int32_t B, int32_t Hin, int32_t Win, int32_t Cin, int32_t Hout, int32_t Wout, int32_t Cout,
                                    int32_t KH, int32_t KW, int32_t stride, int32_t pad_t, int32_t pad_l, int32_t oH,
                                    int32_t oW, int32_t o_sy, int32_t o_sx, int32_t o_oy, int32_t o_ox, int32_t relu, void *stream)
{
    SPIRAL_REQUIRE(x); SPIRAL_REQUIRE(w_planes); SPIRAL_REQUIRE(out);
    if (B <= 0 || Cout <= 0 || Cout > 32 || !lc_cin_ok(Cin) || KH * KW > kLcMaxTaps || !lc_prec_ok(n_planes))
        return spiral_set_error(SPIRAL_EINVAL, "lconv: unsupported shape Cin=%d Cout=%d taps=%d planes=%d", Cin, Cout, KH * KW, n_planes);
    const int CinP = lc_cin_pad(Cin);
    LConvParams p;
    lc_fill(p, B, Hin, Win, CinP, Hout, Wout, Cout, KH, KW, stride, pad_t, pad_l, oH, oW, o_sy, o_sx, o_oy, o_ox);
    p.relu = relu ? 1 : 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int npl = lc_prec_planes(n_planes);
    const size_t n_w = (size_t)p.CoutPad * p.KS;
    for (int mt = 2; mt >= 1; mt--) {
        lc2_tile(p, CinP, mt);
        const size_t smem = (size_t)npl * ((size_t)p.n_in + n_w) * 2 + (size_t)(p.KP / 8) * 4;
        if (smem > kLcSmemMax) {
            if (mt == 1) return spiral_set_error(SPIRAL_EUNSUPPORTED, "lconv: tile does not fit shared memory (%zu bytes)", smem);
            continue;
        }
        const int grid = ((B + p.fr - 1) / p.fr) * p.tiles_x * p.tiles_y;
        LC2_DISPATCH(lconv2_launch_mt, mt, p, x, w_planes, bias, out, grid, smem, st);
        break;
    }
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_lconv_forward");
}

// The four parity classes (c = py * 2 + px) of a transposed convolution's forward, or of a stride-2 convolution's input
// gradient, in ONE launch: 2x2-tap / stride-1 convolutions over the same input that differ in their padding (0 or 1 per
// axis), their kernel and the output pixels they own --
//   out[b, 2 oy + py, 2 ox + px, co] = bias[co] + sum_{ky,kx,ci} x[b, oy - pad_t[c] + ky, ox - pad_l[c] + kx, ci] * packed_c[co][(ky*2 + kx)*Cin + ci]
// for oy < Hout, ox < Wout (pixels outside [oH, oW] are skipped).  The input tile is staged once (with padding 1) for all
// four; w_planes holds the classes' packed kernels back to back (spiral_lconv_pack with the class's four taps).
extern "C" int spiral_lconv_forward_classes(const float *x, const void *w_planes, int32_t n_planes, const float *bias, float *out,
                                            int32_t B, int32_t Hin, int32_t Win, int32_t Cin, int32_t Hout, int32_t Wout,
                                            int32_t Cout, const int32_t *cls_pad_t, const int32_t *cls_pad_l, int32_t oH,
                                            int32_t oW, int32_t relu, void *stream)
{
    SPIRAL_REQUIRE(x); SPIRAL_REQUIRE(w_planes); SPIRAL_REQUIRE(out); SPIRAL_REQUIRE(cls_pad_t); SPIRAL_REQUIRE(cls_pad_l);
    if (B <= 0 || Cout <= 0 || Cout > 32 || !lc_cin_ok(Cin) || !lc_prec_ok(n_planes))
        return spiral_set_error(SPIRAL_EINVAL, "lconv classes: unsupported shape Cin=%d Cout=%d planes=%d", Cin, Cout, n_planes);
    for (int c = 0; c < 4; c++)
        if ((cls_pad_t[c] | cls_pad_l[c]) & ~1) return spiral_set_error(SPIRAL_EINVAL, "lconv classes: paddings must be 0 or 1");
    const int CinP = lc_cin_pad(Cin);
    LConvParams p;
    lc_fill(p, B, Hin, Win, CinP, Hout, Wout, Cout, 2, 2, 1, 1, 1, oH, oW, 2, 2, 0, 0);
    p.relu = relu ? 1 : 0;
    p.n_cls = 4;
    for (int c = 0; c < 4; c++) {
        p.cls_dy[c] = 1 - cls_pad_t[c]; p.cls_dx[c] = 1 - cls_pad_l[c];
        p.cls_oy[c] = c >> 1; p.cls_ox[c] = c & 1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int npl = lc_prec_planes(n_planes);
    const size_t n_w = (size_t)p.CoutPad * p.KS;
    for (int mt = 2; mt >= 1; mt--) {
        lc2_tile(p, CinP, mt);
        const size_t smem = (size_t)npl * ((size_t)p.n_in + 4 * n_w) * 2 + (size_t)4 * (p.KP / 8) * 4;
        if (smem > kLcSmemMax) {
            if (mt == 1) return spiral_set_error(SPIRAL_EUNSUPPORTED, "lconv classes: tile does not fit shared memory (%zu bytes)", smem);
            continue;
        }
        const int grid = ((B + p.fr - 1) / p.fr) * p.tiles_x * p.tiles_y;
        LC2_DISPATCH(lconv2_launch_mt, mt, p, x, w_planes, bias, out, grid, smem, st);
        break;
    }
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_lconv_forward_classes");
}

// A narrow layer of the discriminator (disc_dim 8 / 16 / 32; 5x5 / stride 2 / 'SAME' = pad 1 before, 2 after) on
// lconv2_kernel's discriminator variant instead of the fp32 SIMT tiles (64 x 64 outputs per CTA: at Cout = 16, Cin = 8
// seven eighths of their FMAs are padding; the image layer's kernel is a 25-term scalar loop): kernel HWIO fp32 -> two fp16
// planes (f16x3: 22 operand bits), the producer's BN + leaky-ReLU -- or, for layer 0, the image normalisation -- applied
// while the input tile is staged, + bias, RAW fp32 output, per-CTA batch-norm partials.  Up to 128 output channels as
// slices of 32.  *n_part = rows of the partials (CTAs of one launch).
namespace spiral {
bool lconv_disc_layer_supported(int Cin, int Cout)
{
    return lc_cin_ok(Cin) && Cout % 2 == 0 && (Cout <= 32 || (Cout % 32 == 0 && Cout <= 128));
}
size_t lconv_disc_pack_bytes(int Cin, int Cout)
{
    const int slices = (Cout + 31) / 32;
    return (size_t)slices * 2 * spiral_lconv_packed_elems(Cin, Cout < 32 ? Cout : 32, 25) * sizeof(uint16_t);
}

template <int MT>
static int lconv2x_dispatch(int Cin, const LConvParams &p, const float *x, const void *w, const float *bias, float *out,
                            const LConvExtra &ex, int grid, size_t smem, cudaStream_t st)
{
    switch (Cin) {
    case 1: return lconv2x_launch<1, MT>(p, x, w, bias, out, ex, grid, smem, st);
    case 2: return lconv2x_launch<2, MT>(p, x, w, bias, out, ex, grid, smem, st);
    case 3: return lconv2x_launch<3, MT>(p, x, w, bias, out, ex, grid, smem, st);
    case 6: return lconv2x_launch<6, MT>(p, x, w, bias, out, ex, grid, smem, st);
    case 8: return lconv2x_launch<8, MT>(p, x, w, bias, out, ex, grid, smem, st);
    case 16: return lconv2x_launch<16, MT>(p, x, w, bias, out, ex, grid, smem, st);
    default: return lconv2x_launch<32, MT>(p, x, w, bias, out, ex, grid, smem, st);
    }
}

// in_mode 0: the input as it is; 1: leaky_relu(in * in_scale[c] + in_shift[c]); 2: in * in_a + in_b
int lconv_disc_layer(const float *in, int in_mode, const float *in_scale, const float *in_shift, float in_a, float in_b,
                     const float *w_hwio, const float *bias, float *out, float *bn_partial, void *pack_scratch, int B, int H,
                     int Cin, int Cout, int *n_part, cudaStream_t st)
{
    int tk[25], tx[25];
    for (int i = 0; i < 25; i++) { tk[i] = i / 5; tx[i] = i % 5; }
    const int Ho = H / 2, CinP = lc_cin_pad(Cin);
    for (int c0 = 0; c0 < Cout; c0 += 32) {
        const int cs = Cout - c0 < 32 ? Cout - c0 : 32;
        char *planes = (char *)pack_scratch + (size_t)(c0 / 32) * 2 * spiral_lconv_packed_elems(Cin, Cout < 32 ? Cout : 32, 25) * sizeof(uint16_t);
        int rc = spiral_lconv_pack(w_hwio + c0, planes, 4, cs, Cin, 25, 1, Cout, (int64_t)5 * Cin * Cout, (int64_t)Cin * Cout, tk, tx, st);
        if (rc != SPIRAL_OK) return rc;
        LConvParams p;
        lc_fill(p, B, H, H, CinP, Ho, Ho, cs, 5, 5, 2, 1, 1, Ho, Ho, 1, 1, 0, 0);
        LConvExtra ex;
        ex.in_scale = in_scale; ex.in_shift = in_shift; ex.in_a = in_a; ex.in_b = in_b; ex.in_mode = in_mode;
        ex.bn_partial = bn_partial; ex.out_ld = Cout; ex.out_c0 = c0; ex.bn_ld = Cout;
        const size_t n_w = (size_t)p.CoutPad * p.KS;
        bool done = false;
        for (int mt = 2; mt >= 1 && !done; mt--) {
            lc2_tile(p, CinP, mt);
            const size_t smem = (size_t)2 * ((size_t)p.n_in + n_w) * 2 + (size_t)(p.KP / 8) * 4;
            if (smem > kLcSmemMax - 2048) continue;
            const int grid = ((B + p.fr - 1) / p.fr) * p.tiles_x * p.tiles_y;
            *n_part = grid;
            rc = mt == 2 ? lconv2x_dispatch<2>(Cin, p, in, planes, bias ? bias + c0 : nullptr, out, ex, grid, smem, st)
                         : lconv2x_dispatch<1>(Cin, p, in, planes, bias ? bias + c0 : nullptr, out, ex, grid, smem, st);
            if (rc != SPIRAL_OK) return rc;
            done = true;
        }
        if (!done) return spiral_set_error(SPIRAL_EUNSUPPORTED, "lconv_disc_layer: tile does not fit shared memory");
    }
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "lconv_disc_layer");
}
}  // namespace spiral

// Kernel gradient of the discriminator's IMAGE layer (5x5 / stride 2 / 'SAME' on 1, 2, 3 or 6 channels -> Cout a multiple of
// 32): on the tcgen05 wgrad GEMM (conv_tc.cu) M = Cin pads 3 rows to 128 and the 25-fold im2col^T of the 64 x 64 input costs
// as much as the GEMM; here it is lwgrad2_kernel over 32-channel slices of dy (the padded input chunk is 8 channels wide).
// dw HWIO [5][5][Cin][Cout].
namespace spiral {
bool lwgrad2_image_layer_supported(int Cin, int Cout) { return (Cin == 1 || Cin == 2 || Cin == 3 || Cin == 6) && Cout % 32 == 0; }
size_t lwgrad2_image_layer_workspace_bytes() { return (size_t)148 * 208 * 32 * sizeof(float); }
int lwgrad2_image_layer(const float *x, const float *dy, float *dw, float *workspace, int n_planes, int B, int H, int Cin, int Cout,
                        cudaStream_t st)
{
    const int Ho = H / 2, CinP = 8;
    for (int c0 = 0; c0 < Cout; c0 += 32) {
        LConvParams p;
        lc_fill(p, B, H, H, CinP, Ho, Ho, 32, 5, 5, 2, 1, 1, Ho, Ho, 1, 1, 0, 0);
        p.dy_pitch = Cout; p.dy_c0 = c0;
        lc2_tile(p, CinP, 1);
        p.lw_ks = 8;
        const int npl = lc_prec_planes(n_planes);
        const size_t smem = (size_t)npl * ((size_t)p.n_in + (size_t)kLwPix * p.CoutPad) * 2 + (size_t)(p.KP / 8) * 4;
        if (smem > kLcSmemMax) return spiral_set_error(SPIRAL_EUNSUPPORTED, "lwgrad2_image_layer: tile does not fit shared memory");
        const int n_items = ((B + p.fr - 1) / p.fr) * p.tiles_x * p.tiles_y;
        const int grid = n_items < 148 ? n_items : 148;
        LC2_DISPATCH(lwgrad2_launch, p, x, dy, workspace, n_items, grid, smem, st);
        const int n = 25 * Cin * 32;
        lwgrad_reduce_kernel<<<(n * 8 + 255) / 256, 256, 0, st>>>(workspace, dw, grid, 25 * Cin, p.KP, 32, p.CoutPad, Cin, CinP, Cout, c0);
    }
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "lwgrad2_image_layer");
}
}  // namespace spiral

extern "C" int64_t spiral_lconv_wgrad_workspace_floats(int32_t Cin, int32_t Cout, int32_t n_taps)
{
    const int K = n_taps * lc_cin_pad(Cin), KP = (K + 15) & ~15, CoutPad = (Cout + 7) & ~7;
    const int64_t mma = (int64_t)148 * KP * CoutPad, scalar = (int64_t)148 * 4 * K;
    return mma > scalar ? mma : scalar;
}

extern "C" int spiral_lconv_wgrad(const float *x, const float *dy, float *dw, float *workspace, int32_t n_planes, int32_t B,
                                  int32_t Hin, int32_t Win, int32_t Cin, int32_t Hout, int32_t Wout, int32_t Cout, int32_t KH,
                                  int32_t KW, int32_t stride, int32_t pad_t, int32_t pad_l, void *stream)
{
    SPIRAL_REQUIRE(x); SPIRAL_REQUIRE(dy); SPIRAL_REQUIRE(dw); SPIRAL_REQUIRE(workspace);
    const int CinP = lc_cin_pad(Cin);
    if (B <= 0 || !(Cout == 1 || Cout == 8 || Cout == 16 || Cout == 32) || !lc_cin_ok(Cin) || KH * KW > kLcMaxTaps ||
        KH * KW * CinP > 8 * kLwMaxMt * 16 || !lc_prec_ok(n_planes))
        return spiral_set_error(SPIRAL_EINVAL, "lconv_wgrad: unsupported shape Cin=%d Cout=%d taps=%d planes=%d", Cin, Cout, KH * KW, n_planes);
    LConvParams p;
    lc_fill(p, B, Hin, Win, CinP, Hout, Wout, Cout, KH, KW, stride, pad_t, pad_l, Hout, Wout, 1, 1, 0, 0);
    cudaStream_t st = (cudaStream_t)stream;
    if (Cout == 1 && Cin == 32 && KH == 3 && KW == 3 && stride == 1) {
        // scalar output map: fp32 correlation (lwgrad_scalar_kernel); partial[cta][288] fits the workspace of the MMA path
        const int rows = B * Hin;
        const int grid = rows < 148 * 4 * 8 ? (rows + 7) / 8 : 148 * 4;
        lwgrad_scalar_kernel<3, 3><<<grid, 256, 0, st>>>(x, dy, workspace, B, Hin, Win, Hout, Wout, pad_t, pad_l);
        lwgrad_scalar_reduce_kernel<<<(288 + 255) / 256, 256, 0, st>>>(workspace, dw, grid, 288);
        const cudaError_t e1 = cudaGetLastError();
        return e1 == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e1, "spiral_lconv_wgrad");
    }
    const int npl = lc_prec_planes(n_planes);
    // items of 256 output pixels when there are enough of them to keep 148 persistent CTAs balanced (the per-item cost --
    // two barriers, one staging round trip -- is what the large maps pay), else of 128
    size_t smem = 0;
    int n_items = 0;
    for (int mtw = 2; mtw >= 1; mtw--) {
        lc2_tile(p, CinP, mtw);
        p.lw_ks = 8 * mtw;
        smem = (size_t)npl * ((size_t)p.n_in + (size_t)p.lw_ks * 16 * p.CoutPad) * 2 + (size_t)(p.KP / 8) * 4;
        n_items = ((B + p.fr - 1) / p.fr) * p.tiles_x * p.tiles_y;
        if (mtw == 1 || (smem <= kLcSmemMax && n_items >= 148 * 8)) break;
    }
    if (smem > kLcSmemMax) return spiral_set_error(SPIRAL_EUNSUPPORTED, "lconv_wgrad: tile does not fit shared memory (%zu bytes)", smem);
    const int grid = n_items < 148 ? n_items : 148;
    LC2_DISPATCH(lwgrad2_launch, p, x, dy, workspace, n_items, grid, smem, st);
    const int n = KH * KW * Cin * Cout;
    lwgrad_reduce_kernel<<<(n * 8 + 255) / 256, 256, 0, st>>>(workspace, dw, grid, KH * KW * Cin, p.KP, Cout, p.CoutPad, Cin, CinP);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_lconv_wgrad");
}
