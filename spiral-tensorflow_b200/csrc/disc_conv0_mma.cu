// disc_conv0_mma.cu -- K2, layer 0 on tensor cores.
//
// Layer 0 of the discriminator (models/discriminator.py:61-81, i = 0) is a 5x5 / stride-2 / 'SAME'
// convolution with K = 25 * C (C = 1, 3 or 6 input channels): too narrow for a TMA-fed tcgen05
// pipeline (a 3-channel NHWC pixel is 6 bytes of bf16; TMA boxes need 16-byte inner extents), and as
// a register-tiled SIMT kernel it was the single most expensive kernel of the reward forward (ncu,
// round 1: 7.6 of 27 ms at B = 16384, issue-bound at 21 TFLOP/s).  Here the im2col never exists in
// memory: the normalised input tile is staged once in shared memory as bf16 hi/lo, every thread
// gathers its mma.m16n8k16 A fragment straight from it through a per-k offset table, B fragments come
// from the pre-split, pre-transposed kernel in shared memory, and the same bf16x3 split as the other
// layers (hi*hi + hi*lo + lo*hi, fp32 accumulate) keeps fp32-equivalent accuracy.  Epilogue: + bias,
// leaky-ReLU (no BN at layer 0), hi/lo split, and -- the packed kernel's rows are permuted so that a thread's
// accumulator fragment is 16 consecutive channels -- 128-bit stores straight from registers into the bf16
// NHWC pair layer 1's TMA reads.
//
// CTA = one image, 8 warps, looping over its 8 tiles of 8 output rows x 16 output pixels x 64 channels (the
// packed kernel is staged once per CTA); warp = one row of 16 pixels (M = 16) x all 64 channels (8 n-tiles),
// K padded to a multiple of 16 with zero weights.
#include "disc.cuh"

#include <cuda_bf16.h>

namespace spiral {

constexpr int kC0Tw = 16, kC0Th = 8;                    // output tile
constexpr int kC0Iw = 2 * kC0Tw + 3, kC0Ih = 2 * kC0Th + 3;   // input tile 35 x 19
constexpr int kC0Cout = 64;

// K order: k' = ky * SEG + (kx * CIN + c), every kernel row ky padded from 5*CIN to SEG = 8 / 16 / 32 entries (zero
// weights).  The 5*CIN inputs of one kernel row are CONSECUTIVE elements of the staged NHWC input row, so an MMA
// A-fragment pair (k', k'+1) is one aligned 32-bit shared-memory load at a compile-time offset -- no k -> offset
// table, no 16-bit loads, no packing -- and the pad entries read finite neighbours that meet zero weights.
__host__ __device__ constexpr int c0_seg(int cin) { return cin == 1 ? 8 : (cin <= 3 ? 16 : 32); }   // >= 5 * cin
__host__ __device__ constexpr int c0_kpad(int cin) { return (5 * c0_seg(cin) + 15) / 16 * 16; }
__host__ __device__ constexpr int c0_kstride(int cin) { return c0_kpad(cin) + 8; }   // bf16 per row; conflict-free, 16-byte rows
__host__ __device__ constexpr int c0_rowstride(int cin) { return (kC0Iw * cin + 1) & ~1; }   // staged input row, even

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// Row n = nt * 8 + j of the packed kernel (column j of n-tile nt of the MMA) holds output channel
//   (j >> 1) * 16 + nt * 2 + (j & 1)
// so that the accumulator fragment of thread t (columns 2t, 2t+1 of every n-tile) is the 16 CONSECUTIVE
// channels t*16 .. t*16+15 of its pixel: the epilogue stores 128-bit words straight from registers.
__host__ __device__ constexpr int c0_channel_of_row(int n) { return ((n & 7) >> 1) * 16 + (n >> 3) * 2 + (n & 1); }

__device__ __forceinline__ uint32_t pack2(uint16_t lo, uint16_t hi) { return (uint32_t)lo | ((uint32_t)hi << 16); }

// HWIO [25][CIN][64] fp32 -> [64][KS] bf16 hi / lo in the k' order above (zero padded), followed by the 64 effective
// biases.  With `normalize` (set in the bf16x3 parity mode) the input normalisation (x - 127.5) / 127.5 (envs/base.py:51-52) is folded into the layer:
//   sum_k W_k (x_k - 127.5) / 127.5 + b  =  sum_k (W_k / 127.5) x_k + (b - sum_k W_k)        (padding: x_k = 127.5)
// so the kernel multiplies RAW pixel values -- integers 0..255 and the padding value 127.5 are exact in bf16, the
// activations' lo plane is identically zero and one of the three bf16x3 products disappears (checked per tile).
__global__ void conv0_pack_kernel(const float *__restrict__ w, const float *__restrict__ bias, __nv_bfloat16 *__restrict__ hi,
                                  __nv_bfloat16 *__restrict__ lo, float *__restrict__ bias_eff, int Cin, int normalize)
{
    const int KS = c0_kstride(Cin), SEG = c0_seg(Cin);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kC0Cout * KS; i += gridDim.x * blockDim.x) {
        const int n = i / KS, kp = i % KS;
        const int ch = c0_channel_of_row(n);
        const int ky = kp / SEG, j = kp % SEG;                      // j = kx * Cin + c
        float v = (ky < 5 && j < 5 * Cin) ? w[(size_t)(ky * 5 * Cin + j) * kC0Cout + ch] : 0.0f;
        if (normalize) v = v / 127.5f;
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        hi[i] = h;
        lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch < kC0Cout) {
        double acc = 0.0;
        if (normalize)
            for (int k = 0; k < 25 * Cin; k++) acc += (double)w[(size_t)k * kC0Cout + ch];
        bias_eff[ch] = (float)((double)bias[ch] - acc);
    }
}

template <int CIN>
__global__ void __launch_bounds__(256, 4)
conv0_mma_kernel(const float *__restrict__ x, const __nv_bfloat16 *__restrict__ w_hi, const __nv_bfloat16 *__restrict__ w_lo,
                 const float *__restrict__ bias, __nv_bfloat16 *__restrict__ out_hi, __nv_bfloat16 *__restrict__ out_lo,
                 int B, int H, int normalize, int x3)
{
    constexpr int KP = c0_kpad(CIN), KS = c0_kstride(CIN), SEG = c0_seg(CIN), ROW = c0_rowstride(CIN);
    constexpr int RW = kC0Iw * CIN;                                        // real elements per staged input row
    constexpr int NIN = kC0Ih * RW;                                        // elements fetched per tile
    constexpr int NST = (kC0Ih * ROW + 32 + 7) & ~7;                       // staged plane incl. slack for the pad reads
    extern __shared__ __align__(16) uint8_t sm_raw[];
    uint16_t *s_wh = reinterpret_cast<uint16_t *>(sm_raw);                 // [64][KS]
    uint16_t *s_wl = s_wh + kC0Cout * KS;
    uint16_t *s_ih = s_wl + kC0Cout * KS;                                  // [19][ROW]: rows of 35 NHWC pixels
    uint16_t *s_il = s_ih + NST;
    float *s_raw = reinterpret_cast<float *>(s_il + NST);                  // [NIN] raw input of the NEXT tile

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Ho = H / 2;
    const int tiles_x = Ho / kC0Tw, tiles_y = Ho / kC0Th;
    const int img = blockIdx.x;

    // ---- once per CTA (= one image): the packed kernel (16-byte copies); pad / slack entries of the staged planes
    //      are read (against zero weights) and must be finite ----
    {
        const uint4 *gh = reinterpret_cast<const uint4 *>(w_hi), *gl = reinterpret_cast<const uint4 *>(w_lo);
        uint4 *sh = reinterpret_cast<uint4 *>(s_wh), *sl = reinterpret_cast<uint4 *>(s_wl);
        constexpr int n16 = kC0Cout * KS * 2 / 16;
        for (int i = tid; i < n16; i += 256) {
            sh[i] = __ldg(gh + i);
            sl[i] = __ldg(gl + i);
        }
        for (int i = tid; i < 2 * NST / 2; i += 256) reinterpret_cast<uint32_t *>(s_ih)[i] = 0u;     // both planes
    }
    const int g = lane >> 2, t = lane & 3;
    const int base0 = (2 * warp) * ROW + 2 * g * CIN + 2 * t;            // pixel g, fragment columns 2t, 2t+1
    const int base1 = base0 + 16 * CIN;                                  // pixel g + 8
    __shared__ float s_bias[kC0Cout];
    if (tid < kC0Cout) s_bias[tid] = __ldg(reinterpret_cast<const float *>(w_hi + kC0Cout * KS) + tid);   // effective bias (conv0_pack_kernel)
    const float *bias_r = s_bias + t * 16;                               // the thread's channels t*16 .. t*16+15
    // ldmatrix.x4 row address of this lane: matrices {n-tile 2p: k 0-7, k 8-15, n-tile 2p+1: k 0-7, k 8-15}
    const uint32_t ldm_off = (uint32_t)((((lane >> 4) * 8 + (lane & 7)) * KS + ((lane >> 3) & 1) * 8) * 2);
    const uint32_t s_wh_u32 = (uint32_t)__cvta_generic_to_shared(s_wh) + ldm_off;
    const uint32_t s_wl_u32 = (uint32_t)__cvta_generic_to_shared(s_wl) + ldm_off;

    // Input tile staging, software-pipelined: the raw fp32 values of tile i+1 are fetched with cp.async into a
    // shared staging buffer while tile i's MMAs run, and converted to bf16 hi/lo after them, so the global-load
    // latency (ncu, round 1: 4.1 of 12 stall cycles per issue were long-scoreboard waits in this loop) overlaps the
    // tensor work.  A thread converts exactly the elements it fetched: its own cp.async.wait_group is enough.
    constexpr int NPT = (NIN + 255) / 256;                               // staged elements per thread
    const float *xi = x + (size_t)img * H * H * CIN;
    auto fetch = [&](int tile) {
        const int oy0 = (tile / tiles_x) * kC0Th, ox0 = (tile % tiles_x) * kC0Tw;
#pragma unroll
        for (int e = 0; e < NPT; e++) {
            const int i = tid + e * 256;
            if (i < NIN) {
                const int ly = i / RW, rem = i - ly * RW;                // a staged row is RW consecutive floats of the image
                const int iy = 2 * oy0 - 1 + ly, col = (2 * ox0 - 1) * CIN + rem;
                if (iy >= 0 && iy < H && col >= 0 && col < H * CIN) {
                    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_raw + i);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(xi + (size_t)iy * H * CIN + col) : "memory");
                } else {
                    s_raw[i] = normalize ? 127.5f : 0.0f;                // SAME padding: zero in normalised units
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    fetch(0);
#pragma unroll 1
    for (int tile = 0; tile < tiles_x * tiles_y; tile++) {
    const int oy0 = (tile / tiles_x) * kC0Th, ox0 = (tile % tiles_x) * kC0Tw;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                                     // previous tile's input is consumed
    int lo_nonzero = 0;
#pragma unroll
    for (int e = 0; e < NPT; e++) {
        const int i = tid + e * 256;
        if (i < NIN) {
            float v = s_raw[i];                                          // stays a raw pixel value when the normalisation is folded
            if (normalize && !x3) {
                // plain-bf16 mode keeps the normalised operand (a single rounded product has no room for the fold's
                // cancellation): (v - 127.5) / 127.5 correctly rounded without the division sequence -- with
                // rc = RN(1/127.5), q0 = a*rc, r = fma(-q0, 127.5, a), q = fma(r, rc, q0) equals a / 127.5 for EVERY finite
                // float a with |a| >= 2^-118 (exhaustive check: tools/verify_const_div.c); a = v - 127.5 is 0 or >= 2^-17
                const float a = v - 127.5f, rc = 1.0f / 127.5f;
                const float q0 = a * rc;
                v = fmaf(fmaf(-q0, 127.5f, a), rc, q0);
            }
            const int ly = i / RW, o = ly * ROW + (i - ly * RW);
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            const uint16_t lb = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(h)));
            s_ih[o] = __bfloat16_as_ushort(h);
            s_il[o] = lb;
            lo_nonzero |= lb & 0x7fff;
        }
    }
    // block-uniform: does any staged value of this tile need its lo plane?  (never for 8-bit images)
    const bool need_lo = __syncthreads_or(lo_nonzero) != 0 && x3;
    if (tile + 1 < tiles_x * tiles_y) fetch(tile + 1);

    // ---- main loop: warp = output row `warp` of the tile, 16 pixels x 64 channels ----
    float acc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; nt++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[nt][j] = 0.0f;

#pragma unroll
    for (int ks = 0; ks < KP / 16; ks++) {
        // compile-time offsets of the two 8-wide halves of this k-step inside the staged input
        const int kA = ks * 16, kB = ks * 16 + 8;
        const int kyA = (kA / SEG) < 5 ? (kA / SEG) : 4, kyB = (kB / SEG) < 5 ? (kB / SEG) : 4;   // beyond 5*SEG: zero weights
        const int oA = kyA * ROW + (kA % SEG), oB = kyB * ROW + (kB % SEG);
        uint32_t ah[4], al[4];
        ah[0] = *reinterpret_cast<const uint32_t *>(s_ih + base0 + oA);
        ah[1] = *reinterpret_cast<const uint32_t *>(s_ih + base1 + oA);
        ah[2] = *reinterpret_cast<const uint32_t *>(s_ih + base0 + oB);
        ah[3] = *reinterpret_cast<const uint32_t *>(s_ih + base1 + oB);
        if (need_lo) {
            al[0] = *reinterpret_cast<const uint32_t *>(s_il + base0 + oA);
            al[1] = *reinterpret_cast<const uint32_t *>(s_il + base1 + oA);
            al[2] = *reinterpret_cast<const uint32_t *>(s_il + base0 + oB);
            al[3] = *reinterpret_cast<const uint32_t *>(s_il + base1 + oB);
        }
#pragma unroll
        for (int np = 0; np < 4; np++) {                                 // two n-tiles per ldmatrix.x4
            const uint32_t woff = (uint32_t)((np * 16 * KS + ks * 16) * 2);
            uint32_t bh[4], bl[4];
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                         : "=r"(bh[0]), "=r"(bh[1]), "=r"(bh[2]), "=r"(bh[3]) : "r"(s_wh_u32 + woff));
            if (x3) {
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                             : "=r"(bl[0]), "=r"(bl[1]), "=r"(bl[2]), "=r"(bl[3]) : "r"(s_wl_u32 + woff));
                if (need_lo) {
                    mma_bf16_16816(acc[2 * np], al, *reinterpret_cast<const uint32_t(*)[2]>(&bh[0]));
                    mma_bf16_16816(acc[2 * np + 1], al, *reinterpret_cast<const uint32_t(*)[2]>(&bh[2]));
                }
                mma_bf16_16816(acc[2 * np], ah, *reinterpret_cast<const uint32_t(*)[2]>(&bl[0]));
                mma_bf16_16816(acc[2 * np + 1], ah, *reinterpret_cast<const uint32_t(*)[2]>(&bl[2]));
            }
            mma_bf16_16816(acc[2 * np], ah, *reinterpret_cast<const uint32_t(*)[2]>(&bh[0]));
            mma_bf16_16816(acc[2 * np + 1], ah, *reinterpret_cast<const uint32_t(*)[2]>(&bh[2]));
        }
    }

    // ---- epilogue: + bias, leaky-ReLU, hi/lo split; thread t owns channels t*16 .. t*16+15 of pixels g and g+8 ----
    const size_t opix = ((size_t)img * Ho + oy0 + warp) * Ho + ox0;        // 16 consecutive pixels
#pragma unroll
    for (int half = 0; half < 2; half++) {
        uint32_t h[8], l[8];
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            float v0 = acc[nt][2 * half] + bias_r[nt * 2], v1 = acc[nt][2 * half + 1] + bias_r[nt * 2 + 1];
            v0 = v0 > 0.f ? v0 : 0.2f * v0;
            v1 = v1 > 0.f ? v1 : 0.2f * v1;
            // packed conversions: one cvt.rn.bf16x2.f32 per pair (same round-to-nearest-even as the scalar form)
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(v0, v1);
            const uint32_t hb = *reinterpret_cast<const uint32_t *>(&h2);
            const __nv_bfloat162 l2 = __floats2bfloat162_rn(v0 - __uint_as_float(hb << 16), v1 - __uint_as_float(hb & 0xffff0000u));
            h[nt] = hb;
            l[nt] = *reinterpret_cast<const uint32_t *>(&l2);
        }
        const size_t o = (opix + g + 8 * half) * kC0Cout + t * 16;
        uint4 *gh = reinterpret_cast<uint4 *>(out_hi + o), *gl = reinterpret_cast<uint4 *>(out_lo + o);
        gh[0] = make_uint4(h[0], h[1], h[2], h[3]);
        gh[1] = make_uint4(h[4], h[5], h[6], h[7]);
        gl[0] = make_uint4(l[0], l[1], l[2], l[3]);
        gl[1] = make_uint4(l[4], l[5], l[6], l[7]);
    }
    }   // next tile
}

size_t conv0_mma_pack_bytes(int Cin) { return (size_t)kC0Cout * c0_kstride(Cin) * 2 + kC0Cout * sizeof(float); }   // + effective bias

bool conv0_mma_supported(const SpiralDisc *d)
{
    const int Cin = d->cfg.in_channels, Ho = d->cfg.screen_size / 2;
    return d->cfg.disc_dim == kC0Cout && (Cin == 1 || Cin == 2 || Cin == 3 || Cin == 6) && (Ho % kC0Tw) == 0 && (Ho % kC0Th) == 0;
}

void conv0_mma_pack(const SpiralDisc *d, void *w_hi, void *w_lo, cudaStream_t st)
{
    const int Cin = d->cfg.in_channels;
    float *bias_eff = reinterpret_cast<float *>((char *)w_hi + (size_t)kC0Cout * c0_kstride(Cin) * 2);
    conv0_pack_kernel<<<8, 256, 0, st>>>((const float *)d->w.conv_w[0], (const float *)d->w.conv_b[0], (__nv_bfloat16 *)w_hi,
                                         (__nv_bfloat16 *)w_lo, bias_eff, Cin, d->cfg.normalize && d->cfg.precision == 1);
}

template <int CIN>
static int conv0_launch(const SpiralDisc *d, const float *x, int batch, const void *w_hi, const void *w_lo, void *out_hi,
                        void *out_lo, int x3, cudaStream_t st)
{
    constexpr int KS = c0_kstride(CIN), NIN = kC0Ih * kC0Iw * CIN;
    constexpr int NST = (kC0Ih * c0_rowstride(CIN) + 32 + 7) & ~7;
    constexpr size_t smem = (size_t)2 * kC0Cout * KS * 2 + (size_t)2 * NST * 2 + (size_t)NIN * 4;
    static bool attr = false;
    if (!attr) {
        const cudaError_t e = cudaFuncSetAttribute(conv0_mma_kernel<CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return spiral_cuda_error(e, "cudaFuncSetAttribute(conv0_mma_kernel)");
        attr = true;
    }
    const int H = d->cfg.screen_size, Ho = H / 2;
    const int grid = batch;                         // one CTA per image: the kernel is staged once for its 8 tiles
    conv0_mma_kernel<CIN><<<grid, 256, smem, st>>>(x, (const __nv_bfloat16 *)w_hi, (const __nv_bfloat16 *)w_lo,
                                                   (const float *)d->w.conv_b[0], (__nv_bfloat16 *)out_hi,
                                                   (__nv_bfloat16 *)out_lo, batch, H, d->cfg.normalize, x3);
    return SPIRAL_OK;
}

int conv0_mma_forward(const SpiralDisc *d, const float *x, int batch, const void *w_hi, const void *w_lo, void *out_hi,
                      void *out_lo, int x3, cudaStream_t st)
{
    switch (d->cfg.in_channels) {
    case 1: return conv0_launch<1>(d, x, batch, w_hi, w_lo, out_hi, out_lo, x3, st);
    case 2: return conv0_launch<2>(d, x, batch, w_hi, w_lo, out_hi, out_lo, x3, st);      // conditional D on gray canvases
    case 3: return conv0_launch<3>(d, x, batch, w_hi, w_lo, out_hi, out_lo, x3, st);
    case 6: return conv0_launch<6>(d, x, batch, w_hi, w_lo, out_hi, out_lo, x3, st);
    }
    return spiral_set_error(SPIRAL_EUNSUPPORTED, "conv0_mma: in_channels must be 1, 2, 3 or 6");
}

}  // namespace spiral
