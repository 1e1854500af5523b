// conv_tc.cu -- K2b on the 5th-generation tensor cores: the 5x5 / stride-2 / 'SAME' convolution's
// input gradient (dgrad) and kernel gradient (wgrad) as tcgen05 / TMEM implicit GEMMs, TMA-fed.  Operands are
// fp32 values split into ns bf16 planes (x = p0 + p1 [+ p2]) and every K step issues the products that matter
// into one fp32 TMEM accumulator:  ns = 1 plain bf16;  ns = 2 "bf16x3" (p0*p0 + p0*p1 + p1*p0, 16 operand bits,
// what the reward forward uses);  ns = 3 "bf16x6" (+ p0*p2 + p2*p0 + p1*p1, 24 operand bits = fp32-exact
// products) for gradients: the penalty's double backward through training-mode BN amplifies operand
// round-off ~100x, so the 16-bit split that is ample for the reward is not for the gradient.
// The ns = 3 forward (taps_tc_kernel mode 0) lives here too; ns <= 2 forwards use disc_tc.cu's persistent kernel.
//
// dgrad   dx[n, iy, ix, ci] = sum_{ky,kx,co} dy[n, (iy+1-ky)/2, (ix+1-kx)/2, co] * w[ky,kx,ci,co]
//         A stride-2 transposed convolution decomposes by the parity of (iy, ix) into four stride-1
//         convolutions over dy with 2x2 / 2x3 / 3x2 / 3x3 of the 25 taps:  iy = 2a + py  =>  ky has the
//         parity of py + 1 and the dy row is a + (py + 1 - ky) / 2.   Per class an implicit GEMM
//             D[m, ci] = sum_{tap in class} sum_co  dy[(img, a+sy, b+sx), co] * w[tap][ci][co]
//         M = B*Ha*Ha pixels (Ha = H/2), N = Cin, K = taps*Cout.  The A tile of a (tap, 64-channel chunk)
//         is ONE 4-D TMA box over the NHWC dy tensor whose start coordinate carries the (sx, sy) shift;
//         out-of-range rows/columns are zero-filled by TMA, which is exactly the border behaviour.
//         The B tile comes straight from the HWIO kernel: for a fixed tap, [ci][co] with co contiguous
//         is already K-major (3-D TMA {Cout, Cin, 25}; rows past Cin zero-fill, so Cin = 3 works).
//         The epilogue scatters rows to (2a+py, 2b+px).
//
// wgrad   dw[tap][ci][co] = sum_p x[pix(p, tap), ci] * dy[p, co],  p = (img, oy, ox)
//         The reduction runs over pixels, so both operands must be pixel-contiguous (K-major): a
//         pre-pass writes x^T per tap (im2col transposed, [25][Cin][P]) and dy^T ([Cout][P]) as bf16
//         hi/lo; then per tap a plain GEMM  D[ci, co] = sum_p XT[tap][ci][p] * DYT[co][p]  with the
//         pixel range split over CTAs (partials reduced in a fixed order: deterministic).
//
// Warp roles and the smem ring are those of conv_tc_kernel: warp 0 TMA producer, warp 1 MMA issuer
// (one elected lane) + TMEM allocator, warps 2..5 epilogue (tcgen05.ld -> fp32 stores).
#include "conv_ops.cuh"

#include <cstdio>

#include "../../include/spiral_b200.h"
#include "common.h"
#include "tc_ptx.cuh"

namespace spiral {

// from disc_tc.cu
void conv_tc_pack_weights(const float *w, void *w_hi, void *w_lo, int Cin, int Cout, cudaStream_t st);
int conv_tc_layer(void *a_hi, void *a_lo, void *w_hi, void *w_lo, float *out, int B, int H, int Cin, int Cout,
                  int x3, cudaStream_t st);
// from conv_ops.cu
__global__ void reduce_slices_kernel(const float *part, float *out, size_t n, int n_slices);

constexpr int kM = 128;           // UMMA M (tile rows)
constexpr int kK = 64;            // K per stage: one 128-byte swizzle row of bf16
constexpr int kStagesG = 3;
constexpr int kThreadsG = 192;
constexpr uint32_t kATile = kM * kK * 2;                         // 16 KB

constexpr int kMaxSplit = 3;
constexpr int kMaxTapSlices = 5;   // taps_tc_kernel: CTAs sharing one tile's taps at small batch
// The tensor core adds each MMA into the fp32 TMEM accumulator with truncation, so the accumulation error grows
// with the number of MMAs chained into one accumulator (measured: bf16x6 through ONE accumulator is less accurate
// than bf16x3).  Four accumulators per tile: the dominant p0*p0 products alternate between 0 and 1 (halving their
// chain), every cross / low-order product (2^-8 .. 2^-16 of the result, whose truncation is invisible) goes to 2
// and 3; the epilogue adds the four in registers with round-to-nearest.
constexpr int kAccs = 4;

struct MapSet {                    // one TMA descriptor per bf16 plane of each operand
    CUtensorMap a[kMaxSplit];
    CUtensorMap w[kMaxSplit];
};

template <int BN>
struct Cfg {
    static constexpr uint32_t kWTile = BN * kK * 2;
    __host__ __device__ static constexpr uint32_t stage_bytes(int ns) { return (uint32_t)ns * (kATile + kWTile); }
    // as many stages as fit next to the barriers: 3 for ns <= 2, 2 for ns = 3 (BN = 128)
    __host__ __device__ static constexpr int stages(int ns) { return (int)((196u * 1024u) / stage_bytes(ns)) > 4 ? 4 : (int)((196u * 1024u) / stage_bytes(ns)); }
    __host__ __device__ static constexpr uint32_t smem_bytes(int ns) { return (uint32_t)stages(ns) * stage_bytes(ns) + 1024 + 256; }
};

struct Pipe {
    uint32_t base, full0, empty0, accum_bar;
    uint32_t tmem_acc;
    uint32_t stage_bytes;
    int n_stages, ns;
};

// barriers + TMEM allocation; every thread of the CTA calls this
template <int BN>
__device__ __forceinline__ Pipe pipe_setup(uint8_t *smem_raw, int ns)
{
    Pipe pp;
    pp.ns = ns;
    pp.n_stages = Cfg<BN>::stages(ns);
    pp.stage_bytes = Cfg<BN>::stage_bytes(ns);
    pp.base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *gen = smem_raw + (pp.base - smem_u32(smem_raw));
    const uint32_t ring = (uint32_t)pp.n_stages * pp.stage_bytes;
    const uint32_t bar_base = pp.base + ring;
    pp.full0 = bar_base;
    pp.empty0 = bar_base + 8 * 4;
    pp.accum_bar = bar_base + 16 * 4;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(gen + ring + 16 * 4 + 8);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < pp.n_stages; s++) {
            mbar_init(pp.full0 + 8 * s, 1);
            mbar_init(pp.empty0 + 8 * s, 1);
        }
        mbar_init(pp.accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kAccs * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    pp.tmem_acc = *tmem_slot;
    return pp;
}

// stage layout: A planes 0..ns-1, then W planes 0..ns-1
template <int BN>
__device__ __forceinline__ uint32_t a_tile(const Pipe &pp, int s, int plane) { return pp.base + (uint32_t)s * pp.stage_bytes + (uint32_t)plane * kATile; }
template <int BN>
__device__ __forceinline__ uint32_t w_tile(const Pipe &pp, int s, int plane)
{
    return pp.base + (uint32_t)s * pp.stage_bytes + (uint32_t)pp.ns * kATile + (uint32_t)plane * Cfg<BN>::kWTile;
}

// one elected lane: n_iter stages
template <int BN>
__device__ __forceinline__ void mma_issue(const Pipe &pp, int n_iter)
{
    constexpr uint32_t idesc = umma_idesc(kM, BN);
    // (a plane, w plane) pairs: ns = 1 -> 1 product, ns = 2 -> 3, ns = 3 -> 6
    const int pa[6] = {2, 0, 1, 1, 0, 0}, pw[6] = {0, 2, 1, 0, 1, 0};
    const int first = pp.ns == 3 ? 0 : (pp.ns == 2 ? 3 : 5);
    uint32_t used = 0;                                         // accumulators that already hold a value
    for (int it = 0; it < n_iter; it++) {
        const int s = it % pp.n_stages;
        mbar_wait(pp.full0 + 8 * s, (it / pp.n_stages) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint64_t da[kMaxSplit], dw[kMaxSplit];
        for (int q = 0; q < pp.ns; q++) {
            da[q] = umma_desc(a_tile<BN>(pp, s, q));
            dw[q] = umma_desc(w_tile<BN>(pp, s, q));
        }
#pragma unroll
        for (int k = 0; k < kK / 16; k++) {
            const uint64_t adv = (uint64_t)((k * 32) >> 4);
            for (int j = first; j < 6; j++) {
                const int acc = (j == 5 ? 0 : 2) + (k & 1);
                umma_bf16(pp.tmem_acc + (uint32_t)(acc * BN), da[pa[j]] + adv, dw[pw[j]] + adv, idesc, (used >> acc) & 1u);
                used |= 1u << acc;
            }
        }
        umma_commit(pp.empty0 + 8 * s);
    }
    umma_commit(pp.accum_bar);
}

// epilogue: sum of the tile's accumulators for 32 columns of this thread's row
template <int BN>
__device__ __forceinline__ void load_acc32(const Pipe &pp, int q, int cb, float (&v)[32])
{
    float t[32];
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    tmem_ld32(pp.tmem_acc + lane_base + (uint32_t)cb, v);
    tmem_ld32(pp.tmem_acc + lane_base + (uint32_t)(BN + cb), t);
#pragma unroll
    for (int j = 0; j < 32; j++) v[j] += t[j];
    if (pp.ns >= 2) {
        float u[32];
        tmem_ld32(pp.tmem_acc + lane_base + (uint32_t)(2 * BN + cb), t);
        tmem_ld32(pp.tmem_acc + lane_base + (uint32_t)(3 * BN + cb), u);
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] += t[j] + u[j];
    }
}

template <int BN>
__device__ __forceinline__ void pipe_teardown(const Pipe &pp)
{
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if ((threadIdx.x >> 5) == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(pp.tmem_acc), "n"(kAccs * BN) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// forward (mode 0) and dgrad (mode 1): tap-wise implicit GEMM over an NHWC activation tensor
// ---------------------------------------------------------------------------------------------
struct TapsParams {
    int mode;                 // 0 forward: A = x (stride-2 boxes), tiles over y;  1 dgrad: A = dy (stride-1 boxes), tiles over a parity class of dx
    int B, H, Cin, Cout;      // conv layer shape: x [B,H,H,Cin], y / dy [B,H/2,H/2,Cout]
    int bw, bh, bb;           // A box: pixels, rows, images (bw*bh*bb == 128) on the Ho = H/2 grid
    int M;                    // B * Ho * Ho tile rows (per parity class for dgrad)
    int ns;
    int n_slices;             // the tile's taps are split over n_slices CTAs (gridDim.z = classes * n_slices), each writing its
    size_t slice_stride;      // own fp32 copy of the output, slice_stride floats apart (reduced afterwards in a fixed order)
};

template <int BN>
__global__ void __launch_bounds__(kThreadsG, 1)
taps_tc_kernel(const __grid_constant__ MapSet maps, const __grid_constant__ TapsParams p, float *__restrict__ out)
{
    extern __shared__ uint8_t smem_raw[];
    const Pipe pp = pipe_setup<BN>(smem_raw, p.ns);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tile = blockIdx.x, n_tile = blockIdx.y;
    const int n_classes = p.mode == 0 ? 1 : 4;
    const int cls = blockIdx.z % n_classes, slice = blockIdx.z / n_classes;
    const int py = cls >> 1, px = cls & 1;                      // dgrad parity class (0 for forward)
    const int Ho = p.H / 2;
    out += (size_t)slice * p.slice_stride;
    const int m0 = m_tile * kM;
    const int img0 = m0 / (Ho * Ho);
    const int r0 = (m0 % (Ho * Ho)) / Ho;                       // first tile row on the Ho grid
    const int r_lo = (p.bb > 1) ? 0 : r0, r_hi = (p.bb > 1) ? Ho - 1 : r0 + p.bh - 1;

    // taps that reach at least one real pixel for this tile, with their A-box start coordinates
    int taps[25], n_taps = 0;
    if (p.mode == 0) {
        for (int ky = 0; ky < 5; ky++) {
            if (!((2 * r_hi + ky - 1 >= 0) && (2 * r_lo + ky - 1 < p.H))) continue;
            for (int kx = 0; kx < 5; kx++)
                if ((2 * (Ho - 1) + kx - 1 >= 0) && (kx - 1 < p.H)) taps[n_taps++] = ky * 5 + kx;
        }
    } else {
        for (int ky = (py + 1) & 1; ky < 5; ky += 2) {
            const int sy = (py + 1 - ky) / 2;                     // exact: py + 1 - ky is even
            if (r_hi + sy < 0 || r_lo + sy >= Ho) continue;
            for (int kx = (px + 1) & 1; kx < 5; kx += 2) {
                const int sx = (px + 1 - kx) / 2;
                if (Ho - 1 + sx < 0 || sx >= Ho) continue;
                taps[n_taps++] = ky * 5 + kx;
            }
        }
    }
    const int kdim = p.mode == 0 ? p.Cin : p.Cout;              // channels reduced per tap
    const int chunks = kdim / kK;
    // this CTA's share of the taps (deep layers at small batch have a handful of tiles: without the split 16 - 32 CTAs stream
    // the whole 25 * Cin reduction while the other SMs idle)
    const int per = (n_taps + p.n_slices - 1) / p.n_slices;
    const int t_begin = min(n_taps, slice * per), t_end = min(n_taps, t_begin + per);
    const int n_iter = (t_end - t_begin) * chunks;

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&maps.a[0]);
            tma_prefetch_desc(&maps.w[0]);
            int it = 0;
            for (int t = t_begin; t < t_end; t++) {
                const int ky = taps[t] / 5, kx = taps[t] % 5;
                int cx, cy;
                if (p.mode == 0) { cx = kx - 1; cy = 2 * r0 + ky - 1; }
                else { cx = (px + 1 - kx) / 2; cy = r0 + (py + 1 - ky) / 2; }
                for (int ch = 0; ch < chunks; ch++, it++) {
                    const int s = it % pp.n_stages;
                    if (it >= pp.n_stages) mbar_wait(pp.empty0 + 8 * s, ((it / pp.n_stages) - 1) & 1);
                    const uint32_t full = pp.full0 + 8 * s;
                    mbar_expect_tx(full, pp.stage_bytes);
                    const int c0 = ch * kK;
                    for (int q = 0; q < p.ns; q++) {
                        tma_load_4d(a_tile<BN>(pp, s, q), &maps.a[q], full, c0, cx, cy, img0);
                        if (p.mode == 0) tma_load_3d(w_tile<BN>(pp, s, q), &maps.w[q], full, taps[t] * p.Cin + c0, n_tile * BN, 0);
                        else tma_load_3d(w_tile<BN>(pp, s, q), &maps.w[q], full, c0, n_tile * BN, taps[t]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && n_iter > 0) mma_issue<BN>(pp, n_iter);
    } else {
        const int q = warp & 3;
        if (n_iter > 0) {
            mbar_wait(pp.accum_bar, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const int row = q * 32 + lane;
        const int m = m0 + row;
        const bool row_ok = m < p.M;
        const int mm = row_ok ? m : 0;
        const int ncol_total = p.mode == 0 ? p.Cout : p.Cin;
        float *orow;
        if (p.mode == 0) {
            orow = out + (size_t)mm * p.Cout + (size_t)n_tile * BN;
        } else {
            const int img = mm / (Ho * Ho), r = mm % (Ho * Ho);
            const int iy = 2 * (r / Ho) + py, ix = 2 * (r % Ho) + px;
            orow = out + (((size_t)img * p.H + iy) * p.H + ix) * p.Cin + (size_t)n_tile * BN;
        }
        const int ncols = min(BN, ncol_total - n_tile * BN);
#pragma unroll 1
        for (int cb = 0; cb < BN; cb += 32) {
            float v[32];
            if (n_iter > 0) {                      // CTA-uniform; tcgen05.ld is warp-collective
                load_acc32<BN>(pp, q, cb, v);
            } else {
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] = 0.0f;
            }
            if (!row_ok) continue;
            if (cb + 32 <= ncols && (ncol_total & 3) == 0) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4 *>(orow + cb + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 32; j++)
                    if (cb + j < ncols) orow[cb + j] = v[j];
            }
        }
    }
    pipe_teardown<BN>(pp);
}

// ---------------------------------------------------------------------------------------------
// wgrad
// ---------------------------------------------------------------------------------------------
struct WgradParams {
    int Cin, Cout;
    int n_chunks;             // K chunks of 64 pixels (padded pixel count / 64)
    int n_slices;             // pixel-range slices (gridDim.z = n_valid_taps * n_slices)
    int n_valid;
    int taps[25];             // valid taps
    int ns;
    size_t slice_stride;      // floats between slices of the output (25 * Cin * Cout), 0 when n_slices == 1
};

template <int BN>
__global__ void __launch_bounds__(kThreadsG, 1)
wgrad_tc_kernel(const __grid_constant__ MapSet maps, const __grid_constant__ WgradParams p, float *__restrict__ out)
{
    extern __shared__ uint8_t smem_raw[];
    const Pipe pp = pipe_setup<BN>(smem_raw, p.ns);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tile = blockIdx.x, n_tile = blockIdx.y;
    const int tap = p.taps[blockIdx.z % p.n_valid], slice = blockIdx.z / p.n_valid;
    const int per = (p.n_chunks + p.n_slices - 1) / p.n_slices;
    const int k_begin = slice * per, k_end = min(p.n_chunks, k_begin + per);
    const int n_iter = max(0, k_end - k_begin);

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&maps.a[0]);
            tma_prefetch_desc(&maps.w[0]);
            for (int it = 0; it < n_iter; it++) {
                const int s = it % pp.n_stages;
                if (it >= pp.n_stages) mbar_wait(pp.empty0 + 8 * s, ((it / pp.n_stages) - 1) & 1);
                const uint32_t full = pp.full0 + 8 * s;
                mbar_expect_tx(full, pp.stage_bytes);
                const int k0 = (k_begin + it) * kK;
                for (int q = 0; q < p.ns; q++) {
                    tma_load_3d(a_tile<BN>(pp, s, q), &maps.a[q], full, k0, m_tile * kM, tap);
                    tma_load_2d(w_tile<BN>(pp, s, q), &maps.w[q], full, k0, n_tile * BN);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && n_iter > 0) mma_issue<BN>(pp, n_iter);
    } else {
        const int q = warp & 3;
        if (n_iter > 0) {
            mbar_wait(pp.accum_bar, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const int ci = m_tile * kM + q * 32 + lane;
        float *orow = out + (size_t)slice * p.slice_stride + ((size_t)tap * p.Cin + ci) * p.Cout + (size_t)n_tile * BN;
#pragma unroll 1
        for (int cb = 0; cb < BN; cb += 32) {
            float v[32];
            if (n_iter > 0) {
                load_acc32<BN>(pp, q, cb, v);
            } else {
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] = 0.0f;
            }
            if (ci < p.Cin) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4 *>(orow + cb + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
        }
    }
    pipe_teardown<BN>(pp);
}

// ---------------------------------------------------------------------------------------------
// pre-passes: fp32 -> ns bf16 planes (plane q at element offset q * plane_elems)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_planes(float x, __nv_bfloat16 *dst, size_t idx, size_t plane, int ns)
{
    float r = x;
    for (int q = 0; q < ns; q++) {
        const __nv_bfloat16 h = __float2bfloat16_rn(r);
        dst[(size_t)q * plane + idx] = h;
        r -= __bfloat162float(h);
    }
}

__global__ void split_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ dst, size_t n, size_t plane, int ns)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) split_planes(x[i], dst, i, plane, ns);
}

// HWIO [25][Cin][Cout] fp32 -> K-major [Cout][25*Cin] planes (forward's B operand)
__global__ void pack_kmajor_kernel(const float *__restrict__ w, __nv_bfloat16 *__restrict__ dst, int Cin, int Cout,
                                   size_t plane, int ns)
{
    const size_t total = (size_t)25 * Cin * Cout;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int n = (int)(i / ((size_t)25 * Cin));
        const int k = (int)(i % ((size_t)25 * Cin));
        split_planes(w[(size_t)k * Cout + n], dst, i, plane, ns);
    }
}

// src fp32 [rows of C floats]; dst plane q: out[(z * C + c) * Ppad + p] = src[pix(p, z)][c]
//   TAPS = true : z = tap, pix(p, tap) = input pixel of output pixel p under that tap (or zero)
//   TAPS = false: z = 0,  pix(p) = p                        (plain transpose, for dy)
// 32 x 32 tiles through shared memory: coalesced reads along c, coalesced writes along p.
template <bool TAPS>
__global__ void __launch_bounds__(256)
transpose_split_kernel(const float *__restrict__ src, __nv_bfloat16 *__restrict__ dst, int B, int H, int C, long long P,
                       long long Ppad, size_t plane, int ns)
{
    __shared__ float s[32][33];
    const int tap = blockIdx.z;
    const int ky = tap / 5, kx = tap % 5;
    const int Ho = H / 2;
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
    for (int j = ty; j < 32; j += 8) {
        const long long pidx = p0 + j;
        float v = 0.0f;
        if (pidx < P && c0 + tx < C) {
            long long pix;
            if (TAPS) {
                const int img = (int)(pidx / (Ho * Ho));
                const int r = (int)(pidx % (Ho * Ho));
                const int sy = 2 * (r / Ho) - 1 + ky, sx = 2 * (r % Ho) - 1 + kx;
                pix = (sy >= 0 && sy < H && sx >= 0 && sx < H) ? ((long long)img * H + sy) * H + sx : -1;
            } else {
                pix = pidx;
            }
            if (pix >= 0) v = src[(size_t)pix * C + c0 + tx];
        }
        s[j][tx] = v;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j;
        const long long pidx = p0 + tx;
        if (c < C && pidx < Ppad)
            split_planes(s[tx][j], dst, ((size_t)tap * C + c) * (size_t)Ppad + (size_t)pidx, plane, ns);
    }
}

// zero the taps no pixel reaches (small layers), after the reduction / direct write of the valid ones
__global__ void zero_taps_kernel(float *dw, int Cin, int Cout, uint32_t valid_mask)
{
    const int tap = blockIdx.y;
    if ((valid_mask >> tap) & 1) return;
    const size_t n = (size_t)Cin * Cout;
    float *o = dw + (size_t)tap * n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) o[i] = 0.0f;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static size_t al(size_t v) { return (v + 1023) & ~(size_t)1023; }

static bool pow2(int v) { return v > 0 && !(v & (v - 1)); }

bool conv_tc_supported(int op, int B, int H, int Cin, int Cout)
{
    const int Ho = H / 2;
    if (!pow2(Ho) || Ho > 64 || B < 1) return false;
    if (tc_encode_fn() == nullptr) return false;
    if (op == 0) return (Cin % 64) == 0 && (Cout % 128) == 0;
    if (op == 1) return (Cout % 64) == 0;
    return (Cout % 64) == 0;
}

static long long padded_pixels(int B, int H) { return (((long long)B * (H / 2) * (H / 2)) + 63) / 64 * 64; }

static uint32_t valid_taps(int H, int *list, int *n)
{
    // tap (ky, kx) reaches a real pixel iff 0 <= 2*o - 1 + k < H for some o in [0, H/2)
    uint32_t mask = 0;
    *n = 0;
    for (int ky = 0; ky < 5; ky++)
        for (int kx = 0; kx < 5; kx++) {
            auto ok = [&](int k) { return 2 * (H / 2 - 1) - 1 + k >= 0 && k - 1 < H; };
            if (ok(ky) && ok(kx)) {
                mask |= 1u << (ky * 5 + kx);
                list[(*n)++] = ky * 5 + kx;
            }
        }
    return mask;
}

static int wgrad_tc_slices(int Cin, int Cout, int BN, int n_valid, int n_chunks)
{
    const long long tiles = (long long)((Cin + kM - 1) / kM) * (Cout / BN) * n_valid;
    long long s = (148 * 2 + tiles - 1) / tiles;
    if (s > n_chunks / 4) s = n_chunks / 4;
    if (s < 1) s = 1;
    if (s > 32) s = 32;
    return (int)s;
}

size_t conv_tc_workspace_bytes(int B, int H, int Cin, int Cout)
{
    const size_t ns = kMaxSplit;
    const size_t Ppad = (size_t)padded_pixels(B, H);
    const size_t nx = (size_t)B * H * H * Cin, ny = (size_t)B * (H / 2) * (H / 2) * Cout, nw = (size_t)25 * Cin * Cout;
    // forward: x planes + packed w planes;  dgrad: dy planes + w planes
    const size_t fwd = al(ns * nx * 2) + al(ns * nw * 2) + al((size_t)kMaxTapSlices * ny * 4);
    const size_t dg = al(ns * ny * 2) + al(ns * nw * 2) + al((size_t)kMaxTapSlices * nx * 4);
    // wgrad: x^T per tap planes, dy^T planes, slice partials
    const int BN = (Cout % 128) == 0 ? 128 : 64;
    int list[25], nv;
    valid_taps(H, list, &nv);
    const int slices = wgrad_tc_slices(Cin, Cout, BN, nv, (int)(Ppad / 64));
    const size_t wg = al(ns * 25 * Cin * Ppad * 2) + al(ns * Cout * Ppad * 2) + (slices > 1 ? al((size_t)slices * nw * 4) : 0);
    size_t m = fwd > dg ? fwd : dg;
    return m > wg ? m : wg;
}

template <typename K>
static int set_smem(K kernel, uint32_t bytes, uint32_t *done)
{
    if (*done >= bytes) return SPIRAL_OK;
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return spiral_cuda_error(e, "cudaFuncSetAttribute");
    *done = bytes;
    return SPIRAL_OK;
}

static int encode(CUtensorMap *map, void *ptr, int rank, const cuuint64_t *dims, const cuuint64_t *strides,
                  const cuuint32_t *box, const cuuint32_t *estr)
{
    const CUresult r = tc_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, ptr, dims, strides, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

static void launch_split(const float *x, void *dst, size_t n, int ns, cudaStream_t st)
{
    const size_t want = (n + 255) / 256;
    const int blocks = (int)(want < (size_t)(148 * 8) ? want : (size_t)(148 * 8));
    split_kernel<<<blocks, 256, 0, st>>>(x, (__nv_bfloat16 *)dst, n, n, ns);
}

// mode 0 forward / mode 1 dgrad through taps_tc_kernel; act = the activation-side operand planes
static int taps_slices(int mode, long long ctas)
{
    if (ctas >= 74) return 1;
    int s = (int)((148 + ctas - 1) / ctas);
    const int cap = mode == 0 ? kMaxTapSlices : 3;             // a dgrad parity class has 4 .. 9 taps
    return s > cap ? cap : s;
}

template <int BN>
static int taps_launch(int mode, char *act, size_t act_plane, char *wp, size_t w_plane, float *out, int B, int H, int Cin,
                       int Cout, int ns, cudaStream_t st, float *part)
{
    static uint32_t attr = 0;
    int rc = set_smem(taps_tc_kernel<BN>, Cfg<BN>::smem_bytes(ns), &attr);
    if (rc != SPIRAL_OK) return rc;
    const int Ho = H / 2;
    TapsParams p;
    p.mode = mode; p.B = B; p.H = H; p.Cin = Cin; p.Cout = Cout; p.M = B * Ho * Ho; p.ns = ns;
    if (Ho * Ho >= kM) { p.bw = Ho; p.bh = kM / Ho; p.bb = 1; }
    else { p.bw = Ho; p.bh = Ho; p.bb = kM / (Ho * Ho); }
    MapSet maps;
    int er = 0;
    for (int q = 0; q < ns; q++) {
        if (mode == 0) {
            // x [B,H,H,Cin], traversal stride 2 in W and H: boxDim = elements loaded * elementStride
            const cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)H, (cuuint64_t)H, (cuuint64_t)B};
            const cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)Cin * 2 * H, (cuuint64_t)Cin * 2 * H * H};
            const cuuint32_t box[4] = {(cuuint32_t)kK, (cuuint32_t)(p.bw * 2), (cuuint32_t)(p.bh * 2), (cuuint32_t)p.bb};
            const cuuint32_t estr[4] = {1, 2, 2, 1};
            er |= encode(&maps.a[q], act + q * act_plane, 4, dims, strides, box, estr);
            // packed K-major kernel [Cout][25*Cin]
            const cuuint64_t wd[3] = {(cuuint64_t)25 * Cin, (cuuint64_t)Cout, 1};
            const cuuint64_t wst[2] = {(cuuint64_t)25 * Cin * 2, (cuuint64_t)25 * Cin * 2 * Cout};
            const cuuint32_t wb[3] = {(cuuint32_t)kK, (cuuint32_t)BN, 1};
            const cuuint32_t one[3] = {1, 1, 1};
            er |= encode(&maps.w[q], wp + q * w_plane, 3, wd, wst, wb, one);
        } else {
            // dy [B,Ho,Ho,Cout], stride-1 boxes
            const cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)Ho, (cuuint64_t)Ho, (cuuint64_t)B};
            const cuuint64_t strides[3] = {(cuuint64_t)Cout * 2, (cuuint64_t)Cout * 2 * Ho, (cuuint64_t)Cout * 2 * Ho * Ho};
            const cuuint32_t box[4] = {(cuuint32_t)kK, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bb};
            const cuuint32_t estr[4] = {1, 1, 1, 1};
            er |= encode(&maps.a[q], act + q * act_plane, 4, dims, strides, box, estr);
            // HWIO kernel: for a fixed tap [ci][co] is K-major already
            const cuuint64_t wd[3] = {(cuuint64_t)Cout, (cuuint64_t)Cin, 25};
            const cuuint64_t wst[2] = {(cuuint64_t)Cout * 2, (cuuint64_t)Cout * 2 * Cin};
            const cuuint32_t wb[3] = {(cuuint32_t)kK, (cuuint32_t)BN, 1};
            const cuuint32_t one[3] = {1, 1, 1};
            er |= encode(&maps.w[q], wp + q * w_plane, 3, wd, wst, wb, one);
        }
    }
    for (int q = ns; q < kMaxSplit; q++) { maps.a[q] = maps.a[0]; maps.w[q] = maps.w[0]; }
    if (er) return spiral_set_error(SPIRAL_ECUDA, "cuTensorMapEncodeTiled failed (mode %d, CUresult %d)", mode, er);
    const int ncols = mode == 0 ? Cout : Cin;
    const int n_classes = mode == 0 ? 1 : 4;
    dim3 grid((p.M + kM - 1) / kM, (ncols + BN - 1) / BN, n_classes);
    const size_t n_out = mode == 0 ? (size_t)p.M * Cout : (size_t)B * H * H * Cin;
    p.n_slices = taps_slices(mode, (long long)grid.x * grid.y * grid.z);
    p.slice_stride = p.n_slices > 1 ? n_out : 0;
    grid.z = n_classes * p.n_slices;
    taps_tc_kernel<BN><<<grid, kThreadsG, Cfg<BN>::smem_bytes(ns), st>>>(maps, p, p.n_slices > 1 ? part : out);
    if (p.n_slices > 1) reduce_slices_kernel<<<148 * 2, 256, 0, st>>>(part, out, n_out, p.n_slices);
    return SPIRAL_OK;
}

// The operand planes of a kernel, for the forward (op 0: K-major [Cout][25 * Cin]) or the input gradient (op 1: HWIO as it is --
// [tap][ci][co] is K-major per tap).  The three shared-weight passes of a WGAN-GP step (and the convolutions of its double
// backward) read the same kernels: spiral_conv_pack + the *_packed entry points split each kernel once per step instead of
// once per call.
size_t conv_tc_packed_bytes(int Cin, int Cout) { return al((size_t)kMaxSplit * 25 * Cin * Cout * 2); }

void conv_tc_pack(const float *w, char *wp, int op, int Cin, int Cout, int ns, cudaStream_t st)
{
    const size_t nw = (size_t)25 * Cin * Cout;
    if (op == 1) launch_split(w, wp, nw, ns, st);
    else if (ns <= 2) conv_tc_pack_weights(w, wp, wp + nw * 2, Cin, Cout, st);      // the reward path's layout (disc_tc.cu): planes 0/1 = hi/lo
    else pack_kmajor_kernel<<<148 * 4, 256, 0, st>>>(w, (__nv_bfloat16 *)wp, Cin, Cout, nw, ns);
}

// workspace: [activation planes][kernel planes (unused when `packed` is given)][tap-slice partials]
int conv_tc_forward(const float *x, const float *w, float *y, int B, int H, int Cin, int Cout, int ns, char *ws,
                    cudaStream_t st, const char *packed)
{
    const size_t nx = (size_t)B * H * H * Cin, nw = (size_t)25 * Cin * Cout;
    char *xp = ws, *wp = ws + al((size_t)kMaxSplit * nx * 2);
    float *part = reinterpret_cast<float *>(wp + al((size_t)kMaxSplit * nw * 2));
    launch_split(x, xp, nx, ns, st);
    if (packed) wp = const_cast<char *>(packed);
    else conv_tc_pack(w, wp, 0, Cin, Cout, ns, st);
    if (ns <= 2) return conv_tc_layer(xp, xp + nx * 2, wp, wp + nw * 2, y, B, H, Cin, Cout, ns == 2 ? 1 : 0, st);
    return taps_launch<128>(0, xp, nx * 2, wp, nw * 2, y, B, H, Cin, Cout, ns, st, part);
}

int conv_tc_dgrad(const float *dy, const float *w, float *dx, int B, int H, int Cin, int Cout, int ns, char *ws,
                  cudaStream_t st, const char *packed)
{
    const size_t ny = (size_t)B * (H / 2) * (H / 2) * Cout, nw = (size_t)25 * Cin * Cout;
    char *dyp = ws, *wp = ws + al((size_t)kMaxSplit * ny * 2);
    float *part = reinterpret_cast<float *>(wp + al((size_t)kMaxSplit * nw * 2));
    launch_split(dy, dyp, ny, ns, st);
    if (packed) wp = const_cast<char *>(packed);
    else conv_tc_pack(w, wp, 1, Cin, Cout, ns, st);
    if (Cin >= 128) return taps_launch<128>(1, dyp, ny * 2, wp, nw * 2, dx, B, H, Cin, Cout, ns, st, part);
    return taps_launch<64>(1, dyp, ny * 2, wp, nw * 2, dx, B, H, Cin, Cout, ns, st, part);
}

template <int BN>
static int wgrad_launch(char *xt, size_t xt_plane, char *dyt, size_t dyt_plane, float *part, float *dw, int Cin, int Cout,
                        long long Ppad, const WgradParams &p, cudaStream_t st)
{
    static uint32_t attr = 0;
    int rc = set_smem(wgrad_tc_kernel<BN>, Cfg<BN>::smem_bytes(p.ns), &attr);
    if (rc != SPIRAL_OK) return rc;
    MapSet maps;
    int er = 0;
    for (int q = 0; q < p.ns; q++) {
        const cuuint64_t ad[3] = {(cuuint64_t)Ppad, (cuuint64_t)Cin, 25};
        const cuuint64_t ast[2] = {(cuuint64_t)Ppad * 2, (cuuint64_t)Ppad * 2 * Cin};
        const cuuint32_t ab[3] = {(cuuint32_t)kK, (cuuint32_t)kM, 1};
        const cuuint32_t one[3] = {1, 1, 1};
        er |= encode(&maps.a[q], xt + q * xt_plane, 3, ad, ast, ab, one);
        const cuuint64_t wd[2] = {(cuuint64_t)Ppad, (cuuint64_t)Cout};
        const cuuint64_t wst[1] = {(cuuint64_t)Ppad * 2};
        const cuuint32_t wb[2] = {(cuuint32_t)kK, (cuuint32_t)BN};
        er |= encode(&maps.w[q], dyt + q * dyt_plane, 2, wd, wst, wb, one);
    }
    for (int q = p.ns; q < kMaxSplit; q++) { maps.a[q] = maps.a[0]; maps.w[q] = maps.w[0]; }
    if (er) return spiral_set_error(SPIRAL_ECUDA, "cuTensorMapEncodeTiled failed for wgrad (CUresult %d)", er);
    dim3 grid((Cin + kM - 1) / kM, Cout / BN, p.n_valid * p.n_slices);
    wgrad_tc_kernel<BN><<<grid, kThreadsG, Cfg<BN>::smem_bytes(p.ns), st>>>(maps, p, p.n_slices > 1 ? part : dw);
    return SPIRAL_OK;
}

int conv_tc_wgrad(const float *x, const float *dy, float *dw, int B, int H, int Cin, int Cout, int ns, char *ws,
                  cudaStream_t st)
{
    const long long P = (long long)B * (H / 2) * (H / 2), Ppad = padded_pixels(B, H);
    const size_t nw = (size_t)25 * Cin * Cout;
    const int BN = (Cout % 128) == 0 ? 128 : 64;
    WgradParams p;
    p.Cin = Cin; p.Cout = Cout; p.ns = ns;
    p.n_chunks = (int)(Ppad / 64);
    const uint32_t vmask = valid_taps(H, p.taps, &p.n_valid);
    p.n_slices = wgrad_tc_slices(Cin, Cout, BN, p.n_valid, p.n_chunks);
    p.slice_stride = p.n_slices > 1 ? nw : 0;
    const size_t xt_plane = (size_t)25 * Cin * Ppad, dyt_plane = (size_t)Cout * Ppad;       // elements
    char *xt = ws, *dyt = xt + al((size_t)kMaxSplit * xt_plane * 2);
    float *part = reinterpret_cast<float *>(dyt + al((size_t)kMaxSplit * dyt_plane * 2));
    {
        dim3 g((unsigned)(Ppad / 32), (Cin + 31) / 32, 25);
        transpose_split_kernel<true><<<g, 256, 0, st>>>(x, (__nv_bfloat16 *)xt, B, H, Cin, P, Ppad, xt_plane, ns);
        dim3 g2((unsigned)(Ppad / 32), (Cout + 31) / 32, 1);
        transpose_split_kernel<false><<<g2, 256, 0, st>>>(dy, (__nv_bfloat16 *)dyt, B, H, Cout, P, Ppad, dyt_plane, ns);
    }
    int rc = BN == 128 ? wgrad_launch<128>(xt, xt_plane * 2, dyt, dyt_plane * 2, part, dw, Cin, Cout, Ppad, p, st)
                       : wgrad_launch<64>(xt, xt_plane * 2, dyt, dyt_plane * 2, part, dw, Cin, Cout, Ppad, p, st);
    if (rc != SPIRAL_OK) return rc;
    if (p.n_slices > 1) reduce_slices_kernel<<<148 * 2, 256, 0, st>>>(part, dw, nw, p.n_slices);
    if (p.n_valid < 25) {
        dim3 g(32, 25);
        zero_taps_kernel<<<g, 256, 0, st>>>(dw, Cin, Cout, vmask);
    }
    return SPIRAL_OK;
}

}  // namespace spiral
