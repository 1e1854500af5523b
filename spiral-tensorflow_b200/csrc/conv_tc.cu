// conv_tc.cu -- K2b on the 5th-generation tensor cores: the 5x5 / stride-2 / 'SAME' convolution's
// input gradient (dgrad) and kernel gradient (wgrad) as tcgen05 / TMEM implicit GEMMs, TMA-fed, with the
// same bf16x3 split (hi*hi + hi*lo + lo*hi into an fp32 TMEM accumulator) as the forward kernel in
// disc_tc.cu, so gradients keep fp32-equivalent accuracy.
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

template <int BN>
struct Cfg {
    static constexpr uint32_t kWTile = BN * kK * 2;
    static constexpr uint32_t kStage = 2 * kATile + 2 * kWTile;   // A_hi, A_lo, W_hi, W_lo
    static constexpr uint32_t kSmem = kStagesG * kStage + 1024 + 256;
};

struct Pipe {
    uint32_t base, full0, empty0, accum_bar;
    uint32_t tmem_acc;
};

// barriers + TMEM allocation; every thread of the CTA calls this
template <int BN>
__device__ __forceinline__ Pipe pipe_setup(uint8_t *smem_raw)
{
    Pipe pp;
    pp.base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *gen = smem_raw + (pp.base - smem_u32(smem_raw));
    const uint32_t bar_base = pp.base + kStagesG * Cfg<BN>::kStage;
    pp.full0 = bar_base;
    pp.empty0 = bar_base + 8 * kStagesG;
    pp.accum_bar = bar_base + 16 * kStagesG;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(gen + kStagesG * Cfg<BN>::kStage + 16 * kStagesG + 8);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStagesG; s++) {
            mbar_init(pp.full0 + 8 * s, 1);
            mbar_init(pp.empty0 + 8 * s, 1);
        }
        mbar_init(pp.accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    pp.tmem_acc = *tmem_slot;
    return pp;
}

// one elected lane: n_iter stages of four (or one) UMMAs each
template <int BN>
__device__ __forceinline__ void mma_issue(const Pipe &pp, int n_iter, int x3)
{
    constexpr uint32_t idesc = umma_idesc(kM, BN);
    for (int it = 0; it < n_iter; it++) {
        const int s = it % kStagesG;
        mbar_wait(pp.full0 + 8 * s, (it / kStagesG) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t stage = pp.base + s * Cfg<BN>::kStage;
        const uint64_t a_hi = umma_desc(stage), a_lo = umma_desc(stage + kATile);
        const uint64_t w_hi = umma_desc(stage + 2 * kATile), w_lo = umma_desc(stage + 2 * kATile + Cfg<BN>::kWTile);
#pragma unroll
        for (int k = 0; k < kK / 16; k++) {
            const uint64_t adv = (uint64_t)((k * 32) >> 4);
            if (x3) {
                umma_bf16(pp.tmem_acc, a_lo + adv, w_hi + adv, idesc, (it | k) != 0);
                umma_bf16(pp.tmem_acc, a_hi + adv, w_lo + adv, idesc, 1);
                umma_bf16(pp.tmem_acc, a_hi + adv, w_hi + adv, idesc, 1);
            } else {
                umma_bf16(pp.tmem_acc, a_hi + adv, w_hi + adv, idesc, (it | k) != 0);
            }
        }
        umma_commit(pp.empty0 + 8 * s);
    }
    umma_commit(pp.accum_bar);
}

template <int BN>
__device__ __forceinline__ void pipe_teardown(const Pipe &pp)
{
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if ((threadIdx.x >> 5) == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(pp.tmem_acc), "n"(BN) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// dgrad
// ---------------------------------------------------------------------------------------------
struct DgradParams {
    int B, H, Cin, Cout;      // conv layer shape; dy is [B, H/2, H/2, Cout], dx is [B, H, H, Cin]
    int bw, bh, bb;           // A box: pixels, rows, images (bw*bh*bb == 128) on the Ha = H/2 grid
    int M;                    // B * Ha * Ha rows per parity class
    int x3;
};

template <int BN>
__global__ void __launch_bounds__(kThreadsG, 1)
dgrad_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                const __grid_constant__ DgradParams p, float *__restrict__ dx)
{
    extern __shared__ uint8_t smem_raw[];
    const Pipe pp = pipe_setup<BN>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tile = blockIdx.x, n_tile = blockIdx.y;
    const int py = blockIdx.z >> 1, px = blockIdx.z & 1;
    const int Ha = p.H / 2;
    const int m0 = m_tile * kM;
    const int img0 = m0 / (Ha * Ha);
    const int a0 = (m0 % (Ha * Ha)) / Ha;
    const int a_lo = (p.bb > 1) ? 0 : a0, a_hi = (p.bb > 1) ? Ha - 1 : a0 + p.bh - 1;

    // taps of this parity class that reach at least one real dy pixel for this tile
    int taps[9], n_taps = 0;
    for (int ky = (py + 1) & 1; ky < 5; ky += 2) {
        const int sy = (py + 1 - ky) / 2;                     // exact: py + 1 - ky is even
        if (a_hi + sy < 0 || a_lo + sy >= Ha) continue;
        for (int kx = (px + 1) & 1; kx < 5; kx += 2) {
            const int sx = (px + 1 - kx) / 2;
            if (Ha - 1 + sx < 0 || sx >= Ha) continue;
            taps[n_taps++] = ky * 5 + kx;
        }
    }
    const int chunks = p.Cout / kK;
    const int n_iter = n_taps * chunks;

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&map_a_hi);
            tma_prefetch_desc(&map_w_hi);
            int it = 0;
            for (int t = 0; t < n_taps; t++) {
                const int ky = taps[t] / 5, kx = taps[t] % 5;
                const int sy = (py + 1 - ky) / 2, sx = (px + 1 - kx) / 2;
                for (int ch = 0; ch < chunks; ch++, it++) {
                    const int s = it % kStagesG;
                    if (it >= kStagesG) mbar_wait(pp.empty0 + 8 * s, ((it / kStagesG) - 1) & 1);
                    const uint32_t full = pp.full0 + 8 * s;
                    const uint32_t stage = pp.base + s * Cfg<BN>::kStage;
                    mbar_expect_tx(full, p.x3 ? Cfg<BN>::kStage : kATile + Cfg<BN>::kWTile);
                    const int c0 = ch * kK;
                    tma_load_4d(stage, &map_a_hi, full, c0, sx, a0 + sy, img0);
                    tma_load_3d(stage + 2 * kATile, &map_w_hi, full, c0, n_tile * BN, taps[t]);
                    if (p.x3) {
                        tma_load_4d(stage + kATile, &map_a_lo, full, c0, sx, a0 + sy, img0);
                        tma_load_3d(stage + 2 * kATile + Cfg<BN>::kWTile, &map_w_lo, full, c0, n_tile * BN, taps[t]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && n_iter > 0) mma_issue<BN>(pp, n_iter, p.x3);
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
        const int img = mm / (Ha * Ha), r = mm % (Ha * Ha);
        const int iy = 2 * (r / Ha) + py, ix = 2 * (r % Ha) + px;
        float *orow = dx + (((size_t)img * p.H + iy) * p.H + ix) * p.Cin + (size_t)n_tile * BN;
        const int ncols = min(BN, p.Cin - n_tile * BN);
#pragma unroll 1
        for (int cb = 0; cb < BN; cb += 32) {
            float v[32];
            if (n_iter > 0) {                      // CTA-uniform; tcgen05.ld is warp-collective
                tmem_ld32(pp.tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)cb, v);
            } else {
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] = 0.0f;
            }
            if (!row_ok) continue;
            if (cb + 32 <= ncols && (p.Cin & 3) == 0) {
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
    int x3;
    size_t slice_stride;      // floats between slices of the output (25 * Cin * Cout), 0 when n_slices == 1
};

template <int BN>
__global__ void __launch_bounds__(kThreadsG, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                const __grid_constant__ WgradParams p, float *__restrict__ out)
{
    extern __shared__ uint8_t smem_raw[];
    const Pipe pp = pipe_setup<BN>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tile = blockIdx.x, n_tile = blockIdx.y;
    const int tap = p.taps[blockIdx.z % p.n_valid], slice = blockIdx.z / p.n_valid;
    const int per = (p.n_chunks + p.n_slices - 1) / p.n_slices;
    const int k_begin = slice * per, k_end = min(p.n_chunks, k_begin + per);
    const int n_iter = max(0, k_end - k_begin);

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&map_a_hi);
            tma_prefetch_desc(&map_w_hi);
            for (int it = 0; it < n_iter; it++) {
                const int s = it % kStagesG;
                if (it >= kStagesG) mbar_wait(pp.empty0 + 8 * s, ((it / kStagesG) - 1) & 1);
                const uint32_t full = pp.full0 + 8 * s;
                const uint32_t stage = pp.base + s * Cfg<BN>::kStage;
                mbar_expect_tx(full, p.x3 ? Cfg<BN>::kStage : kATile + Cfg<BN>::kWTile);
                const int k0 = (k_begin + it) * kK;
                tma_load_3d(stage, &map_a_hi, full, k0, m_tile * kM, tap);
                tma_load_2d(stage + 2 * kATile, &map_w_hi, full, k0, n_tile * BN);
                if (p.x3) {
                    tma_load_3d(stage + kATile, &map_a_lo, full, k0, m_tile * kM, tap);
                    tma_load_2d(stage + 2 * kATile + Cfg<BN>::kWTile, &map_w_lo, full, k0, n_tile * BN);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && n_iter > 0) mma_issue<BN>(pp, n_iter, p.x3);
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
                tmem_ld32(pp.tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)cb, v);
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
// pre-passes
// ---------------------------------------------------------------------------------------------
__global__ void split_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ lo,
                             size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) split_bf16(x[i], hi[i], lo[i]);
}

// src fp32 [rows of C floats]; out[(z * C + c) * Ppad + p] = src[pix(p, z)][c] as bf16 hi / lo.
//   TAPS = true : z = tap, pix(p, tap) = input pixel of output pixel p under that tap (or zero)
//   TAPS = false: z = 0,  pix(p) = p                        (plain transpose, for dy)
// 32 x 32 tiles through shared memory: coalesced reads along c, coalesced writes along p.
template <bool TAPS>
__global__ void __launch_bounds__(256)
transpose_split_kernel(const float *__restrict__ src, __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ lo,
                       int B, int H, int C, long long P, long long Ppad)
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
        if (c < C && pidx < Ppad) {
            __nv_bfloat16 h, l;
            split_bf16(s[tx][j], h, l);
            const size_t o = ((size_t)tap * C + c) * (size_t)Ppad + (size_t)pidx;
            hi[o] = h;
            lo[o] = l;
        }
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
    if (op == 0) return (Cin % 64) == 0 && (Cout % 128) == 0 && Ho <= 128;
    if (op == 1) return (Cout % 64) == 0 && Ho <= 128;
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
    const size_t Ppad = (size_t)padded_pixels(B, H);
    const size_t nx = (size_t)B * H * H * Cin, ny = (size_t)B * (H / 2) * (H / 2) * Cout, nw = (size_t)25 * Cin * Cout;
    // forward: x hi/lo + packed w hi/lo;  dgrad: dy hi/lo + w hi/lo
    const size_t fwd = 2 * al(nx * 2) + 2 * al(nw * 2);
    const size_t dg = 2 * al(ny * 2) + 2 * al(nw * 2);
    // wgrad: x^T per tap hi/lo, dy^T hi/lo, slice partials
    const int BN = (Cout % 128) == 0 ? 128 : 64;
    int list[25], nv;
    valid_taps(H, list, &nv);
    const int slices = wgrad_tc_slices(Cin, Cout, BN, nv, (int)(Ppad / 64));
    const size_t wg = 2 * al((size_t)25 * Cin * Ppad * 2) + 2 * al((size_t)Cout * Ppad * 2) +
                      (slices > 1 ? al((size_t)slices * nw * 4) : 0);
    size_t m = fwd > dg ? fwd : dg;
    return m > wg ? m : wg;
}

template <typename K>
static int set_smem(K kernel, uint32_t bytes, bool *done)
{
    if (*done) return SPIRAL_OK;
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return spiral_cuda_error(e, "cudaFuncSetAttribute");
    *done = true;
    return SPIRAL_OK;
}

static int encode(CUtensorMap *map, void *ptr, int rank, const cuuint64_t *dims, const cuuint64_t *strides,
                  const cuuint32_t *box)
{
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = tc_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, ptr, dims, strides, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

static void launch_split(const float *x, void *hi, void *lo, size_t n, cudaStream_t st)
{
    const size_t want = (n + 255) / 256;
    const int blocks = (int)(want < (size_t)(148 * 8) ? want : (size_t)(148 * 8));
    split_kernel<<<blocks, 256, 0, st>>>(x, (__nv_bfloat16 *)hi, (__nv_bfloat16 *)lo, n);
}

int conv_tc_forward(const float *x, const float *w, float *y, int B, int H, int Cin, int Cout, bool x3, char *ws,
                    cudaStream_t st)
{
    const size_t nx = (size_t)B * H * H * Cin, nw = (size_t)25 * Cin * Cout;
    char *xh = ws, *xl = xh + al(nx * 2), *wh = xl + al(nx * 2), *wl = wh + al(nw * 2);
    launch_split(x, xh, xl, nx, st);
    conv_tc_pack_weights(w, wh, wl, Cin, Cout, st);
    return conv_tc_layer(xh, xl, wh, wl, y, B, H, Cin, Cout, x3 ? 1 : 0, st);
}

template <int BN>
static int dgrad_launch(char *dyh, char *dyl, char *wh, char *wl, float *dx, int B, int H, int Cin, int Cout, bool x3,
                        cudaStream_t st)
{
    static bool attr = false;
    int rc = set_smem(dgrad_tc_kernel<BN>, Cfg<BN>::kSmem, &attr);
    if (rc != SPIRAL_OK) return rc;
    const int Ha = H / 2;
    DgradParams p;
    p.B = B; p.H = H; p.Cin = Cin; p.Cout = Cout; p.M = B * Ha * Ha; p.x3 = x3 ? 1 : 0;
    if (Ha * Ha >= kM) { p.bw = Ha; p.bh = kM / Ha; p.bb = 1; }
    else { p.bw = Ha; p.bh = Ha; p.bb = kM / (Ha * Ha); }
    CUtensorMap ma_hi, ma_lo, mw_hi, mw_lo;
    {
        const cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)Ha, (cuuint64_t)Ha, (cuuint64_t)B};
        const cuuint64_t strides[3] = {(cuuint64_t)Cout * 2, (cuuint64_t)Cout * 2 * Ha, (cuuint64_t)Cout * 2 * Ha * Ha};
        const cuuint32_t box[4] = {(cuuint32_t)kK, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bb};
        int er = encode(&ma_hi, dyh, 4, dims, strides, box);
        er |= encode(&ma_lo, dyl, 4, dims, strides, box);
        const cuuint64_t wd[3] = {(cuuint64_t)Cout, (cuuint64_t)Cin, 25};
        const cuuint64_t wst[2] = {(cuuint64_t)Cout * 2, (cuuint64_t)Cout * 2 * Cin};
        const cuuint32_t wb[3] = {(cuuint32_t)kK, (cuuint32_t)BN, 1};
        er |= encode(&mw_hi, wh, 3, wd, wst, wb);
        er |= encode(&mw_lo, wl, 3, wd, wst, wb);
        if (er) return spiral_set_error(SPIRAL_ECUDA, "cuTensorMapEncodeTiled failed for dgrad (CUresult %d)", er);
    }
    dim3 grid((p.M + kM - 1) / kM, (Cin + BN - 1) / BN, 4);
    dgrad_tc_kernel<BN><<<grid, kThreadsG, Cfg<BN>::kSmem, st>>>(ma_hi, ma_lo, mw_hi, mw_lo, p, dx);
    return SPIRAL_OK;
}

int conv_tc_dgrad(const float *dy, const float *w, float *dx, int B, int H, int Cin, int Cout, bool x3, char *ws,
                  cudaStream_t st)
{
    const size_t ny = (size_t)B * (H / 2) * (H / 2) * Cout, nw = (size_t)25 * Cin * Cout;
    char *dyh = ws, *dyl = dyh + al(ny * 2), *wh = dyl + al(ny * 2), *wl = wh + al(nw * 2);
    launch_split(dy, dyh, dyl, ny, st);
    launch_split(w, wh, wl, nw, st);                       // HWIO: [tap][ci][co] is already K-major per tap
    if (Cin >= 128) return dgrad_launch<128>(dyh, dyl, wh, wl, dx, B, H, Cin, Cout, x3, st);
    return dgrad_launch<64>(dyh, dyl, wh, wl, dx, B, H, Cin, Cout, x3, st);
}

template <int BN>
static int wgrad_launch(char *xth, char *xtl, char *dyth, char *dytl, float *part, float *dw, int Cin, int Cout,
                        long long Ppad, const WgradParams &p, cudaStream_t st)
{
    static bool attr = false;
    int rc = set_smem(wgrad_tc_kernel<BN>, Cfg<BN>::kSmem, &attr);
    if (rc != SPIRAL_OK) return rc;
    CUtensorMap ma_hi, ma_lo, mw_hi, mw_lo;
    {
        const cuuint64_t ad[3] = {(cuuint64_t)Ppad, (cuuint64_t)Cin, 25};
        const cuuint64_t ast[2] = {(cuuint64_t)Ppad * 2, (cuuint64_t)Ppad * 2 * Cin};
        const cuuint32_t ab[3] = {(cuuint32_t)kK, (cuuint32_t)kM, 1};
        int er = encode(&ma_hi, xth, 3, ad, ast, ab);
        er |= encode(&ma_lo, xtl, 3, ad, ast, ab);
        const cuuint64_t wd[2] = {(cuuint64_t)Ppad, (cuuint64_t)Cout};
        const cuuint64_t wst[1] = {(cuuint64_t)Ppad * 2};
        const cuuint32_t wb[2] = {(cuuint32_t)kK, (cuuint32_t)BN};
        er |= encode(&mw_hi, dyth, 2, wd, wst, wb);
        er |= encode(&mw_lo, dytl, 2, wd, wst, wb);
        if (er) return spiral_set_error(SPIRAL_ECUDA, "cuTensorMapEncodeTiled failed for wgrad (CUresult %d)", er);
    }
    dim3 grid((Cin + kM - 1) / kM, Cout / BN, p.n_valid * p.n_slices);
    wgrad_tc_kernel<BN><<<grid, kThreadsG, Cfg<BN>::kSmem, st>>>(ma_hi, ma_lo, mw_hi, mw_lo, p, p.n_slices > 1 ? part : dw);
    return SPIRAL_OK;
}

int conv_tc_wgrad(const float *x, const float *dy, float *dw, int B, int H, int Cin, int Cout, bool x3, char *ws,
                  cudaStream_t st)
{
    const long long P = (long long)B * (H / 2) * (H / 2), Ppad = padded_pixels(B, H);
    const size_t nw = (size_t)25 * Cin * Cout;
    const int BN = (Cout % 128) == 0 ? 128 : 64;
    WgradParams p;
    p.Cin = Cin; p.Cout = Cout; p.x3 = x3 ? 1 : 0;
    p.n_chunks = (int)(Ppad / 64);
    const uint32_t vmask = valid_taps(H, p.taps, &p.n_valid);
    p.n_slices = wgrad_tc_slices(Cin, Cout, BN, p.n_valid, p.n_chunks);
    p.slice_stride = p.n_slices > 1 ? nw : 0;
    char *xth = ws, *xtl = xth + al((size_t)25 * Cin * Ppad * 2);
    char *dyth = xtl + al((size_t)25 * Cin * Ppad * 2), *dytl = dyth + al((size_t)Cout * Ppad * 2);
    float *part = reinterpret_cast<float *>(dytl + al((size_t)Cout * Ppad * 2));
    {
        dim3 g((unsigned)(Ppad / 32), (Cin + 31) / 32, 25);
        transpose_split_kernel<true><<<g, 256, 0, st>>>(x, (__nv_bfloat16 *)xth, (__nv_bfloat16 *)xtl, B, H, Cin, P, Ppad);
        dim3 g2((unsigned)(Ppad / 32), (Cout + 31) / 32, 1);
        transpose_split_kernel<false><<<g2, 256, 0, st>>>(dy, (__nv_bfloat16 *)dyth, (__nv_bfloat16 *)dytl, B, H, Cout, P, Ppad);
    }
    int rc = BN == 128 ? wgrad_launch<128>(xth, xtl, dyth, dytl, part, dw, Cin, Cout, Ppad, p, st)
                       : wgrad_launch<64>(xth, xtl, dyth, dytl, part, dw, Cin, Cout, Ppad, p, st);
    if (rc != SPIRAL_OK) return rc;
    if (p.n_slices > 1) reduce_slices_kernel<<<148 * 2, 256, 0, st>>>(part, dw, nw, p.n_slices);
    if (p.n_valid < 25) {
        dim3 g(32, 25);
        zero_taps_kernel<<<g, 256, 0, st>>>(dw, Cin, Cout, vmask);
    }
    return SPIRAL_OK;
}

}  // namespace spiral
