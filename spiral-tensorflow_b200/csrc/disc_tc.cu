// disc_tc.cu -- K2 on the 5th-generation tensor cores: tcgen05.mma + TMEM accumulators, TMA-fed
// implicit GEMM for the discriminator's 5x5 / stride-2 / 'SAME' convolutions (layers >= 1).
//
//   D[M, N] = sum_k A[m, k] * W[n, k]      M = B*Ho*Wo output pixels, N = Cout, K = 25 * Cin
//
// * A is never materialised: for kernel tap (ky, kx) and channel chunk c0 the 128 x 64 operand tile
//   is ONE 4-D TMA box over the NHWC activation tensor, {64 channels, Wo pixels with element
//   stride 2, rows with element stride 2, images}, whose start coordinate may be -1 or run past
//   the edge: TMA zero-fills, which is exactly TF 'SAME' padding (1 before / 2 after).
// * W is packed once per weight update as a K-major [Cout][25*Cin] matrix (2-D TMA).
// * fp32-equivalent arithmetic from bf16 tensor cores: every fp32 value is split as
//   x = hi + lo (hi = bf16(x), lo = bf16(x - hi)) and each K step issues three MMAs
//   (hi*hi + hi*lo + lo*hi) into the same fp32 TMEM accumulator (dropped lo*lo term ~2^-18).
//   precision 2 issues only hi*hi.
// * warp roles per CTA (one 128 x BN output tile): warp 0 = TMA producer, warp 1 = MMA issuer
//   (one elected lane) + TMEM allocator, warps 2..5 = epilogue (tcgen05.ld -> + bias -> fp32 store,
//   per-channel sum / sum-of-squares partials for the training-mode batch norm that follows).
//   smem ring of kStages {A_hi, A_lo, W_hi, W_lo} tiles in the 128-byte-swizzled K-major UMMA
//   layout, full/empty mbarriers, tcgen05.commit releases a stage back to the producer.
// * kernel taps that fall entirely into the padding for a tile are skipped (layer 5 uses 4 of 25).
#include "disc.cuh"
#include "tc_ptx.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

namespace spiral {

constexpr int kBM = 128;          // tile rows (output pixels)  == UMMA M
constexpr int kBN = 128;          // tile cols (output channels) == UMMA N
constexpr int kBK = 64;           // K per stage: 64 bf16 = one 128-byte swizzle row
constexpr int kStages = 3;
constexpr int kTcThreads = 192;
constexpr uint32_t kTileBytes = kBM * kBK * 2;                 // 16 KB (A and W tiles have the same shape)
constexpr uint32_t kStageBytes = 4 * kTileBytes;               // A_hi, A_lo, W_hi, W_lo
constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers: ring, 2 x accumulator full/empty, TMEM slot*/;

struct ConvTcParams {
    int B, H, Cin, Cout;      // input NHWC [B,H,H,Cin]; output [B,H/2,H/2,Cout]
    int bw, bh, bb;           // A box: Wo pixels, rows, images (bw*bh*bb == 128)
    int M;                    // B * Ho * Wo
    int x3;                   // 1: hi*hi + hi*lo + lo*hi ; 0: hi*hi only
    int mt, nt;               // output tiles per work item along M / N: (1,1), (2,1) or (1,2)
    int pos;                  // 1: position-pure tiles -- 128 images at ONE output pixel (small maps, B % 128 == 0)
};

// Where an m-tile sits.  Row-major tiles (pos == 0) are 128 consecutive rows of [B, Ho, Wo]: whole rows of one image
// or whole small images.  On the small maps of the deep layers (Ho <= 8) such a tile mixes every output position, so
// no kernel tap is padding for ALL of its rows and the SAME-padding zeros are multiplied like data (layer 4: 51 % of
// the MMAs).  Position-pure tiles take 128 images at one output pixel -- the TMA box is {64 ch, 1 px, 1 row, 128
// images} -- so the tap mask is exact per tile and the padding taps are neither loaded nor multiplied.
struct TileOrigin { int img0, oy0, ox0, posi; };
__device__ __forceinline__ TileOrigin tile_origin(const ConvTcParams &p, int Ho, int m_tile)
{
    TileOrigin t;
    if (p.pos) {
        const int hw = Ho * Ho;
        t.posi = m_tile % hw;                       // positions interleaved: every CTA sees the same mix of tap counts
        t.img0 = (m_tile / hw) * kBM;
        t.oy0 = t.posi / Ho;
        t.ox0 = t.posi % Ho;
    } else {
        const int m0 = m_tile * kBM;
        t.img0 = m0 / (Ho * Ho);
        t.oy0 = (m0 % (Ho * Ho)) / Ho;
        t.ox0 = 0;
        t.posi = 0;
    }
    return t;
}

// Persistent: gridDim.x CTAs (one per SM) walk the (m_tile, n_tile) space; the TMEM accumulator is double
// buffered (2 x kBN columns), so while the four epilogue warps drain tile i (tcgen05.ld, + bias, fp32 stores,
// BN partials) the MMA warp is already accumulating tile i+1 and the TMA warp is a few stages further --
// the tensor pipe no longer idles through every tile's epilogue and prologue (round-1 ncu: 53-78 % busy).
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void epilogue_bar_sync()       // the 4 epilogue warps only (named barrier 1)
{
    asm volatile("bar.sync 1, 128;" ::: "memory");
}

__device__ __forceinline__ uint32_t conv_tap_mask(const ConvTcParams &p, int Ho, int oy0, int ox0)
{
    // taps (ky, kx) that touch at least one real input pixel for this tile
    uint32_t tap_mask = 0;
    if (p.pos) {
        for (int ky = 0; ky < 5; ky++) {
            const int iy = 2 * oy0 + ky - 1;
            for (int kx = 0; kx < 5; kx++) {
                const int ix = 2 * ox0 + kx - 1;
                if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.H) tap_mask |= 1u << (ky * 5 + kx);
            }
        }
        return tap_mask;
    }
    const int oy_lo = oy0, oy_hi = (p.bb > 1) ? Ho - 1 : oy0 + p.bh - 1;
    for (int ky = 0; ky < 5; ky++) {
        const bool y_ok = (2 * oy_hi + ky - 1 >= 0) && (2 * oy_lo + ky - 1 < p.H);
        for (int kx = 0; kx < 5; kx++) {
            const bool x_ok = (2 * (Ho - 1) + kx - 1 >= 0) && (kx - 1 < p.H);
            if (y_ok && x_ok) tap_mask |= 1u << (ky * 5 + kx);
        }
    }
    return tap_mask;
}

// Work item = mt x nt output tiles that share operands: two n-tiles reuse one A tile (Cout >= 256), or two
// m-tiles reuse one W tile (Cout = 128).  With the bf16x3 split a single 128x128 tile needs 64 KB of operands
// per 12 MMAs (0.42 us of tensor time) -- about 155 GB/s per SM, which three 64 KB stages in flight cannot
// sustain against ~1 us of L2 latency (ncu, round 1: tensor pipe 64-83 % busy).  A pair needs 96 KB per 24
// MMAs: 25 % fewer bytes per flop and twice the tensor time per stage to cover the same latency.
template <int mt, int nt>
__global__ void __launch_bounds__(kTcThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
               const __grid_constant__ ConvTcParams p, const float *__restrict__ bias,
               float *__restrict__ out, float *__restrict__ bn_partial)
{
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128-byte swizzle atoms
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *gen_base = smem_raw + (base - smem_u32(smem_raw));
    constexpr int n_stages = (mt + nt == 2) ? 3 : 2;
    constexpr uint32_t stage_bytes = (uint32_t)(mt + nt) * 2u * kTileBytes;   // {A_hi, A_lo} x mt, {W_hi, W_lo} x nt
    constexpr uint32_t kRing = kStages * kStageBytes;                          // 192 KB either way
    const uint32_t bar_base = base + kRing;
    uint64_t *bar_gen = reinterpret_cast<uint64_t *>(gen_base + kRing);
    const uint32_t full0 = bar_base, empty0 = bar_base + 8 * kStages;
    const uint32_t acc_full0 = bar_base + 16 * kStages, acc_empty0 = acc_full0 + 16;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_gen + 2 * kStages + 4);
    __shared__ float s_red[4][2][kBN];      // per-epilogue-warp column sums
    __shared__ float s_tr[4][32][33];       // transpose staging

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Ho = p.H / 2;
    const int m_tiles = (p.M + kBM - 1) / kBM, n_tiles = p.Cout / kBN;
    const int m_groups = (m_tiles + mt - 1) / mt, n_groups = n_tiles / nt;
    const int n_work = m_groups * n_groups;
    const int chunks = p.Cin / kBK;
    constexpr int per_acc = mt * nt;          // accumulators per buffer

    if (threadIdx.x == 0) {
        for (int s = 0; s < n_stages; s++) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int b = 0; b < 2; b++) {
            mbar_init(acc_full0 + 8 * b, 1);
            mbar_init(acc_empty0 + 8 * b, 4);       // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        // TMEM: 128 lanes x 2 buffers x (mt * nt <= 2) accumulators x kBN fp32 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(4 * kBN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a_hi);
        tma_prefetch_desc(&map_a_lo);
        tma_prefetch_desc(&map_w_hi);
        tma_prefetch_desc(&map_w_lo);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // union of the member tiles' tap masks (a tap that is pure padding for one member just loads zeros for it)
    auto group_taps = [&](int mg, int &mt_here) -> uint32_t {
        mt_here = min(mt, m_tiles - mg * mt);
        uint32_t mask = 0;
        for (int i = 0; i < mt_here; i++) {
            const TileOrigin to = tile_origin(p, Ho, mg * mt + i);
            mask |= conv_tap_mask(p, Ho, to.oy0, to.ox0);
        }
        return mask;
    };

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int it = 0;
            for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
                const int mg = w / n_groups, ng = w % n_groups;       // neighbouring CTAs share the A tiles in L2
                int mt_here;
                const uint32_t tap_mask = group_taps(mg, mt_here);
                const uint32_t tiles_bytes = (uint32_t)(mt_here + nt) * (p.x3 ? 2u : 1u) * kTileBytes;
                for (int tap = 0; tap < 25; tap++) {
                    if (!((tap_mask >> tap) & 1)) continue;
                    const int ky = tap / 5, kx = tap % 5;
                    for (int ch = 0; ch < chunks; ch++, it++) {
                        const int s = it % n_stages;
                        if (it >= n_stages) mbar_wait(empty0 + 8 * s, ((it / n_stages) - 1) & 1);
                        const uint32_t full = full0 + 8 * s;
                        const uint32_t stage = base + s * stage_bytes;
                        mbar_expect_tx(full, tiles_bytes);
                        const int c0 = ch * kBK;
                        for (int i = 0; i < mt_here; i++) {
                            const TileOrigin to = tile_origin(p, Ho, mg * mt + i);
                            const int ix0 = 2 * to.ox0 + kx - 1, iy0 = 2 * to.oy0 + ky - 1;
                            tma_load_4d(stage + (2 * i) * kTileBytes, &map_a_hi, full, c0, ix0, iy0, to.img0);
                            if (p.x3) tma_load_4d(stage + (2 * i + 1) * kTileBytes, &map_a_lo, full, c0, ix0, iy0, to.img0);
                        }
                        for (int j = 0; j < nt; j++) {
                            const int n0 = (ng * nt + j) * kBN;
                            tma_load_2d(stage + (2 * mt + 2 * j) * kTileBytes, &map_w_hi, full, tap * p.Cin + c0, n0);
                            if (p.x3) tma_load_2d(stage + (2 * mt + 2 * j + 1) * kTileBytes, &map_w_lo, full, tap * p.Cin + c0, n0);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(kBM, kBN);
            int it = 0, tile_i = 0;
            for (int w = blockIdx.x; w < n_work; w += gridDim.x, tile_i++) {
                int mt_here;
                const int n_iter = __popc(group_taps(w / n_groups, mt_here)) * chunks;
                const int b = tile_i & 1, use = tile_i >> 1;
                if (use > 0) {                                    // the epilogue has drained this buffer
                    mbar_wait(acc_empty0 + 8 * b, (use - 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                for (int j_it = 0; j_it < n_iter; j_it++, it++) {
                    const int s = it % n_stages;
                    mbar_wait(full0 + 8 * s, (it / n_stages) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t stage = base + s * stage_bytes;
#pragma unroll
                    for (int i = 0; i < mt; i++) {
                        if (i >= mt_here) break;
                        const uint64_t a_hi = umma_desc(stage + (2 * i) * kTileBytes), a_lo = umma_desc(stage + (2 * i + 1) * kTileBytes);
#pragma unroll
                        for (int j = 0; j < nt; j++) {
                            const uint64_t w_hi = umma_desc(stage + (2 * mt + 2 * j) * kTileBytes);
                            const uint64_t w_lo = umma_desc(stage + (2 * mt + 2 * j + 1) * kTileBytes);
                            const uint32_t tmem_acc = tmem_base + (uint32_t)((b * per_acc + i * nt + j) * kBN);
#pragma unroll
                            for (int k = 0; k < kBK / 16; k++) {
                                const uint64_t adv = (uint64_t)((k * 32) >> 4);       // 16 bf16 = 32 bytes along K
                                if (p.x3) {
                                    umma_bf16(tmem_acc, a_lo + adv, w_hi + adv, idesc, (j_it | k) != 0);
                                    umma_bf16(tmem_acc, a_hi + adv, w_lo + adv, idesc, 1);
                                    umma_bf16(tmem_acc, a_hi + adv, w_hi + adv, idesc, 1);
                                } else {
                                    umma_bf16(tmem_acc, a_hi + adv, w_hi + adv, idesc, (j_it | k) != 0);
                                }
                            }
                        }
                    }
                    umma_commit(empty0 + 8 * s);          // stage free once these MMAs have read it
                }
                umma_commit(acc_full0 + 8 * b);           // accumulators complete
            }
        }
    } else {
        // ===== epilogue: warps 2..5 own TMEM lane quarters (warp % 4) =====
        const int q = warp & 3;
        const int et = threadIdx.x - 64;                  // 0..127 among the epilogue threads
        int tile_i = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, tile_i++) {
            const int mg = w / n_groups, ng = w % n_groups;
            const int mt_here = min(mt, m_tiles - mg * mt);
            const int b = tile_i & 1, use = tile_i >> 1;
            mbar_wait(acc_full0 + 8 * b, use & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int i = 0; i < mt_here; i++) {
                for (int j = 0; j < nt; j++) {
                    const int m_tile = mg * mt + i, n_tile = ng * nt + j;
                    const uint32_t tmem_acc = tmem_base + (uint32_t)((b * per_acc + i * nt + j) * kBN);
                    const int row = q * 32 + lane;
                    int m = m_tile * kBM + row;
                    bool row_ok = m < p.M;
                    if (p.pos) {                                   // row = image, all at output pixel posi
                        const TileOrigin to = tile_origin(p, Ho, m_tile);
                        row_ok = to.img0 + row < p.B;
                        m = (to.img0 + row) * (Ho * Ho) + to.posi;
                    }
                    float *orow = out + (size_t)m * p.Cout + (size_t)n_tile * kBN;
#pragma unroll 1
                    for (int cb = 0; cb < kBN; cb += 32) {
                        float v[32];
                        tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)cb, v);
#pragma unroll
                        for (int jj = 0; jj < 32; jj++)
                            v[jj] = row_ok ? v[jj] + (bias ? __ldg(bias + n_tile * kBN + cb + jj) : 0.0f) : 0.0f;
                        if (row_ok) {
#pragma unroll
                            for (int jj = 0; jj < 32; jj += 4)
                                *reinterpret_cast<float4 *>(orow + cb + jj) = make_float4(v[jj], v[jj + 1], v[jj + 2], v[jj + 3]);
                        }
                        if (bn_partial) {
                            // column sums over this warp's 32 rows via an smem transpose
#pragma unroll
                            for (int jj = 0; jj < 32; jj++) s_tr[q][lane][jj] = v[jj];
                            __syncwarp();
                            float sm = 0.f, sq = 0.f;
#pragma unroll
                            for (int r = 0; r < 32; r++) {
                                const float x = s_tr[q][r][lane];
                                sm += x;
                                sq += x * x;
                            }
                            s_red[q][0][cb + lane] = sm;
                            s_red[q][1][cb + lane] = sq;
                            __syncwarp();
                        }
                    }
                    if (bn_partial) {
                        epilogue_bar_sync();
                        if (et < kBN) {
                            const int c = et;
                            const float sm = (s_red[0][0][c] + s_red[1][0][c]) + (s_red[2][0][c] + s_red[3][0][c]);
                            const float sq = (s_red[0][1][c] + s_red[1][1][c]) + (s_red[2][1][c] + s_red[3][1][c]);
                            float *pp = bn_partial + ((size_t)m_tile * p.Cout + n_tile * kBN + c) * 2;
                            pp[0] = sm;
                            pp[1] = sq;
                        }
                        epilogue_bar_sync();                      // s_red is free for the next tile
                    }
                }
            }
            // this warp's TMEM reads are complete (tcgen05.wait::ld inside tmem_ld32): release the buffer
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (lane == 0) mbar_arrive(acc_empty0 + 8 * b);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(4 * kBN) : "memory");
    }
}

static void conv_tc_launch(dim3 grid, cudaStream_t st, const CUtensorMap &ma_hi, const CUtensorMap &ma_lo, const CUtensorMap &mw_hi,
                           const CUtensorMap &mw_lo, const ConvTcParams &p, const float *bias, float *out, float *part)
{
    if (p.nt == 2) conv_tc_kernel<1, 2><<<grid, kTcThreads, kSmemBytes, st>>>(ma_hi, ma_lo, mw_hi, mw_lo, p, bias, out, part);
    else if (p.mt == 2) conv_tc_kernel<2, 1><<<grid, kTcThreads, kSmemBytes, st>>>(ma_hi, ma_lo, mw_hi, mw_lo, p, bias, out, part);
    else conv_tc_kernel<1, 1><<<grid, kTcThreads, kSmemBytes, st>>>(ma_hi, ma_lo, mw_hi, mw_lo, p, bias, out, part);
}

static cudaError_t conv_tc_set_smem()
{
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    return e;
}

// tiles per work item: two n-tiles share A when Cout >= 256, two m-tiles share W when Cout = 128 -- as long as
// enough work items remain to fill the 148 SMs (SPIRAL_CONV_PAIR=0 disables: profiling)
// position-pure tiles (see TileOrigin): maps of 2x2 .. 8x8 output pixels, whole tiles of 128 images
// (SPIRAL_CONV_POS=0 disables: profiling)
static int conv_tc_position_tiles(int batch, int Ho)
{
    static int enabled = -1;
    if (enabled < 0) {
        const char *e = getenv("SPIRAL_CONV_POS");
        enabled = e ? (atoi(e) != 0) : 1;
    }
    return (enabled && Ho >= 2 && Ho <= 8 && batch >= kBM && (batch % kBM) == 0) ? 1 : 0;
}

static void conv_tc_pairing(ConvTcParams *p, int mtiles)
{
    p->mt = 1;
    p->nt = 1;
    static int enabled = -1;
    if (enabled < 0) {
        const char *e = getenv("SPIRAL_CONV_PAIR");
        enabled = e ? (atoi(e) != 0) : 1;
    }
    if (!enabled) return;
    const int ntiles = p->Cout / kBN;
    if (ntiles >= 2 && (ntiles % 2) == 0 && (long long)mtiles * (ntiles / 2) >= 148) p->nt = 2;
    else if (mtiles >= 2 * 148) p->mt = 2;      // with position-pure tiles the pair's tap masks differ: the union is loaded
}

// ---------------------------------------------------------------------------------------------
// elementwise helpers: hi/lo split of activations (with the producer's BN + leaky-ReLU applied)
// and of the packed weights
// ---------------------------------------------------------------------------------------------
// a = lrelu(raw * scale[c] + shift[c]);  raw fp32 NHWC with C channels (C % 4 == 0)
__global__ void act_split_kernel(const float *__restrict__ raw, const float *__restrict__ scale,
                                 const float *__restrict__ shift, __nv_bfloat16 *__restrict__ hi,
                                 __nv_bfloat16 *__restrict__ lo, size_t n4, int C)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = reinterpret_cast<const float4 *>(raw)[i];
        const int c = (int)((i * 4) % (size_t)C);
        const float4 sc = *reinterpret_cast<const float4 *>(scale + c);
        const float4 sh = *reinterpret_cast<const float4 *>(shift + c);
        float a[4] = {fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w)};
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            a[j] = a[j] > 0.f ? a[j] : 0.2f * a[j];
            split_bf16(a[j], h[j], l[j]);
        }
        reinterpret_cast<uint2 *>(hi)[i] = make_uint2(
            (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16),
            (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16));
        reinterpret_cast<uint2 *>(lo)[i] = make_uint2(
            (uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16),
            (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16));
    }
}

// HWIO [25][Cin][Cout] fp32 -> K-major [Cout][25*Cin] bf16 hi / lo
__global__ void weight_pack_kernel(const float *__restrict__ w, __nv_bfloat16 *__restrict__ hi,
                                   __nv_bfloat16 *__restrict__ lo, int Cin, int Cout)
{
    const size_t total = (size_t)25 * Cin * Cout;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int n = (int)(i / ((size_t)25 * Cin));
        const int k = (int)(i % ((size_t)25 * Cin));        // tap * Cin + c
        const float x = w[(size_t)k * Cout + n];
        split_bf16(x, hi[i], lo[i]);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
EncodeTiledFn tc_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        const cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e == cudaSuccess && p && qres == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
    }
    return fn;
}

struct DiscTc {
    EncodeTiledFn encode;
    int first_tc_layer;       // layers >= this use tensor cores (needs Cin % 64 == 0 and Cout % 128 == 0)
};

static size_t tc_layer_act_elems(const DiscPlan &pl, int l, int batch) { return (size_t)batch * pl.H[l] * pl.H[l] * pl.Cin[l]; }

struct TcOffsets {
    size_t act_hi[SPIRAL_DISC_MAX_LAYERS], act_lo[SPIRAL_DISC_MAX_LAYERS];
    size_t w_hi[SPIRAL_DISC_MAX_LAYERS], w_lo[SPIRAL_DISC_MAX_LAYERS];
    size_t w0_hi, w0_lo;      // layer 0: [64][K + pad] bf16 (disc_conv0_mma.cu)
    size_t bytes;
};

static void tc_plan(const SpiralDisc *d, int batch, const DiscPlan &pl, TcOffsets *o)
{
    // the tensor-core region follows the fp32 region; weights first (their offsets must not depend on batch)
    size_t off = 0;
    o->w0_hi = off; off += disc_align(conv0_mma_pack_bytes(d->cfg.in_channels));
    o->w0_lo = off; off += disc_align(conv0_mma_pack_bytes(d->cfg.in_channels));
    for (int l = 1; l < d->n_layers; l++) {
        const size_t wn = (size_t)25 * pl.Cin[l] * pl.Cout[l];
        o->w_hi[l] = off; off += disc_align(wn * 2);
        o->w_lo[l] = off; off += disc_align(wn * 2);
    }
    for (int l = 1; l < d->n_layers; l++) {
        const size_t an = tc_layer_act_elems(pl, l, batch);
        o->act_hi[l] = off; off += disc_align(an * 2);
        o->act_lo[l] = off; off += disc_align(an * 2);
    }
    o->bytes = off;
}

int disc_tc_init(SpiralDisc *d)
{
    if (d->cfg.disc_dim < 64 || (d->cfg.disc_dim % 64) != 0)
        return spiral_set_error(SPIRAL_EUNSUPPORTED,
                                "tensor-core discriminator needs disc_dim %% 64 == 0 (got %d); use precision 0", d->cfg.disc_dim);
    DiscTc *t = new (std::nothrow) DiscTc();
    if (!t) return spiral_set_error(SPIRAL_EINVAL, "out of host memory");
    t->encode = tc_encode_fn();
    if (!t->encode) {
        delete t;
        return spiral_set_error(SPIRAL_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    }
    t->first_tc_layer = 1;
    cudaError_t e = conv_tc_set_smem();
    if (e != cudaSuccess) {
        delete t;
        return spiral_cuda_error(e, "cudaFuncSetAttribute(conv_tc_kernel)");
    }
    d->tc = t;
    return SPIRAL_OK;
}

size_t disc_tc_workspace_bytes(const SpiralDisc *d, int batch)
{
    DiscPlan pl;
    disc_plan(d, batch, &pl);
    TcOffsets o;
    tc_plan(d, batch, pl, &o);
    return o.bytes;
}

// the fp32 region is sized for max_batch so that the tensor-core region (and the packed weights in
// it) sits at a fixed offset regardless of the batch of a particular call
static char *tc_region(const SpiralDisc *d, char *ws)
{
    DiscPlan plmax;
    disc_plan(d, d->cfg.max_batch, &plmax);
    return ws + plmax.simt_bytes;
}

int disc_tc_pack_weights(SpiralDisc *d, char *ws, cudaStream_t st)
{
    DiscPlan pl;
    disc_plan(d, d->cfg.max_batch, &pl);
    TcOffsets o;
    tc_plan(d, d->cfg.max_batch, pl, &o);
    char *tc = tc_region(d, ws);
    if (conv0_mma_supported(d)) conv0_mma_pack(d, tc + o.w0_hi, tc + o.w0_lo, st);
    for (int l = 1; l < d->n_layers; l++) {
        weight_pack_kernel<<<148 * 4, 256, 0, st>>>((const float *)d->w.conv_w[l], (__nv_bfloat16 *)(tc + o.w_hi[l]),
                                                    (__nv_bfloat16 *)(tc + o.w_lo[l]), pl.Cin[l], pl.Cout[l]);
    }
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "disc_tc_pack_weights");
}

static int encode_act_map(const DiscTc *t, CUtensorMap *map, void *ptr, int B, int H, int C, int bw, int bh, int bb)
{
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)H, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * H, (cuuint64_t)C * 2 * H * H};
    // traversal stride 2 in W and H: boxDim = elements loaded * elementStride
    const cuuint32_t box[4] = {(cuuint32_t)kBK, (cuuint32_t)(bw * 2), (cuuint32_t)(bh * 2), (cuuint32_t)bb};
    const cuuint32_t estr[4] = {1, 2, 2, 1};
    const CUresult r = t->encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, ptr, dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

static int encode_w_map(const DiscTc *t, CUtensorMap *map, void *ptr, int K, int N)
{
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)kBN};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = t->encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

int disc_tc_forward(SpiralDisc *d, const float *x, int batch, float *logits, float *probs, float *bn_stats,
                    char *ws, cudaStream_t st, int stage, double count, double *sums)
{
    const DiscTc *t = (const DiscTc *)d->tc;
    DiscPlan pl;
    disc_plan(d, batch, &pl);                       // fp32 raw outputs, partials, scale/shift (offsets for this batch)
    DiscPlan plmax;
    disc_plan(d, d->cfg.max_batch, &plmax);
    TcOffsets o;
    tc_plan(d, d->cfg.max_batch, plmax, &o);        // offsets fixed by max_batch
    char *tc = tc_region(d, ws);
    const SpiralDiscWeights &W = d->w;
    const int maxC = d->cfg.disc_dim << (d->n_layers - 1);
    const int x3 = d->cfg.precision == 1;

    // layer 0 (K = 25 * C, C in {1,3,6}): mma.sync implicit GEMM gathered from a shared-memory input tile
    // (disc_conv0_mma.cu; direct fp32 convolution when disc_dim != 64); the epilogue applies the leaky-ReLU
    // (no BN at layer 0) and writes the bf16 hi/lo pair layer 1 reads
    // staged (global-batch batch norm, spiral_disc_forward_staged): stage s finishes layer s's batch norm from `sums`
    // (s >= 1), runs layer s + 1 (stage 0: layers 0 and 1) and leaves that layer's per-channel sums in `sums`
    const int l_first = stage <= 0 ? 1 : stage + 1;
    const int l_last = stage < 0 ? d->n_layers - 1 : (stage + 1 < d->n_layers ? stage + 1 : d->n_layers - 1);
    if (stage >= 1)
        launch_bn_affine(sums, pl.Cout[stage], count, (const float *)W.bn_gamma[stage], (const float *)W.bn_beta[stage], 1e-3f,
                         (float *)(ws + pl.scale_off[stage]), (float *)(ws + pl.shift_off[stage]),
                         bn_stats ? bn_stats + (size_t)stage * 2 * maxC : nullptr, maxC, st);
    if (stage <= 0) {
        int rc = conv0_mma_supported(d)
                     ? conv0_mma_forward(d, x, batch, tc + o.w0_hi, tc + o.w0_lo, tc + o.act_hi[1], tc + o.act_lo[1], x3, st)
                     : disc_conv0(d, x, batch, nullptr, tc + o.act_hi[1], tc + o.act_lo[1], st);
        if (rc != SPIRAL_OK) return rc;
    }
    for (int l = l_first; l <= l_last; l++) {
        const int H = pl.H[l], Cin = pl.Cin[l], Cout = pl.Cout[l], Ho = H / 2;
        // producer's BN + leaky-ReLU, then hi/lo split of this layer's input
        const size_t n4 = tc_layer_act_elems(pl, l, batch) / 4;
        __nv_bfloat16 *ahi = (__nv_bfloat16 *)(tc + o.act_hi[l]), *alo = (__nv_bfloat16 *)(tc + o.act_lo[l]);
        const int blocks = (int)((n4 + 255) / 256 < (size_t)(148 * 8) ? (n4 + 255) / 256 : (size_t)(148 * 8));
        if (l > 1)
            act_split_kernel<<<blocks, 256, 0, st>>>((const float *)(ws + pl.act_off[l - 1]), (const float *)(ws + pl.scale_off[l - 1]),
                                                  (const float *)(ws + pl.shift_off[l - 1]), ahi, alo, n4, Cin);
        ConvTcParams p;
        p.B = batch; p.H = H; p.Cin = Cin; p.Cout = Cout; p.M = batch * Ho * Ho; p.x3 = x3;
        if (Ho * Ho >= kBM) { p.bw = Ho; p.bh = kBM / Ho; p.bb = 1; }
        else { p.bw = Ho; p.bh = Ho; p.bb = kBM / (Ho * Ho); }
        p.pos = conv_tc_position_tiles(batch, Ho);
        if (p.pos) { p.bw = 1; p.bh = 1; p.bb = kBM; }
        CUtensorMap ma_hi, ma_lo, mw_hi, mw_lo;
        int er = encode_act_map(t, &ma_hi, ahi, batch, H, Cin, p.bw, p.bh, p.bb);
        er |= encode_act_map(t, &ma_lo, alo, batch, H, Cin, p.bw, p.bh, p.bb);
        er |= encode_w_map(t, &mw_hi, tc + o.w_hi[l], 25 * Cin, Cout);
        er |= encode_w_map(t, &mw_lo, tc + o.w_lo[l], 25 * Cin, Cout);
        if (er) return spiral_set_error(SPIRAL_ECUDA, "cuTensorMapEncodeTiled failed for layer %d (CUresult %d)", l, er);
        const int mtiles = (p.M + kBM - 1) / kBM;
        const bool bn = d->cfg.batch_norm != 0;
        float *part = (float *)(ws + pl.part_off[l]);
        conv_tc_pairing(&p, mtiles);
        const int n_work = ((mtiles + p.mt - 1) / p.mt) * ((Cout / kBN) / p.nt);
        dim3 grid(n_work < 148 ? n_work : 148);
        conv_tc_launch(grid, st, ma_hi, ma_lo, mw_hi, mw_lo, p, (const float *)W.conv_b[l], (float *)(ws + pl.act_off[l]),
                       bn ? part : nullptr);
        float *scale = (float *)(ws + pl.scale_off[l]), *shift = (float *)(ws + pl.shift_off[l]);
        if (bn && stage >= 0) {
            launch_bn_sums(part, mtiles, Cout, sums, st);                // the caller all-reduces, the next stage finishes
        } else if (bn) {
            launch_bn_finalize(part, mtiles, Cout, (double)batch * Ho * Ho, (const float *)W.bn_gamma[l],
                               (const float *)W.bn_beta[l], 1e-3f, scale, shift,
                               bn_stats ? bn_stats + (size_t)l * 2 * maxC : nullptr, maxC, st);
        } else {
            fill_affine_kernel<<<(Cout + 127) / 128, 128, 0, st>>>(scale, shift, Cout, 1.0f, 0.0f);
        }
    }
    const int last = d->n_layers - 1;
    if (stage < 0 || stage == last)
        dense_sigmoid_kernel<<<(batch * 32 + 127) / 128, 128, 0, st>>>(
            (const float *)(ws + pl.act_off[last]), (const float *)(ws + pl.scale_off[last]), (const float *)(ws + pl.shift_off[last]),
            (const float *)W.dense_w, (const float *)W.dense_b, logits, probs, batch, pl.Cout[last]);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPIRAL_OK : spiral_cuda_error(e, "spiral_disc_forward(tensor cores)");
}

// ---------------------------------------------------------------------------------------------
// standalone layer (conv_tc.cu / spiral_conv_forward): same kernel, no bias, no BN partials
// ---------------------------------------------------------------------------------------------
void conv_tc_pack_weights(const float *w, void *w_hi, void *w_lo, int Cin, int Cout, cudaStream_t st)
{
    weight_pack_kernel<<<148 * 4, 256, 0, st>>>(w, (__nv_bfloat16 *)w_hi, (__nv_bfloat16 *)w_lo, Cin, Cout);
}

int conv_tc_layer(void *a_hi, void *a_lo, void *w_hi, void *w_lo, float *out, int B, int H, int Cin, int Cout,
                  int x3, cudaStream_t st)
{
    static bool attr_set = false;
    DiscTc t;
    t.encode = tc_encode_fn();
    t.first_tc_layer = 1;
    if (!t.encode) return spiral_set_error(SPIRAL_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    if (!attr_set) {
        const cudaError_t e = conv_tc_set_smem();
        if (e != cudaSuccess) return spiral_cuda_error(e, "cudaFuncSetAttribute(conv_tc_kernel)");
        attr_set = true;
    }
    const int Ho = H / 2;
    ConvTcParams p;
    p.B = B; p.H = H; p.Cin = Cin; p.Cout = Cout; p.M = B * Ho * Ho; p.x3 = x3;
    if (Ho * Ho >= kBM) { p.bw = Ho; p.bh = kBM / Ho; p.bb = 1; }
    else { p.bw = Ho; p.bh = Ho; p.bb = kBM / (Ho * Ho); }
    p.pos = conv_tc_position_tiles(B, Ho);
    if (p.pos) { p.bw = 1; p.bh = 1; p.bb = kBM; }
    CUtensorMap ma_hi, ma_lo, mw_hi, mw_lo;
    int er = encode_act_map(&t, &ma_hi, a_hi, B, H, Cin, p.bw, p.bh, p.bb);
    er |= encode_act_map(&t, &ma_lo, a_lo, B, H, Cin, p.bw, p.bh, p.bb);
    er |= encode_w_map(&t, &mw_hi, w_hi, 25 * Cin, Cout);
    er |= encode_w_map(&t, &mw_lo, w_lo, 25 * Cin, Cout);
    if (er) return spiral_set_error(SPIRAL_ECUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", er);
    const int mtiles = (p.M + kBM - 1) / kBM;
    conv_tc_pairing(&p, mtiles);
    const int n_work = ((mtiles + p.mt - 1) / p.mt) * ((Cout / kBN) / p.nt);
    dim3 grid(n_work < 148 ? n_work : 148);
    conv_tc_launch(grid, st, ma_hi, ma_lo, mw_hi, mw_lo, p, nullptr, out, nullptr);
    return SPIRAL_OK;
}

}  // namespace spiral
