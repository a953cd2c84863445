// sm_100a "tile block" kernels: fused depthwise KSxKS (3x3 / 5x5, stride 1 / 2) -> pointwise blocks of ANY map size and
// channel count (K = Cs_in up to 1024, N up to 256) with
//   * the input halo staged in shared memory by ONE TMA TENSOR copy per 32-channel chunk (cp.async.bulk.tensor.3d through
//     a CUtensorMap over the NHWC activation with the image and row dimensions merged: box = 32 channels x tile width +
//     halo x tile rows + halo; columns / channels outside the tensor are zero-filled by the hardware, completion counted on
//     an mbarrier together with the chunk's depthwise weights; SASS: UTMALDG) - every input element is read from L2 / HBM
//     once per CTA instead of up to KS*KS times by per-thread window loads, and the next chunk (of this tile or of the
//     CTA's next tile) streams in while the current one is being convolved;
//   * the depthwise stage on the CUDA cores as a sliding window (4 horizontally adjacent outputs x one channel quad per
//     thread, packed FFMA2, weights of the chunk in shared memory), written straight into the UMMA K-major operand tile
//     as TF32 hi / lo parts;
//   * the pointwise contraction on tcgen05.mma kind::tf32 (3xTF32, FP32 accumulation in TMEM), pre-split weights
//     streamed per K chunk by TMA bulk copy; the commit is waited for only where the operand tile is overwritten or the
//     accumulator read, and the issuing thread rotates over the warps;
//   * the fused epilogue (bias -> act -> residual [channel-pad, 2x2 max-pool] -> act) from TMEM.
//
// A tile = 128 output pixels = TH "virtual rows" x TW columns, where the virtual rows run over (image, output row) in
// order: on small maps (12x12, 6x6, 3x3 ...) a tile spans several images, so the M = 128 MMA rows stay full whatever the
// map size.  CTAs are PERSISTENT (two per SM, striding over the tiles): TMEM, barriers and - when they fit - the pointwise
// weights are set up once per CTA.  Replaces both the global-window tcgen05 kernel (dwpw_tc_kernel) and the FFMA
// implicit-GEMM tile (conv_gemm_kernel<.., CONV_DWPW, ..>) for these blocks.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstdlib>

#include "conv_common.cuh"
#include "kernels.h"
#include "tc_common.cuh"
#include "tc_epilogue.cuh"

namespace zb {
namespace {

using namespace tc;

constexpr int TCB_M = 128;            // UMMA M
constexpr int TCB_CK = 32;            // channels per K chunk
constexpr int TCB_KQC = TCB_CK / 4;   // channel quads per chunk
constexpr int TCB_AROWS = TCB_M + 1;  // padded chunk stride of the A tile: conflict-free 128-bit stores
// floats reserved for the A tile (hi + lo); the epilogue's staging tile (128 x TCE_STRIDE) aliases it once the MMAs are done
constexpr int TCB_A_FLOATS = 2 * TCB_KQC * TCB_AROWS * 4 > TCB_M * TCE_STRIDE ? 2 * TCB_KQC * TCB_AROWS * 4 : TCB_M * TCE_STRIDE;

constexpr int TCB_ZERO_FLOATS = ((4 - 1) * 2 + 5) * TCB_CK;   // widest window row: SPAN = (PPT - 1) * S + KS channel-chunk pixels

#ifndef TCB_GEMM_CTAS
#define TCB_GEMM_CTAS 3
#endif

struct TcbGeom {
    int TW, TH;            // tile = TH virtual rows x TW columns, TW * TH = 128
    int tiles_x, tiles_y;
    int vrows;             // images * Ho
    int WBOX;              // staged input columns per row = (TW - 1) * S + KS
    int rows_max;          // worst-case staged input rows of one tile
    int nin;               // input staging buffers (1 or 2)
    int nbw;               // pointwise-weight buffers (1 or 2)
    int tw_shift;          // log2(TW)
    unsigned ho_magic, tx_magic;   // magic multipliers of the per-tile divisions by Ho and tiles_x (0: divisor is 1)
};

// n / d for 0 <= n, n * d < 2^32, with magic = floor(2^32 / d) + 1 (d >= 2) or 0 (d == 1)
__device__ __forceinline__ int fast_div(int n, unsigned magic) { return magic ? (int)__umulhi((unsigned)n, magic) : n; }
__host__ unsigned div_magic(int d) { return d <= 1 ? 0u : (unsigned)((1ull << 32) / (unsigned)d + 1ull); }
// the same with an escape: magic == ~0u means "n * d may reach 2^32: divide for real" (div_magic_checked decides on the host)
__device__ __forceinline__ int fast_div_or(int n, int d, unsigned magic) { return magic == 0xffffffffu ? n / d : fast_div(n, magic); }
__host__ unsigned div_magic_checked(int d, unsigned long long n_max) {
    return (n_max * (unsigned long long)d < (1ull << 32)) ? div_magic(d) : 0xffffffffu;
}

__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *tmap, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}

// ------------------------------------------------------------------------------------------------
// PPT = horizontally adjacent outputs per producer thread: 4 (256 threads, 2 CTAs per SM at <= 128 registers) or 2 (512 lighter
// threads, 2 CTAs per SM at <= 64 registers: twice the warps to hide shared-memory and barrier latency behind)
template <int KS, int S, int PPT>
__global__ void __launch_bounds__(1024 / PPT, 2) tcb_dwpw_kernel(const __grid_constant__ CUtensorMap tmap, const ConvDev p,
                                                          const float *__restrict__ w_hi, const float *__restrict__ w_lo, int NP,
                                                          int nchunks, const TcbGeom g) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    constexpr int TAPS = KS * KS, SPAN = (PPT - 1) * S + KS, NT = 1024 / PPT;
    const int halo_floats = g.rows_max * g.WBOX * TCB_CK;                    // staged halo of one chunk (multiple of 32 floats = 128 B)
    const int in_floats = halo_floats + (TAPS + 1) * TCB_CK;                 // ... followed by the chunk's depthwise weights + bias
    float *s_in0 = reinterpret_cast<float *>(smem_raw);                      // nin x {[rows][WBOX][32], [TAPS + 1][32]}   (TMA destinations)
    float *sA_hi = s_in0 + (size_t)g.nin * in_floats;                        // [8][129][4]
    float *sA_lo = sA_hi + TCB_KQC * TCB_AROWS * 4;
    float *sB0 = sA_hi + TCB_A_FLOATS;                                       // nbw x {hi [8][NP][4], lo [8][NP][4]}   (bulk-copy destinations)
    const int b_floats = 2 * TCB_KQC * NP * 4;
    __shared__ __align__(8) uint64_t mbar_in[2], mbar_b[2], mbar_mma;
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) TceRow s_rowinfo[TCB_M];
    // a window row of zeros: tap rows outside the image read it instead of branching around the row, which keeps the whole
    // window in one basic block (the per-row branch stopped the compiler from issuing row ky + 1's loads above row ky's FMAs)
    // (it lives behind the weight buffers in dynamic shared memory: static shared memory is not part of the launch's budget)
    float *s_zero = sB0 + (size_t)g.nbw * b_floats;                          // [TCB_ZERO_FLOATS]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < TCB_ZERO_FLOATS; i += 1024 / PPT) s_zero[i] = 0.0f;
    const uint32_t ncols = tmem_cols_for(NP);
    const uint32_t b_bytes = (uint32_t)TCB_KQC * NP * 16;
    const uint32_t in_bytes = (uint32_t)in_floats * 4;                        // halo box + depthwise weights: one mbarrier phase
    // PERSISTENT: this CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...; a "step" is one (tile, K chunk) pair and the
    // input / weight rings run across tile boundaries, so the next tile's first halo is in flight during this tile's
    // epilogue, TMEM and the barriers are set up once, and layers whose weight chunks all fit in the weight buffers
    // (K <= 32 * nbw) fetch them once per CTA instead of once per tile.
    const int total_tiles = g.tiles_x * g.tiles_y;
    const int my_tiles = ((int)blockIdx.x < total_tiles) ? (total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int total_steps = my_tiles * nchunks;
    const bool b_resident = nchunks <= g.nbw;

    // pointwise weights of step s (chunk s % nchunks) -> weight buffer s % nbw
    auto load_b = [&](int s) {
        const int b = s % g.nbw, c = s % nchunks;
        float *dst = sB0 + (size_t)b * b_floats;
        mbar_expect_tx(&mbar_b[b], 2 * b_bytes);
        bulk_copy_g2s(dst, w_hi + (size_t)c * TCB_KQC * NP * 4, b_bytes, &mbar_b[b]);
        bulk_copy_g2s(dst + TCB_KQC * NP * 4, w_lo + (size_t)c * TCB_KQC * NP * 4, b_bytes, &mbar_b[b]);
    };
    // stage the halo of step s into input buffer s % nin with ONE tensor copy: the (image, row) dimensions of the activation
    // are merged (images are contiguous, H == Ho * S), so a tile's input rows are one contiguous range even when it spans
    // several images; tap rows that belong to a neighbouring image read the row of zeros instead (window loop below).  (One copy per staged
    // row - the first version - paid the TMA unit's per-instruction cost ~20 times per chunk.)
    auto issue_in = [&](int s) {
        if (lane == 0 && s < total_steps) {
            const int tile = (int)blockIdx.x + (s / nchunks) * (int)gridDim.x, c = s % nchunks;
            const int tx = tile % g.tiles_x, ty = tile / g.tiles_x;
            float *dst = s_in0 + (size_t)(s % g.nin) * in_floats;
            uint64_t *bar = &mbar_in[s % g.nin];
            mbar_expect_tx(bar, in_bytes);
            tma_load_3d(dst, &tmap, c * TCB_CK, tx * g.TW * S - p.pl, ty * g.TH * S - p.pt, bar);
            // the chunk's depthwise weights + bias, pre-packed per chunk (zero beyond the true channel count), ride along
            bulk_copy_g2s(dst + halo_floats, p.dw_c + (size_t)c * (TAPS + 1) * TCB_CK, (uint32_t)(TAPS + 1) * TCB_CK * 4, bar);
        }
    };
    if (warp == 0) tmem_alloc(&tmem_slot, ncols);
    if (tid == 0) {
        mbar_init(&mbar_in[0], 1);
        mbar_init(&mbar_in[1], 1);
        mbar_init(&mbar_b[0], 1);
        mbar_init(&mbar_b[1], 1);
        mbar_init(&mbar_mma, 1);
        for (int s0 = 0; s0 < g.nbw && s0 < total_steps; s0++) load_b(s0);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    pdl_trigger();                                                            // (TMEM is allocated: the next kernel's CTAs may move in)
    pdl_wait();                                                               // the previous kernel's activations are complete and visible
    if (warp == 0) issue_in(0);

    // this thread's producer item: strip of PPT horizontally adjacent outputs x channel quad
    const int quad = tid & 7, strip = tid >> 3;
    const int strips_x = g.TW / PPT;
    const int pr = strip / strips_x, pcol = (strip - pr * strips_x) * PPT;     // virtual row offset, first column of the strip
    const int pm = pr * g.TW + pcol;                                          // tile row (MMA row) of the strip's first pixel
    const int origin = S * pr * g.WBOX + pcol * S;                            // staged pixel index of the window's corner
    const uint32_t idesc = make_idesc_tf32(TCB_M, NP);
    const uint64_t ad_hi = make_smem_desc(smem_u32(sA_hi), TCB_AROWS * 16, 128), ad_lo = make_smem_desc(smem_u32(sA_lo), TCB_AROWS * 16, 128);

    // The MMAs of step s are only WAITED FOR where their result or their operand buffer is needed: before the A tile is
    // overwritten by step s + 1 (after that step's depthwise FMAs, which only touch registers) and before the epilogue.
    // (ncu on the first version, which waited right after the commit: a third of all warp samples sat on that barrier.)
    bool mma_pending = false;
    int step = 0;
    auto mma_done = [&]() {                                                   // all threads; `step` = steps issued so far
        if (!mma_pending) return;
        mbar_wait(&mbar_mma, (step - 1) & 1);
        tc_fence_after();
        mma_pending = false;
        if (tid == 0 && !b_resident && step - 1 + g.nbw < total_steps) load_b(step - 1 + g.nbw);   // that weight buffer is free again
    };
    const bool prelu_mid = p.act_mid.kind == ACT_PRELU;
    // The epilogue (quarter-local barriers, tc_epilogue.cuh; outputs on the all-128-bit path only: tcb_dwpw_supported) ends
    // WITHOUT a CTA barrier, so the first step of the next tile places one before it overwrites the row table and the A tile.
    int it = 0;                                                               // tiles done by this CTA
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
        const int tile_y = fast_div(tile, g.tx_magic), tile_x = tile - tile_y * g.tiles_x;
        const int vr0 = tile_y * g.TH, ox0 = tile_x * g.TW;
        const int n_vr = min(g.TH, g.vrows - vr0);                           // valid virtual rows of this tile (>= 1)
        const bool p_valid = pr < n_vr && ox0 + pcol < p.Wo;
        const int p_vr = vr0 + pr;
        const int p_iy0 = (p_vr - fast_div(p_vr, g.ho_magic) * p.Ho) * S - p.pt;          // input row of the window's first tap row
        unsigned kymask = 0;                                                  // bit ky: that tap row lies inside the image
#pragma unroll
        for (int ky = 0; ky < KS; ky++) kymask |= (p_iy0 + ky >= 0 && p_iy0 + ky < p.H) ? (1u << ky) : 0u;
        for (int c = 0; c < nchunks; c++, step++) {
            const int buf = g.nin == 2 ? (step & 1) : 0;
            mbar_wait(&mbar_in[buf], (g.nin == 2 ? (step >> 1) : step) & 1);  // this step's halo + depthwise weights have landed
            // buffer (step + 1) % 2 was last read by step - 1, and everybody passed that step's barrier
            if (g.nin == 2 && warp == 0) issue_in(step + 1);                  // next step streams in behind this one

            float4 v[PPT];
            const int k = c * TCB_CK + quad * 4;
            if (p_valid) {
                const float *s_w = s_in0 + (size_t)buf * in_floats + halo_floats;
                const float *s_in = s_in0 + (size_t)buf * in_floats + (size_t)origin * TCB_CK + quad * 4;
                const float4 bias = *reinterpret_cast<const float4 *>(s_w + TAPS * TCB_CK + quad * 4);
#pragma unroll
                for (int i = 0; i < PPT; i++) v[i] = bias;
#pragma unroll
                for (int ky = 0; ky < KS; ky++) {
                    float4 x[SPAN];
                    const float *rowp = ((kymask >> ky) & 1u) ? s_in + (size_t)ky * g.WBOX * TCB_CK : s_zero + quad * 4;
#pragma unroll
                    for (int j = 0; j < SPAN; j++) x[j] = *reinterpret_cast<const float4 *>(rowp + j * TCB_CK);
#pragma unroll
                    for (int kx = 0; kx < KS; kx++) {
                        const float4 wv = *reinterpret_cast<const float4 *>(s_w + (ky * KS + kx) * TCB_CK + quad * 4);
#pragma unroll
                        for (int i = 0; i < PPT; i++) fma4(v[i], x[i * S + kx], wv);   // packed FFMA2, rounds like fmaf
                    }
                }
                if (k < p.K && p.act_mid.kind != ACT_NONE) {
                    if (prelu_mid) {
                        const float4 sl = ldg4(p.act_mid.slope + k);
#pragma unroll
                        for (int i = 0; i < PPT; i++) {
                            v[i].x = v[i].x < 0.0f ? v[i].x * sl.x : v[i].x;
                            v[i].y = v[i].y < 0.0f ? v[i].y * sl.y : v[i].y;
                            v[i].z = v[i].z < 0.0f ? v[i].z * sl.z : v[i].z;
                            v[i].w = v[i].w < 0.0f ? v[i].w * sl.w : v[i].w;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < PPT; i++) act4(v[i], p.act_mid, k);
                    }
                }
            }
            mma_done();                                                       // the previous step's MMAs have read the A tile
            if (c == 0) {
                if (it > 0) __syncthreads();                                  // the previous tile's epilogue is over for everybody
                if (tid < TCB_M) {                                            // row table of this tile's epilogue
                    const int er = tid >> g.tw_shift, ecol = tid - (er << g.tw_shift);
                    const int evr = vr0 + er, ox = ox0 + ecol;
                    TceRow ri;
                    ri.out_off = -1, ri.res_off = 0;
                    if (er < n_vr && ox < p.Wo) {
                        const int img = fast_div(evr, g.ho_magic);
                        ri = tce_row(p, img, evr - img * p.Ho, ox);
                    }
                    s_rowinfo[tid] = ri;
                }
            }
            if (p_valid) {
#pragma unroll
                for (int i = 0; i < PPT; i++) {
                    float4 hi, lo;
                    split_tf32_fast(v[i].x, hi.x, lo.x);
                    split_tf32_fast(v[i].y, hi.y, lo.y);
                    split_tf32_fast(v[i].z, hi.z, lo.z);
                    split_tf32_fast(v[i].w, hi.w, lo.w);
                    *reinterpret_cast<float4 *>(sA_hi + ((size_t)quad * TCB_AROWS + pm + i) * 4) = hi;
                    *reinterpret_cast<float4 *>(sA_lo + ((size_t)quad * TCB_AROWS + pm + i) * 4) = lo;
                }
            }
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
            if (g.nin == 1 && warp == 0) issue_in(step + 1);                  // single buffer: refill behind the MMAs
            // the issuing thread rotates over the warps: the ~100 dependent instructions of descriptor arithmetic and MMA
            // issue sit on a different warp's critical path every step (every warp does the same producer work)
            if (tid == ((step & (NT / 32 - 1)) << 5)) {
                tc_fence_after();
                uint32_t acc_flag = c > 0;
                const int bb = b_resident ? c : step % g.nbw;
                mbar_wait(&mbar_b[bb], b_resident ? 0 : ((step / g.nbw) & 1)); // this chunk's pointwise weights have landed
                const float *sB = sB0 + (size_t)bb * b_floats;
                const uint64_t bd_hi = make_smem_desc(smem_u32(sB), (uint32_t)NP * 16, 128);
                const uint64_t bd_lo = make_smem_desc(smem_u32(sB + TCB_KQC * NP * 4), (uint32_t)NP * 16, 128);
#pragma unroll 1
                for (int pass = 0; pass < 3; pass++) {                        // lo*hi, hi*lo, hi*hi (small terms first)
                    uint64_t ad = pass == 0 ? ad_lo : ad_hi;
                    uint64_t bd = pass == 1 ? bd_lo : bd_hi;
#pragma unroll
                    for (int j = 0; j < TCB_CK / 8; j++) {
                        umma_tf32(tmem, ad, bd, idesc, acc_flag);
                        acc_flag = 1;
                        ad += (uint64_t)(2 * TCB_AROWS);
                        bd += (uint64_t)(2 * NP);
                    }
                }
                umma_commit(&mbar_mma);
            }
            mma_pending = true;
        }
        // --- epilogue through shared memory (coalesced residual reads / stores); the staging tile aliases the A tile, the
        // next tile's first halo is already on its way.  `step` now counts this tile's last step as issued. ---
        // the wait for the last MMAs sits inside, behind the epilogue's residual / bias loads
        tc_epilogue_tile_quarters<NT>(p, tmem, 0, NP, s_rowinfo, sA_hi, tid, [&]() { mma_done(); });
        tc_fence_before();                                                    // accumulator read before the next tile's MMAs overwrite it
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, ncols);
}

// ------------------------------------------------------------------------------------------------
// Plain convolutions as tcgen05 GEMMs ("direct" operand: no depthwise stage): 1x1 convs, Gemm / dense heads, and
// windowed convs (2x2 stride 2, 3x3 stride 3 ...) whose im2col row is a concatenation of whole pixels (the tap of a
// channel quad is worked out per thread, so a 32-wide K chunk may span several taps when Cs_in < 32).  One CTA = 128 consecutive output pixels x one
// tile of up to 256 output channels (blockIdx.y); the A chunk is gathered with coalesced 128-bit loads, TF32-split and
// stored in the UMMA K-major layout; the loads of chunk c + 1 are in flight while the MMAs of chunk c run.
// Weights: [N tile][Kpad / 4][NT][4], TF32 hi / lo, one TMA bulk copy per chunk and half.
// ------------------------------------------------------------------------------------------------
// Three CTAs per SM (<= 85 registers): ncu had a third of this kernel's warp samples on the barrier behind the MMA wait at two.
// The row -> (image, y, x) divisions use host-computed magic multipliers (17 % of the executed instructions were divisions).
__global__ void __launch_bounds__(256, TCB_GEMM_CTAS) tcb_gemm_kernel(const ConvDev p, const float *__restrict__ w_hi, const float *__restrict__ w_lo,
                                                          int NP, int nchunks, int kpad, unsigned howo_magic, unsigned wo_magic) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int n0 = blockIdx.y * 256;
    const int NT = min(256, NP - n0);                                        // output channels of this CTA (multiple of 16)
    float *sA_hi = reinterpret_cast<float *>(smem_raw);                      // [8][129][4]
    float *sA_lo = sA_hi + TCB_KQC * TCB_AROWS * 4;
    float *sB_hi = sA_hi + TCB_A_FLOATS;                                     // 2 x [8][NT][4]: double-buffered weight chunks
    float *sB_lo = sB_hi + 2 * TCB_KQC * NT * 4;
    __shared__ __align__(8) uint64_t mbar_b[2], mbar_mma;
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) TceRow s_rowinfo[TCB_M];

    const int tid = threadIdx.x, warp = tid >> 5;
    const int m0 = blockIdx.x * TCB_M;
    const int HoWo = p.Ho * p.Wo;
    const uint32_t ncols = tmem_cols_for(NT);
    const uint32_t b_bytes = (uint32_t)TCB_KQC * NT * 16;
    // this N tile's weights: tiles of 256 columns precede it
    const float *wt_hi = w_hi + (size_t)blockIdx.y * kpad * 256, *wt_lo = w_lo + (size_t)blockIdx.y * kpad * 256;
    auto load_b = [&](int c) {
        const int b = c & 1;
        mbar_expect_tx(&mbar_b[b], 2 * b_bytes);
        bulk_copy_g2s(sB_hi + (size_t)b * TCB_KQC * NT * 4, wt_hi + (size_t)c * TCB_KQC * NT * 4, b_bytes, &mbar_b[b]);
        bulk_copy_g2s(sB_lo + (size_t)b * TCB_KQC * NT * 4, wt_lo + (size_t)c * TCB_KQC * NT * 4, b_bytes, &mbar_b[b]);
    };
    if (warp == 0) tmem_alloc(&tmem_slot, ncols);
    if (tid == 0) {
        mbar_init(&mbar_b[0], 1);
        mbar_init(&mbar_b[1], 1);
        mbar_init(&mbar_mma, 1);
        load_b(0);
        if (nchunks > 1) load_b(1);
    }
    // this thread's 4 operand rows (row = (tid >> 3) + 32 i) and channel quad
    const int quad = tid & 7;
    long long roff[4];
    int riy[4], rix[4];
    // one true division per CTA tile: (image, offset) of the tile's first row; the rows of a tile are consecutive, so row
    // m0 + d sits at offset r_b + d < HoWo + 128 of image img_b, and the small quotients come from magic multipliers
    const int img_b = m0 / HoWo, r_b = m0 - img_b * HoWo;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int d = (tid >> 3) + 32 * i;
        const int gm = m0 + d;
        roff[i] = -1;
        riy[i] = rix[i] = 0;
        if (gm < p.M) {
            const int wrap = fast_div_or(r_b + d, HoWo, howo_magic);
            const int img = img_b + wrap;
            const int r = r_b + d - wrap * HoWo;
            const int oy = fast_div_or(r, p.Wo, wo_magic), ox = r - oy * p.Wo;
            roff[i] = (long long)img * p.in_img_stride;
            riy[i] = oy * p.sh - p.pt;
            rix[i] = ox * p.sw - p.pl;
        }
    }
    const bool multi_tap = p.kh * p.kw > 1;
    auto fetch = [&](int c, float4 (&x)[4]) {
        int k = c * TCB_CK + quad * 4, ky = 0, kx = 0;
        const bool kok = k < p.K;                                             // (K is padded to the chunk width with zero weights)
        if (multi_tap) {
            // per thread: with Cs_in < 32 a chunk spans several taps (a quad never does: Cs_in % 4 == 0)
            const int tap = k / p.Cs_in;
            k -= tap * p.Cs_in;
            ky = tap / p.kw, kx = tap - ky * p.kw;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int iy = riy[i] + ky, ix = rix[i] + kx;
            x[i] = (roff[i] >= 0 && kok && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)
                       ? ldg4(p.in + roff[i] + ((long long)iy * p.W + ix) * p.Cs_in + k)
                       : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    float4 x[4];
    pdl_wait();                                                               // the previous kernel's activations are complete and visible
    fetch(0, x);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_trigger();                                                            // (after the TMEM allocation)
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = make_idesc_tf32(TCB_M, NT);
    const uint64_t ad_hi = make_smem_desc(smem_u32(sA_hi), TCB_AROWS * 16, 128), ad_lo = make_smem_desc(smem_u32(sA_lo), TCB_AROWS * 16, 128);
    uint32_t acc_flag = 0;

    for (int c = 0; c < nchunks; c++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int m = (tid >> 3) + 32 * i;
            float4 hi, lo;
            split_tf32_fast(x[i].x, hi.x, lo.x);
            split_tf32_fast(x[i].y, hi.y, lo.y);
            split_tf32_fast(x[i].z, hi.z, lo.z);
            split_tf32_fast(x[i].w, hi.w, lo.w);
            *reinterpret_cast<float4 *>(sA_hi + ((size_t)quad * TCB_AROWS + m) * 4) = hi;
            *reinterpret_cast<float4 *>(sA_lo + ((size_t)quad * TCB_AROWS + m) * 4) = lo;
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (c + 1 < nchunks) fetch(c + 1, x);                                 // next chunk's loads fly behind the MMAs
        if (tid == 0) {
            const int b = c & 1;
            mbar_wait(&mbar_b[b], (c >> 1) & 1);
            const uint64_t bd_hi = make_smem_desc(smem_u32(sB_hi + (size_t)b * TCB_KQC * NT * 4), (uint32_t)NT * 16, 128);
            const uint64_t bd_lo = make_smem_desc(smem_u32(sB_lo + (size_t)b * TCB_KQC * NT * 4), (uint32_t)NT * 16, 128);
#pragma unroll 1
            for (int pass = 0; pass < 3; pass++) {
                uint64_t ad = pass == 0 ? ad_lo : ad_hi;
                uint64_t bd = pass == 1 ? bd_lo : bd_hi;
#pragma unroll
                for (int j = 0; j < TCB_CK / 8; j++) {
                    umma_tf32(tmem, ad, bd, idesc, acc_flag);
                    acc_flag = 1;
                    ad += (uint64_t)(2 * TCB_AROWS);
                    bd += (uint64_t)(2 * NT);
                }
            }
            umma_commit(&mbar_mma);
            mbar_wait(&mbar_mma, c & 1);
            if (c + 2 < nchunks) load_b(c + 2);                               // this weight buffer is free again
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }

    // --- epilogue through shared memory (the staging tile aliases the A tile) -----------------------------------------
    if (tid < TCB_M) {
        const int m = m0 + tid;
        TceRow ri;
        ri.out_off = -1, ri.res_off = 0;
        if (m < p.M) {
            const int wrap = fast_div_or(r_b + tid, HoWo, howo_magic);
            const int img = img_b + wrap;
            const int r = r_b + tid - wrap * HoWo;
            const int oy = fast_div_or(r, p.Wo, wo_magic);
            ri = tce_row(p, img, oy, r - oy * p.Wo);
        }
        s_rowinfo[tid] = ri;
    }
    __syncthreads();
    tc_epilogue_tile<256, 0>(p, tmem, n0, NT, s_rowinfo, sA_hi, tid);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, ncols);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return (EncodeTiledFn) nullptr;
        }
        return (EncodeTiledFn)f;
    }();
    return fn;
}

// NHWC activation as a 3-D tensor (channel, x, image * H + y); box = 32 channels x WBOX columns x BR rows; OOB -> 0.
bool make_input_map(const ConvDev &p, int images, int WBOX, int BR, CUtensorMap *out) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)p.Cs_in, (cuuint64_t)p.W, (cuuint64_t)p.H * images};
    const cuuint64_t strides[2] = {(cuuint64_t)p.Cs_in * 4, (cuuint64_t)p.W * p.Cs_in * 4};
    const cuuint32_t box[3] = {(cuuint32_t)TCB_CK, (cuuint32_t)WBOX, (cuuint32_t)BR};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(p.in), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

TcbGeom choose_geom(const ConvDev &p, int KS, int S) {
    const int images = p.M / (p.Ho * p.Wo);
    TcbGeom best{};
    double best_cost = 1e30;
    static const int force_tw = getenv("ZB_TCB_TW") ? atoi(getenv("ZB_TCB_TW")) : 0;
    for (int TW = 4; TW <= (S == 2 ? 64 : 128); TW *= 2) {
        if (force_tw && TW != force_tw) continue;
        const int TH = TCB_M / TW;
        const int tiles_x = (p.Wo + TW - 1) / TW;
        const int WBOX = (TW - 1) * S + KS;
        if (WBOX > 256) continue;
        const double rows = (TH - 1) * S + KS;
        const double valid = (double)TH * p.Wo / tiles_x;
        const double cost = rows * WBOX / valid + 0.5 * TCB_M / valid;   // staged pixels per output + half-weighted MMA / producer waste
        if (cost < best_cost) {
            best_cost = cost;
            best.TW = TW, best.TH = TH, best.tiles_x = tiles_x, best.WBOX = WBOX;
        }
    }
    best.vrows = images * p.Ho;
    best.tiles_y = (best.vrows + best.TH - 1) / best.TH;
    best.rows_max = (best.TH - 1) * S + KS;    // one contiguous row range per tile (merged image / row dimension)
    best.tw_shift = 0;
    while ((1 << best.tw_shift) < best.TW) best.tw_shift++;
    // the magic divisions are exact while n * d < 2^32: n < vrows (resp. tiles), d = Ho (resp. tiles_x)
    const bool ho_ok = (unsigned long long)best.vrows * (unsigned)p.Ho < (1ull << 32);
    const bool tx_ok = (unsigned long long)best.tiles_x * best.tiles_y * (unsigned)best.tiles_x < (1ull << 32);
    if (!ho_ok || !tx_ok) best.TW = 0;         // (never for the networks here: 65535 frames x 96 rows is 6e6 virtual rows)
    best.ho_magic = div_magic(p.Ho), best.tx_magic = div_magic(best.tiles_x);
    return best;
}

size_t tcb_smem(const TcbGeom &g, int KS, int NP, int nin, int nbw = 1) {
    return (size_t)nin * ((size_t)g.rows_max * g.WBOX * TCB_CK + (size_t)(KS * KS + 1) * TCB_CK) * 4 + (size_t)TCB_A_FLOATS * 4 +
           (size_t)nbw * 2 * TCB_KQC * NP * 16 + (size_t)TCB_ZERO_FLOATS * 4;   // (static smem is 3 KB: the 1024-byte alignment of the base costs nothing)
}

template <int KS, int S, int PPT>
bool launch_tcb_cfg(const ConvDev &p, const float *w_hi, const float *w_lo, int NP, cudaStream_t s) {
    TcbGeom g = choose_geom(p, KS, S);
    if (g.TW == 0) return false;
    // two staging buffers when two CTAs still fit on an SM with them (the next chunk streams in behind the convolution of
    // this one); otherwise one buffer, refilled behind the MMAs.  A layer whose smallest footprint already excludes a second
    // CTA (stride-2 halos, N = 256 weight chunks) has the rest of the SM's shared memory for free: it takes the second weight
    // buffer first (a 64 KB chunk fetched only after the previous MMAs retire is a serial L2 round trip per step), then the
    // second staging buffer.
    static const int force_nin = getenv("ZB_TCB_NIN") ? atoi(getenv("ZB_TCB_NIN")) : 0;
    static const int force_nbw = getenv("ZB_TCB_NBW") ? atoi(getenv("ZB_TCB_NBW")) : 0;
    static const bool fill_sm = !(getenv("ZB_TCB_FILL") && atoi(getenv("ZB_TCB_FILL")) == 0);
    constexpr size_t TWO_CTAS = 110 * 1024, ONE_CTA = 220 * 1024;
    const int nchunks = (p.K + TCB_CK - 1) / TCB_CK;
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    // one CTA per SM whatever the buffering - or fewer tiles than SMs (the deep layers of small batches: every CTA has its SM to
    // itself and walks a long chain of K chunks, so the next chunk's halo and weights should already be on their way)
    const bool alone = fill_sm && (tcb_smem(g, KS, NP, 1, 1) > TWO_CTAS || g.tiles_x * g.tiles_y <= num_sms);
    int nin, nbw;
    if (alone) {
        nbw = (nchunks > 1 && tcb_smem(g, KS, NP, 1, 2) <= ONE_CTA) ? 2 : 1;
        nin = tcb_smem(g, KS, NP, 2, nbw) <= ONE_CTA ? 2 : 1;
    } else {
        nin = tcb_smem(g, KS, NP, 2) <= TWO_CTAS ? 2 : 1;
        nbw = tcb_smem(g, KS, NP, nin, 2) <= TWO_CTAS ? 2 : 1;              // second weight buffer only while two CTAs still fit per SM
    }
    if (force_nin == 1 || force_nin == 2) nin = force_nin;
    if (tcb_smem(g, KS, NP, nin) > ONE_CTA) nin = 1;
    if (force_nbw == 1 || force_nbw == 2) nbw = force_nbw;
    if (tcb_smem(g, KS, NP, nin, nbw) > ONE_CTA) nbw = 1;
    const size_t smem = tcb_smem(g, KS, NP, nin, nbw);
    if (smem > 220 * 1024) return false;
    g.nin = nin;
    g.nbw = nbw;
    const int images = p.M / (p.Ho * p.Wo);
    CUtensorMap tmap;
    if (!make_input_map(p, images, g.WBOX, g.rows_max, &tmap)) return false;
    auto kern = tcb_dwpw_kernel<KS, S, PPT>;
    static SmemOptIn opt_in;
    if (!opt_in.ensure(kern, smem)) return false;
    // persistent CTAs: two per SM (what registers and shared memory allow), each striding over the tiles
    static const int ctas_per_sm = getenv("ZB_TCB_CTAS_PER_SM") ? atoi(getenv("ZB_TCB_CTAS_PER_SM")) : 2;
    // (a footprint that leaves room for one CTA per SM gets one: a second "wave" of persistent CTAs would only queue up)
    const int resident = (fill_sm && smem > TWO_CTAS) ? 1 : ctas_per_sm;
    const int grid = ctas_per_sm > 0 ? std::min(g.tiles_x * g.tiles_y, resident * num_sms) : g.tiles_x * g.tiles_y;
    ZB_KNAME("tcb_dwpw_kernel", KS, S, PPT);
    launch_pdl(1, kern, dim3((unsigned)grid), dim3(1024 / PPT), smem, s, tmap, p, w_hi, w_lo, NP, nchunks, g);
    return true;
}

}  // namespace

bool tcb_dwpw_supported(const ConvDev &p, int NP) {
    static const bool disabled = getenv("ZB_NO_TCB") && atoi(getenv("ZB_NO_TCB")) != 0;
    if (disabled) return false;
    if (!((p.kh == 3 && p.kw == 3) || (p.kh == 5 && p.kw == 5))) return false;
    if (!((p.sh == 1 && p.sw == 1) || (p.sh == 2 && p.sw == 2))) return false;
    if (p.K != p.Cs_in || p.K % 8 || p.K < 8 || p.K > 1024 || NP % 16 || NP < 16 || NP > 256 || !p.dw_c) return false;
    if (p.M % (p.Ho * p.Wo) || p.pt < 0 || p.pl < 0) return false;
    if (((uintptr_t)p.in) % 16 || (p.in_img_stride % 4)) return false;
    if (p.in_img_stride != (long long)p.H * p.W * p.Cs_in || p.H != p.Ho * p.sh) return false;   // merged (image, row) staging
    // the epilogue is the all-128-bit one (every block-to-block tensor qualifies; odd-width graph outputs go elsewhere)
    if (p.out_pix_stride % 4 || p.Nstore % 4 || p.out_img_stride % 4 || p.Ns % 4 || ((uintptr_t)p.out) % 16 ||
        (p.epi.res && p.epi.res_Cs % 4))
        return false;
    return encode_fn() != nullptr;
}

bool tcb_gemm_supported(const ConvDev &p, int NP) {
    static const bool disabled = getenv("ZB_NO_TCB_GEMM") && atoi(getenv("ZB_NO_TCB_GEMM")) != 0;
    if (disabled) return false;
    if (p.Cs_in % 8 || p.Cs_in < 8 || NP % 16 || NP < 16) return false;
    if (p.K != p.kh * p.kw * p.Cs_in || p.K > 8192) return false;
    if (p.M % (p.Ho * p.Wo) || ((uintptr_t)p.in) % 16 || (p.in_img_stride % 4)) return false;
    return true;
}

bool launch_tcb_gemm(const ConvDev &p, const float *w_hi, const float *w_lo, int NP, int kpad, cudaStream_t s) {
    if (!tcb_gemm_supported(p, NP)) return false;
    const int ntiles = (NP + 255) / 256;
    const int nt_max = std::min(NP, 256);
    const size_t smem = (size_t)TCB_A_FLOATS * 4 + 4 * (size_t)TCB_KQC * nt_max * 16 + 1024;
    static SmemOptIn opt_in;
    if (!opt_in.ensure(tcb_gemm_kernel, smem)) return false;
    g_launch_count++;
    const int nchunks = (p.K + TCB_CK - 1) / TCB_CK;
    ZB_KNAME("tcb_gemm_kernel");
    launch_pdl(2, tcb_gemm_kernel, dim3((unsigned)((p.M + TCB_M - 1) / TCB_M), (unsigned)ntiles), dim3(256), smem, s, p, w_hi, w_lo, NP, nchunks, kpad,
               div_magic_checked(p.Ho * p.Wo, (unsigned long long)p.Ho * p.Wo + TCB_M), div_magic_checked(p.Wo, (unsigned long long)p.Ho * p.Wo));
    return true;
}

bool launch_tcb_dwpw(const ConvDev &p, const float *w_hi, const float *w_lo, int NP, cudaStream_t s) {
    if (!tcb_dwpw_supported(p, NP)) return false;
    g_launch_count++;
    bool ok;
    // measured per layer (profiles/): 512 lighter threads win on the 3x3 stride-2 blocks, 256 threads of 4 outputs elsewhere
    static const int ppt_env = getenv("ZB_TCB_PPT") ? atoi(getenv("ZB_TCB_PPT")) : 0;
    const int ppt = ppt_env ? ppt_env : (p.kh == 3 && p.sh == 2) ? 2 : 4;
    if (ppt == 2) {
        if (p.kh == 3) ok = p.sh == 1 ? launch_tcb_cfg<3, 1, 2>(p, w_hi, w_lo, NP, s) : launch_tcb_cfg<3, 2, 2>(p, w_hi, w_lo, NP, s);
        else ok = p.sh == 1 ? launch_tcb_cfg<5, 1, 2>(p, w_hi, w_lo, NP, s) : launch_tcb_cfg<5, 2, 2>(p, w_hi, w_lo, NP, s);
    } else {
        if (p.kh == 3) ok = p.sh == 1 ? launch_tcb_cfg<3, 1, 4>(p, w_hi, w_lo, NP, s) : launch_tcb_cfg<3, 2, 4>(p, w_hi, w_lo, NP, s);
        else ok = p.sh == 1 ? launch_tcb_cfg<5, 1, 4>(p, w_hi, w_lo, NP, s) : launch_tcb_cfg<5, 2, 4>(p, w_hi, w_lo, NP, s);
    }
    if (!ok) g_launch_count--;
    return ok;
}

}  // namespace zb
