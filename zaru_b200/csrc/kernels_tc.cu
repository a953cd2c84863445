// sm_100a tensor-core path: the pointwise (1x1) contraction of the fused depthwise->pointwise blocks on
// tcgen05.mma (kind::tf32) with the accumulator in TMEM.
//
//   dwpw_tc_kernel: one CTA = 128 output pixels (UMMA M = 128) x all output channels.
//     1. 256 threads produce the depthwise result for their pixels straight into shared memory in the UMMA
//        K-major canonical layout (tc_common.cuh), split into TF32 hi/lo parts;
//     2. the pre-split pointwise weights (hi/lo, same layout, packed at load time) are copied next to it;
//     3. ONE thread issues the K/8 x 3 tcgen05.mma instructions (3xTF32: lo*hi + hi*lo + hi*hi, FP32
//        accumulation in TMEM) and commits them to an mbarrier;
//     4. all 8 warps read their TMEM lanes back (tcgen05.ld 32x32b) and run the fused epilogue
//        (bias -> act1 -> +residual [channel-pad, 2x2 max-pool] -> act2) with 128-bit stores.
//   3xTF32 keeps the contraction at FP32-level accuracy (the reference is pure f32; plain TF32/BF16 operands
//   move BlazeFace logits by up to 1e-3), at 3x a negligible MMA cost: these layers are bound by
//   the depthwise producer and the epilogue, not by the tensor pipe.
//
//   tc_gemm_test_kernel: D[128,N] = A[128,K] * B[N,K]^T through the same descriptors (unit test of the path).
#include <cuda_runtime.h>

#include <climits>
#include <cstdlib>

#include "conv_common.cuh"
#include "kernels.h"
#include "tc_common.cuh"
#include "tc_epilogue.cuh"

namespace zb {
namespace {

using namespace tc;

constexpr int TC_M = 128;

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) tc_gemm_test_kernel(const float *__restrict__ A, const float *__restrict__ B,
                                                           float *__restrict__ D, int N, int K, int nsplit) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int KQ = K / 4;
    float *sA_hi = reinterpret_cast<float *>(smem_raw);
    float *sA_lo = sA_hi + (size_t)KQ * TC_M * 4;
    float *sB_hi = sA_lo + (size_t)KQ * TC_M * 4;
    float *sB_lo = sB_hi + (size_t)KQ * N * 4;
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t ncols = tmem_cols_for(N);
    if (warp == 0) tmem_alloc(&tmem_slot, ncols);
    if (tid == 0) mbar_init(&mbar, 1);
    for (int e = tid; e < TC_M * K; e += 128) {
        const int m = e / K, k = e - m * K;
        float hi, lo;
        split_tf32(A[e], hi, lo);
        if (nsplit == 1) hi = A[e];   // raw FP32 bits: shows what the hardware does with un-rounded operands
        sA_hi[((size_t)(k / 4) * TC_M + m) * 4 + (k & 3)] = hi;
        sA_lo[((size_t)(k / 4) * TC_M + m) * 4 + (k & 3)] = lo;
    }
    for (int e = tid; e < N * K; e += 128) {
        const int n = e / K, k = e - n * K;
        float hi, lo;
        split_tf32(B[e], hi, lo);
        if (nsplit == 1) hi = B[e];
        sB_hi[((size_t)(k / 4) * N + n) * 4 + (k & 3)] = hi;
        sB_lo[((size_t)(k / 4) * N + n) * 4 + (k & 3)] = lo;
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_tf32(TC_M, N);
        uint32_t acc = 0;
        for (int pass = 0; pass < nsplit; pass++) {
            // pass order: lo*hi, hi*lo, hi*hi (small terms first); nsplit == 1: hi*hi only
            const float *a = (nsplit == 3 && pass == 0) ? sA_lo : sA_hi;
            const float *b = (nsplit == 3 && pass == 1) ? sB_lo : sB_hi;
            for (int j = 0; j < K / 8; j++) {
                const uint64_t ad = make_smem_desc(smem_u32(a) + (uint32_t)(2 * j) * TC_M * 16, TC_M * 16, 128);
                const uint64_t bd = make_smem_desc(smem_u32(b) + (uint32_t)(2 * j) * N * 16, (uint32_t)N * 16, 128);
                umma_tf32(tmem, ad, bd, idesc, acc);
                acc = 1;
            }
        }
        umma_commit(&mbar);
    }
    mbar_wait(&mbar, 0);
    tc_fence_after();
    const int row = warp * 32 + (tid & 31);
    for (int c0 = 0; c0 < N; c0 += 8) {
        float v[8];
        tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (c0 + i < N) D[(size_t)row * N + c0 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, ncols);
}

// ------------------------------------------------------------------------------------------------
// Fused depthwise KSxKS -> pointwise on tcgen05.  p.K = Cs_in (multiple of 8), NP = N padded to 16.
// w_hi / w_lo: [K/4][NP][4] TF32-split pointwise weights.  KC = K-chunk resident in smem per MMA batch.
// ------------------------------------------------------------------------------------------------
// STRIP (3x3, stride 1, Wo % 4 == 0): the producer's work item is 4 horizontally adjacent pixels x one channel
// quad, so the 3x6 input window is loaded once for 4 outputs (4.5 instead of 9 global loads per output quad).
template <int KS, bool STRIP>
__global__ void __launch_bounds__(256, 3) dwpw_tc_kernel(const ConvDev p, const float *__restrict__ w_hi,
                                                      const float *__restrict__ w_lo, int NP, int KC, int Kpad) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int KQC = KC / 4;
    constexpr int A_ROWS = TC_M + 1;                             // chunk stride padded by one row: conflict-free stores
    float *sA_hi = reinterpret_cast<float *>(smem_raw);          // [KQC][129][4]
    float *sA_lo = sA_hi + (size_t)KQC * A_ROWS * 4;
    float *sB_hi = sA_lo + (size_t)KQC * A_ROWS * 4;             // [KQC][NP][4]
    float *sB_lo = sB_hi + (size_t)KQC * NP * 4;
    __shared__ __align__(16) int4 rowinfo[TC_M];                 // {img offset lo, hi, iy0, ix0}; iy0 == INT_MIN: no pixel
    __shared__ __align__(8) uint64_t mbar_mma, mbar_b;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * TC_M;
    const int HoWo = p.Ho * p.Wo;
    const uint32_t ncols = tmem_cols_for(NP);
    const uint32_t b_bytes = (uint32_t)KQC * NP * 16;            // one K chunk of one weight half
    if (warp == 0) tmem_alloc(&tmem_slot, ncols);
    if (tid == 0) {
        mbar_init(&mbar_mma, 1);
        mbar_init(&mbar_b, 1);
        // first weight chunk: two TMA bulk copies (already in UMMA layout, already TF32-split)
        mbar_expect_tx(&mbar_b, 2 * b_bytes);
        bulk_copy_g2s(sB_hi, w_hi, b_bytes, &mbar_b);
        bulk_copy_g2s(sB_lo, w_lo, b_bytes, &mbar_b);
    }
    if (tid < TC_M) {
        const int gm = m0 + tid;
        int4 ri = make_int4(0, 0, INT_MIN, 0);
        if (gm < p.M) {
            const int img = gm / HoWo;
            const int r = gm - img * HoWo;
            const int oy = r / p.Wo, ox = r - oy * p.Wo;
            const long long off = (long long)img * p.in_img_stride;
            ri = make_int4((int)(off & 0xffffffffll), (int)(off >> 32), oy * p.sh - p.pt, ox * p.sw - p.pl);
        }
        rowinfo[tid] = ri;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = make_idesc_tf32(TC_M, NP);

    // depthwise producer for element e of the current chunk: 8 rows x consecutive channel quads per warp, so both
    // the 128-bit global loads (8 pixels x 64 B) and the 128-bit smem stores (8 rows x 16 B = one core matrix)
    // are contiguous.
    auto produce = [&](int e, int kc0) -> float4 {
        const int r8 = e & 7, g = e >> 3;
        const int kq = g % KQC, mg = g / KQC;
        const int m = mg * 8 + r8;
        const int k = kc0 + kq * 4;
        const int4 ri = rowinfo[m];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ri.z != INT_MIN && k < p.K) {
            const long long off = ((long long)ri.y << 32) | (unsigned)ri.x;
            v = dw_window<KS>(p, p.in + off, ri.z, ri.w, k);
            act4(v, p.act_mid, k);
        }
        return v;
    };
    auto store_split = [&](int e, const float4 &v) {
        const int r8 = e & 7, g = e >> 3;
        const int kq = g % KQC, mg = g / KQC;
        const int m = mg * 8 + r8;
        float4 hi, lo;
        split_tf32_fast(v.x, hi.x, lo.x);
        split_tf32_fast(v.y, hi.y, lo.y);
        split_tf32_fast(v.z, hi.z, lo.z);
        split_tf32_fast(v.w, hi.w, lo.w);
        *reinterpret_cast<float4 *>(sA_hi + ((size_t)kq * A_ROWS + m) * 4) = hi;
        *reinterpret_cast<float4 *>(sA_lo + ((size_t)kq * A_ROWS + m) * 4) = lo;
    };
    // strip item: rows m..m+3 are 4 consecutive pixels of one image row (Wo % 4 == 0), channel quad kq
    auto produce_strip = [&](int it, int kc0) {
        const int kq = it % KQC, strip = it / KQC;
        const int m = strip * 4;
        const int k = kc0 + kq * 4;
        const int4 ri = rowinfo[m];
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ri.z != INT_MIN && k < p.K) {
            const long long off = ((long long)ri.y << 32) | (unsigned)ri.x;
            const float *base = p.in + off + k;
            const float4 bias = ldg4(p.dw_b + k);
#pragma unroll
            for (int i = 0; i < 4; i++) v[i] = bias;
#pragma unroll
            for (int ky = 0; ky < 3; ky++) {
                const int iy = ri.z + ky;
                const bool rowok = iy >= 0 && iy < p.H;
                float4 x[6];
#pragma unroll
                for (int c = 0; c < 6; c++) {
                    const int ix = ri.w + c;
                    x[c] = (rowok && ix >= 0 && ix < p.W) ? ldg4(base + ((long long)iy * p.W + ix) * p.Cs_in)
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int kx = 0; kx < 3; kx++) {
                    const float4 wv = ldg4(p.dw_w + (ky * 3 + kx) * p.Cs_in + k);
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        v[i].x = fmaf(x[i + kx].x, wv.x, v[i].x);
                        v[i].y = fmaf(x[i + kx].y, wv.y, v[i].y);
                        v[i].z = fmaf(x[i + kx].z, wv.z, v[i].z);
                        v[i].w = fmaf(x[i + kx].w, wv.w, v[i].w);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; i++) act4(v[i], p.act_mid, k);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float4 hi, lo;
            split_tf32_fast(v[i].x, hi.x, lo.x);
            split_tf32_fast(v[i].y, hi.y, lo.y);
            split_tf32_fast(v[i].z, hi.z, lo.z);
            split_tf32_fast(v[i].w, hi.w, lo.w);
            *reinterpret_cast<float4 *>(sA_hi + ((size_t)kq * A_ROWS + m + i) * 4) = hi;
            *reinterpret_cast<float4 *>(sA_lo + ((size_t)kq * A_ROWS + m + i) * 4) = lo;
        }
    };

    uint32_t acc_flag = 0, phase = 0;
    const int n_elem = TC_M * KQC;
    const uint64_t ad_hi = make_smem_desc(smem_u32(sA_hi), A_ROWS * 16, 128), ad_lo = make_smem_desc(smem_u32(sA_lo), A_ROWS * 16, 128);
    const uint64_t bd_hi = make_smem_desc(smem_u32(sB_hi), (uint32_t)NP * 16, 128), bd_lo = make_smem_desc(smem_u32(sB_lo), (uint32_t)NP * 16, 128);
    for (int kc0 = 0; kc0 < Kpad; kc0 += KC) {
        if (STRIP) {
            for (int it = tid; it < (TC_M / 4) * KQC; it += 256) produce_strip(it, kc0);
        } else {
            // --- A chunk: two elements in flight per thread (all their loads are issued before the FMAs) ----
            int e = tid;
            for (; e + 256 < n_elem; e += 512) {
                const float4 v0 = produce(e, kc0);
                const float4 v1 = produce(e + 256, kc0);
                store_split(e, v0);
                store_split(e + 256, v1);
            }
            if (e < n_elem) store_split(e, produce(e, kc0));
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (tid == 0) {
            mbar_wait(&mbar_b, phase);                       // weights of this chunk have landed (TMA)
            // descriptors are built once; each K step only adds its byte offset (>> 4) to the address field, so the
            // single issuing thread spends ~3 instructions per MMA (it is latency-bound on its own instruction stream)
#pragma unroll 1
            for (int pass = 0; pass < 3; pass++) {
                uint64_t ad = pass == 0 ? ad_lo : ad_hi;
                uint64_t bd = pass == 1 ? bd_lo : bd_hi;
#pragma unroll 4
                for (int j = 0; j < KC / 8; j++) {
                    umma_tf32(tmem, ad, bd, idesc, acc_flag);
                    acc_flag = 1;
                    ad += (uint64_t)(2 * A_ROWS);          // 2 K-quads * A_ROWS * 16 B, in 16-byte units
                    bd += (uint64_t)(2 * NP);
                }
            }
            umma_commit(&mbar_mma);
        }
        // ONE thread polls the mbarrier (a 256-thread try_wait spin loop burns the issue slots the other resident
        // CTAs need); everybody else parks at the hardware barrier until the MMAs have completed.
        if (tid == 0) mbar_wait(&mbar_mma, phase);
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        phase ^= 1;
        if (tid == 0 && kc0 + KC < Kpad) {                   // prefetch the next weight chunk behind the next producer pass
            mbar_expect_tx(&mbar_b, 2 * b_bytes);
            bulk_copy_g2s(sB_hi, w_hi + (size_t)((kc0 + KC) / 4) * NP * 4, b_bytes, &mbar_b);
            bulk_copy_g2s(sB_lo, w_lo + (size_t)((kc0 + KC) / 4) * NP * 4, b_bytes, &mbar_b);
        }
    }

    // --- epilogue: warp w owns TMEM lanes 32*(w%4).., columns [half*NP/2, (half+1)*NP/2) ---------------------
    const int row = (warp & 3) * 32 + lane;
    const int m = m0 + row;
    const int half = warp >> 2;
    const int cbeg = half * (NP / 2), cend = cbeg + NP / 2;
    int img = 0, oy = 0, ox = 0, r = 0;
    const bool rowok = m < p.M;
    if (rowok) {
        img = m / HoWo;
        r = m - img * HoWo;
        oy = r / p.Wo;
        ox = r - oy * p.Wo;
    }
    float *orow = p.out + (long long)img * p.out_img_stride + (long long)r * p.out_pix_stride;
    const bool vec_ok = (p.out_pix_stride % 4 == 0) && (p.Nstore % 4 == 0) && (p.out_img_stride % 4 == 0);
    const uint32_t tbase = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    int c0 = cbeg;
#pragma unroll 1
    for (; c0 + 16 <= cend; c0 += 16) {
        float v[16];
        ResidualPrefetch<16> pre;
        if (rowok) tc_prefetch_residual<16>(p, c0, img, oy, ox, pre);
        tmem_ld16(tbase + (uint32_t)c0, v);
        if (rowok) tc_epilogue_cols<16>(p, v, c0, img, oy, ox, orow, vec_ok, pre);
    }
    for (; c0 + 8 <= cend; c0 += 8) {
        float v[8];
        ResidualPrefetch<8> pre;
        if (rowok) tc_prefetch_residual<8>(p, c0, img, oy, ox, pre);
        tmem_ld8(tbase + (uint32_t)c0, v);
        if (rowok) tc_epilogue_cols<8>(p, v, c0, img, oy, ox, orow, vec_ok, pre);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, ncols);
}

// ------------------------------------------------------------------------------------------------
// Thin fused block on the tensor core ("tile-tc"): depthwise 3x3 (stride 1, pads 1) -> pointwise -> epilogue for
// Cs_in <= 48 on large maps.  One CTA = a 32 x TH output tile of one image (M = 32*TH = 256 or 128 rows).
//   1. input halo tile -> shared memory by cp.async (each input element read from L2/HBM once per CTA);
//   2. depthwise by SLIDING WINDOW: one work item = 4 horizontally adjacent pixels x one channel quad, so the
//      3x6 window is loaded once for 4 outputs (18 instead of 36 LDS.128) and the 9 weight quads once;
//   3. results are TF32-split and stored straight into the UMMA K-major A tile ([k/4][row][4], chunk stride
//      padded to (M+1)*16 B through the descriptor's LBO so the stores are bank-conflict free);
//   4. pointwise weights (pre-split, UMMA layout) arrive by TMA bulk copy; one thread issues the 3xTF32
//      tcgen05.mma chain per 128-row half, accumulators in TMEM;
//   5. every thread reads its own pixel's row back (tcgen05.ld) and applies bias / residual (from the staged
//      tile) / activation, storing N contiguous floats.
// ------------------------------------------------------------------------------------------------
template <int CS, int TH, int NBUF, int TW>
__global__ void __launch_bounds__(256, 4) dwpw_ttc_kernel(const ConvDev p, const float *__restrict__ w_hi,
                                                       const float *__restrict__ w_lo, int NP, int tiles_x, int tiles_y,
                                                       int total_tiles, int poll_all) {
    constexpr int M = TW * TH, NT = 256;
    // pixel stride of the staged tile: CS + 4 words keeps per-pixel 128-bit reads conflict free; for CS == 32 the
    // pad is replaced by an XOR swizzle of the channel quad with the pixel index (3 KB less shared memory, which
    // is what lets a third CTA fit on the SM)
    constexpr bool SWZ = CS == 32;
    constexpr int PS = SWZ ? CS : CS + 4, IW = TW + 2, IH = TH + 2, CQ = CS / 4;
    constexpr int HALVES = M / 128;
    constexpr int IN_TILE = IH * IW * PS;                            // floats per staged input tile
    constexpr uint32_t LBO_A = (M + 1) * 16;                         // padded chunk stride of the A tile (bytes)
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float *sA_hi = reinterpret_cast<float *>(smem_raw);              // [CQ][(M+1)][4]
    float *sA_lo = sA_hi + CQ * (M + 1) * 4;
    float *sB_hi = sA_lo + CQ * (M + 1) * 4;                         // [CQ][NP][4]  (TMA destination, 16 B aligned)
    float *sB_lo = sB_hi + CQ * NP * 4;
    float *s_in0 = sB_lo + CQ * NP * 4;                              // NBUF x [IH][IW][PS]
    float *s_dww = s_in0 + NBUF * IN_TILE;                           // [9][CS]
    float *s_dwb = s_dww + 9 * CS;                                   // [CS]
    float *s_pb = s_dwb + CS;                                        // [NP]
    float *s_sl = s_pb + NP;                                         // [NP]
    __shared__ __align__(8) uint64_t mbar_mma, mbar_b;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t ncols = tmem_cols_for(HALVES * NP);
    const uint32_t b_bytes = (uint32_t)CQ * NP * 16;

    // ---- once per (persistent) CTA: TMEM, barriers, pointwise weights by TMA, depthwise weights, bias, slopes ----
    if (warp == 0) tmem_alloc(&tmem_slot, ncols);
    if (tid == 0) {
        mbar_init(&mbar_mma, 1);
        mbar_init(&mbar_b, 1);
        mbar_expect_tx(&mbar_b, 2 * b_bytes);
        bulk_copy_g2s(sB_hi, w_hi, b_bytes, &mbar_b);
        bulk_copy_g2s(sB_lo, w_lo, b_bytes, &mbar_b);
    }
    for (int e = tid; e < 9 * CQ; e += NT) cp_async16(s_dww + e * 4, p.dw_w + e * 4, 16);
    for (int e = tid; e < CQ; e += NT) cp_async16(s_dwb + e * 4, p.dw_b + e * 4, 16);
    for (int e = tid; e < NP; e += NT) {
        s_pb[e] = e < p.Ns ? __ldg(p.epi.bias + e) : 0.f;
        s_sl[e] = (p.epi.act2.kind == ACT_PRELU && e < p.Ns) ? __ldg(p.epi.act2.slope + e) : 0.f;
    }

    // stage the input halo tile of `tile` into buffer `buf` (cp.async: returns immediately)
    auto tile_coords = [&](int tile, int &tile_x, int &tile_y, int &img) {
        if (NBUF == 1) {
            tile_x = blockIdx.x, tile_y = blockIdx.y, img = blockIdx.z;
        } else {
            int t = tile;
            tile_x = t % tiles_x;
            t /= tiles_x;
            tile_y = t % tiles_y;
            img = t / tiles_y;
        }
    };
    // stage the input halo tile of `tile` into buffer `buf` (cp.async: returns immediately)
    auto issue_fill = [&](int tile, int buf) {
        int tile_x, tile_y, img;
        tile_coords(tile, tile_x, tile_y, img);
        const int oy0 = tile_y * TH, ox0 = tile_x * TW;
        const float *in_img = p.in + (long long)img * p.in_img_stride;
        float *dst = s_in0 + buf * IN_TILE;
        // row by row with the per-thread (column, quad) arithmetic hoisted out of the row loop (the per-chunk
        // divisions were a third of the kernel's instructions)
        constexpr int ROW_CHUNKS = IW * CQ, CPT = (ROW_CHUNKS + NT - 1) / NT;
#pragma unroll
        for (int u = 0; u < CPT; u++) {
            const int c = tid + u * NT;
            if (c < ROW_CHUNKS) {
                const int tx = c / CQ, q = c - tx * CQ;
                const int ix = ox0 - 1 + tx;
                const bool col_ok = ix >= 0 && ix < p.W;
                const float *src = in_img + ((long long)(oy0 - 1) * p.W + ix) * CS + q * 4;
#pragma unroll 2
                for (int ty = 0; ty < IH; ty++) {
                    const int iy = oy0 - 1 + ty, pix = ty * IW + tx;
                    const bool ok = col_ok && iy >= 0 && iy < p.H;
                    cp_async16(dst + pix * PS + (SWZ ? (q ^ (pix & 7)) : q) * 4, ok ? src : in_img, ok ? 16 : 0);
                    src += (long long)p.W * CS;
                }
            }
        }
    };

    // NBUF == 1: 3-D grid (tile_x, tile_y, image), one tile per CTA, no index divisions;
    // NBUF == 2: 1-D persistent grid looping over linear tile ids.
    int tile = NBUF == 1 ? ((int)blockIdx.z * tiles_y + (int)blockIdx.y) * tiles_x + (int)blockIdx.x : (int)blockIdx.x;
    const int tile_step = NBUF == 1 ? total_tiles : (int)gridDim.x;
    int buf = 0;
    uint32_t phase = 0;
    pdl_wait();                                // the previous kernel's activations are complete and visible
    if (tile < total_tiles) issue_fill(tile, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_trigger();                             // (after the TMEM allocation)
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = make_idesc_tf32(128, NP);
    const uint64_t ad_hi = make_smem_desc(smem_u32(sA_hi), LBO_A, 128), ad_lo = make_smem_desc(smem_u32(sA_lo), LBO_A, 128);
    const uint64_t bd_hi = make_smem_desc(smem_u32(sB_hi), (uint32_t)NP * 16, 128), bd_lo = make_smem_desc(smem_u32(sB_lo), (uint32_t)NP * 16, 128);
    const EpiDev &e = p.epi;
    const bool res_smem = e.res == p.in && !e.res_pool;
    bool weights_ready = false;

    for (; tile < total_tiles; tile += tile_step, buf ^= (NBUF - 1)) {
        const int next = NBUF == 2 ? tile + tile_step : total_tiles;
        cp_async_wait_all();                   // this tile's input has landed ...
        tc_fence_before();
        __syncthreads();                       // ... for everybody; previous tile's epilogue is finished
        tc_fence_after();
        if (next < total_tiles) issue_fill(next, buf ^ 1);   // (persistent variant) prefetch behind this tile's compute
        const float *s_in = s_in0 + buf * IN_TILE;

        // --- depthwise, sliding window: item = (strip of 4 pixels along x, channel quad) -----------------------
        constexpr int STRIPS = TH * (TW / 4);
        for (int it = tid; it < STRIPS * CQ; it += NT) {
            const int q = it % CQ, strip = it / CQ;
            const int sy = strip / (TW / 4), sx = (strip % (TW / 4)) * 4;
            const float4 bias = *reinterpret_cast<const float4 *>(s_dwb + q * 4);
            float4 v[4] = {bias, bias, bias, bias};
#pragma unroll
            for (int ky = 0; ky < 3; ky++) {
                float4 x[6];
                const int pix0 = (sy + ky) * IW + sx;
#pragma unroll
                for (int c = 0; c < 6; c++)
                    x[c] = *reinterpret_cast<const float4 *>(s_in + (pix0 + c) * PS + (SWZ ? (q ^ ((pix0 + c) & 7)) : q) * 4);
#pragma unroll
                for (int kx = 0; kx < 3; kx++) {
                    const float4 wv = *reinterpret_cast<const float4 *>(s_dww + (ky * 3 + kx) * CS + q * 4);
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        v[i].x = fmaf(x[i + kx].x, wv.x, v[i].x);
                        v[i].y = fmaf(x[i + kx].y, wv.y, v[i].y);
                        v[i].z = fmaf(x[i + kx].z, wv.z, v[i].z);
                        v[i].w = fmaf(x[i + kx].w, wv.w, v[i].w);
                    }
                }
            }
            const int m = sy * TW + sx;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                act4(v[i], p.act_mid, q * 4);
                float4 hi, lo;
                split_tf32_fast(v[i].x, hi.x, lo.x);
                split_tf32_fast(v[i].y, hi.y, lo.y);
                split_tf32_fast(v[i].z, hi.z, lo.z);
                split_tf32_fast(v[i].w, hi.w, lo.w);
                *reinterpret_cast<float4 *>(sA_hi + (q * (M + 1) + m + i) * 4) = hi;
                *reinterpret_cast<float4 *>(sA_lo + (q * (M + 1) + m + i) * 4) = lo;
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (tid == 0) {
            if (!weights_ready) mbar_wait(&mbar_b, 0), weights_ready = true;   // weights landed (first tile only)
#pragma unroll
            for (int half = 0; half < HALVES; half++) {
                uint32_t acc_flag = 0;
#pragma unroll
                for (int pass = 0; pass < 3; pass++) {
                    // prebuilt descriptors + 16-byte-unit offsets: ~3 instructions per MMA on the issuing thread
                    uint64_t ad = (pass == 0 ? ad_lo : ad_hi) + (uint64_t)(half * 128);
                    uint64_t bd = pass == 1 ? bd_lo : bd_hi;
#pragma unroll
                    for (int j = 0; j < CS / 8; j++) {
                        umma_tf32(tmem + (uint32_t)half * NP, ad, bd, idesc, acc_flag);
                        acc_flag = 1;
                        ad += (uint64_t)(2 * (LBO_A / 16));
                        bd += (uint64_t)(2 * NP);
                    }
                }
            }
            umma_commit(&mbar_mma);
        }
        if (warp == 0) __syncwarp();                 // lanes 1-31 must not spin (and suspend the warp) while lane 0 issues
        if (poll_all) {
            mbar_wait(&mbar_mma, phase);             // every thread polls the mbarrier itself
        } else {
            if (tid == 0) mbar_wait(&mbar_mma, phase);   // one poller; the rest park at bar.sync
            tc_fence_before();
            __syncthreads();
        }
        phase ^= 1;
        tc_fence_after();

        // --- epilogue: thread t owns row (t % 128) of half (t / 128); with M = 128 the upper warps take the
        // upper half of the columns of the same rows.
        int tile_x, tile_y, img;
        tile_coords(tile, tile_x, tile_y, img);
        const int row = (warp & 3) * 32 + lane;
        const int half = HALVES == 2 ? (warp >> 2) : 0;
        const int m = half * 128 + row;
        const int ty = m / TW, tx = m % TW;
        const int oy = tile_y * TH + ty, ox = tile_x * TW + tx;
        const bool ok = oy < p.Ho && ox < p.Wo;
        const int cbeg = HALVES == 2 ? 0 : (warp >> 2) * (NP / 2);
        const int cend = HALVES == 2 ? NP : cbeg + NP / 2;
        float *orow = p.out + (long long)img * p.out_img_stride + ((long long)oy * p.Wo + ox) * p.out_pix_stride;
        const int pix_c = (ty + 1) * IW + tx + 1;
        const float *s_center = s_in + pix_c * PS;
        const uint32_t tbase = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)half * NP;
        for (int c0 = cbeg; c0 < cend; c0 += 8) {
            float v8[8];
            tmem_ld8(tbase + (uint32_t)c0, v8);
            if (!ok) continue;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int n = c0 + 4 * h;
                if (n >= p.Nstore) continue;
                const float4 bv = *reinterpret_cast<const float4 *>(s_pb + n);
                float v[4] = {v8[4 * h] + bv.x, v8[4 * h + 1] + bv.y, v8[4 * h + 2] + bv.z, v8[4 * h + 3] + bv.w};
                act4(v, e.act1, n);
                if (e.res) {
                    float4 rr = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (res_smem) {
                        if (n < CS) rr = *reinterpret_cast<const float4 *>(s_center + (SWZ ? ((n >> 2) ^ (pix_c & 7)) * 4 : n));
                    } else {
                        rr = residual4_at(e, img, oy, ox, n);
                    }
                    v[0] += rr.x, v[1] += rr.y, v[2] += rr.z, v[3] += rr.w;
                }
                if (e.act2.kind == ACT_PRELU) {
                    const float4 sl = *reinterpret_cast<const float4 *>(s_sl + n);
                    v[0] = v[0] < 0.f ? v[0] * sl.x : v[0];
                    v[1] = v[1] < 0.f ? v[1] * sl.y : v[1];
                    v[2] = v[2] < 0.f ? v[2] * sl.z : v[2];
                    v[3] = v[3] < 0.f ? v[3] * sl.w : v[3];
                } else {
                    act4(v, e.act2, n);
                }
                *reinterpret_cast<float4 *>(orow + n) = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
    }
    cp_async_wait_all();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, ncols);
}

template <int CS, int TH, int NBUF, int TW>
size_t ttc_smem(int NP) {
    constexpr int M = TW * TH, CQ = CS / 4, PS = CS == 32 ? CS : CS + 4;
    return sizeof(float) * (2 * (size_t)CQ * (M + 1) * 4 + 2 * (size_t)CQ * NP * 4 + NBUF * (size_t)(TH + 2) * (TW + 2) * PS +
                            10 * CS + 2 * (size_t)NP) + 128;
}

template <int CS, int TH, int NBUF, int TW>
bool launch_ttc_cfg(const ConvDev &p, const float *w_hi, const float *w_lo, int NP, cudaStream_t s) {
    const size_t smem = ttc_smem<CS, TH, NBUF, TW>(NP);
    auto kern = dwpw_ttc_kernel<CS, TH, NBUF, TW>;
    static SmemOptIn opt_in;
    if (!opt_in.ensure(kern, smem)) return false;
    const int tiles_x = (p.Wo + TW - 1) / TW, tiles_y = (p.Ho + TH - 1) / TH;
    const int images = p.M / (p.Ho * p.Wo);
    const int total = tiles_x * tiles_y * images;
    // persistent CTAs: as many as fit on the 148 SMs, each looping over tiles
    static int per_sm = 0, num_sms = 0;
    static size_t per_sm_smem = 0;
    if (per_sm == 0 || per_sm_smem != smem) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem);
        if (per_sm < 1) per_sm = 1;
        per_sm_smem = smem;
    }
    // NBUF == 2: persistent CTAs (as many as fit on the SMs) looping over tiles with the next input tile
    // prefetched; NBUF == 1: one tile per CTA (measured faster: more resident CTAs overlap the serial phases)
    static const int poll_all = getenv("ZB_TC_POLL_ALL") ? atoi(getenv("ZB_TC_POLL_ALL")) : 1;
    ZB_KNAME("dwpw_ttc_kernel", CS, TH, NBUF, TW);
    if (NBUF == 1) {
        launch_pdl(8, kern, dim3(tiles_x, tiles_y, images), dim3(256), smem, s, p, w_hi, w_lo, NP, tiles_x, tiles_y, total, poll_all);
    } else {
        const int grid = total < num_sms * per_sm ? total : num_sms * per_sm;
        launch_pdl(8, kern, dim3((unsigned)grid), dim3(256), smem, s, p, w_hi, w_lo, NP, tiles_x, tiles_y, total, poll_all);
    }
    return true;
}

template <int CS>
bool launch_ttc_cs(const ConvDev &p, const float *w_hi, const float *w_lo, int NP, cudaStream_t s) {
    // tile height: 8 rows (M = 256) when three CTAs still fit in one SM's shared memory, else 4 rows (M = 128)
    static const bool persist = getenv("ZB_TTC_PERSIST") && atoi(getenv("ZB_TTC_PERSIST")) != 0;
    if (persist) {
        if (ttc_smem<CS, 8, 2, 32>(NP) <= 110 * 1024 && p.Ho % 8 == 0) return launch_ttc_cfg<CS, 8, 2, 32>(p, w_hi, w_lo, NP, s);
        return launch_ttc_cfg<CS, 4, 2, 32>(p, w_hi, w_lo, NP, s);
    }
    // maps whose width is a multiple of 16 but not of 32 (48x48): 16 x 8 tiles, no half-empty tile column
    if (p.Wo % 32 != 0 && p.Wo % 16 == 0 && p.Ho % 8 == 0) return launch_ttc_cfg<CS, 8, 1, 16>(p, w_hi, w_lo, NP, s);
    if (ttc_smem<CS, 8, 1, 32>(NP) <= 74 * 1024 && p.Ho % 8 == 0) return launch_ttc_cfg<CS, 8, 1, 32>(p, w_hi, w_lo, NP, s);
    return launch_ttc_cfg<CS, 4, 1, 32>(p, w_hi, w_lo, NP, s);
}

// ------------------------------------------------------------------------------------------------
// Micro-benchmark: back-to-back tcgen05.mma (kind::tf32, M = 128, K = 8) issue rate for a given operand
// layout (LBO / SBO / start offset of A).  One thread issues `iters` MMAs and one commit; cycles by clock64.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) tc_mma_rate_kernel(int N, int lbo_a, int sbo_a, int a_off, int iters, int ksteps,
                                                          long long *cycles) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<float *>(smem_raw)[i] = 0.f;
    const uint32_t ncols = tmem_cols_for(N * ksteps);
    if (warp == 0) tmem_alloc(&tmem_slot, ncols);
    if (tid == 0) mbar_init(&mbar, 1);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        const uint32_t idesc = make_idesc_tf32(128, N);
        const uint32_t a0 = smem_u32(smem_raw) + (uint32_t)a_off, b0 = smem_u32(smem_raw) + 40 * 1024;
        const long long t0 = clock64();
        // ksteps = number of accumulators used round-robin (power of two): 1 = every MMA depends on the previous one
        const uint64_t ad0 = make_smem_desc(a0, lbo_a, sbo_a), ad1 = make_smem_desc(a0 + 2 * lbo_a, lbo_a, sbo_a);
        const uint64_t bd = make_smem_desc(b0, N * 16, 128);
#pragma unroll 4
        for (int i = 0; i < iters; i++) umma_tf32(tmem_slot + (uint32_t)((i & (ksteps - 1)) * N), (i & 1) ? ad1 : ad0, bd, idesc, 1);
        umma_commit(&mbar);
        mbar_wait(&mbar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0) cycles[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_slot, ncols);
}

size_t dwpw_tc_smem(int KC, int NP) { return sizeof(float) * 2 * ((size_t)KC * (TC_M + 1) + (size_t)KC * NP) + 1024; }

template <int KS, bool STRIP>
bool launch_dwpw_tc_ks(const ConvDev &p, const float *w_hi, const float *w_lo, int NP, cudaStream_t s) {
    static const int kc_max = getenv("ZB_TC_KC") ? atoi(getenv("ZB_TC_KC")) : 32;
    const int KC = p.K <= kc_max ? p.K : kc_max;               // K chunk resident in shared memory
    const int Kpad = (p.K + KC - 1) / KC * KC;                 // packed weights are zero-padded to this many rows
    const size_t smem = dwpw_tc_smem(KC, NP);
    if (smem > 220 * 1024) return false;
    auto kern = dwpw_tc_kernel<KS, STRIP>;
    static SmemOptIn opt_in;
    if (!opt_in.ensure(kern, smem)) return false;
    ZB_KNAME("dwpw_tc_kernel", KS, STRIP ? 1 : 0);
    kern<<<(unsigned)((p.M + TC_M - 1) / TC_M), 256, smem, s>>>(p, w_hi, w_lo, NP, KC, Kpad);
    return true;
}

}  // namespace

bool dwpw_tc_supported(const ConvDev &p, int NP) {
    if (!((p.kh == 3 && p.kw == 3) || (p.kh == 5 && p.kw == 5))) return false;
    if (p.K % 8 || p.K < 8 || p.K > 1024 || NP % 16 || NP < 16 || NP > 256) return false;
    return true;
}

bool launch_dwpw_tc(const ConvDev &p, const float *w_hi, const float *w_lo, int NP, cudaStream_t s) {
    if (!dwpw_tc_supported(p, NP)) return false;
    g_launch_count++;
    static const bool no_strip = getenv("ZB_TC_NO_STRIP") && atoi(getenv("ZB_TC_NO_STRIP")) != 0;
    if (p.kh == 5) return launch_dwpw_tc_ks<5, false>(p, w_hi, w_lo, NP, s);
    const bool strip = !no_strip && p.sh == 1 && p.sw == 1 && p.Wo % 4 == 0 && (p.Ho * p.Wo) % 4 == 0;
    return strip ? launch_dwpw_tc_ks<3, true>(p, w_hi, w_lo, NP, s) : launch_dwpw_tc_ks<3, false>(p, w_hi, w_lo, NP, s);
}

bool dwpw_ttc_supported(const ConvDev &p, int NP) {
    static const bool disabled = getenv("ZB_NO_TTC") && atoi(getenv("ZB_NO_TTC")) != 0;
    if (disabled) return false;
    if (p.kh != 3 || p.kw != 3 || p.sh != 1 || p.sw != 1 || p.pt != 1 || p.pl != 1) return false;
    // measured per layer against the SIMT thin kernel (same box, batch 1024): Cs 16 loses (1.17 vs 1.07 ms on the
    // 96x96x16 blocks), Cs 24 ties, Cs >= 32 wins (0.78 vs 0.92 ms on 48x48x32, 0.18 vs 0.23 ms on 32x32x36)
    static const int min_cs = getenv("ZB_TTC_MIN_CS") ? atoi(getenv("ZB_TTC_MIN_CS")) : 32;
    if (!(p.Cs_in == 16 || p.Cs_in == 24 || p.Cs_in == 32 || p.Cs_in == 40 || p.Cs_in == 48) || p.Cs_in < min_cs) return false;
    if (p.K != p.Cs_in || NP % 16 || NP < 16 || NP > 64 || p.Ns % 4 || p.Nstore != p.Ns || p.out_pix_stride != p.Ns) return false;
    if (p.Ho * p.Wo < 1024 || p.Wo < 32 || p.Ho % 4 || p.M % (p.Ho * p.Wo)) return false;
    if (p.epi.res && p.epi.res == p.in && p.epi.res_Cs != p.Cs_in) return false;
    return true;
}

bool launch_dwpw_ttc(const ConvDev &p, const float *w_hi, const float *w_lo, int NP, cudaStream_t s) {
    if (!dwpw_ttc_supported(p, NP)) return false;
    g_launch_count++;
    switch (p.Cs_in) {
        case 16: return launch_ttc_cs<16>(p, w_hi, w_lo, NP, s);
        case 24: return launch_ttc_cs<24>(p, w_hi, w_lo, NP, s);
        case 32: return launch_ttc_cs<32>(p, w_hi, w_lo, NP, s);
        case 40: return launch_ttc_cs<40>(p, w_hi, w_lo, NP, s);
        default: return launch_ttc_cs<48>(p, w_hi, w_lo, NP, s);
    }
}

// D[128,N] = A[128,K] * B[N,K]^T on tcgen05 (nsplit 1: raw TF32, 3: 3xTF32).  Device pointers.
bool launch_tc_mma_rate(int N, int lbo_a, int sbo_a, int a_off, int iters, int ksteps, int ctas, long long *cycles_dev, cudaStream_t s) {
    if (N % 8 || N < 8 || N > 256 || iters < 1 || ksteps < 1 || (ksteps & (ksteps - 1)) || N * ksteps > 512) return false;
    static SmemOptIn opt_in;
    if (!opt_in.ensure(tc_mma_rate_kernel, 48 * 1024)) return false;
    g_launch_count++;
    tc_mma_rate_kernel<<<ctas, 128, 48 * 1024, s>>>(N, lbo_a, sbo_a, a_off, iters, ksteps, cycles_dev);
    return true;
}

bool launch_tc_gemm_test(const float *A, const float *B, float *D, int N, int K, int nsplit, cudaStream_t s) {
    if (N % 16 || N < 16 || N > 256 || K % 8 || K < 8) return false;
    const size_t smem = sizeof(float) * 2 * ((size_t)K * TC_M + (size_t)K * N) + 1024;
    if (smem > 220 * 1024) return false;
    static SmemOptIn opt_in;
    if (!opt_in.ensure(tc_gemm_test_kernel, smem)) return false;
    g_launch_count++;
    tc_gemm_test_kernel<<<1, 128, smem, s>>>(A, B, D, N, K, nsplit);
    return true;
}

}  // namespace zb
