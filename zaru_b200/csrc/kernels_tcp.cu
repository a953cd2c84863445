// sm_100a PERSISTENT, WARP-SPECIALISED tile-block kernel: fused depthwise KSxKS (3x3 / 5x5, stride 1 / 2) -> pointwise
// blocks, the same operation and the same tile geometry as tcb_dwpw_kernel (kernels_tcb.cu), restructured so that nothing
// in a CTA waits for anything else:
//
//   warp 12 (TMA)      streams, for every (tile, 32-channel chunk) step, the input halo (one cp.async.bulk.tensor.3d, image and
//                      row dimensions merged) and the chunk's depthwise weights (one bulk copy) into a
//                      ring of input slots, and the pointwise weight chunks into a ring of weight slots (or once, if the
//                      whole layer's weights fit in shared memory);
//   warps 0-7 (dw)     convolve a landed chunk (sliding window: 4 adjacent outputs x one channel quad per thread) and write
//                      the TF32 hi / lo operand tile into a ring of A slots;
//   warp 13 (MMA)      one thread issues the 3xTF32 tcgen05.mma chain of a step as soon as its A slot is full, commits the
//                      slot back to the producers and, after a tile's last chunk, the accumulator to the epilogue;
//   warps 8-11 (epi)   read a finished accumulator (one of TWO in TMEM, so the MMAs of the next tile run meanwhile), stage it
//                      in shared memory and run the coalesced epilogue (bias, activation, residual, stores).
//
// One CTA per SM, each walking tiles blockIdx.x, blockIdx.x + gridDim.x, ...  All hand-offs are mbarriers (full / empty per
// ring slot); the only CTA-wide barriers are at start-up and tear-down.
//
// STATUS: parity-green, OFF by default (ZB_TCP=1 turns it on).  Measured on the face-mesh layers at batch 1024
// (profiles/README.md): 1.3-2x SLOWER than tcb_dwpw_kernel.  With one CTA per SM only 8 warps do the depthwise stage, and
// that stage - shared-memory loads and FMAs with little instruction-level parallelism - needs 16 to 32 warps per SM to
// hide its latencies; the hand-offs this design removes were not the bound.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "conv_common.cuh"
#include "kernels.h"
#include "tc_common.cuh"
#include "tc_epilogue.cuh"

namespace zb {
namespace {

using namespace tc;

constexpr int TCP_M = 128, TCP_CK = 32, TCP_KQC = 8, TCP_AROWS = TCP_M + 1;
constexpr int TCP_A_SLOT = 2 * TCP_KQC * TCP_AROWS * 4;        // floats per A slot (hi + lo)
constexpr int TCP_PRODUCERS = 256, TCP_EPI = 128, TCP_THREADS = 448;
constexpr int TCP_MAX_IN = 4, TCP_MAX_A = 2, TCP_MAX_B = 2;

struct TcpCfg {
    int TW, TH, tiles_x, tiles_y, vrows, WBOX, rows_max;
    int nin, na, nb;       // ring depths: input slots, A slots, weight slots (nb == nchunks: weights resident)
    int resident;          // pointwise weights of all chunks stay in shared memory
    int total_tiles;
};

__device__ __forceinline__ void tcp_tma_load_3d(void *smem_dst, const CUtensorMap *tmap, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}


// geometry of one tile (see kernels_tcb.cu): TH virtual rows (image, output row) x TW columns
struct TileGeo {
    int vr0, ox0, n_vr;
};
__device__ __forceinline__ TileGeo tile_geo(int tile, const TcpCfg &g) {
    TileGeo t;
    const int tile_x = tile % g.tiles_x, tile_y = tile / g.tiles_x;
    t.vr0 = tile_y * g.TH;
    t.ox0 = tile_x * g.TW;
    t.n_vr = min(g.TH, g.vrows - t.vr0);
    return t;
}

template <int KS, int S>
__global__ void __launch_bounds__(TCP_THREADS, 1) tcp_dwpw_kernel(const __grid_constant__ CUtensorMap tmap, const ConvDev p,
                                                                  const float *__restrict__ w_hi, const float *__restrict__ w_lo, int NP,
                                                                  int nchunks, const TcpCfg g) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    constexpr int TAPS = KS * KS, SPAN = 3 * S + KS;
    const int in_tile_floats = g.rows_max * g.WBOX * TCP_CK;
    const int in_slot_floats = in_tile_floats + (TAPS + 1) * TCP_CK;         // halo rows, then the chunk's dw weights + bias
    const int b_slot_floats = 2 * TCP_KQC * NP * 4;                          // hi, then lo
    float *s_in0 = reinterpret_cast<float *>(smem_raw);
    float *s_a0 = s_in0 + (size_t)g.nin * in_slot_floats;
    float *s_b0 = s_a0 + (size_t)g.na * TCP_A_SLOT;
    float *s_stage = s_b0 + (size_t)g.nb * b_slot_floats;                    // 128 x TCE_STRIDE
    __shared__ __align__(8) uint64_t in_full[TCP_MAX_IN], in_empty[TCP_MAX_IN], a_full[TCP_MAX_A], a_empty[TCP_MAX_A];
    __shared__ __align__(8) uint64_t b_full[32], b_empty[TCP_MAX_B], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) TceRow s_rowinfo[TCP_M];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t ncols = tmem_cols_for(2 * NP);
    const uint32_t b_bytes = (uint32_t)TCP_KQC * NP * 16;

    if (warp == 0) tmem_alloc(&tmem_slot, ncols);
    if (tid == 0) {
        for (int i = 0; i < g.nin; i++) mbar_init(&in_full[i], 1), mbar_init(&in_empty[i], TCP_PRODUCERS / 32);
        for (int i = 0; i < g.na; i++) mbar_init(&a_full[i], TCP_PRODUCERS / 32), mbar_init(&a_empty[i], 1);
        for (int i = 0; i < (g.resident ? nchunks : g.nb); i++) mbar_init(&b_full[i], 1);
        for (int i = 0; i < TCP_MAX_B; i++) mbar_init(&b_empty[i], 1);
        for (int i = 0; i < 2; i++) mbar_init(&acc_full[i], 1), mbar_init(&acc_empty[i], TCP_EPI / 32);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 12) {
        // ================================ TMA warp ================================
        if (g.resident && lane == 0) {
            for (int c = 0; c < nchunks; c++) {
                mbar_expect_tx(&b_full[c], 2 * b_bytes);
                bulk_copy_g2s(s_b0 + (size_t)c * b_slot_floats, w_hi + (size_t)c * TCP_KQC * NP * 4, b_bytes, &b_full[c]);
                bulk_copy_g2s(s_b0 + (size_t)c * b_slot_floats + TCP_KQC * NP * 4, w_lo + (size_t)c * TCP_KQC * NP * 4, b_bytes, &b_full[c]);
            }
        }
        int step = 0;
        for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
            const TileGeo t = tile_geo(tile, g);
            // ONE tensor copy per step: the (image, row) dimensions of the activation are merged into one (images are
            // contiguous and H == Ho * S), so the halo rows of a tile are one contiguous row range even when the tile
            // spans several images; rows that belong to a neighbouring image are masked by the producers
            const uint32_t in_bytes = (uint32_t)in_tile_floats * 4 + (uint32_t)(TAPS + 1) * TCP_CK * 4;
            for (int c = 0; c < nchunks; c++, step++) {
                const int slot = step % g.nin;
                mbar_wait(&in_empty[slot], ((step / g.nin) & 1) ^ 1);       // producers have drained this slot
                float *dst = s_in0 + (size_t)slot * in_slot_floats;
                if (lane == 0) {
                    mbar_expect_tx(&in_full[slot], in_bytes);
                    bulk_copy_g2s(dst + in_tile_floats, p.dw_c + (size_t)c * (TAPS + 1) * TCP_CK, (uint32_t)(TAPS + 1) * TCP_CK * 4, &in_full[slot]);
                }
                if (lane == 0) tcp_tma_load_3d(dst, &tmap, c * TCP_CK, t.ox0 * S - p.pl, t.vr0 * S - p.pt, &in_full[slot]);
                if (!g.resident && lane == 0) {                             // this step's pointwise weights
                    const int bs = step % g.nb;
                    mbar_wait(&b_empty[bs], ((step / g.nb) & 1) ^ 1);
                    mbar_expect_tx(&b_full[bs], 2 * b_bytes);
                    bulk_copy_g2s(s_b0 + (size_t)bs * b_slot_floats, w_hi + (size_t)c * TCP_KQC * NP * 4, b_bytes, &b_full[bs]);
                    bulk_copy_g2s(s_b0 + (size_t)bs * b_slot_floats + TCP_KQC * NP * 4, w_lo + (size_t)c * TCP_KQC * NP * 4, b_bytes, &b_full[bs]);
                }
                __syncwarp();
            }
        }
    } else if (warp == 13) {
        // ================================ MMA thread ================================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(TCP_M, NP);
            int step = 0, it = 0;
            for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, it++) {
                const int acc = it & 1;
                mbar_wait(&acc_empty[acc], ((it >> 1) & 1) ^ 1);             // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem + (uint32_t)acc * NP;
                uint32_t acc_flag = 0;
                for (int c = 0; c < nchunks; c++, step++) {
                    const int as = step % g.na;
                    const int bs = g.resident ? c : step % g.nb;
                    mbar_wait(&b_full[bs], g.resident ? 0 : ((step / g.nb) & 1));
                    mbar_wait(&a_full[as], (step / g.na) & 1);
                    tc_fence_after();
                    const float *sA = s_a0 + (size_t)as * TCP_A_SLOT, *sB = s_b0 + (size_t)bs * b_slot_floats;
                    const uint64_t ad_hi = make_smem_desc(smem_u32(sA), TCP_AROWS * 16, 128);
                    const uint64_t ad_lo = make_smem_desc(smem_u32(sA + TCP_KQC * TCP_AROWS * 4), TCP_AROWS * 16, 128);
                    const uint64_t bd_hi = make_smem_desc(smem_u32(sB), (uint32_t)NP * 16, 128);
                    const uint64_t bd_lo = make_smem_desc(smem_u32(sB + TCP_KQC * NP * 4), (uint32_t)NP * 16, 128);
#pragma unroll 1
                    for (int pass = 0; pass < 3; pass++) {                  // lo*hi, hi*lo, hi*hi
                        uint64_t ad = pass == 0 ? ad_lo : ad_hi;
                        uint64_t bd = pass == 1 ? bd_lo : bd_hi;
#pragma unroll
                        for (int j = 0; j < TCP_CK / 8; j++) {
                            umma_tf32(d_tmem, ad, bd, idesc, acc_flag);
                            acc_flag = 1;
                            ad += (uint64_t)(2 * TCP_AROWS);
                            bd += (uint64_t)(2 * NP);
                        }
                    }
                    umma_commit(&a_empty[as]);                              // A slot free once these MMAs have read it
                    if (!g.resident) umma_commit(&b_empty[bs]);
                }
                umma_commit(&acc_full[acc]);                                // accumulator complete -> epilogue
            }
        }
    } else if (warp < 8) {
        // ================================ depthwise producers ================================
        const int quad = tid & 7, strip = tid >> 3;
        const int strips_x = g.TW >> 2;
        const int pr = strip / strips_x, pcol = (strip - pr * strips_x) << 2;
        const int pm = pr * g.TW + pcol;
        int step = 0;
        for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
            const TileGeo t = tile_geo(tile, g);
            const bool p_valid = pr < t.n_vr && t.ox0 + pcol < p.Wo;
            const int origin = S * pr * g.WBOX + pcol * S;
            // taps whose input row lies outside this output row's own image contribute zero (the staged row belongs to
            // the neighbouring image there): bit ky of the mask = row ky of the window is inside the image
            const int p_vr = t.vr0 + pr;
            const int p_iy0 = (p_vr - (p_vr / p.Ho) * p.Ho) * S - p.pt;
            unsigned kymask = 0;
#pragma unroll
            for (int ky = 0; ky < KS; ky++) kymask |= (p_iy0 + ky >= 0 && p_iy0 + ky < p.H) ? (1u << ky) : 0u;
            for (int c = 0; c < nchunks; c++, step++) {
                const int slot = step % g.nin, as = step % g.na;
                mbar_wait(&in_full[slot], (step / g.nin) & 1);               // halo rows + dw weights have landed
                mbar_wait(&a_empty[as], ((step / g.na) & 1) ^ 1);            // the MMAs that read this A slot are done
                if (p_valid) {
                    const float *s_slot = s_in0 + (size_t)slot * in_slot_floats;
                    const float *s_in = s_slot + (size_t)origin * TCP_CK + quad * 4;
                    const float *s_w = s_slot + in_tile_floats;
                    float *sA_hi = s_a0 + (size_t)as * TCP_A_SLOT, *sA_lo = sA_hi + TCP_KQC * TCP_AROWS * 4;
                    const float4 bias = *reinterpret_cast<const float4 *>(s_w + TAPS * TCP_CK + quad * 4);
                    float4 v[4] = {bias, bias, bias, bias};
#pragma unroll
                    for (int ky = 0; ky < KS; ky++) {
                        if (!((kymask >> ky) & 1u)) continue;
                        float4 x[SPAN];
                        const float *rowp = s_in + (size_t)ky * g.WBOX * TCP_CK;
#pragma unroll
                        for (int j = 0; j < SPAN; j++) x[j] = *reinterpret_cast<const float4 *>(rowp + j * TCP_CK);
#pragma unroll
                        for (int kx = 0; kx < KS; kx++) {
                            const float4 wv = *reinterpret_cast<const float4 *>(s_w + (ky * KS + kx) * TCP_CK + quad * 4);
#pragma unroll
                            for (int i = 0; i < 4; i++) {
                                v[i].x = fmaf(x[i * S + kx].x, wv.x, v[i].x);
                                v[i].y = fmaf(x[i * S + kx].y, wv.y, v[i].y);
                                v[i].z = fmaf(x[i * S + kx].z, wv.z, v[i].z);
                                v[i].w = fmaf(x[i * S + kx].w, wv.w, v[i].w);
                            }
                        }
                    }
                    const int k = c * TCP_CK + quad * 4;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        if (k < p.K) act4(v[i], p.act_mid, k);
                        float4 hi, lo;
                        split_tf32_fast(v[i].x, hi.x, lo.x);
                        split_tf32_fast(v[i].y, hi.y, lo.y);
                        split_tf32_fast(v[i].z, hi.z, lo.z);
                        split_tf32_fast(v[i].w, hi.w, lo.w);
                        *reinterpret_cast<float4 *>(sA_hi + ((size_t)quad * TCP_AROWS + pm + i) * 4) = hi;
                        *reinterpret_cast<float4 *>(sA_lo + ((size_t)quad * TCP_AROWS + pm + i) * 4) = lo;
                    }
                }
                fence_async_smem();                                         // operand tile -> visible to the tensor core
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&a_full[as]);
                    mbar_arrive(&in_empty[slot]);
                }
            }
        }
    } else if (warp < 12) {
        // ================================ epilogue warps ================================
        const int etid = tid - 256;
        int it = 0;
        for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, it++) {
            const int acc = it & 1;
            const TileGeo t = tile_geo(tile, g);
            {   // row table of this tile
                const int er = etid / g.TW, ecol = etid - er * g.TW;
                const int evr = t.vr0 + er, ox = t.ox0 + ecol;
                TceRow ri;
                ri.out_off = -1, ri.res_off = 0;
                if (er < t.n_vr && ox < p.Wo) {
                    const int img = evr / p.Ho;
                    ri = tce_row(p, img, evr - img * p.Ho, ox);
                }
                s_rowinfo[etid] = ri;
            }
            mbar_wait(&acc_full[acc], (it >> 1) & 1);
            tc_fence_after();
            tc_epilogue_tile<TCP_EPI, 1>(p, tmem + (uint32_t)acc * NP, 0, NP, s_rowinfo, s_stage, etid);
            tc_fence_before();                                              // accumulator read: hand it back to the MMA thread
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, ncols);
}

// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn tcp_encode_fn() {
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

bool tcp_input_map(const ConvDev &p, int images, int WBOX, int BR, CUtensorMap *out) {
    EncodeTiledFn enc = tcp_encode_fn();
    if (!enc) return false;
    // (channel, x, merged image * H + y): images are contiguous, so the merged dimension has the row stride
    const cuuint64_t dims[3] = {(cuuint64_t)p.Cs_in, (cuuint64_t)p.W, (cuuint64_t)p.H * images};
    const cuuint64_t strides[2] = {(cuuint64_t)p.Cs_in * 4, (cuuint64_t)p.W * p.Cs_in * 4};
    const cuuint32_t box[3] = {(cuuint32_t)TCP_CK, (cuuint32_t)WBOX, (cuuint32_t)BR};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(p.in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

void tcp_geometry(const ConvDev &p, int KS, int S, TcpCfg &g) {
    const int images = p.M / (p.Ho * p.Wo);
    double best_cost = 1e30;
    static const int force_tw = getenv("ZB_TCB_TW") ? atoi(getenv("ZB_TCB_TW")) : 0;
    for (int TW = 4; TW <= (S == 2 ? 64 : 128); TW *= 2) {
        if (force_tw && TW != force_tw) continue;
        const int TH = TCP_M / TW;
        const int tiles_x = (p.Wo + TW - 1) / TW;
        const int WBOX = (TW - 1) * S + KS;
        if (WBOX > 256) continue;
        const double rows = (TH - 1) * S + KS;
        const double valid = (double)TH * p.Wo / tiles_x;
        const double cost = rows * WBOX / valid + 0.5 * TCP_M / valid;
        if (cost < best_cost) best_cost = cost, g.TW = TW, g.TH = TH, g.tiles_x = tiles_x, g.WBOX = WBOX;
    }
    g.vrows = images * p.Ho;
    g.tiles_y = (g.vrows + g.TH - 1) / g.TH;
    g.rows_max = (g.TH - 1) * S + KS;          // one contiguous row range per tile (merged image / row dimension)
    g.total_tiles = g.tiles_x * g.tiles_y;
}

size_t tcp_smem(const TcpCfg &g, int KS, int NP) {
    const size_t in_slot = ((size_t)g.rows_max * g.WBOX * TCP_CK + (size_t)(KS * KS + 1) * TCP_CK) * 4;
    return g.nin * in_slot + (size_t)g.na * TCP_A_SLOT * 4 + (size_t)g.nb * 2 * TCP_KQC * NP * 16 + (size_t)TCP_M * TCE_STRIDE * 4 + 1024;
}

template <int KS, int S>
bool launch_tcp_cfg(const ConvDev &p, const float *w_hi, const float *w_lo, int NP, cudaStream_t s) {
    TcpCfg g{};
    tcp_geometry(p, KS, S, g);
    if (g.TW == 0) return false;
    const int nchunks = (p.K + TCP_CK - 1) / TCP_CK;
    constexpr size_t BUDGET = 224 * 1024;
    // ring depths by shared-memory budget: prefer resident weights, then two A slots, then deeper input rings
    static const int force_res = getenv("ZB_TCP_RESIDENT") ? atoi(getenv("ZB_TCP_RESIDENT")) : -1;
    bool found = false;
    for (int pref = 0; pref < 2 && !found; pref++) {
        const bool resident = pref == 0;
        if (resident && (nchunks > 32 || force_res == 0)) continue;
        if (!resident && force_res == 1) continue;
        for (int na = 2; na >= 1 && !found; na--)
            for (int nin = 3; nin >= 1 && !found; nin--)
                for (int nb = resident ? nchunks : 2; nb >= (resident ? nchunks : 1) && !found; nb--) {
                    g.nin = nin, g.na = na, g.nb = nb, g.resident = resident;
                    if (resident && na == 1 && nin == 1) continue;      // a starved pipeline is worse than streaming weights
                    if (tcp_smem(g, KS, NP) <= BUDGET) found = true;
                }
    }
    if (!found) return false;
    const size_t smem = tcp_smem(g, KS, NP);
    const int images = p.M / (p.Ho * p.Wo);
    CUtensorMap tmap;
    if (!tcp_input_map(p, images, g.WBOX, g.rows_max, &tmap)) return false;
    auto kern = tcp_dwpw_kernel<KS, S>;
    static SmemOptIn opt_in;
    if (!opt_in.ensure(kern, smem)) return false;
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int grid = std::min(g.total_tiles, num_sms);
    ZB_KNAME("tcp_dwpw_kernel", KS, S);
    kern<<<(unsigned)grid, TCP_THREADS, smem, s>>>(tmap, p, w_hi, w_lo, NP, nchunks, g);
    return true;
}

}  // namespace

bool tcp_dwpw_supported(const ConvDev &p, int NP) {
    static const bool disabled = getenv("ZB_NO_TCP") && atoi(getenv("ZB_NO_TCP")) != 0;
    if (disabled || !p.dw_c) return false;
    if (!((p.kh == 3 && p.kw == 3) || (p.kh == 5 && p.kw == 5))) return false;
    if (!((p.sh == 1 && p.sw == 1) || (p.sh == 2 && p.sw == 2))) return false;
    if (p.K != p.Cs_in || p.K % 8 || p.K < 8 || p.K > 1024 || NP % 16 || NP < 16 || NP > 256) return false;
    if (p.M % (p.Ho * p.Wo) || p.pt < 0 || p.pl < 0) return false;
    if (((uintptr_t)p.in) % 16 || (p.in_img_stride % 4)) return false;
    // merged (image, row) staging: images back to back and an input height that is exactly stride x output height
    if (p.in_img_stride != (long long)p.H * p.W * p.Cs_in || p.H != p.Ho * p.sh) return false;
    return tcp_encode_fn() != nullptr;
}

bool launch_tcp_dwpw(const ConvDev &p, const float *w_hi, const float *w_lo, int NP, cudaStream_t s) {
    if (!tcp_dwpw_supported(p, NP)) return false;
    g_launch_count++;
    bool ok;
    if (p.kh == 3) ok = p.sh == 1 ? launch_tcp_cfg<3, 1>(p, w_hi, w_lo, NP, s) : launch_tcp_cfg<3, 2>(p, w_hi, w_lo, NP, s);
    else ok = p.sh == 1 ? launch_tcp_cfg<5, 1>(p, w_hi, w_lo, NP, s) : launch_tcp_cfg<5, 2>(p, w_hi, w_lo, NP, s);
    if (!ok) g_launch_count--;
    return ok;
}

}  // namespace zb
